#!/bin/bash
# same-box A/B over library variants for chain work: config 2 shape (8 channels of one stream) and config 3 shape, isolated and overlapped launches
for lib in "$@"; do
  c2=$(VDL2_OVERLAP=1 VDL2_LIB=$PWD/vdlm2dec_b200/$lib python tools/perf_probe.py 8 4194000 5 8 bursts 2>&1 | grep -E "^rep 4|^overlap" | awk '{print $1 $2, $3}' | tr '\n' ' ')
  c3=$(VDL2_OVERLAP=1 VDL2_LIB=$PWD/vdlm2dec_b200/$lib python tools/perf_probe.py 1024 4194000 4 1 bursts 2>&1 | grep -E "^rep 3|^overlap" | awk '{print $1 $2, $3}' | tr '\n' ' ')
  echo "$lib cfg2: $c2 | cfg3: $c3"
done
