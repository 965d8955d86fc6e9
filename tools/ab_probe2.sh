#!/bin/bash
# same-box A/B over library variants: noise probe (idle-mode worst case), burst probe (bench-like), overlapped burst launches
for lib in "$@"; do
  n=$(VDL2_LIB=$PWD/vdlm2dec_b200/$lib python tools/perf_probe.py 1024 2097152 4 2>&1 | grep "^rep 3" | awk '{print $3}')
  b=$(VDL2_LIB=$PWD/vdlm2dec_b200/$lib python tools/perf_probe.py 1024 4194000 4 1 bursts 2>&1 | grep "^rep 3" | awk '{print $3}')
  o=$(VDL2_OVERLAP=1 VDL2_LIB=$PWD/vdlm2dec_b200/$lib python tools/perf_probe.py 1024 4194000 4 1 bursts 2>&1 | grep "^overlap" | awk '{print $2}')
  echo "$lib noise_ms=$n bursts_ms=$b overlap_ms=$o"
done
