#!/bin/bash
# burst probe only, with the environment given in $ENVV
for lib in "$@"; do
  b=$(VDL2_LIB=$PWD/vdlm2dec_b200/$lib env $ENVV python tools/perf_probe.py 1024 4194000 4 1 bursts 2>&1 | grep "^rep 3" | awk '{print $3}')
  echo "$lib [$ENVV] bursts_ms=$b"
done
