#!/bin/bash
# same-box A/B of the 8-bit mixers: int8 tensor-core (default) vs IDP.4A (VDL2_DP4A_MIX=1); noise probe and burst probe
for mode in mma dp4a; do
  if [ $mode = dp4a ]; then export VDL2_DP4A_MIX=1; else unset VDL2_DP4A_MIX; fi
  n=$(python tools/perf_probe.py 1024 2097152 4 2>&1 | grep "^rep 3" | awk '{print $3}')
  b=$(python tools/perf_probe.py 1024 4194000 4 1 bursts 2>&1 | grep "^rep 3" | awk '{print $3}')
  o=$(VDL2_OVERLAP=1 python tools/perf_probe.py 1024 4194000 4 1 bursts 2>&1 | grep "^overlap" )
  echo "$mode noise_ms=$n bursts_ms=$b $o"
done
