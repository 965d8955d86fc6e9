#!/bin/bash
# where the time of the small-N shapes goes: 8 channels of one stream, noise only (every tile idle, speculation always right) against bursts
export VDL2_PRE_STATS=1
for args in "8 4194000 4 8" "8 4194000 4 8 bursts" "8 8388000 4 8 bursts" "64 4194000 4 8 bursts" "64 4194000 4 1 bursts" "1024 4194000 4 1 bursts"; do
  echo "== $args"; python tools/perf_probe.py $args 2>&1 | grep -E "^rep [23]|bursts placed|blocks|prepass" | tail -5
done
