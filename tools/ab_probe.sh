#!/bin/bash
# same-box A/B: run the noise probe and the bench for each library variant given as argument
for lib in "$@"; do
  echo "== $lib"
  VDL2_LIB=$PWD/vdlm2dec_b200/$lib python tools/perf_probe.py 1024 2097152 4 2>&1 | tail -2 | head -1
  VDL2_LIB=$PWD/vdlm2dec_b200/$lib python bench.py --no-cpu --no-e2e --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['config']['blocks_decoded_per_step'])"
done
