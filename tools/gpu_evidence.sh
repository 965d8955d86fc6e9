#!/bin/bash
# evidence for the kernel as built: whole GPU tier, bench lines (config 3 default, config 2, reference arm), launch list
# usage: tools/gpu_evidence.sh <tag>     (files land in gpurun_out/<tag>_*)
tag=${1:-r2_v18}
(python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/${tag}_gpu_tests.log 2>&1; cat gpurun_out/${tag}_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
python bench.py --config 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; head -c 400 gpurun_out/${tag}_bench_cfg2.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vdl2_ -c 80 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches_${tag}.csv
