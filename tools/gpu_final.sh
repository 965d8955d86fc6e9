#!/bin/bash
# final evidence of the round: whole GPU tier, consolidated fuzz, ncu capture + traffic, launch list, bench lines, sweeps
(python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r2_gpu_tests_final.log 2>&1; cat gpurun_out/r2_gpu_tests_final.log
python tools/fuzz_gpu.py 401 512 --json gpurun_out/fuzz_gpu_401.json | tail -1 | cut -c1-700
python tools/fuzz_gpu.py 501 512 lowsnr --json gpurun_out/fuzz_gpu_501_lowsnr.json | tail -1 | cut -c1-700
ncu --set full --clock-control none --import-source on -k regex:vdl2_frontend -c 1 -o gpurun_out/r2_v17 python tools/perf_probe.py 1024 4194000 1 1 bursts > gpurun_out/ncu_v17.log 2>&1; tail -1 gpurun_out/ncu_v17.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 400 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; head -c 700 gpurun_out/r2_bench_ref.json; tail -2 gpurun_out/r2_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vdl2_ -c 80 --csv --log-file gpurun_out/launches_r2_v17.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python tools/f3_probe.py 2>&1 | tail -1; python tools/f3_probe.py 128 8 2>&1 | tail -1
