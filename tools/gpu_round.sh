#!/bin/bash
# one GPU call: f3 + parity tests, DRAM traffic of the front-end kernel, bench line, launch list, sweeps (outputs under gpurun_out/)
(python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2_gpu_tests_3.log 2>&1; cat gpurun_out/r2_gpu_tests_3.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:vdl2_frontend -c 1 python tools/perf_probe.py 1024 4194000 1 1 bursts 2>&1 | grep -E "dram__|gpu__time" | tee gpurun_out/r2_traffic_v16.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_2.json 2> gpurun_out/r2_bench_2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity_checked']['ok'], d['link']['drain_frames_ms_host'], d['link']['avlc'])"
tail -2 gpurun_out/r2_bench_2.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vdl2_ -c 80 --csv --log-file gpurun_out/launches_r2_v16.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
grep -c vdl2_ gpurun_out/launches_r2_v16.csv
python tools/sweep.py > gpurun_out/sweep_r2.log 2>&1; tail -28 gpurun_out/sweep_r2.log | cut -c1-330
