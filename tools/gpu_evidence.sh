#!/bin/bash
# evidence for the kernel as built: whole GPU tier, ncu capture + launch list, bench lines (config 3 default, reference arm, config 2,
# config 5), randomized parity sweep through the C ABI, shape sweep, channeliser probes
# usage: tools/gpu_evidence.sh <tag> [quick]     (files land in gpurun_out/<tag>_*)
tag=${1:-r2_v21}
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/${tag}_gpu_tests.log 2>&1; cat gpurun_out/${tag}_gpu_tests.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:vdl2_frontend -c 1 -o gpurun_out/${tag} python tools/perf_probe.py 1024 4194000 1 1 bursts > gpurun_out/ncu_${tag}.log 2>&1; tail -1 gpurun_out/ncu_${tag}.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-330 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-300 gpurun_out/${tag}_bench_ref.json
timeout 200 python bench.py --config 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; cut -c1-300 gpurun_out/${tag}_bench_cfg2.json; echo
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg5.json 2> gpurun_out/${tag}_bench_cfg5.err; cut -c1-300 gpurun_out/${tag}_bench_cfg5.json; echo
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vdl2_ -c 80 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches_${tag}.csv | cut -c1-300
if [ "$2" != "quick" ]; then
timeout 300 python tools/fuzz_gpu.py 801 384 --json gpurun_out/fuzz_gpu_${tag}_seed801.json | tail -1 | cut -c1-600
timeout 300 python tools/fuzz_gpu.py 901 384 lowsnr --json gpurun_out/fuzz_gpu_${tag}_seed901_lowsnr.json | tail -1 | cut -c1-600
timeout 500 python tools/sweep.py > gpurun_out/sweep_${tag}.log 2>&1; mv gpurun_out/sweep_r2.json gpurun_out/sweep_${tag}.json; tail -3 gpurun_out/sweep_${tag}.log | cut -c1-300
timeout 100 python tools/f3_probe.py 2>&1 | tail -1; timeout 100 python tools/f3_probe.py 128 8 2>&1 | tail -1
fi
