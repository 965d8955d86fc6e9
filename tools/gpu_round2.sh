#!/bin/bash
# TMA L2-promotion A/B (DRAM bytes of the front-end kernel), packed-drain test + timing
for p in 0 1 2 3; do echo "promo $p"; VDL2_TMA_PROMO=$p ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:vdl2_frontend -c 1 python tools/perf_probe.py 1024 4194000 1 1 bursts 2>&1 | grep -E "dram__|gpu__time"; done
for p in 0 3; do echo "promo $p timing"; VDL2_TMA_PROMO=$p VDL2_OVERLAP=1 python tools/perf_probe.py 1024 4194000 4 1 bursts 2>&1 | grep -E "^rep 3|^overlap"; done
python -m pytest tests/test_gpu_link.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['link']['drain_frames_ms_host'], d['link']['avlc']['kernel_ms'])"
