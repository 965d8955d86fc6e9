#!/bin/bash
# one GPU round for chain work: per-kind chain statistics (debug build), parity tests, same-box probes, config 2 bench
out=${1:-gpurun_out/chain_round.txt}
{
export VDL2_PRE_STATS=1
for a in "8 4194000 3 8 bursts" "1024 4194000 3 1 bursts"; do echo "== $a"; VDL2_LIB=$PWD/vdlm2dec_b200/libvdl2gpu_cs.so python tools/perf_probe.py $a 2>&1 | grep -E "^rep 2|chain|prepass" | tail -9; done
unset VDL2_PRE_STATS
python -m pytest tests/test_gpu_parity.py tests/test_gpu_link.py -m gpu -x -q 2>&1 | tail -3
bash tools/ab_probe2.sh libvdl2gpu.so
python bench.py --config 2 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('cfg2',d['value'],d['ms_per_step'],d['parity_checked']['ok'],d['roofline']['kernel_ms_isolated'])"
} > $out 2>&1
cat $out
