#!/bin/bash
# A/B of the L2 set-aside for the per-warp scratch (VDL2_L2_PERSIST_MB) and of the input's eviction class: time per step and DRAM
# bytes per launch (ncu, 3 launches each).  usage: tools/ab_l2persist.sh "<mb list>" [lib ...]
mbs=${1:-"0 64"}; shift
libs=${@:-libvdl2gpu.so}
for lib in $libs; do for mb in $mbs; do
  export VDL2_L2_PERSIST_MB=$mb VDL2_LIB=$PWD/vdlm2dec_b200/$lib
  t=$(VDL2_OVERLAP=1 python tools/perf_probe.py 1024 4194000 4 1 bursts 2>&1 | grep -E "^overlap" | awk '{print $2}')
  n=$(python tools/perf_probe.py 1024 2097152 4 2>&1 | grep "^rep 3" | awk '{print $3}')
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:vdl2_frontend -c 3 --csv --log-file gpurun_out/l2p.csv python tools/perf_probe.py 1024 4194000 3 1 bursts > /dev/null 2>&1
  d=$(grep -E "dram__bytes" gpurun_out/l2p.csv | awk -F'","' '{printf "%.3f ", $NF/1e9}' | tr -d '"')
  echo "$lib persist_mb=$mb overlap_ms=$t noise_ms=$n dram GB (read write x3): $d"
done; done
