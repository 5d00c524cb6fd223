#!/bin/bash
# tcgen05 local correlation: parity tests, graph-timed comparison, debug-bit ablations and (LC_NCU=1) one ncu capture
timeout 300 python -m pytest tests/test_gpu_warp.py -x -q -m gpu -k "local_corr" 2>&1 | tail -4
timeout 200 python scripts/bench_local_corr.py 2>&1 | tail -14
for d in ${LC_DEBUG_BITS:-8 15}; do echo "EEM_LC_DEBUG=$d"; EEM_LC_DEBUG=$d BENCH_LC_ONLY=mvsec timeout 100 python scripts/bench_local_corr.py 2>&1 | grep -E "80x96|40x48"; done
if [ -n "$LC_NCU" ]; then
  mkdir -p gpurun_out/r02
  BENCH_LC_ONLY=mvsec timeout 300 ncu --set full --clock-control none --import-source on -k regex:local_corr_tf32_kernel -s 12 -c 1 -o gpurun_out/r02/prof_local_corr_tf32_kernel -f python scripts/bench_local_corr.py > /dev/null 2>&1
fi
