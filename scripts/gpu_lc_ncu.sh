#!/bin/bash
mkdir -p gpurun_out/r02
for s in 80x96 40x48; do
BENCH_LC_ONLY=mvsec BENCH_LC_SHAPE=$s timeout 300 ncu --set full --clock-control none --import-source on -k regex:local_corr_tf32_kernel -s 3 -c 1 -o gpurun_out/r02/prof_local_corr_tf32_$s -f python scripts/bench_local_corr.py > /dev/null 2>&1
done
ls -la gpurun_out/r02 | tail -3
