#!/bin/bash
mkdir -p gpurun_out
EEM_LOOKUP_PACKED_PIPE=0 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_packed_kernel -s 5 -c 1 -o gpurun_out/prof_lookup_packed_np python scripts/bench_packed.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr_lookup_packed_pipe -s 5 -c 1 -o gpurun_out/prof_lookup_packed_pipe python scripts/bench_packed.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pool_pyramid_packed -s 5 -c 1 -o gpurun_out/prof_pool_packed python scripts/bench_packed.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
