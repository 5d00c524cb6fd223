#!/bin/bash
# one ncu --set full capture of the packed lookup inside the bench command; raw page + per-line stall samples
mkdir -p gpurun_out/r02c
B="python bench.py --steps 2 --warmup 1 --workloads none --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:corr_lookup_packed_kernel -s 6 -c 1 -o gpurun_out/r02c/prof_lookup_packed_v2 -f $B > /dev/null 2>&1
ncu -i gpurun_out/r02c/prof_lookup_packed_v2.ncu-rep --page raw --csv > gpurun_out/r02c/prof_lookup_packed_v2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02c/prof_lookup_packed_v2.ncu-rep --page source --csv > gpurun_out/r02c/prof_lookup_packed_v2_source.csv 2>/dev/null
ls -la gpurun_out/r02c
