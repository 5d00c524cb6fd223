#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_packed.csv python scripts/bench_packed.py > gpurun_out/ncu_packed_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr_lookup_packed -s 25 -c 1 -o gpurun_out/prof_lookup_packed python scripts/bench_packed.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr_tf32_pair_kernel -s 25 -c 1 -o gpurun_out/prof_pair_packed python scripts/bench_packed.py > /dev/null 2>&1
ls -la gpurun_out | tail -8
