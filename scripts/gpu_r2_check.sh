#!/bin/bash
# Round-2 GPU check: parity tests, then the 1-GPU bench line.  Run under gpurun from the repo root.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
tail -c 6000 gpurun_out/bench_r2.json
tail -5 gpurun_out/bench_r2.err
