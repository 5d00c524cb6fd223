#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_packed.log
cat gpurun_out/pytest_packed.log
timeout 300 python scripts/bench_packed.py 2>&1 | tee gpurun_out/bench_packed.log
