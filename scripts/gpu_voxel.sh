#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_voxel.log
cat gpurun_out/pytest_voxel.log | cut -c1-300
timeout 600 python scripts/bench_voxel.py 2>&1 | tee gpurun_out/bench_voxel.log
