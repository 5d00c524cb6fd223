#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_corr.py -m gpu -x -q -k "packed" 2>&1 | tail -8
timeout 300 python scripts/bench_packed.py 2>&1 | tee gpurun_out/bench_packed.log
EEM_LOOKUP_PACKED_PIPE=0 timeout 300 python scripts/bench_packed.py 2>&1 | tee gpurun_out/bench_packed_nopipe.log
