#!/bin/bash
# What the driver runs at round end, in one call: GPU tests, smoke, both bench arms.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_final.log
cat gpurun_out/pytest_gpu_final.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -2 gpurun_out/bench_r02.err
