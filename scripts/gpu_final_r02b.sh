#!/bin/bash
# Last validation of round 2 (what the driver runs at round end): full GPU tests, smoke, the bench arm, the reference arm.
mkdir -p gpurun_out/r02d
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | cut -c1-250 > gpurun_out/r02d/pytest_gpu_final.log
cat gpurun_out/r02d/pytest_gpu_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 170 python bench.py > gpurun_out/r02d/bench_r02.json 2> gpurun_out/r02d/bench_r02.err
tail -c 300 gpurun_out/r02d/bench_r02.err; head -c 400 gpurun_out/r02d/bench_r02.json
timeout 100 python bench.py --impl reference > gpurun_out/r02d/bench_ref_r02.json 2> gpurun_out/r02d/bench_ref_r02.err
head -c 300 gpurun_out/r02d/bench_ref_r02.json
