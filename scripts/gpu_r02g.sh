#!/bin/bash
# one normalisation launch over all windows + vectorised pooling pre-pass: parity, then the bench
mkdir -p gpurun_out/r02c
timeout 900 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_corr.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workloads none --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02c/bench_g.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c/bench_g.json").read())
print(round(d["value"]), d["ms_per_step"], d["roofline"]["family_ms_per_step"], d["roofline"]["avg_launch_ms"])
PY
