#!/bin/bash
mkdir -p gpurun_out/r02c
timeout 900 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_corr.py -x -q -m gpu -k "not full_size" 2>&1 | tail -4
python bench.py --workloads none --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02c/bench_f.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c/bench_f.json").read())
print(round(d["value"]), d["ms_per_step"], d["roofline"]["family_ms_per_step"], d["roofline"]["avg_launch_ms"])
PY
