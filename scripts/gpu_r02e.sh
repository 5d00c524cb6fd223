#!/bin/bash
# row-walking multi-map resize: parity tests, then the bench with and without it
mkdir -p gpurun_out/r02c
timeout 900 python -m pytest tests/test_gpu_warp.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -4
python bench.py --workloads none --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02c/bench_rows.json
EEM_RESIZE_ROWS=0 python bench.py --workloads none --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02c/bench_norows.json
python - <<'PY'
import json
for n in ("rows", "norows"):
    d = json.loads(open(f"gpurun_out/r02c/bench_{n}.json").read())
    print(n, round(d["value"]), d["ms_per_step"], d["roofline"]["family_ms_per_step"], d["roofline"]["avg_launch_ms"])
PY
