#!/bin/bash
# cluster-resident normalisation: voxel tests, then the bench with it
mkdir -p gpurun_out/r02c
timeout 900 python -m pytest tests/test_gpu_voxel.py -x -q -m gpu -k "normal or golden or oracle_parity or ragged" 2>&1 | tail -4
python bench.py --workloads none --no-cpu-baseline --no-e2e --local-corr tf32 2>/dev/null | tail -1 > gpurun_out/r02c/bench_normcluster.json
python - <<'PY'
import json
for n in ("normcluster",):
    d = json.loads(open(f"gpurun_out/r02c/bench_{n}.json").read())
    print(n, round(d["value"]), d["ms_per_step"], d["roofline"]["family_ms_per_step"], d["roofline"]["avg_launch_ms"])
PY
