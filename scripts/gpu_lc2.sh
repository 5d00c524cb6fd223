#!/bin/bash
# tcgen05 local correlation after the sector-aligned f2 boxes / contiguous job ranges; lookup retest; bench with both
mkdir -p gpurun_out/r02c
timeout 600 python -m pytest tests/test_gpu_corr.py tests/test_gpu_warp.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python scripts/bench_local_corr.py 2>&1 | tail -14
for d in ${LC_DEBUG_BITS:-8 15}; do echo "EEM_LC_DEBUG=$d"; EEM_LC_DEBUG=$d BENCH_LC_ONLY=mvsec timeout 100 python scripts/bench_local_corr.py 2>&1 | grep -E "80x96|40x48"; done
python bench.py --workloads none --no-cpu-baseline --no-e2e --local-corr tf32 2>/dev/null | tail -1 > gpurun_out/r02c/bench_tf32lc_b.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c/bench_tf32lc_b.json").read())
print(round(d["value"]), d["ms_per_step"], d["roofline"]["family_ms_per_step"], d["roofline"]["avg_launch_ms"])
PY
