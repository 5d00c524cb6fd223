#!/bin/bash
# round 2, third session: leaner packed-lookup gather, small-map local correlation, tf32 local correlation in the bench
mkdir -p gpurun_out/r02c
timeout 900 python -m pytest tests/test_gpu_corr.py tests/test_gpu_warp.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -6
timeout 200 python scripts/bench_local_corr.py 2>&1 | tail -14
EEM_LC_SMALL=0 BENCH_LC_ONLY=mvsec timeout 100 python scripts/bench_local_corr.py 2>&1 | grep -E "5x6|10x12"
python bench.py --workloads none --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02c/bench_fp32lc.json
python bench.py --workloads none --no-cpu-baseline --no-e2e --local-corr tf32 2>/dev/null | tail -1 > gpurun_out/r02c/bench_tf32lc.json
python - <<'PY'
import json
for n in ("fp32lc", "tf32lc"):
    d = json.loads(open(f"gpurun_out/r02c/bench_{n}.json").read())
    print(n, round(d["value"]), d["ms_per_step"], d["roofline"]["family_ms_per_step"], d["roofline"]["avg_launch_ms"])
PY
