#!/bin/bash
# warp-specialised lookup inside the step bench (MVSEC headline + HREM workloads), against the one-batch kernel
mkdir -p gpurun_out/r02c
for cfg in ${WS_CFGS:-0 3x1x8x5}; do
  EEM_LOOKUP_PACKED_WS=$cfg python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02c/bench_ws_$cfg.json
done
python - <<'PY'
import json, os
for n in os.environ.get("WS_CFGS", "0 3x1x8x5").split():
    d = json.loads(open(f"gpurun_out/r02c/bench_ws_{n}.json").read())
    print(n, round(d["value"]), round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["roofline"]["family_ms_per_step"].items()},
          {k: (round(w["value"], 1), round(w["ms_per_step"], 4), round(w["roofline"]["family_ms_per_step"]["corr_lookup"], 4)) for k, w in d["workloads"].items()})
PY
