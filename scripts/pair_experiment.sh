#!/bin/bash
for pair in 0 1; do
  EEM_TF32_PAIR=$pair timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PAIR $pair mvsec:', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['family_ms_per_step'].items()})"
  EEM_TF32_PAIR=$pair timeout 300 python bench.py --workload hrem_dt1 --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PAIR $pair hrem :', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['family_ms_per_step'].items()})"
done
EEM_TF32_PAIR=1 timeout 300 python -m pytest tests/test_gpu_corr.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -2
