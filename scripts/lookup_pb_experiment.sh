#!/bin/bash
for pb in 32 16; do
  EEM_LOOKUP_PB=$pb timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PB $pb:', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['family_ms_per_step'].items()}, 'lookup us', round(d['roofline']['avg_launch_ms']*1e3,2), 'frac', round(d['roofline']['frac'],3))"
done
