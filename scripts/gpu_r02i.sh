#!/bin/bash
# persistent lookup on fewer SMs: does leaving SMs to the other graph branches shorten the step?
for n in 148 144 132 120; do
  EEM_LOOKUP_WS_CTAS=$n python bench.py --workloads none --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n CTAs', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['family_ms_per_step']['corr_lookup'],4))"
done
