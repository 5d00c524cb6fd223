#!/bin/bash
for mb in 0 120 60 40 30 20; do
  EEM_VOXEL_GROUP_MB=$mb timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('group MB $mb:', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['family_ms_per_step'].items()}, d['gpu_launches']//20)"
done
