#!/bin/bash
# warp-specialised persistent packed lookup: timing per configuration (stages x CTAs per SM x producer warps)
for cfg in ${WS_CFGS:-3x1x8 3x1x12 2x2x8 4x1x8 2x1x8}; do
  echo -n "EEM_LOOKUP_PACKED_WS=$cfg  "
  EEM_LOOKUP_PACKED_WS=$cfg ABLATE=0 timeout 120 python scripts/lookup_ablation.py 2>&1 | tail -1
done
