#!/bin/bash
# warp-specialised persistent packed lookup: parity, then timing per configuration (stages x CTAs/SM x geometry warps x copy warps)
timeout 600 python -m pytest tests/test_gpu_corr.py -x -q -m gpu -k "packed or f16" 2>&1 | tail -2
for cfg in ${WS_CFGS:-3x1x8x10 3x1x4x5 3x1x4x10 3x1x8x5 4x1x8x10 2x2x4x5 0}; do
  echo -n "EEM_LOOKUP_PACKED_WS=$cfg  "
  EEM_LOOKUP_PACKED_WS=$cfg ABLATE=0 timeout 120 python scripts/lookup_ablation.py 2>&1 | tail -1
done
