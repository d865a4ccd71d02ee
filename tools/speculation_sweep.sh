#!/bin/bash
# Warp-per-sequence kernel: speculative rows (whole sequences / blocks of 64 rows) against exact rows, per model length.
cd "$(dirname "$0")/.."
for m in "$@"; do for s in whole blocks none; do
  MSV_CUDA_SPECULATION=$s python tools/sweep_geometry.py --model $m --sequences 100000 --steps 4 --geometries default | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$m', '$s', d['chosen']['columns_per_lane'], d['chosen']['tensor_columns_per_lane'], d['plan'], d['gcups'], d['mismatches'])"
done; done
