#!/bin/bash
# Eight lanes per sequence with speculative rows beyond K = 56 (LENG 500 .. 700) against the default (warp-per-sequence) plan.
cd "$(dirname "$0")/.."
for n in 100000 1000000; do
python tools/sweep_geometry.py --model 500.hmm --sequences $n --steps 4 --geometries default 8,64
python tools/sweep_geometry.py --model 600.hmm --sequences $n --steps 4 --geometries default 8,76
python tools/sweep_geometry.py --model 700.hmm --sequences $n --steps 4 --geometries default 8,88
done | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['model'], d['sequences'], d['geometry'], d['plan'], d['gcups'], d['mismatches'])"
