#!/bin/bash
# Tensor-memory columns per lane (and the row-ahead variant) re-measured for the K of the fixture models that have alternatives.
cd "$(dirname "$0")/.."
run() { python tools/sweep_geometry.py --model $1 --sequences 100000 --steps 4 --geometries default "${@:2}" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['model'], d['geometry'], d['chosen']['columns_per_lane'], d['chosen']['tensor_columns_per_lane'], d['gcups'], d['mismatches'])"; }
run 600.hmm 32,20,8 32,20,16 32,20,16,0,1
run 800.hmm 32,28,16 32,28,16,0,1 32,28,24,0,1
run 1001.hmm 32,32,16 32,32,16,0,1 32,32,24 
run 1100.hmm 32,36,16,0,1 32,36,24
run 1509.hmm 32,48,16 32,48,24 32,48,24,0,1
run 1600.hmm 32,52,16 32,52,24 32,52,24,0,1
run 1799.hmm 32,60,16,0,1 32,60,24,0,1
