cd "$(dirname "$0")/.."
run() { python tools/sweep_geometry.py --model $1 --sequences $2 --steps 4 --slots ${@:4} --geometries $3 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['model'], d['sequences'], d['geometry'], d['slots'], d['gcups'], d['mismatches'])"; }
run 200.hmm 1000000 "4,52 8,26" 0
run 200.hmm 100000 "4,52 8,26" 0 64 96
run 100.hmm 1000000 "4,26 8,14" 0
run 100.hmm 100000 "4,26 8,14" 0 64 96
run 300.hmm 100000 "8,38" 0 48 64 80
run 400.hmm 100000 "8,52" 0 48 64
