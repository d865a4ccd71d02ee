#!/bin/bash
# Lane-group plans for the short models: lanes per sequence (4 / 8) x sequences in flight per CTA, 100 k and 1 M sequences.
cd "$(dirname "$0")/.."
for n in 100000 1000000; do
  python tools/sweep_geometry.py --model 100.hmm --sequences $n --steps 3 --geometries default 4,28 8,16 --slots 0 192 128 96 64 48
  python tools/sweep_geometry.py --model 200.hmm --sequences $n --steps 3 --geometries default 4,52 8,28 --slots 0 128 96 64 48
  python tools/sweep_geometry.py --model 300.hmm --sequences $n --steps 3 --geometries default 8,40 --slots 0 96 64 48 32
  python tools/sweep_geometry.py --model 400.hmm --sequences $n --steps 3 --geometries default 8,52 --slots 0 80 64 48 32
done
