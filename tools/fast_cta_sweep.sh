#!/bin/bash
# Lane-group plans of the short models on the config-3 database (100 k sequences): sequences in flight per full CTA (--slots)
# x the longest sequences on a few fast CTAs ("ctas,warps,rows" = MSV_CUDA_FAST_CTAS; "off" = the slot cut of round 1).
cd "$(dirname "$0")/.."
python tools/sweep_geometry.py --model 100.hmm --sequences 100000 --steps 5 --slots 64 96 128 --fast off 2,8,1900 4,8,1500 8,8,1200 4,4,1900
python tools/sweep_geometry.py --model 200.hmm --sequences 100000 --steps 5 --slots 64 96 128 --fast off 2,8,1900 4,8,1500 8,8,1200 4,4,1900
python tools/sweep_geometry.py --model 300.hmm --sequences 100000 --steps 5 --slots 64 80 96 --fast off 1,8,2200 2,8,1900 4,8,1500 2,4,1900
python tools/sweep_geometry.py --model 400.hmm --sequences 100000 --steps 5 --slots 64 80 --fast off 1,8,2400 2,8,2000 4,8,1600 2,4,2000
