#!/bin/bash
# Lane-group plans of the short models on the config-3 database (100 k sequences): sequences in flight per full CTA (--slots)
# x the longest sequences on a few fast CTAs ("ctas,warps,rows" = MSV_CUDA_FAST_CTAS; "off" = the slot cut of round 1).
cd "$(dirname "$0")/.."
python tools/sweep_geometry.py --model 100.hmm --sequences 100000 --steps 5 --slots 0 --fast off 32,8,700 40,8,700 48,8,700 40,8,600 56,8,500 64,8,500 48,12,600 64,12,500
python tools/sweep_geometry.py --model 200.hmm --sequences 100000 --steps 5 --slots 0 --fast off 24,8,800 32,8,800 40,8,700 48,8,600 40,12,700
python tools/sweep_geometry.py --model 300.hmm --sequences 100000 --steps 5 --slots 0 --fast off 8,8,1500 16,8,1200 24,8,1000 32,8,900
python tools/sweep_geometry.py --model 400.hmm --sequences 100000 --steps 5 --slots 0 --fast off 8,8,1500 16,8,1200 24,8,1000
