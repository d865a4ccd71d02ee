#!/bin/bash
# Lane-group plans of the short models on the config-3 database (100 k sequences): sequences in flight per full CTA (--slots)
# x the longest sequences on a few fast CTAs ("ctas,warps,rows" = MSV_CUDA_FAST_CTAS; "off" = the slot cut of round 1).
cd "$(dirname "$0")/.."
python tools/sweep_geometry.py --model 100.hmm --sequences 100000 --steps 5 --slots 0 --fast off auto 8,8,1200 12,8,1000 16,8,1000 24,8,800 32,8,700 40,8,700 48,8,600
python tools/sweep_geometry.py --model 100.hmm --sequences 100000 --steps 5 --slots 128 --fast 8,8,1200 12,8,1000 16,8,1000 24,8,800
python tools/sweep_geometry.py --model 200.hmm --sequences 100000 --steps 5 --slots 0 --fast off auto 8,8,1200 12,8,1000 16,8,1000 24,8,800 32,8,800
python tools/sweep_geometry.py --model 300.hmm --sequences 100000 --steps 5 --slots 0 --fast off auto 2,8,1900 4,8,1500 8,8,1500 16,8,1200
python tools/sweep_geometry.py --model 400.hmm --sequences 100000 --steps 5 --slots 0 --fast off auto 2,8,2000 4,8,1600 8,8,1500
