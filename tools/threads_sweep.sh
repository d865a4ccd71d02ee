#!/bin/bash
# Threads per CTA of the warp kernel (= warps per SM and registers per thread) for the mid-length models.
# Needs hmm_fasta_viterbi_b200/variants/libmsv_threads.so (MSV_QUICK_BUILD + MSV_QUICK_EXTRA build).
cd "$(dirname "$0")/.."
run() { MSV_CUDA_LIBRARY=$PWD/hmm_fasta_viterbi_b200/variants/libmsv_threads.so python tools/sweep_geometry.py --model $1 --sequences 100000 --steps 4 --geometries "${@:2}" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['model'], d['geometry'], d['chosen']['threads_per_cta'], d['gcups'], d['mismatches'])"; }
run 500.hmm 32,16,16,1024,1 32,16,16,768,1 32,16,16,640,1
run 600.hmm 32,20,16,1024,1 32,20,16,768,1 32,20,16,640,1
run 700.hmm 32,22,18,1024,1 32,22,18,768,1 32,22,18,640,1
run 800.hmm 32,26,18,768,1 32,26,18,640,1 32,26,18,512,1
run 900.hmm 32,30,18,768,1 32,30,18,640,1 32,30,18,512,1
run 1001.hmm 32,32,24,768,1 32,32,24,640,1 32,32,24,512,1
run 1100.hmm 32,36,24,640,1 32,36,24,512,1
run 1200.hmm 32,38,18,640,1 32,38,18,512,1
run 1301.hmm 32,42,18,512,1 32,42,18,448,1
run 1509.hmm 32,48,16,512,1 32,48,16,448,1
run 1600.hmm 32,52,16,512,1 32,52,16,448,1
