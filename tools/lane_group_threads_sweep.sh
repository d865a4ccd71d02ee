#!/bin/bash
# Lane-group kernels compiled for fewer threads per CTA (more registers per thread) against the standard ones launched with the
# same number of threads.  Needs hmm_fasta_viterbi_b200/variants/libmsv_groups.so (MSV_QUICK_BUILD + MSV_QUICK_EXTRA build).
cd "$(dirname "$0")/.."
run() { MSV_CUDA_LIBRARY=$PWD/hmm_fasta_viterbi_b200/variants/libmsv_groups.so python tools/sweep_geometry.py --model $1 --sequences $2 --steps 5 --slots $3 --geometries "${@:4}" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['model'], d['sequences'], d['geometry'], d['chosen']['threads_per_cta'], d['slots'], d['gcups'], d['mismatches'])"; }
run 100.hmm 100000 64 4,26 4,26,-1,256
run 100.hmm 100000 96 4,26 4,26,-1,384
run 200.hmm 100000 64 4,52 4,52,-1,256
run 200.hmm 100000 96 4,52 4,52,-1,384
run 300.hmm 100000 64 8,38 8,38,-1,512
run 300.hmm 100000 48 8,38 8,38,-1,384
run 100.hmm 1000000 64 4,26 4,26,-1,256
