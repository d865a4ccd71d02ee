#!/bin/bash
# Residue-word loop of the warp kernel: 4 rows (MSV_FORCE_UNROLL=1) against 8 rows (=2) per iteration, for the geometry each fixture
# model runs with.  Needs hmm_fasta_viterbi_b200/variants/libmsv_unroll{1,2}.so (MSV_QUICK_BUILD + MSV_QUICK_EXTRA builds).
cd "$(dirname "$0")/.."
for mg in "500.hmm 32,16,16,0,1" "600.hmm 32,20,16" "700.hmm 32,22,18,0,1" "800.hmm 32,26,18,0,1" "900.hmm 32,30,18,0,1" "1001.hmm 32,32,24,0,1" \
          "1100.hmm 32,36,24,0,1" "1200.hmm 32,38,18,0,1" "1301.hmm 32,42,18,0,1" "1400.hmm 32,44,24,0,1" "1509.hmm 32,48,16,0,1" "1600.hmm 32,52,16,0,1" \
          "1705.hmm 32,54,18,0,1" "1799.hmm 32,58,18,0,1" "1901.hmm 32,60,24,0,1" "2138.hmm 32,68,24,0,1" "2207.hmm 32,72,24,0,1" "2405.hmm 32,76,24,0,1"; do
  set -- $mg
  for u in 1 2; do
    MSV_CUDA_LIBRARY=$PWD/hmm_fasta_viterbi_b200/variants/libmsv_unroll$u.so python tools/sweep_geometry.py --model $1 --sequences 100000 --steps 4 --geometries $2 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('unroll$u', d['model'], d['geometry'], d['gcups'], d['mismatches'])"
  done
done
