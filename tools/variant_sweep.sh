#!/bin/bash
# Kernel-development aid: the same measurement (1400.hmm x 300 k sequences, resident scan) with every library under
# hmm_fasta_viterbi_b200/variants/ (built with MSV_QUICK_BUILD and different macros), and with the shipped library.
cd "$(dirname "$0")/.."
for lib in "" hmm_fasta_viterbi_b200/variants/*.so; do
  echo "== ${lib:-shipped}"
  MSV_CUDA_LIBRARY=${lib:+$PWD/$lib} python tools/sweep_geometry.py --model 1400.hmm --sequences 300000 --steps 5 --geometries default
done
