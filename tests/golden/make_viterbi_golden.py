"""Writes tests/golden/viterbi_scores.json: IEEE-754 bit patterns of the Plan-7 local Viterbi score of every fixture
model against the 7 fixture sequences, computed by oracle/viterbi_oracle.c.

These are NOT reference outputs (the reference has no Viterbi); they freeze the oracle so that an accidental edit, a
compiler flag or a libm change shows up.  Run from the repo root:  python tests/golden/make_viterbi_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import fasta_path, hmm_path, model_files  # noqa: E402
from oracle_lib import Oracle, encode  # noqa: E402

oracle = Oracle()
seqs = oracle.load_fasta(fasta_path("fasta_like_example.fsa")) + oracle.load_fasta(fasta_path("random_FASTA.fsa"))
scores = {}
for name in model_files():
    h = oracle.load_hmm(hmm_path(name))
    table, tr3 = oracle.prepare(h["match_emissions"])
    logtr = oracle.viterbi_prepare(h["transitions"])
    scores[name] = [format(int(oracle.viterbi_score_codes(table, logtr, tr3, encode(s[1:])).view(np.uint32)), "08x") for s in seqs]
with open(os.path.join(HERE, "viterbi_scores.json"), "w") as f:
    json.dump({"generator": "oracle/viterbi_oracle.c via tests/golden/make_viterbi_golden.py (not reference outputs)",
               "sequences": "fasta_like_example.fsa (4) + random_FASTA.fsa (3)", "scores": scores}, f, indent=1)
print("wrote", len(scores), "models")
