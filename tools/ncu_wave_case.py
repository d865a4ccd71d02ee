#!/usr/bin/env python
"""A handful of single-sequence calls (1400.hmm x 3500 residues) for `ncu -k regex:msv_wave`: the latency kernel alone."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import hmm_fasta_viterbi_b200 as msv  # noqa: E402
from hmm_fasta_viterbi_b200 import _cabi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "1400.hmm"
prof = msv.Profile_HMM(os.path.join(REPO, "fixtures", "profile_HMMs", name))
model = msv.Model(_cabi.emission_table(prof.match_emissions), *_cabi.model_transitions(prof.model_length))
rng = np.random.default_rng(5)
s = rng.integers(0, 20, size=3500, dtype=np.uint8)
for _ in range(6):
    model.score_sequence(s)
print(model.wave_geometry)
