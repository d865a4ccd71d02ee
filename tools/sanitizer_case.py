#!/usr/bin/env python
"""Small, fast workload for compute-sanitizer (memcheck / racecheck): every kernel family on a few hundred short
sequences, checked against the oracle.  Usage: compute-sanitizer --tool racecheck python tools/sanitizer_case.py"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import hmm_fasta_viterbi_b200 as msv  # noqa: E402
from hmm_fasta_viterbi_b200 import _cabi  # noqa: E402
from oracle_lib import Oracle, pack  # noqa: E402

oracle = Oracle()
rng = np.random.default_rng(0)
seqs = [rng.integers(0, 20, size=int(n), dtype=np.uint8) for n in rng.integers(0, 60, size=96)]
codes, offsets = pack(seqs)
for name, geometries in (("100.hmm", ["default", "8,16", "32,4,0", "128,4,0"]), ("300.hmm", ["32,12,8", "16,20", "128,8,8"]),
                         ("1400.hmm", ["32,44,16", "128,12,8", "32,44"])):
    path = os.path.join(REPO, "fixtures", "profile_HMMs", name)
    h = oracle.load_hmm(path)
    table, tr3 = oracle.prepare(h["match_emissions"])
    want = oracle.score_batch(table, tr3, codes, offsets)
    for geo in geometries:
        if geo == "default":
            os.environ.pop("MSV_CUDA_GEOMETRY", None)
        else:
            os.environ["MSV_CUDA_GEOMETRY"] = geo
        model = msv.Model(_cabi.emission_table(h["match_emissions"]), *_cabi.model_transitions(h["model_length"]))
        got = model.score_batch(codes, offsets)
        ok = got.view(np.uint32).tolist() == want.view(np.uint32).tolist()
        print(name, geo, model.geometry["lanes_per_sequence"], "ok" if ok else "MISMATCH", flush=True)
        assert ok
        model.close()
print("sanitizer case done")
