#!/usr/bin/env python
"""One-off, exhaustive parity check at the headline size: every one of the 1 000 000 scores of the bench workload
(1400.hmm x synthetic Swiss-Prot-like database, seed 20261018) is compared bit-for-bit with the CPU checker (the
reference's own code from oracle/_ref when present, else the C restatement).  Takes ~1.5 minutes on 16 host threads."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import hmm_fasta_viterbi_b200 as msv  # noqa: E402
from hmm_fasta_viterbi_b200 import _cabi  # noqa: E402
from oracle_lib import Oracle, RefLib  # noqa: E402

# usage: full_parity_check.py [sequences] [model.hmm] [long|viterbi]
#   "long": config-5 style sequences of 10-35 k residues;  "viterbi": the Plan-7 local Viterbi scan against oracle/viterbi_oracle.c
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
model_name = sys.argv[2] if len(sys.argv) > 2 else "1400.hmm"
long_sequences = len(sys.argv) > 3 and sys.argv[3] == "long"
model_path = os.path.join(REPO, "fixtures", "profile_HMMs", model_name)
if len(sys.argv) > 3 and sys.argv[3] == "viterbi":
    oracle = Oracle()
    h = oracle.load_hmm(model_path)
    vit = msv.ViterbiModel(_cabi.emission_table(h["match_emissions"]), _cabi.viterbi_transitions(h["transitions"]),
                           *_cabi.model_transitions(h["model_length"]))
    db = msv.Packed_sequences.synthetic_swissprot_like(n, 20261018)
    gpu = vit.score_batch(db.residues, db.offsets)
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    table, tr3 = oracle.prepare(h["match_emissions"])
    cpu = oracle.viterbi_score_batch(table, oracle.viterbi_prepare(h["transitions"]), tr3, db.residues, db.offsets, threads)
    seconds = time.perf_counter() - t0
    mismatches = int((gpu.view(np.uint32) != cpu.view(np.uint32)).sum())
    print(json.dumps({"scan": "viterbi", "sequences": n, "residues": int(db.total_residues), "model": model_name,
                      "checker": "oracle/viterbi_oracle.c (parity unpinned: no reference implementation exists)", "cpu_threads": threads,
                      "cpu_seconds": round(seconds, 1), "mismatches": mismatches, "geometry": vit.geometry}))
    sys.exit(1 if mismatches else 0)
prof = msv.Profile_HMM(model_path)
model = msv.Model(_cabi.emission_table(prof.match_emissions), *_cabi.model_transitions(prof.model_length))
db = (msv.Packed_sequences.synthetic_long_uniform(n, 2405, 10_000, 35_000) if long_sequences
      else msv.Packed_sequences.synthetic_swissprot_like(n, 20261018))
gpu = model.score_batch(db.residues, db.offsets)
threads = os.cpu_count() or 1
t0 = time.perf_counter()
if RefLib.available():
    kind = "reference (oracle/_ref)"
    cpu = RefLib().model(model_path).run_batch(db.residues, db.offsets, threads)
else:
    kind = "port (oracle/msv_oracle.c)"
    oracle = Oracle()
    table, tr3 = oracle.prepare(oracle.load_hmm(model_path)["match_emissions"])
    cpu = oracle.score_batch(table, tr3, db.residues, db.offsets, threads)
seconds = time.perf_counter() - t0
mismatches = int((gpu.view(np.uint32) != np.asarray(cpu, np.float32).view(np.uint32)).sum())
print(json.dumps({"sequences": n, "residues": int(db.total_residues), "model": model_name, "long_sequences": long_sequences, "checker": kind, "cpu_threads": threads,
                  "cpu_seconds": round(seconds, 1), "mismatches": mismatches, "geometry": model.geometry}))
sys.exit(1 if mismatches else 0)
