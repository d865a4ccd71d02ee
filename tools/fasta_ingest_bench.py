#!/usr/bin/env python
"""Throughput of the FASTA readers on a synthetic 200k-sequence file.
Compares Packed_sequences::from_fasta_file (parallel mmap, straight to the packed layout), this implementation's
FASTA_protein_sequences (strings) and, when oracle/_ref is present, the reference's own reader."""
import json
import os
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
import hmm_fasta_viterbi_b200 as msv  # noqa: E402
from hmm_fasta_viterbi_b200 import host  # noqa: E402
from oracle_lib import LETTERS, RefLib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
db = msv.Packed_sequences.synthetic_swissprot_like(n, 5)
codes, off = db.residues, db.offsets
lut = np.frombuffer(LETTERS.encode(), dtype=np.uint8)
with tempfile.NamedTemporaryFile(suffix=".fsa", delete=False) as f:
    path = f.name
    for q in range(len(db)):
        s = lut[codes[int(off[q]):int(off[q + 1])]].tobytes()
        f.write(b">seq%d synthetic\n" % q)
        f.write(b"\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + b"\n")
size = os.path.getsize(path)
out = {"file_mb": size / 1e6, "sequences": n, "host_threads": os.cpu_count()}
best = 1e9
for _ in range(3):
    t = time.perf_counter()
    p = msv.Packed_sequences.from_fasta_file(path)
    best = min(best, time.perf_counter() - t)
assert (p.residues == codes).all() and (p.offsets == off).all()
out["packed_reader_s"] = best
out["packed_reader_mb_s"] = size / 1e6 / best
t = time.perf_counter()
h = host.lib.msvh_fasta_load(path.encode())
out["string_reader_s"] = time.perf_counter() - t
host.lib.msvh_fasta_free(h)
if RefLib.available():
    ref = RefLib()
    t = time.perf_counter()
    r = ref.lib.ref_fasta_load(path.encode())
    out["reference_reader_s"] = time.perf_counter() - t
    ref.lib.ref_fasta_free(r)
os.unlink(path)
print(json.dumps(out))
