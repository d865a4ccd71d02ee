#!/usr/bin/env python
"""End-to-end call (msv_cuda_score_batch: pinned host buffers in, host scores out) against the resident scan, for database
sizes from one GPU's share of an 8-GPU job (125 k sequences) to the whole config-4 database, under the tuning switches of the
pipelined upload: MSV_CUDA_FIRST_STAGE_MB, MSV_CUDA_STAGE_GROWTH and MSV_CUDA_ONE_COMPUTE_STREAM.  One JSON line per (size, setting)."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

import hmm_fasta_viterbi_b200 as msv  # noqa: E402
from hmm_fasta_viterbi_b200 import _cabi  # noqa: E402

model_name = sys.argv[1] if len(sys.argv) > 1 else "1400.hmm"
profile = msv.Profile_HMM(os.path.join(REPO, "fixtures", "profile_HMMs", model_name))
model = msv.Model(_cabi.emission_table(profile.match_emissions), *_cabi.model_transitions(profile.model_length))
for n in (125_000, 250_000, 1_000_000):
    packed = msv.Packed_sequences.synthetic_swissprot_like(n, 20261018)
    codes, offsets = np.ascontiguousarray(packed.residues), np.ascontiguousarray(packed.offsets)
    out = np.empty(n, np.float32)
    for a in (codes, offsets, out):
        _cabi.check(_cabi.lib.msv_cuda_host_register(a.ctypes.data, a.nbytes))
    cells = float(offsets[-1]) * (profile.model_length - 1)
    db = msv.Database(codes, offsets)
    want = db.score(model)
    t0 = time.perf_counter()
    for _ in range(5):
        db.score(model)
    resident_ms = (time.perf_counter() - t0) / 5 * 1e3  # includes the 4 n-byte download
    for stage_mb, one_stream, growth in ((2, True, 4), (8, True, 4), (8, False, 4), (8, False, 0), (8, False, 2), (8, False, 1.5), (4, False, 0)):
        os.environ["MSV_CUDA_FIRST_STAGE_MB"] = str(stage_mb)
        if growth:
            os.environ["MSV_CUDA_STAGE_GROWTH"] = str(growth)
        else:
            os.environ.pop("MSV_CUDA_STAGE_GROWTH", None)  # the library's own choice from the model length
        if one_stream:
            os.environ["MSV_CUDA_ONE_COMPUTE_STREAM"] = "1"
        else:
            os.environ.pop("MSV_CUDA_ONE_COMPUTE_STREAM", None)
        for _ in range(3):
            model.score_batch(codes, offsets, out)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            model.score_batch(codes, offsets, out)
        ms = (time.perf_counter() - t0) / reps * 1e3
        print(json.dumps({"model": model_name, "sequences": n, "first_stage_mb": stage_mb, "compute_streams": 1 if one_stream else 2, "stage_growth": growth or "auto",
                          "e2e_ms": round(ms, 3), "e2e_gcups": round(cells / ms / 1e6, 1), "resident_scan_plus_download_ms": round(resident_ms, 3),
                          "same_bits": bool((out.view(np.uint32) == want.view(np.uint32)).all())}), flush=True)
    for a in (codes, offsets, out):
        _cabi.lib.msv_cuda_host_unregister(a.ctypes.data)
    db.close()
