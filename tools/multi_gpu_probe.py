#!/usr/bin/env python
"""Where does the time of msv_cuda_multi_score_batch go?  (needs >= 2 GPUs; no torch, no torchrun)
Compares, on the config-4 database cut in `ngpu` slices: the multi call itself (MSV_MULTI_TRACE prints per-GPU wall times),
the same slices through msv_cuda_score_batch from Python threads (one per GPU), and each slice alone."""
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

import hmm_fasta_viterbi_b200 as msv  # noqa: E402
from hmm_fasta_viterbi_b200 import _cabi  # noqa: E402

ngpu = min(int(sys.argv[1]) if len(sys.argv) > 1 else 2, _cabi.device_count())
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
profile = msv.Profile_HMM(os.path.join(REPO, "fixtures", "profile_HMMs", "1400.hmm"))
packed = msv.Packed_sequences.synthetic_swissprot_like(n, 20261018)
codes, offsets = np.ascontiguousarray(packed.residues), np.ascontiguousarray(packed.offsets)
out = np.empty(n, np.float32)
for a in (codes, offsets, out):
    _cabi.check(_cabi.lib.msv_cuda_host_register(a.ctypes.data, a.nbytes))
cells = float(offsets[-1]) * (profile.model_length - 1)
models = [msv.Model(_cabi.emission_table(profile.match_emissions), *_cabi.model_transitions(profile.model_length), device=g) for g in range(ngpu)]
multi = _cabi.MultiGpu(models)
report = {"ngpu": ngpu, "sequences": n}


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


for name, mode in (("host", _cabi.GATHER_HOST), ("peer", _cabi.GATHER_PEER)):
    ms = timed(lambda: multi.score_batch(codes, offsets, out, gather=mode))
    report[f"multi_{name}_ms"] = round(ms, 3)
    report[f"multi_{name}_gcups"] = round(cells / ms / 1e6, 1)
whole = timed(lambda: models[0].score_batch(codes, offsets, out))
report["one_gpu_whole_database_ms"] = round(whole, 3)

bounds = _cabi.partition_by_cells(offsets, ngpu)
slices = []
for g in range(ngpu):
    a, b = int(bounds[g]), int(bounds[g + 1])
    local = np.ascontiguousarray(offsets[a:b + 1] - offsets[a])
    _cabi.check(_cabi.lib.msv_cuda_host_register(local.ctypes.data, local.nbytes))
    slices.append((codes[int(offsets[a]):int(offsets[b])], local, out[a:b]))
report["slice_alone_ms"] = [round(timed(lambda g=g: models[g].score_batch(*slices[g])), 3) for g in range(ngpu)]


def together():
    threads = [threading.Thread(target=lambda g=g: models[g].score_batch(*slices[g])) for g in range(1, ngpu)]
    for t in threads:
        t.start()
    models[0].score_batch(*slices[0])
    for t in threads:
        t.join()


report["slices_from_python_threads_ms"] = round(timed(together), 3)
os.environ["MSV_MULTI_TRACE"] = "1"
multi.score_batch(codes, offsets, out, gather=_cabi.GATHER_HOST)
print(json.dumps(report))
