#!/usr/bin/env python
"""Single-sequence latency (BASELINE.json config 2) of msv_cuda_score_sequence for several chain geometries of the
wavefront kernel: MSV_CUDA_WAVE_K = columns per lane (0 = wavefront kernel off: the exact four-warp kernel).
Each setting runs in its own process (the geometry is fixed when the model is created).  One JSON line per setting."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
import numpy as np
sys.path[:0] = [%(repo)r, os.path.join(%(repo)r, "tests")]
import hmm_fasta_viterbi_b200 as msv
from hmm_fasta_viterbi_b200 import _cabi
from oracle_lib import Oracle
name, length = sys.argv[1], int(sys.argv[2])
oracle = Oracle()
h = oracle.load_hmm(os.path.join(%(repo)r, "fixtures", "profile_HMMs", name))
table, tr3 = oracle.prepare(h["match_emissions"])
model = msv.Model(_cabi.emission_table(h["match_emissions"]), *_cabi.model_transitions(h["model_length"]))
rng = np.random.default_rng(5)
seqs = [rng.integers(0, 20, size=length, dtype=np.uint8) for _ in range(8)]
want = [oracle.score_codes(table, tr3, s) for s in seqs]
got = [model.score_sequence(s) for s in seqs]
ok = [int(np.float32(a).view(np.uint32)) == int(np.float32(b).view(np.uint32)) for a, b in zip(got, want)]
times = []
for _ in range(40):
    for s in seqs:
        t0 = time.perf_counter()
        model.score_sequence(s)
        times.append(time.perf_counter() - t0)
print(json.dumps({"model": name, "length": length, "wave": model.wave_geometry, "us_median": round(float(np.median(times)) * 1e6, 1),
                  "us_min": round(float(np.min(times)) * 1e6, 1), "bit_exact": all(ok), "calls": len(times)}))
''' % {"repo": REPO}

cases = [("1400.hmm", 3500)] if len(sys.argv) < 2 else [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]]
for name, length in cases:
    for k in ("off", "2", "4", "6", "8", "12", "16", "default"):
        env = dict(os.environ)
        env.pop("MSV_CUDA_WAVE_K", None)
        env.pop("MSV_CUDA_NO_WAVE", None)
        if k == "off":
            env["MSV_CUDA_NO_WAVE"] = "1"
        elif k != "default":
            env["MSV_CUDA_WAVE_K"] = k
        out = subprocess.run([sys.executable, "-c", CHILD, name, str(length)], env=env, capture_output=True, text=True, timeout=600)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else json.dumps({"error": out.stderr[-400:]})
        print(json.dumps({"setting": k}) [:-1] + ", " + line[1:], flush=True)
