#!/usr/bin/env python
"""Times the Plan-7 local Viterbi scan on a device-resident synthetic database and checks a sample against the oracle.

    python tools/viterbi_bench.py --models 100.hmm 1400.hmm 2405.hmm --sequences 100000

One JSON line per model: GCUPS = LENG x residues / time (the MSV metric's definition, so the two scans compare directly),
cells per clock per SM, mismatches against oracle/viterbi_oracle.c, and the oracle's own speed on the host cores.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", nargs="+", default=["1400.hmm"])
    ap.add_argument("--sequences", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", type=int, default=64)
    args = ap.parse_args()

    import torch

    import hmm_fasta_viterbi_b200 as msv
    from hmm_fasta_viterbi_b200 import _cabi
    from oracle_lib import Oracle, pack

    oracle = Oracle()
    packed = msv.Packed_sequences.synthetic_swissprot_like(args.sequences, 20261018)
    codes, offsets = packed.residues, packed.offsets
    db = msv.Database(codes, offsets)
    scores = torch.empty(len(packed), dtype=torch.float32, device="cuda")
    rng = np.random.default_rng(0)
    sample = rng.choice(len(packed), size=min(args.check, len(packed)), replace=False)
    sc, so = pack([codes[int(offsets[q]):int(offsets[q + 1])] for q in sample])
    stream = torch.cuda.current_stream()
    cores = os.cpu_count() or 1
    for name in args.models:
        h = oracle.load_hmm(os.path.join(REPO, "fixtures", "profile_HMMs", name))
        leng = h["model_length"] - 1
        table, logtr = _cabi.emission_table(h["match_emissions"]), _cabi.viterbi_transitions(h["transitions"])
        model = msv.ViterbiModel(table, logtr, *_cabi.model_transitions(h["model_length"]))
        otable, otr3 = oracle.prepare(h["match_emissions"])
        t_cpu = time.perf_counter()
        want = oracle.viterbi_score_batch(otable, oracle.viterbi_prepare(h["transitions"]), otr3, sc, so, threads=cores)
        t_cpu = time.perf_counter() - t_cpu
        for _ in range(2):
            db.viterbi_device(model, scores, stream.cuda_stream)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for _ in range(args.steps):
            db.viterbi_device(model, scores, stream.cuda_stream)
        t1.record(stream)
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / args.steps
        got = scores.cpu().numpy()[sample]
        cells = leng * float(offsets[-1])
        print(json.dumps({"model": name, "geometry": model.geometry, "sequences": len(packed), "ms": round(ms, 3),
                          "gcups": round(cells / ms / 1e6, 1), "cells_per_clk_per_sm": round(cells / (ms * 1e-3) / 148 / 1.965e9, 2),
                          "groups_env": os.environ.get("MSV_CUDA_VITERBI_GROUPS"),
                          "mismatches": int((got.view(np.uint32) != want.view(np.uint32)).sum()), "checked": len(sample),
                          "oracle_gcups": round(leng * float(so[-1]) / t_cpu / 1e9, 3), "oracle_threads": cores}), flush=True)
        model.close()


if __name__ == "__main__":
    main()
