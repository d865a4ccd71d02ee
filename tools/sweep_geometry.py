#!/usr/bin/env python
"""Tuning aid: time the scan kernel for several kernel geometries on one model and check each against the oracle.

    python tools/sweep_geometry.py --model 1400.hmm --sequences 300000 --geometries 32,44 32,44,0 32,44,16 32,44,24

A geometry is "G,K" (generic kernel) or "G,K,KT" (warp kernel with KT tensor-memory columns per lane); it is forced
through the MSV_CUDA_GEOMETRY environment variable, which msv_cuda_model_create reads.
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="1400.hmm")
    ap.add_argument("--sequences", type=int, default=300_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--geometries", nargs="+", default=["default"])
    ap.add_argument("--long", action="store_true", help="config-5 style database (L ~ U[10000, 35000])")
    ap.add_argument("--check", type=int, default=200, help="sequences compared with the oracle per geometry")
    ap.add_argument("--slots", nargs="+", type=int, default=[0], help="sequences in flight per CTA (MSV_CUDA_BULK_SLOTS; 0 = the plan's own)")
    ap.add_argument("--fast", nargs="+", default=["off"],
                    help='lane-group plans, long sequences on fast CTAs: "auto", "off" or "ctas,warps,rows" (MSV_CUDA_FAST_CTAS)')
    args = ap.parse_args()

    import torch

    import hmm_fasta_viterbi_b200 as msv
    from hmm_fasta_viterbi_b200 import _cabi
    from oracle_lib import Oracle, pack

    model_path = os.path.join(REPO, "fixtures", "profile_HMMs", args.model)
    profile = msv.Profile_HMM(model_path)
    leng = profile.model_length - 1
    table = _cabi.emission_table(profile.match_emissions)
    tr = _cabi.model_transitions(profile.model_length)
    if args.long:
        packed = msv.Packed_sequences.synthetic_long_uniform(args.sequences, 2405, 10_000, 35_000)
    else:
        packed = msv.Packed_sequences.synthetic_swissprot_like(args.sequences, 20261018)
    codes, offsets = packed.residues, packed.offsets
    if any(f != "off" for f in args.fast):
        os.environ["MSV_CUDA_FAST_CTAS"] = "auto"  # the database keeps its length profile only when the experiment is on
    db = msv.Database(codes, offsets)
    cells = leng * float(offsets[-1])
    scores = torch.empty(len(packed), dtype=torch.float32, device="cuda")

    oracle = Oracle()
    otable, otr3 = oracle.prepare(oracle.load_hmm(model_path)["match_emissions"])
    rng = np.random.default_rng(0)
    sample = rng.choice(len(packed), size=min(args.check, len(packed)), replace=False)
    if args.long:
        sample = sample[:4]
    sc, so = pack([codes[int(offsets[q]):int(offsets[q + 1])] for q in sample])
    want = oracle.score_batch(otable, otr3, sc, so, threads=os.cpu_count() or 1)

    stream = torch.cuda.current_stream()
    for geo in args.geometries:
        if geo == "default":
            os.environ.pop("MSV_CUDA_GEOMETRY", None)
        else:
            os.environ["MSV_CUDA_GEOMETRY"] = geo
        try:
            model = msv.Model(table, *tr)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"geometry": geo, "error": str(e)}))
            continue
        for slots, fast in [(s, f) for s in args.slots for f in args.fast]:
            if slots:
                os.environ["MSV_CUDA_BULK_SLOTS"] = str(slots)
            else:
                os.environ.pop("MSV_CUDA_BULK_SLOTS", None)
            os.environ.pop("MSV_CUDA_FAST_CTAS", None)
            os.environ.pop("MSV_CUDA_NO_FAST_CTAS", None)
            if fast != "off":
                os.environ["MSV_CUDA_FAST_CTAS"] = fast
            for _ in range(2):
                db.score_device(model, scores, stream.cuda_stream)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            for _ in range(args.steps):
                db.score_device(model, scores, stream.cuda_stream)
            t1.record(stream)
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / args.steps
            got = scores.cpu().numpy()[sample]
            bad = int((got.view(np.uint32) != want.view(np.uint32)).sum())
            print(json.dumps({"model": args.model, "sequences": args.sequences, "geometry": geo, "slots": slots, "fast": fast, "plan": model.plan(db),
                              "chosen": model.geometry, "ms": round(ms, 3), "gcups": round(cells / ms / 1e6, 1),
                              "cells_per_clk_per_sm": round(cells / (ms * 1e-3) / 148 / 1.965e9, 2), "mismatches": bad}), flush=True)
        for name in ("MSV_CUDA_BULK_SLOTS", "MSV_CUDA_FAST_CTAS", "MSV_CUDA_NO_FAST_CTAS"):
            os.environ.pop(name, None)
        model.close()


if __name__ == "__main__":
    main()
