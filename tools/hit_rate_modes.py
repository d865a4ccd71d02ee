#!/usr/bin/env python
"""Which row variant should a hit-rich database get?  1400.hmm (and others) x 200 k sequences with a consensus segment planted
in 0 / 2 / 10 / 50 % of them, scanned with the warp kernel forced to whole-sequence speculation, block-wise speculation and
exact rows (MSV_CUDA_SPECULATION), and with the library's own choice ("auto", after two settling scans)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import hmm_fasta_viterbi_b200 as msv  # noqa: E402
from hmm_fasta_viterbi_b200 import _cabi  # noqa: E402

stream = torch.cuda.current_stream()
for name in sys.argv[1:] or ["1400.hmm"]:
    prof, _ = bench.load_model(msv, _cabi, name, 0)
    leng = prof.model_length - 1
    consensus = np.argmax(_cabi.emission_table(prof.match_emissions)[:, 1:], axis=0).astype(np.uint8)
    base = msv.Packed_sequences.synthetic_swissprot_like(200_000, 14)
    offsets = np.ascontiguousarray(base.offsets)
    for fraction in (0.0, 0.02, 0.10, 0.50):
        codes = bench.plant_hits(base.residues, offsets, consensus, fraction, 99)
        db = msv.Database(codes, offsets)
        scores = torch.empty(len(base), dtype=torch.float32, device="cuda")
        row = {"model": name, "hits": fraction}
        for mode in ("whole", "blocks", "none", "auto"):
            if mode == "auto":
                os.environ.pop("MSV_CUDA_SPECULATION", None)
            else:
                os.environ["MSV_CUDA_SPECULATION"] = mode
            _, model = bench.load_model(msv, _cabi, name, 0)
            ms = bench.event_ms(torch, stream, lambda: db.score_device(model, scores, stream.cuda_stream), 3, warm=3)
            row[mode] = round(leng * float(offsets[-1]) / ms / 1e6, 1)
            if mode == "auto":
                row["auto_state"] = model.speculation
            model.close()
        print(json.dumps(row), flush=True)
        db.close()
