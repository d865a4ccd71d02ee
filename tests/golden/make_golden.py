#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference's own unmodified code (oracle/_ref/libmsv_ref.so).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py

Outputs
  msv_scores.json   -- for each of the 24 fixture models: the IEEE-754 bit pattern of
                       MSV_HMM::run_on_sequence (MSV_HMM.cpp:74-113) on the 4 sequences of
                       fasta_like_example.fsa and the 3 of random_FASTA.fsa, plus extra seeded sequences
                       (L = 0, 1, 2, 31, 32, 33, 257, 1000) so that edge lengths are pinned too.
  model_tables.json -- per model: model_length, the three model transition scores (bits) and a CRC32 of the
                       [20][model_length] fp32 emission table built by MSV_HMM::MSV_HMM (MSV_HMM.cpp:35-57).
  readers.json      -- Profile_HMM / FASTA_protein_sequences known answers (name, model_length, stats bits, CRC32 of
                       the three probability matrices; the parsed FASTA records).
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "tests"))
from oracle_lib import LETTERS, RefLib, build_oracle  # noqa: E402

FIX = os.path.join(REPO, "fixtures")
EXTRA_LENGTHS = [0, 1, 2, 31, 32, 33, 257, 1000]


def bits(x) -> str:
    return format(int(np.float32(x).view(np.uint32)), "08x")


def extra_sequences() -> list[str]:
    rng = np.random.Generator(np.random.PCG64(20261018))
    out = []
    for n in EXTRA_LENGTHS:
        out.append("#" + "".join(LETTERS[i] for i in rng.integers(0, 20, size=n)))
    return out


def main() -> None:
    build_oracle()
    ref = RefLib()
    ex = ref.load_fasta(os.path.join(FIX, "FASTA_files", "fasta_like_example.fsa"))
    rnd = ref.load_fasta(os.path.join(FIX, "FASTA_files", "random_FASTA.fsa"))
    extra = extra_sequences()
    models = sorted((f for f in os.listdir(os.path.join(FIX, "profile_HMMs")) if f.endswith(".hmm")),
                    key=lambda s: int(s.split(".")[0]))
    scores, tables, readers = {}, {}, {"hmm": {}, "fasta": {}}
    for name in models:
        path = os.path.join(FIX, "profile_HMMs", name)
        m = ref.model(path)
        scores[name] = {
            "example": [bits(m.run_on_sequence(s)) for s in ex],
            "random": [bits(m.run_on_sequence(s)) for s in rnd],
            "extra": [bits(m.run_on_sequence(s)) for s in extra],
        }
        t, tr3 = m.table()
        tables[name] = {"model_length": m.model_length, "tr_B_Mk": bits(tr3[0]), "tr_E_C": bits(tr3[1]),
                        "tr_E_J": bits(tr3[2]), "table_crc32": format(zlib.crc32(t.tobytes()), "08x")}
        h = ref.load_hmm(path)
        readers["hmm"][name] = {
            "name": h["name"], "model_length": h["model_length"], "stats": [bits(v) for v in h["stats"]],
            "match_crc32": format(zlib.crc32(h["match_emissions"].tobytes()), "08x"),
            "insert_crc32": format(zlib.crc32(h["insert_emissions"].tobytes()), "08x"),
            "transitions_crc32": format(zlib.crc32(h["transitions"].tobytes()), "08x"),
        }
        print(name, scores[name]["example"][0], scores[name]["random"][0], flush=True)
    readers["fasta"] = {"fasta_like_example.fsa": ex, "random_FASTA.fsa": rnd}
    meta = {"generator": "tests/golden/make_golden.py", "source": "reference MSV_HMM::run_on_sequence via oracle/_ref",
            "extra_lengths": EXTRA_LENGTHS, "extra_sequences": extra}
    for fname, obj in (("msv_scores.json", {"meta": meta, "scores": scores}), ("model_tables.json", tables),
                       ("readers.json", readers)):
        with open(os.path.join(HERE, fname), "w") as f:
            json.dump(obj, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
