import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hmm_fasta_viterbi_b200 as msv
from hmm_fasta_viterbi_b200 import _cabi
name = "100.hmm"
prof = msv.Profile_HMM(os.path.join("" + os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/fixtures/profile_HMMs", name))
model = msv.Model(_cabi.emission_table(prof.match_emissions), *_cabi.model_transitions(prof.model_length))
stream = torch.cuda.current_stream()
for label, packed in (("100k swissprot-like (mean 350)", msv.Packed_sequences.synthetic_swissprot_like(100_000, 1)),
                      ("25k uniform 1000-1800", msv.Packed_sequences.synthetic_long_uniform(25_000, 1, 1000, 1800)),
                      ("400k uniform 60-115", msv.Packed_sequences.synthetic_long_uniform(400_000, 1, 60, 115)),
                      ("100k uniform 300-400", msv.Packed_sequences.synthetic_long_uniform(100_000, 1, 300, 400))):
    db = msv.Database(packed.residues, packed.offsets)
    scores = torch.empty(len(packed), dtype=torch.float32, device="cuda")
    cells = float(packed.offsets[-1]) * (prof.model_length - 1)
    for slots in (64, 96, 192):
        os.environ["MSV_CUDA_BULK_SLOTS"] = str(slots)
        for _ in range(3): db.score_device(model, scores, stream.cuda_stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5): db.score_device(model, scores, stream.cuda_stream)
        b.record(stream); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(json.dumps({"db": label, "plan": model.plan(db), "slots": slots, "ms": round(ms, 3), "gcups": round(cells / ms / 1e6, 1)}), flush=True)
