#!/usr/bin/env python
"""What the fused gather costs, piece by piece (torchrun, N >= 2 GPUs of one NVLink domain): the plain scan, the scan
storing into every rank's symmetric buffer, the symmetric-memory barrier, and their combinations, for a strong-scaling
slice and a weak-scaling shard of the bench workload.  One JSON line per rank and mode (ms per step).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/gather_microbench.py
"""
import os, sys, json, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmm_fasta_viterbi_b200 as msv
from hmm_fasta_viterbi_b200 import _cabi, sharded
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prof=msv.Profile_HMM("fixtures/profile_HMMs/1400.hmm")
model=msv.Model(_cabi.emission_table(prof.match_emissions), *_cabi.model_transitions(prof.model_length), device=local)
packed=msv.Packed_sequences.synthetic_swissprot_like(1_000_000, 20261018)
for mode in ("strong","weak"):
    if mode=="strong": codes, offsets, _, _ = sharded.local_slice(packed.residues, packed.offsets, rank, world)
    else: codes, offsets = packed.residues, packed.offsets
    n=len(offsets)-1
    db=msv.Database(codes, offsets, device=local)
    t=torch.tensor([n],device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); n_max=int(t.item())
    fused=sharded.FusedGather(n_max, torch.device("cuda",local))
    scores=torch.empty(n_max, dtype=torch.float32, device="cuda")
    stream=torch.cuda.current_stream()
    def timeit(fn, k=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(k): fn()
        b.record(stream); torch.cuda.synchronize(); dist.barrier()
        return a.elapsed_time(b)/k
    r={}
    r["plain"]=timeit(lambda: db.score_device(model, scores, stream.cuda_stream))
    r["plain+barrier"]=timeit(lambda: (db.score_device(model, scores, stream.cuda_stream), fused._handle.barrier(channel=0)))
    r["gather_nobarrier"]=timeit(lambda: db.score_gather(model, fused._copies, rank*fused.slot, stream.cuda_stream))
    r["gather_own_only"]=timeit(lambda: db.score_gather(model, fused._copies[:1], rank*fused.slot, stream.cuda_stream))
    r["gather+barrier"]=timeit(lambda: fused.scan(model, db, stream.cuda_stream))
    r["barrier_only"]=timeit(lambda: fused._handle.barrier(channel=0))
    print(json.dumps({"mode":mode,"rank":rank,"n":n, **{k:round(v,3) for k,v in r.items()}}), flush=True)
dist.barrier(); dist.destroy_process_group()
