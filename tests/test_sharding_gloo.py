"""World-size-2 test of the multi-rank host logic on CPU (gloo): cell-balanced contiguous slices, per-rank scoring,
all-gather of ragged score vectors, global order.  The scorer here is the CPU checker (the GPU scan itself is covered
by the -m gpu tests); what is under test is hmm_fasta_viterbi_b200/sharded.py and msv_host_partition_by_cells."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import hmm_path


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    from hmm_fasta_viterbi_b200 import Packed_sequences, sharded
    from oracle_lib import Oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = Oracle()
        table, tr3 = oracle.prepare(oracle.load_hmm(hmm_path("200.hmm"))["match_emissions"])
        db = Packed_sequences.synthetic_swissprot_like(400, 77)  # every rank builds the same database
        residues, offsets = db.residues, db.offsets
        res, off, lo, hi = sharded.local_slice(residues, offsets, rank, world)
        local = oracle.score_batch(table, tr3, res, off) if hi > lo else np.zeros(0, np.float32)
        bounds = sharded.shard_bounds(offsets, world)
        counts = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
        everything = sharded.gather_scores(torch.from_numpy(local), counts)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), everything.numpy())
        np.save(os.path.join(out_dir, f"bounds{rank}.npy"), bounds)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_and_gather(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from hmm_fasta_viterbi_b200 import Packed_sequences

    db = Packed_sequences.synthetic_swissprot_like(400, 77)
    table, tr3 = oracle.prepare(oracle.load_hmm(hmm_path("200.hmm"))["match_emissions"])
    want = oracle.score_batch(table, tr3, db.residues, db.offsets, threads=4)
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npy")
        assert got.view(np.uint32).tolist() == want.view(np.uint32).tolist()
    bounds = np.load(tmp_path / "bounds0.npy")
    assert bounds[0] == 0 and bounds[-1] == 400
    cells = [int(db.offsets[bounds[r + 1]] - db.offsets[bounds[r]]) for r in range(world)]
    assert abs(cells[0] - cells[1]) <= 3000  # balanced to within one sequence
