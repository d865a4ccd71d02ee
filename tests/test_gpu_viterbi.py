"""GPU parity tests of the Plan-7 local Viterbi scan (pytest -m gpu), through the C ABI (msv_cuda_viterbi_*) and through
the C++ Viterbi_HMM class.  Checker: oracle/viterbi_oracle.c (parity unpinned against the reference, which has no
Viterbi; pinned as far as tests/test_viterbi_oracle.py can).  The bar is bit-exact fp32: every add has the oracle's
operands, and the delete chain is evaluated as a maximum over left-to-right summed paths (see viterbi_kernels.cuh)."""
import os

import numpy as np
import pytest

import hmm_fasta_viterbi_b200 as msv
from conftest import fasta_path, hmm_path, load_golden, model_files
from hmm_fasta_viterbi_b200 import _cabi
from oracle_lib import pack
from test_viterbi_oracle import DD, DM, II, IM, MD, MI, MM, random_model

pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 1


def ubits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def viterbi_model(oracle, match, transitions, device=0):
    table, tr3 = oracle.prepare(match)
    logtr = oracle.viterbi_prepare(transitions)
    mine_table, mine_logtr = _cabi.emission_table(match), _cabi.viterbi_transitions(transitions)
    assert ubits(mine_table).tolist() == ubits(table).tolist() and ubits(mine_logtr).tolist() == ubits(logtr).tolist()
    return msv.ViterbiModel(mine_table, mine_logtr, *_cabi.model_transitions(match.shape[0]), device=device), table, logtr, tr3


def random_db(rng, n, lo, hi):
    seqs = [rng.integers(0, 20, size=int(k), dtype=np.uint8) for k in rng.integers(lo, hi + 1, size=n)]
    return pack(seqs)


@pytest.mark.parametrize("name", model_files())
def test_fixture_models_match_committed_vectors(name):
    """Viterbi_HMM (C++ class) on the 7 fixture sequences, one call per sequence and as one batch."""
    want = load_golden("viterbi_scores.json")["scores"][name]
    model = msv.Viterbi_HMM(msv.Profile_HMM(hmm_path(name)))
    seqs = msv.FASTA_protein_sequences(fasta_path("fasta_like_example.fsa")).sequences + \
        msv.FASTA_protein_sequences(fasta_path("random_FASTA.fsa")).sequences
    fmt = lambda v: format(int(np.float32(v).view(np.uint32)), "08x")
    assert [fmt(model.parallel_run_on_sequence(s)) for s in seqs[:5]] == want[:5]
    packed = msv.Packed_sequences.from_arrays(*pack([_cabi.encode(s[1:]) for s in seqs]))
    assert [fmt(v) for v in model.parallel_run_on_sequences(packed)] == want
    assert [fmt(v) for v in model.parallel_run_on_sequences(msv.Device_database(packed))] == want


@pytest.mark.parametrize("name,n,hi", [("100.hmm", 3000, 300), ("300.hmm", 1500, 300), ("700.hmm", 1000, 300), ("1400.hmm", 800, 400),
                                        ("1901.hmm", 500, 300), ("2405.hmm", 400, 300)])
def test_batch_matches_oracle(oracle, name, n, hi):
    h = oracle.load_hmm(hmm_path(name))
    model, table, logtr, tr3 = viterbi_model(oracle, h["match_emissions"], h["transitions"])
    rng = np.random.default_rng(int(name.split(".")[0]) + 7)
    codes, offsets = random_db(rng, n, 0, hi)
    want = oracle.viterbi_score_batch(table, logtr, tr3, codes, offsets, threads=CORES)
    got = model.score_batch(codes, offsets)
    assert ubits(got).tolist() == ubits(want).tolist()
    db = msv.Database(codes, offsets)
    assert ubits(db.viterbi(model)).tolist() == ubits(want).tolist()
    assert ubits(db.viterbi(model)).tolist() == ubits(want).tolist()  # the work queue is reset between launches


@pytest.mark.parametrize("leng", [1, 2, 31, 32, 127, 128, 129, 1023, 1024, 2559])
def test_model_lengths_around_lane_boundaries(oracle, leng):
    """Random models whose length sits on either side of a columns-per-lane step; the model is right-aligned in the
    warp, so these exercise the dummy slots on the left."""
    rng = np.random.default_rng(leng)
    match, tr = random_model(rng, leng, spread=0.5)
    model, table, logtr, tr3 = viterbi_model(oracle, match, tr)
    assert model.geometry["columns_per_lane"] == max(4, -(-(leng + 1) // 128) * 4)
    codes, offsets = random_db(rng, 300, 0, 120)
    want = oracle.viterbi_score_batch(table, logtr, tr3, codes, offsets, threads=CORES)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()


@pytest.mark.parametrize("leng", [200, 1400])
def test_long_delete_chains_cross_lanes(oracle, leng):
    """Delete-friendly transitions (d->d ~ 1, cheap m->d) and sequences that match the two ends of the model: the best
    path deletes hundreds of columns, so the carried-in delete path crosses many lanes (the repeated propagation step)."""
    rng = np.random.default_rng(leng)
    match = np.full((leng + 1, 20), 0.002, np.float32)
    match[0] = 0
    consensus = rng.integers(0, 20, size=leng + 1)
    match[np.arange(1, leng + 1), consensus[1:]] = 0.962
    tr = np.zeros((leng + 1, 7), np.float32)
    tr[:, [MM, MI, MD]] = (0.6, 0.05, 0.35)
    tr[:, [IM, II]] = (0.6, 0.4)
    tr[:, [DM, DD]] = (0.02, 0.98)
    model, table, logtr, tr3 = viterbi_model(oracle, match, tr)
    seqs = []
    for _ in range(200):
        a, b = int(rng.integers(5, 30)), int(rng.integers(5, 30))
        gap = int(rng.integers(leng // 3, leng - 60))
        start = int(rng.integers(1, leng - gap - a - b))
        s = np.concatenate([consensus[start:start + a], consensus[start + a + gap:start + a + gap + b]]).astype(np.uint8)
        seqs.append(s)
    codes, offsets = pack(seqs)
    want = oracle.viterbi_score_batch(table, logtr, tr3, codes, offsets, threads=CORES)
    got = model.score_batch(codes, offsets)
    assert ubits(got).tolist() == ubits(want).tolist()
    # the deletes matter: forbidding them changes the scores
    no_delete = tr.copy()
    no_delete[:, MD] = 0
    assert (oracle.viterbi_score_batch(table, oracle.viterbi_prepare(no_delete), tr3, codes, offsets, threads=CORES) < want).mean() > 0.1


def test_long_insert_runs(oracle):
    """Insert-friendly transitions and sequences with long foreign runs inside a consensus match."""
    leng = 300
    rng = np.random.default_rng(3)
    match = np.full((leng + 1, 20), 0.002, np.float32)
    match[0] = 0
    consensus = rng.integers(0, 20, size=leng + 1)
    match[np.arange(1, leng + 1), consensus[1:]] = 0.962
    tr = np.zeros((leng + 1, 7), np.float32)
    tr[:, [MM, MI, MD]] = (0.7, 0.25, 0.05)
    tr[:, [IM, II]] = (0.1, 0.9)
    tr[:, [DM, DD]] = (0.5, 0.5)
    model, table, logtr, tr3 = viterbi_model(oracle, match, tr)
    seqs = []
    for _ in range(200):
        start, a, b = int(rng.integers(1, 200)), int(rng.integers(10, 40)), int(rng.integers(10, 40))
        run = rng.integers(0, 20, size=int(rng.integers(1, 60)))
        seqs.append(np.concatenate([consensus[start:start + a], run, consensus[start + a:start + a + b]]).astype(np.uint8))
    codes, offsets = pack(seqs)
    want = oracle.viterbi_score_batch(table, logtr, tr3, codes, offsets, threads=CORES)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()


def test_collapses_to_msv_on_device(oracle):
    """m->m = 1 and nothing else: the Viterbi kernel must return the MSV kernel's bits (which are the reference's)."""
    h = oracle.load_hmm(hmm_path("1400.hmm"))
    tr = np.zeros_like(h["transitions"])
    tr[:, MM] = 1.0
    vit, table, logtr, tr3 = viterbi_model(oracle, h["match_emissions"], tr)
    msv_model = msv.Model(_cabi.emission_table(h["match_emissions"]), *_cabi.model_transitions(h["model_length"]))
    rng = np.random.default_rng(9)
    codes, offsets = random_db(rng, 1000, 0, 400)
    db = msv.Database(codes, offsets)
    assert ubits(db.viterbi(vit)).tolist() == ubits(db.score(msv_model)).tolist()


def test_edge_cases_and_errors(oracle):
    h = oracle.load_hmm(hmm_path("100.hmm"))
    model, table, logtr, tr3 = viterbi_model(oracle, h["match_emissions"], h["transitions"])
    # empty database, empty sequences, a single residue
    assert model.score_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64)).size == 0
    codes, offsets = pack([np.zeros(0, np.uint8), np.array([3], np.uint8), np.zeros(0, np.uint8)])
    got = model.score_batch(codes, offsets)
    assert np.isneginf(got[0]) and np.isneginf(got[2])
    assert ubits(got[1:2]).tolist() == ubits([oracle.viterbi_score_codes(table, logtr, tr3, np.array([3], np.uint8))]).tolist()
    # a residue code outside the alphabet is reported, not scored
    with pytest.raises(msv.MsvCudaError) as err:
        model.score_batch(np.array([1, 2, 20, 3], np.uint8), np.array([0, 4], np.uint64))
    assert err.value.status == -4
    # positive "log probabilities" and over-long models are refused
    bad = logtr.copy()
    bad[5, 0] = 0.5
    with pytest.raises(msv.MsvCudaError) as err:
        msv.ViterbiModel(table, bad, *tr3)
    assert err.value.status == -1
    long_match, long_tr = random_model(np.random.default_rng(0), 2560)
    with pytest.raises(msv.MsvCudaError) as err:
        msv.ViterbiModel(_cabi.emission_table(long_match), _cabi.viterbi_transitions(long_tr), *_cabi.model_transitions(2561))
    assert err.value.status == -5


@pytest.mark.parametrize("name", ["100.hmm", "200.hmm", "300.hmm", "400.hmm"])
def test_lane_group_kernel_for_short_models(oracle, name, monkeypatch):
    """Eight lanes per sequence, four sequences per warp (viterbi_scan_group_kernel): forced on, it must give the oracle's
    bits on ragged inputs (empty sequences, single residues, hits with deletions) -- and so must the warp kernel."""
    h = oracle.load_hmm(hmm_path(name))
    model, table, logtr, tr3 = viterbi_model(oracle, h["match_emissions"], h["transitions"])
    leng = h["model_length"] - 1
    consensus = np.argmax(h["match_emissions"][1:], axis=1).astype(np.uint8)
    rng = np.random.default_rng(leng + 3)
    seqs = [rng.integers(0, 20, size=int(k), dtype=np.uint8) for k in rng.integers(0, 300, size=1500)]
    for _ in range(300):  # hits, most with a deletion inside
        a = int(rng.integers(0, leng - 40))
        b = int(rng.integers(a + 10, min(leng - 10, a + 50)))
        c = int(rng.integers(b, min(leng - 5, b + 30)))
        seqs.append(np.concatenate([consensus[a:b], consensus[c:min(leng, c + 40)]]))
    seqs += [np.zeros(0, np.uint8), np.array([7], np.uint8)]
    codes, offsets = pack(seqs)
    want = oracle.viterbi_score_batch(table, logtr, tr3, codes, offsets, threads=CORES)
    db = msv.Database(codes, offsets)
    monkeypatch.setenv("MSV_CUDA_VITERBI_GROUPS", "1")
    assert ubits(db.viterbi(model)).tolist() == ubits(want).tolist()
    assert ubits(db.viterbi(model)).tolist() == ubits(want).tolist()
    monkeypatch.setenv("MSV_CUDA_VITERBI_GROUPS", "0")
    assert ubits(db.viterbi(model)).tolist() == ubits(want).tolist()


def test_speculative_rows_equal_exact_rows(oracle, monkeypatch):
    """The kernel assumes J <= N (B = N + move) and verifies once per sequence; hits (consensus-derived sequences) must be
    rescanned exactly, random sequences not; both kernels must return the oracle's bits."""
    h = oracle.load_hmm(hmm_path("700.hmm"))
    model, table, logtr, tr3 = viterbi_model(oracle, h["match_emissions"], h["transitions"])
    consensus = np.argmax(h["match_emissions"][1:], axis=1).astype(np.uint8)
    rng = np.random.default_rng(70)
    seqs = []
    for q in range(600):
        if q % 2:
            a = int(rng.integers(0, 600))
            seqs.append(np.concatenate([rng.integers(0, 20, size=20, dtype=np.uint8), consensus[a:a + int(rng.integers(8, 90))],
                                        rng.integers(0, 20, size=int(rng.integers(0, 50)), dtype=np.uint8)]))
        else:
            seqs.append(rng.integers(0, 20, size=int(rng.integers(0, 400)), dtype=np.uint8))
    seqs.append(rng.integers(0, 20, size=5000, dtype=np.uint8))  # longer than the kernel speculates on
    codes, offsets = pack(seqs)
    want = oracle.viterbi_score_batch(table, logtr, tr3, codes, offsets, threads=CORES)
    assert 0.3 < (want > 0).mean() < 0.6
    db = msv.Database(codes, offsets)
    assert ubits(db.viterbi(model)).tolist() == ubits(want).tolist()
    monkeypatch.setenv("MSV_CUDA_NO_SPECULATION", "1")
    assert ubits(db.viterbi(model)).tolist() == ubits(want).tolist()


def test_full_size_properties(oracle):
    """1400.hmm x 20 000 Swiss-Prot-like sequences: run-to-run bit determinism, a seeded sample against the oracle, and
    invariance under permutation of the database (scores belong to sequences, not to queue positions)."""
    h = oracle.load_hmm(hmm_path("1400.hmm"))
    model, table, logtr, tr3 = viterbi_model(oracle, h["match_emissions"], h["transitions"])
    database = msv.Packed_sequences.synthetic_swissprot_like(20_000, 11)
    resident = msv.Database(database.residues, database.offsets)
    first, second = resident.viterbi(model), resident.viterbi(model)
    assert ubits(first).tolist() == ubits(second).tolist()
    rng = np.random.default_rng(2)
    sample = rng.choice(len(database), size=64, replace=False)
    off = database.offsets
    seqs = [database.residues[int(off[q]):int(off[q + 1])] for q in sample]
    sc, so = pack(seqs)
    want = oracle.viterbi_score_batch(table, logtr, tr3, sc, so, threads=CORES)
    assert ubits(first[sample]).tolist() == ubits(want).tolist()
    perm = rng.permutation(2000)
    all_seqs = [database.residues[int(off[q]):int(off[q + 1])] for q in range(2000)]
    pc, po = pack([all_seqs[q] for q in perm])
    assert ubits(model.score_batch(pc, po)).tolist() == ubits(first[:2000][perm]).tolist()


def test_viterbi_filter_statistics(oracle):
    """Bit score against the null length model and Gumbel P-value with STATS LOCAL VITERBI (HMMER3's second filter stage):
    raw scores bit-exact, statistics within 1e-6 relative of an fp64 host evaluation (they are ordinary floating point)."""
    import math
    name = "400.hmm"
    h = oracle.load_hmm(hmm_path(name))
    mu, lam = float(h["stats"][2]), float(h["stats"][3])
    model, table, logtr, tr3 = viterbi_model(oracle, h["match_emissions"], h["transitions"])
    rng = np.random.default_rng(21)
    codes, offsets = random_db(rng, 500, 1, 300)
    db = msv.Database(codes, offsets)
    scores, bits, p = db.viterbi_filter(model, mu, lam)
    want = oracle.viterbi_score_batch(table, logtr, tr3, codes, offsets, threads=CORES)
    assert ubits(scores).tolist() == ubits(want).tolist()
    for q in range(0, 500, 7):
        L = int(offsets[q + 1] - offsets[q])
        null1 = L * math.log(L / (L + 1.0)) + math.log(1.0 / (L + 1.0))
        b = (float(want[q]) - null1) / math.log(2.0)
        pv = -math.expm1(-math.exp(-lam * (b - mu)))
        assert abs(bits[q] - b) <= 1e-6 * max(1.0, abs(b))
        assert abs(p[q] - pv) <= 1e-5 * max(pv, 1e-30)
    # the C++ class reads mu / lambda from the profile and returns the same hits
    prof = msv.Profile_HMM(hmm_path(name))
    assert prof.stats_local_viterbi_mu == np.float32(mu) and prof.stats_local_viterbi_lambda == np.float32(lam)
    hits = msv.Viterbi_HMM(prof).viterbi_filter(msv.Device_database(msv.Packed_sequences.from_arrays(codes, offsets)), 0.5)
    keep = np.flatnonzero(p <= 0.5)
    assert hits["index"].tolist() == keep.tolist() and ubits(hits["score"]).tolist() == ubits(want[keep]).tolist()
    assert np.array_equal(hits["bits"], bits[keep]) and np.array_equal(hits["p_value"], p[keep])


def test_msv_scan_two_stage_pipeline(tmp_path):
    """tools/msv_scan --viterbi: MSV filter, then the survivors rescored by Viterbi_HMM (first two stages of HMMER3's
    pipeline).  A sequence sampled from the model's consensus must survive both stages; random sequences must not."""
    import subprocess
    from conftest import REPO
    exe = os.path.join(REPO, "build", "msv_scan")
    if not os.path.exists(exe):
        pytest.skip("build/msv_scan not built")
    prof = msv.Profile_HMM(hmm_path("300.hmm"))
    consensus = "".join("ACDEFGHIKLMNPQRSTVWY"[int(np.argmax(row))] for row in prof.match_emissions[1:])
    rng = np.random.default_rng(8)
    fasta = tmp_path / "db.fsa"
    with open(fasta, "w") as f:
        for q in range(200):
            f.write(f">r{q}\n" + "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWY"), size=int(rng.integers(50, 400)))) + "\n")
        f.write(">homolog\n" + consensus[40:120] + consensus[150:260] + "\n")  # a deletion of 30 columns inside the hit
    out = subprocess.run([exe, "--viterbi", hmm_path("300.hmm"), str(fasta)], capture_output=True, text=True, check=True)
    rows = [line.split("\t") for line in out.stdout.splitlines() if not line.startswith("#")]
    assert [r[1] for r in rows] == ["200"], out.stdout + out.stderr
    assert float(rows[0][5]) < 1e-6 and float(rows[0][8]) < 1e-6 and float(rows[0][6]) > 100.0


@pytest.mark.parametrize("seed", range(6))
def test_random_models_both_scans(oracle, seed):
    """Fuzz: random model lengths (every columns-per-lane step up to 1300 columns gets hit over the seeds), emission rows
    with impossible residues (probability 0 -> -inf scores), sparse transitions (probability 0), ragged databases.  MSV and
    Viterbi through every default plan against their oracles, bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    for leng in (int(rng.integers(1, 60)), int(rng.integers(60, 450)), int(rng.integers(450, 1300))):
        match, tr = random_model(rng, leng, spread=0.4)
        match[rng.random(match.shape) < 0.05] = 0.0   # impossible residues
        match[0] = 0
        tr[rng.random(tr.shape) < 0.03] = 0.0         # impossible transitions
        n = 5000 if leng < 450 else 1200
        seqs = [rng.integers(0, 20, size=int(k), dtype=np.uint8) for k in rng.integers(0, 260, size=n)]
        codes, offsets = pack(seqs)
        table, tr3 = oracle.prepare(match)
        logtr = oracle.viterbi_prepare(tr)
        db = msv.Database(codes, offsets)
        vit = msv.ViterbiModel(_cabi.emission_table(match), _cabi.viterbi_transitions(tr), *_cabi.model_transitions(leng + 1))
        want = oracle.viterbi_score_batch(table, logtr, tr3, codes, offsets, threads=CORES)
        assert ubits(db.viterbi(vit)).tolist() == ubits(want).tolist(), ("viterbi", leng)
        model = msv.Model(_cabi.emission_table(match), *_cabi.model_transitions(leng + 1))
        want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
        assert ubits(db.score(model)).tolist() == ubits(want).tolist(), ("msv", leng)
