"""GPU parity tests proper (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(include/msv_cuda.h), either directly (hmm_fasta_viterbi_b200._cabi) or via the C++ MSV_HMM class that sits on it.
The bar is bit-exact fp32 (tolerance 0 ULP): each cell performs the reference's add on the reference's operands, and
max is exact, so any difference is a bug."""
import os
import subprocess

import numpy as np
import pytest

import hmm_fasta_viterbi_b200 as msv
from conftest import REPO, fasta_path, hmm_path, model_files
from hmm_fasta_viterbi_b200 import _cabi
from oracle_lib import LETTERS, pack

pytestmark = pytest.mark.gpu
CORES = os.cpu_count() or 1


def ubits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def bits(x) -> str:
    return format(int(np.float32(x).view(np.uint32)), "08x")


def device_model(oracle, name, device=0):
    h = oracle.load_hmm(hmm_path(name))
    table, tr3 = oracle.prepare(h["match_emissions"])
    mine = _cabi.emission_table(h["match_emissions"])
    assert ubits(mine).tolist() == ubits(table).tolist()
    return msv.Model(mine, *_cabi.model_transitions(h["model_length"]), device=device), table, tr3


def random_db(rng, n, lo, hi, background=None):
    lens = rng.integers(lo, hi + 1, size=n)
    if background is None:
        seqs = [rng.integers(0, 20, size=int(k), dtype=np.uint8) for k in lens]
    else:
        seqs = [rng.choice(20, size=int(k), p=background).astype(np.uint8) for k in lens]
    return seqs, *pack(seqs)


# ---- configuration 1 and 2 inputs: every fixture model against the committed reference bits ---------------------
@pytest.mark.parametrize("name", model_files())
def test_golden_scores_single_sequence_api(golden_scores, golden_readers, name):
    """MSV_HMM::parallel_run_on_sequence (reference MSV_HMM.hpp:23), both values of should_specialize."""
    model = msv.MSV_HMM(msv.Profile_HMM(hmm_path(name)))
    g = golden_scores["scores"][name]
    seqs = {"example": golden_readers["fasta"]["fasta_like_example.fsa"], "random": golden_readers["fasta"]["random_FASTA.fsa"],
            "extra": golden_scores["meta"]["extra_sequences"]}
    for key, want in g.items():
        assert [bits(model.parallel_run_on_sequence(s)) for s in seqs[key]] == want, (name, key)
        assert [bits(model.parallel_run_on_sequence(s, True)) for s in seqs[key]] == want, (name, key, "specialised")
    # the batch entry point returns the same bits in input order
    everything = seqs["example"] + seqs["random"] + seqs["extra"]
    packed = msv.Packed_sequences.from_arrays(*pack([np.array([LETTERS.index(c) for c in s[1:]], np.uint8) for s in everything]))
    got = model.parallel_run_on_sequences(packed)
    assert [bits(v) for v in got] == g["example"] + g["random"] + g["extra"]
    # seq (host) == par (device), the invariant of the reference's test_MSV.cpp:26 with tolerance 0
    assert [bits(model.run_on_sequence(s)) for s in seqs["example"]] == g["example"]


# ---- batches against the oracle -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,hi", [("100.hmm", 6000, 400), ("500.hmm", 3000, 400), ("1001.hmm", 1500, 400),
                                        ("1400.hmm", 1500, 500), ("2405.hmm", 800, 500)])
def test_batch_matches_oracle(oracle, name, n, hi):
    model, table, tr3 = device_model(oracle, name)
    rng = np.random.default_rng(int(name.split(".")[0]))
    seqs, codes, offsets = random_db(rng, n, 0, hi)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    got = model.score_batch(codes, offsets)
    assert ubits(got).tolist() == ubits(want).tolist()
    # resident database path gives the same bits, twice (the work queue is reset between launches)
    db = msv.Database(codes, offsets)
    assert db.info() == {"n": n, "total_residues": int(offsets[-1]), "longest": int(np.diff(offsets.astype(np.int64)).max())}
    assert ubits(db.score(model)).tolist() == ubits(want).tolist()
    assert ubits(db.score(model)).tolist() == ubits(want).tolist()


def test_homologous_sequences_exercise_the_J_state(oracle):
    """Sequences sampled from the model itself score high and make J (multi-hit) win over N in B (MSV_HMM.cpp:110)."""
    name = "300.hmm"
    model, table, tr3 = device_model(oracle, name)
    h = oracle.load_hmm(hmm_path(name))
    rng = np.random.default_rng(5)
    cons = h["match_emissions"][1:].argmax(axis=1).astype(np.uint8)
    seqs = []
    for _ in range(200):
        a, b = sorted(rng.integers(0, len(cons), size=2))
        noise = rng.integers(0, 20, size=int(rng.integers(0, 50)), dtype=np.uint8)
        seqs.append(np.concatenate([noise, cons[a:b], noise[::-1], cons[a // 2:b]]))
    codes, offsets = pack(seqs)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert (want > 0).sum() > 50
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()


@pytest.mark.parametrize("geometry", ["4,8", "4,28", "4,52", "8,4", "8,16", "8,64", "16,8", "16,32", "16,88", "32,4", "32,20", "32,24", "32,44",
                                       "32,56", "32,60", "32,88",
                                       # lane groups whose two highest columns are a pair read with LDS.64 (K % 4 == 2)
                                       "4,6", "4,26", "4,38", "4,54", "8,6", "8,14", "8,26", "8,38", "8,54",
                                       # warp kernels: shared memory + KT tensor-memory columns per lane
                                       "32,4,0", "32,8,0", "32,8,8", "32,20,8", "32,24,16", "32,28,16", "32,44,0", "32,44,8",
                                       "32,44,16", "32,44,24", "32,44,16,640", "32,64,16", "32,76,16", "32,76,24", "32,88,16",
                                       # warp kernels with the tensor-memory columns loaded a row ahead (variant 1)
                                       "32,8,8,0,1", "32,16,16,0,1", "32,28,16,0,1", "32,36,24,0,1", "32,44,16,0,1",
                                       "32,60,24,0,1", "32,76,24,0,1", "32,88,24,0,1",
                                       # columns per lane in steps of two (tensor part 6/10/14/18 = sums of 16, 8, 4, 2)
                                       "32,6,6,0,1", "32,10,10,0,1", "32,14,14,0,1", "32,18,18,0,1", "32,22,18,0,1", "32,42,18,0,1",
                                       "32,58,18,0,1",
                                       # quad kernels: four warps (128 lanes) per sequence
                                       "128,4,0", "128,8,8", "128,12,8", "128,16,16", "128,20,16", "128,28,16", "128,36,16",
                                       "128,40,24", "128,44,24"])
def test_every_kernel_geometry(oracle, geometry, monkeypatch):
    """Lanes-per-sequence x columns-per-lane (x tensor-memory columns) variants, forced through MSV_CUDA_GEOMETRY, all
    give reference bits."""
    parts = [int(v) for v in geometry.split(",")]
    G, K = parts[0], parts[1]
    KT = parts[2] if len(parts) > 2 else -1
    limit = G * K - 1  # every kernel needs one padding column at the end of the row
    fits = [n for n in model_files() if int(n.split(".")[0]) <= limit]
    monkeypatch.setenv("MSV_CUDA_GEOMETRY", geometry)
    if fits:
        name = fits[-1]
        model, table, tr3 = device_model(oracle, name)
    else:  # no fixture is this short (8 lanes x 4 columns): a synthetic model that fills the geometry
        name = f"{limit}.synthetic"
        rng = np.random.default_rng(limit)
        match = np.zeros((limit + 1, 20), np.float32)
        match[1:] = rng.dirichlet(np.full(20, 0.4), size=limit).astype(np.float32)
        table, tr3 = oracle.prepare(match)
        mine = _cabi.emission_table(match)
        assert ubits(mine).tolist() == ubits(table).tolist()
        model = msv.Model(mine, *_cabi.model_transitions(limit + 1))
    geo = model.geometry
    assert (geo["lanes_per_sequence"], geo["columns_per_lane"], geo["tensor_columns_per_lane"]) == (G, K, KT)
    rng = np.random.default_rng(G * 1000 + K)
    n = max(64, min(4000, int(4e7 / (int(name.split(".")[0]) * 150))))
    seqs, codes, offsets = random_db(rng, n, 0, 300)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()


def test_general_E_transitions_keep_C_and_J_apart(oracle):
    """tr_E_C != tr_E_J selects the kernel variant that carries C separately (the reference always has them equal)."""
    h = oracle.load_hmm(hmm_path("1400.hmm"))
    table, tr3 = oracle.prepare(h["match_emissions"])
    tr3 = tr3.copy()
    tr3[1] = np.float32(-1.25)  # E -> C
    tr3[2] = np.float32(-0.40)  # E -> J
    model = msv.Model(table, tr3[0], tr3[1], tr3[2])
    rng = np.random.default_rng(9)
    seqs, codes, offsets = random_db(rng, 600, 0, 300)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()


def test_all_models_default_geometry_small_batch(oracle):
    rng = np.random.default_rng(11)
    seqs, codes, offsets = random_db(rng, 300, 0, 200)
    for name in model_files():
        model, table, tr3 = device_model(oracle, name)
        want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
        assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist(), name


def synthetic_model(rng, leng):
    """A random model of `leng` columns in the reference's table layout [20][leng+1] (column 0 = -inf)."""
    match = rng.dirichlet(np.full(20, 0.3), size=leng + 1).astype(np.float32)
    match[0] = 0.0
    return match


@pytest.mark.parametrize("leng", [2816, 3000, 5631])
def test_models_longer_than_one_warp(oracle, leng):
    """LENG > 2815 does not fit one warp's registers: the four-warps-per-sequence kernel scans it, with the emission
    table distributed over shared memory and tensor memory (20 x 5631 x 4 B = 450 KB at the upper limit)."""
    rng = np.random.default_rng(leng)
    match = synthetic_model(rng, leng)
    table, tr3 = oracle.prepare(match)
    model = msv.Model(_cabi.emission_table(match), *_cabi.model_transitions(leng + 1))
    assert model.geometry["lanes_per_sequence"] == 128
    seqs, codes, offsets = random_db(rng, 700 if leng < 5000 else 400, 0, 250)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()


@pytest.mark.parametrize("leng,seed", [(1, 0), (2, 1), (31, 2), (32, 3), (33, 4), (127, 5), (128, 6), (129, 7), (447, 8), (448, 9), (1023, 10)])
def test_synthetic_models_with_impossible_emissions(oracle, leng, seed):
    """Random models whose match emissions contain exact zeros (log-odds -inf) and ones, at lengths that sit on the
    kernel-geometry boundaries; the eight-lane, warp and four-warp plans all have to propagate -inf exactly."""
    rng = np.random.default_rng(seed)
    match = synthetic_model(rng, leng)
    match[1:][rng.random((leng, 20)) < 0.15] = 0.0   # impossible residues
    match[1:][rng.random((leng, 20)) < 0.02] = 1.0   # "*" fields of a .hmm file parse as probability 1.0
    table, tr3 = oracle.prepare(match)
    model = msv.Model(_cabi.emission_table(match), *_cabi.model_transitions(leng + 1))
    seqs, codes, offsets = random_db(rng, 400, 0, 120)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()
    big = msv.Packed_sequences.synthetic_swissprot_like(30_000, seed)     # large enough for the bulk / eight-lane plans
    got = msv.Database(big.residues, big.offsets).score(model)
    sample = rng.choice(len(big), size=120, replace=False)
    off = big.offsets
    sc, so = pack([big.residues[int(off[q]):int(off[q + 1])] for q in sample])
    want = oracle.score_batch(table, tr3, sc, so, threads=CORES)
    assert ubits(got[sample]).tolist() == ubits(want).tolist()


def test_model_beyond_on_chip_capacity_is_refused(oracle):
    rng = np.random.default_rng(1)
    match = synthetic_model(rng, 5700)
    with pytest.raises(_cabi.MsvCudaError) as err:
        msv.Model(_cabi.emission_table(match), *_cabi.model_transitions(5701))
    assert err.value.status == _cabi.MSV_ERR_MODEL_TOO_LONG


def test_launch_planner_choices(oracle):
    """The planner picks the kernel family and occupancy from the shape of the database (msv_cuda_model_plan)."""
    big, _, _ = device_model(oracle, "1400.hmm")
    short, _, _ = device_model(oracle, "300.hmm")
    one = msv.Database(np.zeros(3500, np.uint8), np.array([0, 3500], np.uint64))
    assert big.plan(one) == {"lanes_per_sequence": 128, "sequences_per_cta": 1}          # one sequence: four warps, one SM
    many = msv.Packed_sequences.synthetic_swissprot_like(150_000, 1)
    many_db = msv.Database(many.residues, many.offsets)
    assert big.plan(many_db) == {"lanes_per_sequence": 32, "sequences_per_cta": 16}      # bulk: a warp each, 16 warps per SM
    assert short.plan(many_db)["lanes_per_sequence"] == 8                                # short model, many sequences: 8 lanes each
    shortest, _, _ = device_model(oracle, "100.hmm")
    plan = shortest.plan(many_db)                                                        # the shortest models: 4 lanes each, and
    assert plan["lanes_per_sequence"] == 4 and 64 <= plan["sequences_per_cta"] < 192     # fewer slots than the maximum at this size
    titin = msv.Packed_sequences.synthetic_long_uniform(2048, 2405, 10_000, 35_000)     # config 5: fewer sequences than warp slots
    plan = big.plan(msv.Database(titin.residues, titin.offsets))
    assert plan["lanes_per_sequence"] == 32 and plan["sequences_per_cta"] in (8, 12)     # a warp each at reduced occupancy
    few_short = msv.Packed_sequences.synthetic_swissprot_like(500, 2)
    assert big.plan(msv.Database(few_short.residues, few_short.offsets))["lanes_per_sequence"] == 128


def test_few_long_sequences_use_four_warps_each(oracle):
    """Below two sequences per warp slot the library switches to the quad plan by itself; same bits either way."""
    model, table, tr3 = device_model(oracle, "1400.hmm")
    rng = np.random.default_rng(6)
    seqs = [rng.integers(0, 20, size=int(k), dtype=np.uint8) for k in rng.integers(500, 3000, size=40)]
    codes, offsets = pack(seqs)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()
    db = msv.Database(codes, offsets)
    assert ubits(db.score(model)).tolist() == ubits(want).tolist()


@pytest.mark.parametrize("name", ["100.hmm", "200.hmm", "300.hmm", "400.hmm"])
def test_long_sequences_on_fast_ctas(oracle, name, monkeypatch):
    """Lane-group plans (short models): the longest sequences of a database are handed to a few CTAs with fewer, faster slots
    (next_ticket in msv_device.cuh).  Same bits as the plain queue, as the oracle, through the resident and the pipelined
    end-to-end path (whose upload stages are planned one by one), and for a forced, lopsided split."""
    model, table, tr3 = device_model(oracle, name)
    packed = msv.Packed_sequences.synthetic_swissprot_like(120_000, 77)
    codes, offsets = packed.residues, packed.offsets
    plain_db = msv.Database(codes, offsets)
    assert "fast_ctas" not in model.plan(plain_db)  # an experiment that did not pay: off unless asked for
    plain = plain_db.score(model)
    monkeypatch.setenv("MSV_CUDA_FAST_CTAS", "auto")
    db = msv.Database(codes, offsets)  # (the length profile is taken when the database is created)
    plan = model.plan(db)
    assert plan["lanes_per_sequence"] in (4, 8) and plan.get("fast_ctas", 0) > 0, plan
    got = db.score(model)
    rng = np.random.default_rng(5)
    longest = np.argsort(np.diff(offsets.astype(np.int64)))[-40:]
    sample = np.concatenate([rng.choice(len(packed), size=400, replace=False), longest])
    sc, so = pack([codes[int(offsets[q]):int(offsets[q + 1])] for q in sample])
    want = oracle.score_batch(table, tr3, sc, so, threads=CORES)
    assert ubits(got[sample]).tolist() == ubits(want).tolist()
    assert (ubits(model.score_batch(codes, offsets)) == ubits(got)).all()
    monkeypatch.setenv("MSV_CUDA_FAST_CTAS", "40,4,128")  # 40 CTAs, one warp per scheduler, everything from 128 rows up is "long"
    assert (ubits(db.score(model)) == ubits(got)).all()
    monkeypatch.setenv("MSV_CUDA_FAST_CTAS", "1,8,2048")  # one fast CTA: the full CTAs have to help with the long ones at the end
    assert (ubits(db.score(model)) == ubits(got)).all()
    monkeypatch.delenv("MSV_CUDA_FAST_CTAS")
    assert (ubits(db.score(model)) == ubits(got)).all() and (ubits(plain) == ubits(got)).all()


def test_short_model_large_database_uses_eight_lanes_per_sequence(oracle):
    """300.hmm x 150k sequences: enough rows per slot for the eight-lane plan to be picked automatically."""
    model, table, tr3 = device_model(oracle, "300.hmm")
    packed = msv.Packed_sequences.synthetic_swissprot_like(150_000, 300)
    codes, offsets = packed.residues, packed.offsets
    got = msv.Database(codes, offsets).score(model)
    rng = np.random.default_rng(300)
    sample = rng.choice(len(packed), size=500, replace=False)
    sc, so = pack([codes[int(offsets[q]):int(offsets[q + 1])] for q in sample])
    want = oracle.score_batch(table, tr3, sc, so, threads=CORES)
    assert ubits(got[sample]).tolist() == ubits(want).tolist()
    assert (ubits(model.score_batch(codes, offsets)) == ubits(got)).all()


def test_resident_database_scanned_by_many_models(oracle):
    """Device_database: one upload, every fixture model scans it (the benchmark_MSV.cpp pattern); bits equal the
    host-buffer path and the oracle."""
    packed = msv.Packed_sequences.synthetic_swissprot_like(3000, 123)
    resident = msv.Device_database(packed)
    for name in ("100.hmm", "800.hmm", "1400.hmm", "2405.hmm"):
        model = msv.MSV_HMM(msv.Profile_HMM(hmm_path(name)))
        got = model.parallel_run_on_sequences(resident)
        assert (ubits(got) == ubits(model.parallel_run_on_sequences(packed))).all()
        h = oracle.load_hmm(hmm_path(name))
        table, tr3 = oracle.prepare(h["match_emissions"])
        want = oracle.score_batch(table, tr3, packed.residues[: int(packed.offsets[200])], packed.offsets[:201].copy(), threads=CORES)
        assert ubits(got[:200]).tolist() == ubits(want).tolist()


def test_msv_filter_keeps_homologs_and_drops_noise(oracle):
    """MSV_HMM::msv_filter: sequences built from the model's consensus pass the P <= 0.02 filter, almost all random
    sequences do not; the reported numbers equal an fp64 evaluation of the same formulas."""
    name = "300.hmm"
    prof = msv.Profile_HMM(hmm_path(name))
    model = msv.MSV_HMM(prof)
    rng = np.random.default_rng(17)
    cons = prof.match_emissions[1:].argmax(axis=1).astype(np.uint8)
    noise = [rng.integers(0, 20, size=int(k), dtype=np.uint8) for k in rng.integers(100, 600, size=2000)]
    homologs = [np.concatenate([rng.integers(0, 20, size=40, dtype=np.uint8), cons[20:280], rng.integers(0, 20, size=40, dtype=np.uint8)])
                for _ in range(20)]
    packed = msv.Packed_sequences.from_arrays(*pack(noise + homologs))
    hits = model.msv_filter(msv.Device_database(packed))
    found = set(int(i) for i in hits["index"])
    assert set(range(2000, 2020)) <= found          # every homolog passes
    assert len(found - set(range(2000, 2020))) < 120  # ~2 % of the noise is expected to pass by chance
    raw = model.parallel_run_on_sequences(packed).astype(np.float64)
    L = np.diff(packed.offsets.astype(np.int64)).astype(np.float64)
    bits = (raw - (L * np.log(L / (L + 1.0)) + np.log(1.0 / (L + 1.0)))) / np.log(2.0)
    ey = -np.exp(-float(prof.stats_local_msv_lambda) * (bits - float(prof.stats_local_msv_mu)))
    pv = np.where(np.abs(ey) < 5e-9, -ey, 1.0 - np.exp(ey))
    idx = hits["index"].astype(np.int64)
    np.testing.assert_allclose(hits["bits"], bits[idx], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(hits["p_value"], pv[idx], rtol=1e-5, atol=1e-12)
    assert (pv[idx] <= 0.02 * (1 + 1e-6)).all() and (np.delete(pv, idx) > 0.02 * (1 - 1e-6)).all()


def test_handles_release_device_memory(oracle):
    """Creating, using and destroying models / databases / workspaces repeatedly leaves the free device memory where
    it was (every plan's table, the workspace buffers, streams and events are released)."""
    import torch

    h = oracle.load_hmm(hmm_path("1400.hmm"))
    table = _cabi.emission_table(h["match_emissions"])
    tr = _cabi.model_transitions(h["model_length"])
    packed = msv.Packed_sequences.synthetic_swissprot_like(5000, 77)

    def cycle():
        model = msv.Model(table, *tr)
        model.score_batch(packed.residues, packed.offsets)
        db = msv.Database(packed.residues, packed.offsets)
        db.score(model)
        db.close()
        model.close()

    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    before = torch.cuda.mem_get_info()[0]
    for _ in range(60):
        cycle()
    torch.cuda.synchronize()
    after = torch.cuda.mem_get_info()[0]
    assert before - after < 8 * 2**20, (before, after)


def test_msv_scan_command_line(oracle):
    """build/msv_scan --all: the printed raw scores are the reference bits for the fixture FASTA file."""
    exe = os.path.join(REPO, "build", "msv_scan")
    if not os.path.exists(exe):
        pytest.skip("build/msv_scan not built")
    run = subprocess.run([exe, "--all", hmm_path("100.hmm"), hmm_path("1400.hmm"), fasta_path("fasta_like_example.fsa")],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    rows = [ln.split("\t") for ln in run.stdout.splitlines() if not ln.startswith("#")]
    assert len(rows) == 8
    h = oracle.load_hmm(hmm_path("100.hmm"))
    table, tr3 = oracle.prepare(h["match_emissions"])
    seqs = oracle.load_fasta(fasta_path("fasta_like_example.fsa"))
    for row, seq in zip(rows[:4], seqs):
        assert row[0] == "Pfam-B_229" and int(row[2]) == len(seq) - 1
        assert abs(float(row[3]) - float(oracle.score_string(table, tr3, seq))) < 1e-4 * abs(float(row[3]))
    missing = subprocess.run([exe, "/nonexistent.hmm", fasta_path("fasta_like_example.fsa")], capture_output=True, text=True)
    assert missing.returncode == 1


def test_host_register_round_trip(oracle):
    """msv_cuda_host_register / _unregister: uploads from a page-locked caller buffer give the same bits."""
    model, table, tr3 = device_model(oracle, "600.hmm")
    packed = msv.Packed_sequences.synthetic_swissprot_like(20_000, 31)
    codes = np.ascontiguousarray(packed.residues).copy()
    offsets = np.ascontiguousarray(packed.offsets).copy()
    plain = model.score_batch(codes, offsets)
    _cabi.check(_cabi.lib.msv_cuda_host_register(codes.ctypes.data, codes.nbytes))
    _cabi.check(_cabi.lib.msv_cuda_host_register(offsets.ctypes.data, offsets.nbytes))
    try:
        pinned = model.score_batch(codes, offsets)
    finally:
        _cabi.check(_cabi.lib.msv_cuda_host_unregister(codes.ctypes.data))
        _cabi.check(_cabi.lib.msv_cuda_host_unregister(offsets.ctypes.data))
    assert (ubits(plain) == ubits(pinned)).all()


def test_single_process_multi_device_driver(oracle):
    """MSV_HMM::parallel_run_on_sequences(db, devices): two slices scored concurrently (here both on GPU 0 when the box
    has one GPU) equal the single-call result."""
    model = msv.MSV_HMM(msv.Profile_HMM(hmm_path("900.hmm")))
    packed = msv.Packed_sequences.synthetic_swissprot_like(20_000, 9)
    one = model.parallel_run_on_sequences(packed)
    devices = [0, 1] if _cabi.device_count() > 1 else [0, 0]
    for gather in ("host", "peer", "nccl"):  # NCCL cannot put two ranks on one GPU: only where the box has several
        if gather == "nccl" and _cabi.device_count() < 3:
            continue
        two = model.parallel_run_on_sequences(packed, devices=devices + [0 if _cabi.device_count() < 3 else 2], gather=gather)
        assert (ubits(one) == ubits(two)).all(), gather


def test_multi_gpu_c_abi(oracle):
    """msv_cuda_multi_score_batch straight through the C ABI: the cell-balanced cut, one host thread per GPU, and the three
    ways of bringing the scores together give the bits of a single-GPU call; the gathered array stays on the first GPU."""
    import ctypes as C
    devices = list(range(min(_cabi.device_count(), 4)))
    if len(devices) == 1:
        devices = [0, 0, 0]  # three slices, one GPU: still exercises the partitioner, the threads and the peer-store path
    h = oracle.load_hmm(hmm_path("500.hmm"))
    table = _cabi.emission_table(h["match_emissions"])
    models = [msv.Model(table, *_cabi.model_transitions(h["model_length"]), device=d) for d in devices]
    packed = msv.Packed_sequences.synthetic_swissprot_like(30_000, 12)
    want = models[0].score_batch(packed.residues, packed.offsets)
    multi = _cabi.MultiGpu(models)
    modes = [_cabi.GATHER_HOST, _cabi.GATHER_PEER] + ([_cabi.GATHER_NCCL] if len(set(devices)) == len(devices) and len(devices) > 1 else [])
    for mode in modes:
        got = multi.score_batch(packed.residues, packed.offsets, gather=mode)
        assert (ubits(got) == ubits(want)).all(), mode
        if mode != _cabi.GATHER_HOST:
            ptr, n, dev = multi.gathered()
            assert n == len(packed) and dev == devices[0] and ptr
            back = np.empty(n, np.float32)
            C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(back.ctypes.data), C.c_void_p(ptr), C.c_size_t(4 * n), 2)
            assert (ubits(back) == ubits(want)).all()
    with pytest.raises(_cabi.MsvCudaError):
        multi.score_batch(packed.residues, packed.offsets, gather=7)
    assert multi.score_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64)).size == 0
    multi.close()


def test_filter_statistics_bits_and_pvalues(oracle):
    """Bit score and Gumbel P-value of the MSV filter (floating point, tolerance 1e-6 relative vs fp64 numpy)."""
    import torch

    prof = msv.Profile_HMM(hmm_path("500.hmm"))
    model, table, tr3 = device_model(oracle, "500.hmm")
    packed = msv.Packed_sequences.synthetic_swissprot_like(5000, 55)
    db = msv.Database(packed.residues, packed.offsets)
    scores = torch.empty(len(packed), dtype=torch.float32, device="cuda")
    bits_d, p_d = torch.empty_like(scores), torch.empty_like(scores)
    stream = torch.cuda.current_stream().cuda_stream
    db.score_device(model, scores, stream)
    db.filter_device(scores, prof.stats_local_msv_mu, prof.stats_local_msv_lambda, bits_d, p_d, stream)
    torch.cuda.synchronize()
    raw = scores.cpu().numpy().astype(np.float64)
    L = np.diff(packed.offsets.astype(np.int64)).astype(np.float64)
    null1 = L * np.log(L / (L + 1.0)) + np.log(1.0 / (L + 1.0))
    bits = (raw - null1) / np.log(2.0)
    ey = -np.exp(-float(prof.stats_local_msv_lambda) * (bits - float(prof.stats_local_msv_mu)))
    pv = np.where(np.abs(ey) < 5e-9, -ey, 1.0 - np.exp(ey))
    np.testing.assert_allclose(bits_d.cpu().numpy(), bits, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(p_d.cpu().numpy(), pv, rtol=1e-5, atol=1e-12)
    assert 0.0 <= pv.min() and pv.max() <= 1.0


# ---- edge cases --------------------------------------------------------------------------------------------------
def test_edge_cases(oracle):
    model, table, tr3 = device_model(oracle, "1400.hmm")
    # empty database
    assert model.score_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64)).size == 0
    # only empty sequences: score is -inf (tr_loop = log(0) = -inf, MSV_HMM.cpp:59-64)
    got = model.score_batch(np.zeros(0, np.uint8), np.zeros(6, np.uint64))
    assert np.isneginf(got).all() and got.size == 5
    # one residue, and ragged mixtures with empties in between
    seqs = [np.array([3], np.uint8), np.zeros(0, np.uint8), np.arange(20, dtype=np.uint8), np.zeros(0, np.uint8),
            np.full(1000, 19, np.uint8), np.array([0, 19], np.uint8)]
    codes, offsets = pack(seqs)
    want = oracle.score_batch(table, tr3, codes, offsets)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()
    assert bits(model.score_sequence(seqs[4])) == bits(want[4])
    assert np.isneginf(model.score_sequence(np.zeros(0, np.uint8)))


def test_bad_residue_code_is_reported_not_scored(oracle):
    model, _, _ = device_model(oracle, "100.hmm")
    codes = np.array([1, 2, 3, 20, 4], np.uint8)
    with pytest.raises(_cabi.MsvCudaError) as err:
        model.score_batch(codes, np.array([0, 5], np.uint64))
    assert err.value.status == _cabi.MSV_ERR_BAD_RESIDUE and "position 3" in str(err.value)
    with pytest.raises(_cabi.MsvCudaError):
        msv.Database(np.array([255], np.uint8), np.array([0, 1], np.uint64))
    with pytest.raises(_cabi.MsvCudaError):
        model.score_batch(codes, np.array([0, 3, 2], np.uint64))  # non-monotonic offsets
    with pytest.raises(KeyError):
        msv.MSV_HMM(msv.Profile_HMM(hmm_path("100.hmm"))).parallel_run_on_sequence("#ACDZ")


def test_bad_residue_at_an_upload_stage_boundary(oracle):
    """msv_cuda_score_batch scans stage s while stage s+1 is still uploading.  The last sequence of a stage prefetches the
    tensor-memory row of "the residue after its last one" -- the first byte of the next stage, which may be unvalidated
    (here: code 255).  It must come back as MSV_ERR_BAD_RESIDUE, not as a fault, and the workspace must be reusable."""
    model, table, tr3 = device_model(oracle, "1400.hmm")
    packed = msv.Packed_sequences.synthetic_swissprot_like(60_000, 5)
    codes, offsets = packed.residues.copy(), packed.offsets.copy()
    good = model.score_batch(codes, offsets)
    # the first cut of the pipelined upload is the first sequence boundary at or after 2 MiB
    q = int(np.searchsorted(offsets, 2 << 20))
    for position in (int(offsets[q]), int(offsets[q]) + 1, int(offsets[-1]) - 1):
        broken = codes.copy()
        broken[position] = 255
        with pytest.raises(_cabi.MsvCudaError) as err:
            model.score_batch(broken, offsets)
        assert err.value.status == _cabi.MSV_ERR_BAD_RESIDUE and f"position {position}" in str(err.value)
        again = model.score_batch(codes, offsets)  # same workspace, right after the failed batch
        assert (ubits(again) == ubits(good)).all()
    sample = np.arange(q - 3, q + 3)
    sc, so = pack([codes[int(offsets[i]):int(offsets[i + 1])] for i in sample])
    assert ubits(good[sample]).tolist() == ubits(oracle.score_batch(table, tr3, sc, so, threads=CORES)).tolist()


def test_long_sequence_and_misaligned_offsets(oracle):
    """One titin-like sequence (config 5 shape, shortened) plus neighbours that start at every byte alignment."""
    model, table, tr3 = device_model(oracle, "2405.hmm")
    rng = np.random.default_rng(2405)
    seqs = [rng.integers(0, 20, size=k, dtype=np.uint8) for k in (1, 2, 3, 5, 12001, 7, 1, 6, 9, 4)]
    codes, offsets = pack(seqs)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()


def test_pipelined_upload_sizes_and_reuse(oracle):
    """msv_cuda_score_batch uploads in stages that overlap with the scan; the workspace is reused and regrown between
    calls.  Every size must equal the resident-database path bit for bit (and the oracle on a sample)."""
    model, table, tr3 = device_model(oracle, "700.hmm")
    rng = np.random.default_rng(21)
    for n in (1, 7, 1000, 60_000, 300, 120_000, 5):
        packed = msv.Packed_sequences.synthetic_swissprot_like(n, 1000 + n)
        codes, offsets = packed.residues, packed.offsets
        got = model.score_batch(codes, offsets)
        resident = msv.Database(codes, offsets).score(model)
        assert (ubits(got) == ubits(resident)).all(), n
        sample = rng.choice(n, size=min(n, 50), replace=False)
        sc, so = pack([codes[int(offsets[q]):int(offsets[q + 1])] for q in sample])
        want = oracle.score_batch(table, tr3, sc, so, threads=CORES)
        assert ubits(got[sample]).tolist() == ubits(want).tolist(), n


# ---- size-independent properties at BASELINE sizes ----------------------------------------------------------------
def test_full_size_properties_config4(oracle):
    """1400.hmm x 1M synthetic sequences (config 4): permutation equivariance, duplicate consistency, run-to-run
    determinism, and a seeded sample checked against the oracle."""
    model, table, tr3 = device_model(oracle, "1400.hmm")
    db = msv.Packed_sequences.synthetic_swissprot_like(1_000_000, 20261018)
    codes, offsets = db.residues, db.offsets
    first = model.score_batch(codes, offsets)
    second = msv.Database(codes, offsets).score(model)
    assert (ubits(first) == ubits(second)).all()
    assert np.isfinite(first).all()
    rng = np.random.default_rng(4)
    sample = rng.choice(len(db), size=400, replace=False)
    seqs = [codes[int(offsets[q]):int(offsets[q + 1])] for q in sample]
    sc, so = pack(seqs)
    want = oracle.score_batch(table, tr3, sc, so, threads=CORES)
    assert ubits(first[sample]).tolist() == ubits(want).tolist()
    # a reversed copy of a slice scores identically, sequence by sequence
    sl = np.arange(200_000, 230_000)[::-1]
    rc, ro = pack([codes[int(offsets[q]):int(offsets[q + 1])] for q in sl])
    assert (ubits(model.score_batch(rc, ro)) == ubits(first[sl])).all()


@pytest.mark.timeout(1500)
def test_full_size_exhaustive_config4(oracle):
    """Slow (about a minute and a half of host time on 16 threads): EVERY one of the 1 000 000 scores of the bench workload
    (1400.hmm x the seed-20261018 database, through the end-to-end call) against the reference's own compiled
    run_on_sequence (oracle/_ref), or the C restatement where that library is absent.  0 ULP."""
    from oracle_lib import RefLib
    model, table, tr3 = device_model(oracle, "1400.hmm")
    packed = msv.Packed_sequences.synthetic_swissprot_like(1_000_000, 20261018)
    got = model.score_batch(packed.residues, packed.offsets)
    if RefLib.available():
        want = RefLib().model(hmm_path("1400.hmm")).run_batch(packed.residues, packed.offsets, CORES)
    else:
        want = oracle.score_batch(table, tr3, packed.residues, packed.offsets, threads=CORES)
    assert int((ubits(got) != ubits(want)).sum()) == 0


def test_full_size_sample_parity_config3(oracle):
    """Every fixture model x 100 000 synthetic sequences (config 3), default plan selection: a seeded sample of each
    scan is compared with the oracle, and the resident and host-buffer paths agree on all 100 000 scores."""
    packed = msv.Packed_sequences.synthetic_swissprot_like(100_000, 1400)
    codes, offsets = packed.residues, packed.offsets
    resident = msv.Database(codes, offsets)
    rng = np.random.default_rng(3)
    sample = rng.choice(len(packed), size=48, replace=False)
    sc, so = pack([codes[int(offsets[q]):int(offsets[q + 1])] for q in sample])
    for name in model_files():
        model, table, tr3 = device_model(oracle, name)
        got = resident.score(model)
        want = oracle.score_batch(table, tr3, sc, so, threads=CORES)
        assert ubits(got[sample]).tolist() == ubits(want).tolist(), name
        if name in ("100.hmm", "400.hmm", "900.hmm", "2405.hmm"):
            assert (ubits(model.score_batch(codes, offsets)) == ubits(got)).all(), name
        model.close()


def test_full_size_properties_config5(oracle):
    """2405.hmm x long sequences (config 5 shape, 256 sequences): determinism and a sample against the oracle."""
    model, table, tr3 = device_model(oracle, "2405.hmm")
    db = msv.Packed_sequences.synthetic_long_uniform(256, 2405, 10_000, 35_000)
    codes, offsets = db.residues, db.offsets
    first = model.score_batch(codes, offsets)
    assert (ubits(first) == ubits(model.score_batch(codes, offsets))).all()
    sample = [0, 17, 255]
    sc, so = pack([codes[int(offsets[q]):int(offsets[q + 1])] for q in sample])
    want = oracle.score_batch(table, tr3, sc, so, threads=CORES)
    assert ubits(first[sample]).tolist() == ubits(want).tolist()


# ---- the reference's own programs, unchanged, against this implementation -----------------------------------------
@pytest.mark.parametrize("prog,cwd", [("test_hmm_parsing", "data_readers"), ("test_fasta_parsing", "data_readers"),
                                       ("test_MSV", "algorithms"), ("benchmark_MSV_1400", "algorithms")])
def test_reference_programs_run_unchanged(prog, cwd):
    exe = os.path.join(REPO, "build", cwd, prog)
    if not os.path.exists(exe):
        pytest.skip("build/ not populated (tools/build_reference_programs.sh needs /root/reference at build time)")
    run = subprocess.run([exe], cwd=os.path.join(REPO, "build", cwd), capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "failed" not in run.stdout


# ---- scan fused with the gather (msv_cuda_db_score_gather) ---------------------------------------------------------------
def test_score_gather_writes_every_copy(oracle):
    """One GPU standing in for several: three 'copies' of the gathered array on the same device.  Every copy must hold
    this shard's scores at first_index.. and stay untouched elsewhere."""
    import torch
    model, table, tr3 = device_model(oracle, "1400.hmm")
    rng = np.random.default_rng(77)
    seqs, codes, offsets = random_db(rng, 3000, 0, 300)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    db = msv.Database(codes, offsets)
    first, total = 1234, 6000
    copies = [torch.full((total,), float("nan"), dtype=torch.float32, device="cuda") for _ in range(3)]
    db.score_gather(model, copies, first)
    torch.cuda.synchronize()
    for c in copies:
        got = c.cpu().numpy()
        assert ubits(got[first:first + 3000]).tolist() == ubits(want).tolist()
        assert np.isnan(got[:first]).all() and np.isnan(got[first + 3000:]).all()
    with pytest.raises(msv.MsvCudaError):
        db.score_gather(model, copies * 3, 0)  # more than 8 copies


def test_fused_gather_two_gpus(tmp_path):
    """torchrun x 2: each rank scans its shard and stores straight into both ranks' symmetric buffers; the result must
    equal the NCCL all-gather of the plain scan on every rank (bench.py measures both for N > 1)."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run(
        ["python", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29731", os.path.join(REPO, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3", "--sequences", "50000"],
        capture_output=True, text=True, cwd=REPO, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and "peer_gather_unavailable" not in line, line.get("peer_gather_unavailable")
    assert line["scaling"] == "strong" and line["warmup"] == 3
    assert line["gather"].startswith("fused into the scan kernel")
    assert line["nccl_gather"]["same_bits_as_fused_gather_on_every_rank"] is True
    # the end-to-end leg leaves the WHOLE job's scores in one host buffer on rank 0, checked against the CPU reference
    assert line["parity"]["e2e_job_buffer_has_no_gaps"] and line["parity"]["e2e_job_buffer_equals_own_device_scan"]
    checked = [v for k, v in line["parity"].items() if k.startswith("gathered_job_buffer_vs_")]
    assert checked and checked[0]["mismatches"] == 0 and checked[0]["checked"] >= 1000
    # ... and the single-process driver behind the C ABI (msv_cuda_multi_score_batch) gives the same bits in every gather mode
    # both forms of the gather (per-sequence stores from the scan kernel / a push kernel behind it) leave the whole job on every rank
    for form in ("stores", "push"):
        assert line["gather_forms"][form]["every_rank_holds_the_whole_job"] is True, line["gather_forms"]
    single = line["single_process_multi_gpu"]
    assert single["same_bits_as_multi_process_job_buffer"] is True, single
    for mode in ("peer", "nccl"):
        assert single[mode].get("same_bits_as_first_mode") is True and single["host"]["e2e_gcups"] > 0, single


# ---- speculative rows (B = N + move while J <= N) and their exact fallback ---------------------------------------------
@pytest.mark.parametrize("name", ["100.hmm", "400.hmm", "1400.hmm", "2405.hmm"])
def test_speculation_falls_back_exactly_on_hits(oracle, name, monkeypatch):
    """The warp kernel assumes J <= N (no hit worth more than the entry cost) and verifies once per sequence.  Sequences
    built from the model's consensus -- one, two or three strong segments, i.e. J far above N and re-entry through J --
    must take the exact path; random ones must not need it; both must give the oracle's bits, with and without the
    speculation compiled in."""
    h = oracle.load_hmm(hmm_path(name))
    model, table, tr3 = device_model(oracle, name)
    leng = h["model_length"] - 1
    consensus = np.argmax(h["match_emissions"][1:], axis=1).astype(np.uint8)
    rng = np.random.default_rng(leng)
    seqs = []
    for q in range(6000):  # enough sequences for the launch planner to take the warp-per-sequence (or lane-group) plan
        kind = q % 5
        if kind == 0:
            seqs.append(rng.integers(0, 20, size=int(rng.integers(0, 400)), dtype=np.uint8))
        else:
            parts = []
            for _ in range(kind if kind < 4 else 1):
                a = int(rng.integers(0, max(1, leng - 8)))
                b = int(rng.integers(a + 4, min(leng, a + 120) + 1))
                parts += [consensus[a:b], rng.integers(0, 20, size=int(rng.integers(0, 40)), dtype=np.uint8)]
            if kind == 4:  # a borderline hit: a short consensus stretch with mutations
                seg = parts[0].copy()
                seg[rng.random(seg.size) < 0.4] = rng.integers(0, 20)
                parts[0] = seg
            seqs.append(np.concatenate(parts))
    for q in range(12):  # longer than the kernel is willing to speculate on (4096 rows), with and without a hit inside
        long_one = rng.integers(0, 20, size=int(rng.integers(4090, 7000)), dtype=np.uint8)
        if q % 2:
            a = int(rng.integers(0, max(1, leng - 60)))
            long_one[1000:1000 + min(60, leng - a)] = consensus[a:a + 60]
        seqs.append(long_one)
    codes, offsets = pack(seqs)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert (want > 0).mean() > 0.4  # most of these really are hits
    db = msv.Database(codes, offsets)
    assert ubits(db.score(model)).tolist() == ubits(want).tolist()
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()
    monkeypatch.setenv("MSV_CUDA_NO_SPECULATION", "1")
    assert ubits(db.score(model)).tolist() == ubits(want).tolist()
    monkeypatch.delenv("MSV_CUDA_NO_SPECULATION")
    # both speculating kernels (whole sequences / checkpointed blocks of 64 rows) and the one that never speculates
    for mode in ("whole", "blocks", "none"):
        monkeypatch.setenv("MSV_CUDA_SPECULATION", mode)
        assert ubits(db.score(model)).tolist() == ubits(want).tolist(), mode
        assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist(), mode
    monkeypatch.delenv("MSV_CUDA_SPECULATION")
    # a database of long sequences only (mean length above the launch-level threshold) takes the block-wise kernel
    long_codes, long_offsets = pack(seqs[-12:])
    assert ubits(model.score_batch(long_codes, long_offsets)).tolist() == ubits(want[-12:]).tolist()
    # feedback: after enough sequences of this hit-rich database the library switches to exact rows by itself (they also count
    # the sequences that would have failed the speculation) ...
    if model.plan(db)["lanes_per_sequence"] == 32 and model.geometry["columns_per_lane"] <= 44:
        for _ in range(3):
            assert ubits(db.score(model)).tolist() == ubits(want).tolist()
        state = model.speculation
        assert state["scanned"] >= 4096 and state["failed"] > state["scanned"] // 4 and state["rows_next"] == "exact", state
        # ... and back to speculation on whole sequences on a database without hits
        noise = [rng.integers(0, 20, size=int(rng.integers(50, 400)), dtype=np.uint8) for _ in range(5000)]
        noise_codes, noise_offsets = pack(noise)
        noise_want = oracle.score_batch(table, tr3, noise_codes, noise_offsets, threads=CORES)
        noise_db = msv.Database(noise_codes, noise_offsets)
        for _ in range(3):
            assert ubits(noise_db.score(model)).tolist() == ubits(noise_want).tolist()
        assert model.speculation["rows_next"] == "whole", model.speculation


@pytest.mark.parametrize("name,geometry", [("100.hmm", "8,16"), ("200.hmm", "8,28"), ("100.hmm", "4,28"), ("200.hmm", "4,52"),
                                           ("100.hmm", "4,26"), ("100.hmm", "8,14"), ("200.hmm", "8,26"), ("300.hmm", "8,38"),
                                           ("400.hmm", "8,52")])
def test_group_speculation_and_its_exact_pass(oracle, name, geometry, monkeypatch):
    """Eight / four lanes per sequence, speculative rows: a sequence whose speculation fails (consensus-derived hits) is
    scanned again at once, exactly, by the same lane group inside the same launch; while it is, the other groups of its warp
    execute exact rows too."""
    monkeypatch.setenv("MSV_CUDA_GEOMETRY", geometry)
    h = oracle.load_hmm(hmm_path(name))
    model, table, tr3 = device_model(oracle, name)
    assert model.geometry["lanes_per_sequence"] == int(geometry.split(",")[0])
    leng = h["model_length"] - 1
    consensus = np.argmax(h["match_emissions"][1:], axis=1).astype(np.uint8)
    rng = np.random.default_rng(leng + 1)
    seqs = []
    for q in range(6000):  # >= 4096 sequences: below that the launch does not speculate
        if q % 3 == 0:
            a = int(rng.integers(0, leng - 20))
            b = int(rng.integers(a + 10, min(leng, a + 90) + 1))
            seqs.append(np.concatenate([rng.integers(0, 20, size=int(rng.integers(0, 30)), dtype=np.uint8), consensus[a:b],
                                        rng.integers(0, 20, size=int(rng.integers(0, 30)), dtype=np.uint8)]))
        else:
            seqs.append(rng.integers(0, 20, size=int(rng.integers(0, 350)), dtype=np.uint8))
    codes, offsets = pack(seqs)
    want = oracle.score_batch(table, tr3, codes, offsets, threads=CORES)
    assert 0.2 < (want > 0).mean() < 0.5
    db = msv.Database(codes, offsets)
    _cabi.launch_count(reset=True)
    assert ubits(db.score(model)).tolist() == ubits(want).tolist()
    assert _cabi.launch_count() == 1  # one launch: sequences whose speculation fails are repeated exactly inside it
    assert ubits(db.score(model)).tolist() == ubits(want).tolist()  # counters are reset between scans
    assert ubits(model.score_batch(codes, offsets)).tolist() == ubits(want).tolist()
    monkeypatch.setenv("MSV_CUDA_NO_SPECULATION", "1")
    _cabi.launch_count(reset=True)
    assert ubits(db.score(model)).tolist() == ubits(want).tolist()
    assert _cabi.launch_count() == 1


# ---- the single-sequence wavefront kernel (msv_wave_kernels.cuh) behind msv_cuda_score_sequence -------------------------
@pytest.mark.parametrize("name,wave_k", [("100.hmm", None), ("500.hmm", None), ("1400.hmm", None), ("1400.hmm", "2"), ("1400.hmm", "4"),
                                         ("1400.hmm", "6"), ("1400.hmm", "8"), ("1400.hmm", "12"), ("1400.hmm", "16"), ("2405.hmm", None),
                                         ("2405.hmm", "4")])
def test_single_sequence_wavefront_kernel(oracle, name, wave_k, monkeypatch):
    """One sequence per call through the chain of warps: every chunk boundary (lengths around multiples of 4 and of the
    16-chunk ring), the longest sequence that rides in the kernel parameters and the first that does not, long sequences
    that wrap the ring many times; random sequences (speculation holds), consensus-derived hits (it fails: the exact kernel
    re-scores) and everything in between -- reference bits every time."""
    if wave_k:  # the chain kernel with this many columns per lane; otherwise the default, the diagonal-worker kernel
        monkeypatch.setenv("MSV_CUDA_WAVE_K", wave_k)
        monkeypatch.setenv("MSV_CUDA_NO_DIAGONAL", "1")
    h = oracle.load_hmm(hmm_path(name))
    model, table, tr3 = device_model(oracle, name)
    geo = model.wave_geometry
    assert (geo["diagonal_ctas"] == 0) == bool(wave_k)
    assert geo["columns_per_lane"] == (int(wave_k) if wave_k else geo["columns_per_lane"]) and geo["columns_per_lane"] > 0
    assert geo["warps"] * 32 * geo["columns_per_lane"] >= h["model_length"] - 1
    rng = np.random.default_rng(h["model_length"])
    consensus = np.argmax(h["match_emissions"][1:], axis=1).astype(np.uint8)
    lengths = [0, 1, 2, 3, 4, 5, 7, 8, 9, 63, 64, 65, 66, 67, 68, 130, 255, 256, 257, 1000, 3500, 3967, 3968, 3969, 3970, 5000, 12001]
    seqs = [rng.integers(0, 20, size=n, dtype=np.uint8) for n in lengths]
    for n in (40, 200, 3500, 6000):  # hits: a consensus stretch somewhere inside
        s = rng.integers(0, 20, size=n, dtype=np.uint8)
        seg = min(n, 100, consensus.size)
        at = int(rng.integers(0, n - seg + 1))
        s[at:at + seg] = consensus[:seg]
        seqs.append(s)
    want = [oracle.score_codes(table, tr3, s) for s in seqs]
    assert sum(w > 0 for w in want) >= 3  # the planted ones really are hits
    for repeat in range(2):  # twice: the device-side accumulators must be clean again after every call
        got = [model.score_sequence(s) for s in seqs]
        assert [bits(v) for v in got] == [bits(v) for v in want], (name, wave_k, repeat)
    with pytest.raises(_cabi.MsvCudaError) as err:
        model.score_sequence(np.array([1, 2, 3, 4, 5, 6, 20, 7], np.uint8))
    assert err.value.status == _cabi.MSV_ERR_BAD_RESIDUE and "position 6" in str(err.value)
    with pytest.raises(_cabi.MsvCudaError):
        bad = rng.integers(0, 20, size=5000, dtype=np.uint8)
        bad[4999] = 200
        model.score_sequence(bad)
    assert bits(model.score_sequence(seqs[20])) == bits(want[20])  # and the next call is unaffected


def test_single_sequence_wavefront_kernel_long_models(oracle):
    """Synthetic models up to the on-chip limit (5631 columns): the chain then spans a whole 8-CTA cluster."""
    for leng in (2816, 4096, 5631):
        rng = np.random.default_rng(leng)
        match = np.zeros((leng + 1, 20), np.float32)
        match[1:] = rng.dirichlet(np.full(20, 0.4), size=leng).astype(np.float32)
        table, tr3 = oracle.prepare(match)
        model = msv.Model(_cabi.emission_table(match), *_cabi.model_transitions(leng + 1))
        assert model.wave_geometry["columns_per_lane"] > 0 and model.wave_geometry["ctas"] <= 8
        assert model.wave_geometry["diagonal_ctas"] == 0  # beyond one SM's shared memory: the chain kernel is in charge
        for n in (0, 5, 333, 2000):
            s = rng.integers(0, 20, size=n, dtype=np.uint8)
            assert bits(model.score_sequence(s)) == bits(oracle.score_codes(table, tr3, s)), (leng, n)


# ---- filter stages kept on the device (csrc/filter_cuda.cu) -------------------------------------------------------------
def test_filter_pipeline_selects_on_the_device(oracle):
    """MSV filter -> Viterbi filter on the survivors, HMMER3's pipeline order, with the selection and the survivors' index list
    on the GPU: the hits of both stages must be exactly those a host-side selection over the full per-sequence arrays gives,
    in database order, with the same numbers; a capacity that is too small truncates and still reports the true count."""
    import ctypes as C
    name = "400.hmm"
    prof = msv.Profile_HMM(hmm_path(name))
    rng = np.random.default_rng(23)
    cons = prof.match_emissions[1:].argmax(axis=1).astype(np.uint8)
    seqs = [rng.integers(0, 20, size=int(k), dtype=np.uint8) for k in rng.integers(50, 700, size=30_000)]
    for q in rng.choice(len(seqs), size=300, replace=False):  # homologs of varying strength, scattered over the database
        a = int(rng.integers(0, 200))
        seg = cons[a:a + int(rng.integers(30, 200))].copy()
        seg[rng.random(seg.size) < rng.uniform(0.0, 0.5)] = rng.integers(0, 20)
        seqs[q] = np.concatenate([seqs[q][:30], seg, seqs[q][30:]])
    packed = msv.Packed_sequences.from_arrays(*pack(seqs))
    database = msv.Device_database(packed)
    msv_model, vit_model = msv.MSV_HMM(prof), msv.Viterbi_HMM(prof)

    # reference selection on the host from the full arrays (the round-1 route)
    raw = msv_model.parallel_run_on_sequences(database)
    h = oracle.load_hmm(hmm_path(name))
    dev_model = msv.Model(_cabi.emission_table(h["match_emissions"]), *_cabi.model_transitions(h["model_length"]))
    db = msv.Database(packed.residues, packed.offsets)
    import torch
    scores = torch.from_numpy(raw.copy()).cuda()
    bits_d, p_d = torch.empty_like(scores), torch.empty_like(scores)
    db.filter_device(scores, float(prof.stats_local_msv_mu), float(prof.stats_local_msv_lambda), bits_d, p_d)
    torch.cuda.synchronize()
    p_all, bits_all = p_d.cpu().numpy(), bits_d.cpu().numpy()
    want_stage1 = np.flatnonzero(p_all <= np.float32(0.02))

    hits = msv_model.msv_filter(database, 0.02)
    assert hits["index"].astype(np.int64).tolist() == want_stage1.tolist() and len(want_stage1) > 300
    assert ubits(hits["score"]).tolist() == ubits(raw[want_stage1]).tolist()
    assert ubits(hits["bits"]).tolist() == ubits(bits_all[want_stage1]).tolist()
    assert ubits(hits["p_value"]).tolist() == ubits(p_all[want_stage1]).tolist()

    # second stage over the survivors only
    vit_all = vit_model.parallel_run_on_sequences(database)
    full = vit_model.viterbi_filter(database, 1e-3)  # statistics of every sequence, selected on the host
    want_stage2 = [int(i) for i in full["index"] if int(i) in set(want_stage1.tolist())]
    survivors = vit_model.viterbi_filter_survivors(database, 1e-3)
    assert survivors["index"].astype(np.int64).tolist() == want_stage2 and 100 < len(want_stage2) < len(want_stage1)
    assert ubits(survivors["score"]).tolist() == ubits(vit_all[np.array(want_stage2)]).tolist()
    lookup = {int(i): k for k, i in enumerate(full["index"])}
    rows = [lookup[i] for i in want_stage2]
    assert ubits(survivors["p_value"]).tolist() == ubits(full["p_value"][rows]).tolist()

    # raw C entry point: small capacity
    idx = np.zeros(5, np.uint32)
    sc, bt, pv = (np.zeros(5, np.float32) for _ in range(3))
    found = C.c_size_t()
    _cabi.check(_cabi.lib.msv_cuda_db_msv_filter(dev_model.handle, db.handle, float(prof.stats_local_msv_mu), float(prof.stats_local_msv_lambda),
                                                 0.02, idx.ctypes.data, sc.ctypes.data, bt.ctypes.data, pv.ctypes.data, 5, C.byref(found)))
    assert found.value == len(want_stage1) and idx.tolist() == want_stage1[:5].tolist()
    # a database nobody passes
    none = msv_model.msv_filter(database, -1.0)
    assert len(none["index"]) == 0 and len(vit_model.viterbi_filter_survivors(database, 1e-3)["index"]) == 0
