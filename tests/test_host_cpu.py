"""CPU-side tests of the product's host layer (no GPU needed): the C ABI library loads and exports every declared
symbol, the C++ readers and the host helpers agree bit-for-bit with the oracle / golden vectors, and the GPU entry
points fail loudly (never fall back) when there is no device."""
import ctypes as C
import os
import re
import subprocess
import zlib

import numpy as np
import pytest

import hmm_fasta_viterbi_b200 as msv
from conftest import REPO, fasta_path, hmm_path, model_files
from hmm_fasta_viterbi_b200 import _cabi
from oracle_lib import LETTERS, pack, synthetic_database


def bits(x) -> str:
    return format(int(np.float32(x).view(np.uint32)), "08x")


def crc(a) -> str:
    return format(zlib.crc32(np.ascontiguousarray(a, np.float32).tobytes()), "08x")


NO_GPU = _cabi.device_count() == 0


# ---- the C ABI library ------------------------------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "msv_cuda.h")).read()
    declared = set(re.findall(r"\b(msv_(?:cuda|host)_[a-z_]+)\s*\(", header))
    assert declared == set(_cabi.DECLARED_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    assert declared <= exported
    assert _cabi.lib.msv_cuda_abi_version() == 3


def test_cabi_has_no_torch_or_oracle_dependency():
    out = subprocess.run(["ldd", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "torch" not in out and "oracle" not in out and "msv_ref" not in out
    out = subprocess.run(["ldd", msv.host.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "oracle" not in out and "msv_ref" not in out


def test_cabi_carries_sm100a_code_with_tma():
    sass = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass       # cp.async.bulk: the emission table is staged by the TMA unit
    assert "CREDUX.MAX.F32" in sass  # warp-wide fp32 max of the E reduction
    assert "FMNMX3" in sass
    assert "LDTM" in sass and "STTM" in sass  # tcgen05.ld / tcgen05.st: part of the emission table lives in tensor memory


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a box without a GPU")
def test_gpu_entry_points_fail_loudly_without_a_device(oracle):
    h = oracle.load_hmm(hmm_path("100.hmm"))
    table = _cabi.emission_table(h["match_emissions"])
    with pytest.raises(_cabi.MsvCudaError) as err:
        msv.Model(table, *_cabi.model_transitions(h["model_length"]))
    assert err.value.status == _cabi.MSV_ERR_NO_DEVICE
    with pytest.raises(_cabi.MsvCudaError):
        msv.Database(np.zeros(4, np.uint8), np.array([0, 4], np.uint64))
    model = msv.MSV_HMM(msv.Profile_HMM(hmm_path("100.hmm")))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        model.parallel_run_on_sequence("#ACDEF")
    # the Viterbi scan has no host implementation either
    with pytest.raises(_cabi.MsvCudaError) as err:
        msv.ViterbiModel(table, _cabi.viterbi_transitions(h["transitions"]), *_cabi.model_transitions(h["model_length"]))
    assert err.value.status == _cabi.MSV_ERR_NO_DEVICE
    with pytest.raises(RuntimeError, match="no CUDA device"):
        msv.Viterbi_HMM(msv.Profile_HMM(hmm_path("100.hmm"))).parallel_run_on_sequence("#ACDEF")


def test_viterbi_host_transitions_match_oracle(oracle):
    for name in ("100.hmm", "2405.hmm"):
        h = oracle.load_hmm(hmm_path(name))
        mine, theirs = _cabi.viterbi_transitions(h["transitions"]), oracle.viterbi_prepare(h["transitions"])
        assert mine.view(np.uint32).tolist() == theirs.view(np.uint32).tolist()
        assert np.all(mine <= 0)


# ---- host helpers vs oracle -------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", model_files())
def test_host_model_arithmetic_matches_golden(golden_tables, golden_readers, name):
    prof = msv.Profile_HMM(hmm_path(name))
    g = golden_readers["hmm"][name]
    assert prof.name == g["name"] and prof.model_length == g["model_length"]
    assert crc(prof.match_emissions) == g["match_crc32"]
    assert crc(prof.insert_emissions) == g["insert_crc32"]
    assert crc(prof.transitions) == g["transitions_crc32"]
    stats = [prof.stats_local_msv_mu, prof.stats_local_msv_lambda, prof.stats_local_viterbi_mu,
             prof.stats_local_viterbi_lambda, prof.stats_local_forward_theta, prof.stats_local_forward_lambda]
    assert [bits(v) for v in stats] == g["stats"]
    t = golden_tables[name]
    assert crc(_cabi.emission_table(prof.match_emissions)) == t["table_crc32"]
    assert [bits(v) for v in _cabi.model_transitions(prof.model_length)] == [t["tr_B_Mk"], t["tr_E_C"], t["tr_E_J"]]


def test_length_transitions_match_oracle(oracle):
    for n in list(range(0, 70)) + [130, 347, 3500, 35000, 10**6]:
        assert [bits(v) for v in _cabi.length_transitions(n)] == [bits(v) for v in oracle.length_transitions(n)]


def test_fasta_readers(golden_readers, tmp_path):
    for fname, want in golden_readers["fasta"].items():
        fa = msv.FASTA_protein_sequences(fasta_path(fname))
        assert fa.sequences == want
        packed = msv.Packed_sequences.from_fasta_file(fasta_path(fname))
        assert packed.rejected == 0
        assert len(packed) == len(want)
        codes = packed.residues
        off = packed.offsets
        for q, seq in enumerate(want):
            assert "".join(LETTERS[c] for c in codes[int(off[q]):int(off[q + 1])]) == seq[1:]
        again = msv.Packed_sequences.from_fasta(fa)
        assert again.offsets.tolist() == off.tolist() and again.residues.tolist() == codes.tolist()
    p = tmp_path / "mixed.fsa"
    p.write_text(">a\nACDE\nFGH\n>b has X\nACXDE\n>c\n\nWYW\n>d lower\nacd\n>e\n")
    assert msv.FASTA_protein_sequences(str(p)).sequences == ["#ACDEFGH", "#WYW", "#"]
    packed = msv.Packed_sequences.from_fasta_file(str(p))
    assert packed.rejected == 2 and packed.offsets.tolist() == [0, 7, 10, 10]


def test_missing_files():
    prof = msv.Profile_HMM("/nonexistent/x.hmm")  # the reference prints and leaves the object empty
    assert prof.model_length == 0 and prof.match_emissions.shape[0] == 0
    assert msv.FASTA_protein_sequences("/nonexistent/x.fsa").sequences == []
    with pytest.raises(RuntimeError):
        msv.Packed_sequences.from_fasta_file("/nonexistent/x.fsa")


def test_encode_and_foreign_letters():
    assert _cabi.encode(LETTERS).tolist() == list(range(20))
    with pytest.raises(KeyError):
        _cabi.encode("ACDX")
    model = msv.MSV_HMM(msv.Profile_HMM(hmm_path("100.hmm")))
    with pytest.raises(KeyError):  # std::out_of_range in C++, as the reference's .at() (MSV_HMM.cpp:101)
        model.run_on_sequence("#ACDB")


@pytest.mark.parametrize("name", ["100.hmm", "700.hmm", "1400.hmm", "2405.hmm"])
def test_run_on_sequence_matches_golden(golden_scores, golden_readers, name):
    """MSV_HMM::run_on_sequence (the API's CPU entry point, host/algorithms/MSV_HMM.cpp) against the reference bits."""
    model = msv.MSV_HMM(msv.Profile_HMM(hmm_path(name)))
    g = golden_scores["scores"][name]
    seqs = {"example": golden_readers["fasta"]["fasta_like_example.fsa"], "random": golden_readers["fasta"]["random_FASTA.fsa"],
            "extra": golden_scores["meta"]["extra_sequences"]}
    for key, want in g.items():
        assert [bits(model.run_on_sequence(s)) for s in seqs[key]] == want


@pytest.mark.parametrize("prog", ["test_hmm_parsing", "test_fasta_parsing"])
def test_reference_reader_tests_pass_unchanged(prog):
    """The reference's own reader tests (data_readers/test_hmm_parsing.cpp, test_fasta_parsing.cpp), compiled unchanged
    against this implementation's readers with asserts enabled (tools/build_reference_programs.sh)."""
    exe = os.path.join(REPO, "build", "data_readers", prog)
    if not os.path.exists(exe):
        pytest.skip("build/ not populated (needs /root/reference at build time)")
    run = subprocess.run([exe], cwd=os.path.dirname(exe), capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr


@pytest.mark.timeout(600)
def test_host_layer_under_address_and_ub_sanitizers(tmp_path):
    """tests/host_selftest.cpp, compiled together with the host sources with -fsanitize=address,undefined."""
    host = os.path.join(REPO, "hmm_fasta_viterbi_b200", "host")
    sources = [os.path.join(host, "data_readers", f) for f in
               ("Profile_HMM.cpp", "FASTA_protein_sequences.cpp", "Packed_sequences.cpp", "Synthetic_database.cpp")]
    sources += [os.path.join(host, "algorithms", "MSV_HMM.cpp"), os.path.join(REPO, "tests", "host_selftest.cpp")]
    exe = str(tmp_path / "host_selftest")
    pkg = os.path.join(REPO, "hmm_fasta_viterbi_b200")
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                    f"-I{host}/data_readers", f"-I{host}/algorithms", f"-I{REPO}/include", *sources, "-o", exe,
                    f"-L{pkg}", "-lmsv_cuda", "-lpthread", f"-Wl,-rpath,{pkg}"], check=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0", UBSAN_OPTIONS="print_stacktrace=1")
    run = subprocess.run([exe, os.path.join(REPO, "fixtures"), str(tmp_path)], capture_output=True, text=True, env=env, timeout=500)
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-3000:]
    assert "host selftest ok" in run.stdout


def test_partition_by_cells():
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 500, size=1000)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for parts in (1, 2, 3, 4, 8):
        b = _cabi.partition_by_cells(offsets, parts)
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
        cells = [int(offsets[b[i + 1]] - offsets[b[i]]) for i in range(parts)]
        assert max(cells) - min(cells) <= 2 * 500
    assert _cabi.partition_by_cells(np.array([0], np.uint64), 4).tolist() == [0, 0, 0, 0, 0]


def test_synthetic_databases_are_seeded_and_shaped():
    a = msv.Packed_sequences.synthetic_swissprot_like(5000, 1400)
    b = msv.Packed_sequences.synthetic_swissprot_like(5000, 1400)
    c = msv.Packed_sequences.synthetic_swissprot_like(5000, 1401)
    assert a.offsets.tolist() == b.offsets.tolist() and (a.residues == b.residues).all()
    assert a.offsets.tolist() != c.offsets.tolist()
    lens = np.diff(a.offsets.astype(np.int64))
    assert lens.min() >= 30 and lens.max() <= 3000 and 300 < lens.mean() < 400
    assert a.residues.max() < 20
    t = msv.Packed_sequences.synthetic_long_uniform(16, 2405, 10000, 35000)
    lens = np.diff(t.offsets.astype(np.int64))
    assert lens.min() >= 10000 and lens.max() <= 35000


def test_checker_side_generator_builds_the_same_databases():
    """bench.py's --impl reference arm builds its workload from oracle/synthetic_db.cpp so that it never loads the product
    libraries; it must be the same database the product generator hands to the GPU arm."""
    ours = msv.Packed_sequences.synthetic_swissprot_like(3000, 20261018)
    codes, offsets = synthetic_database("swissprot_like", 3000, 20261018)
    assert offsets.tolist() == ours.offsets.tolist() and (codes == ours.residues).all()
    ours = msv.Packed_sequences.synthetic_long_uniform(8, 2405, 10000, 35000)
    codes, offsets = synthetic_database("long_uniform", 8, 2405, 10000, 35000)
    assert offsets.tolist() == ours.offsets.tolist() and (codes == ours.residues).all()
