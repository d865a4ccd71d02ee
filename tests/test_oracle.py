"""Pins oracle/msv_oracle.c (the CPU checker) before anything trusts it.

1. against the committed golden vectors produced by the reference's own code (tests/golden/make_golden.py),
2. against the reader known-answer tests the reference itself holds
   (data_readers/test_hmm_parsing.cpp:23-36, data_readers/test_fasta_parsing.cpp:8-14),
3. against oracle/_ref (the compiled reference) on seeded random inputs, when that library is present.
All comparisons are on IEEE-754 bit patterns.
"""
import math
import zlib

import numpy as np
import pytest

from conftest import fasta_path, hmm_path, model_files
from oracle_lib import LETTERS, encode, pack


def bits(x) -> str:
    return format(int(np.float32(x).view(np.uint32)), "08x")


def crc(a: np.ndarray) -> str:
    return format(zlib.crc32(np.ascontiguousarray(a, np.float32).tobytes()), "08x")


# ---- readers -----------------------------------------------------------------------------------------------------
def test_hmm_reader_reference_kats(oracle):
    """The asserts of data_readers/test_hmm_parsing.cpp:23-36, same 5-ULP tolerance (:9-15)."""
    h = oracle.load_hmm(hmm_path("100.hmm"))

    def almost(x, y, ulp=5):
        x, y = np.float32(x), np.float32(y)
        return abs(x - y) <= np.finfo(np.float32).eps * abs(x + y) * ulp or abs(x - y) < np.finfo(np.float32).tiny

    prob = lambda v: np.exp(np.float32(-1) * np.float32(v))
    assert h["model_length"] == 101
    assert h["name"] == "Pfam-B_229"
    assert almost(h["stats"][0], np.float32(-9.5678))
    assert almost(h["stats"][5], np.float32(0.71755))
    assert almost(h["insert_emissions"][0][0], prob(2.68618))
    assert almost(h["transitions"][0][6], prob(0.0))  # "*" parses as probability 1.0
    assert almost(h["match_emissions"][1][0], prob(2.66211))
    assert almost(h["match_emissions"][100][19], prob(4.01014))
    assert almost(h["insert_emissions"][1][19], prob(3.61503))
    assert almost(h["transitions"][1][1], prob(4.09464))
    assert almost(h["insert_emissions"][100][19], prob(3.61503))
    assert almost(h["transitions"][100][5], prob(0.0))
    assert almost(h["transitions"][100][6], prob(0.0))
    assert not h["match_emissions"][0].any()  # dummy node 0 is zero-filled (Profile_HMM.cpp:110-111)


def test_fasta_reader_reference_kat(oracle):
    """data_readers/test_fasta_parsing.cpp:8-14."""
    got = oracle.load_fasta(fasta_path("fasta_like_example.fsa"))
    assert got == [
        "#ACDEFGHIKLMNPQTVWY",
        "#ACDKLMNPQTVWYEFGHI",
        "#EFMNRGHIKLMNPQT",
        "#MKMRFFSSPCGKAAVDPADRCKEVQQIRDQHPSKIPVIIERYKGEKQLPVLDKTKFLVPDHVNMSELVKI"
        "IRRRLQLNPTQAFFLLVNQHSMVSVSTPIADIYEQEKDEDGFLYMVYASQETFGFIRENE",
    ]


@pytest.mark.parametrize("name", model_files())
def test_hmm_reader_golden(oracle, golden_readers, name):
    h = oracle.load_hmm(hmm_path(name))
    g = golden_readers["hmm"][name]
    assert h["name"] == g["name"]
    assert h["model_length"] == g["model_length"]
    assert [bits(v) for v in h["stats"]] == g["stats"]
    assert crc(h["match_emissions"]) == g["match_crc32"]
    assert crc(h["insert_emissions"]) == g["insert_crc32"]
    assert crc(h["transitions"]) == g["transitions_crc32"]


def test_fasta_reader_golden(oracle, golden_readers):
    for fname, want in golden_readers["fasta"].items():
        assert oracle.load_fasta(fasta_path(fname)) == want


def test_fasta_reader_rejects_whole_record(oracle, tmp_path):
    """FASTA_protein_sequences.cpp:26-41: a record with a foreign letter disappears; survivors keep their order."""
    p = tmp_path / "mixed.fsa"
    p.write_text(">a\nACDE\nFGH\n>b has X\nACXDE\n>c\n\nWYW\n>d lower\nacd\n>e\n")
    assert oracle.load_fasta(str(p)) == ["#ACDEFGH", "#WYW", "#"]


# ---- model preparation ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", model_files())
def test_model_table_golden(oracle, golden_tables, name):
    h = oracle.load_hmm(hmm_path(name))
    table, tr3 = oracle.prepare(h["match_emissions"])
    g = golden_tables[name]
    assert table.shape == (20, g["model_length"])
    assert crc(table) == g["table_crc32"]
    assert [bits(v) for v in tr3] == [g["tr_B_Mk"], g["tr_E_C"], g["tr_E_J"]]
    assert np.isneginf(table[:, 0]).all()  # dummy column M0 (MSV_HMM.cpp:36-45)


def test_length_transitions(oracle):
    """MSV_HMM.cpp:59-64, incl. the empty sequence: tr_loop = -inf, tr_move = 0."""
    loop, move = oracle.length_transitions(0)
    assert np.isneginf(loop) and move == 0.0
    loop, move = oracle.length_transitions(18)
    assert loop < 0 and move < 0


# ---- the recurrence ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", model_files())
def test_scores_golden(oracle, golden_scores, golden_readers, name):
    h = oracle.load_hmm(hmm_path(name))
    table, tr3 = oracle.prepare(h["match_emissions"])
    g = golden_scores["scores"][name]
    seqs = {
        "example": golden_readers["fasta"]["fasta_like_example.fsa"],
        "random": golden_readers["fasta"]["random_FASTA.fsa"],
        "extra": golden_scores["meta"]["extra_sequences"],
    }
    for key, want in g.items():
        got = [bits(oracle.score_string(table, tr3, s)) for s in seqs[key]]
        assert got == want, (name, key)


def test_appendix_a_spot_values(oracle):
    """SURVEY.md Appendix A decimal spot checks."""
    h = oracle.load_hmm(hmm_path("100.hmm"))
    table, tr3 = oracle.prepare(h["match_emissions"])
    ex = oracle.load_fasta(fasta_path("fasta_like_example.fsa"))
    assert bits(oracle.score_string(table, tr3, ex[0])) == "c114d20b"
    assert bits(oracle.score_string(table, tr3, ex[3])) == "c13f46b5"


def test_empty_sequence_scores_minus_inf(oracle):
    h = oracle.load_hmm(hmm_path("100.hmm"))
    table, tr3 = oracle.prepare(h["match_emissions"])
    assert np.isneginf(oracle.score_string(table, tr3, "#"))


def test_foreign_letter_is_an_error(oracle):
    h = oracle.load_hmm(hmm_path("100.hmm"))
    table, tr3 = oracle.prepare(h["match_emissions"])
    with pytest.raises(KeyError):
        oracle.score_string(table, tr3, "#ACDX")


def test_batch_equals_single_and_threads(oracle):
    h = oracle.load_hmm(hmm_path("300.hmm"))
    table, tr3 = oracle.prepare(h["match_emissions"])
    rng = np.random.default_rng(7)
    seqs = [rng.integers(0, 20, size=int(n), dtype=np.uint8) for n in rng.integers(0, 200, size=64)]
    codes, offsets = pack(seqs)
    one = np.array([oracle.score_codes(table, tr3, s) for s in seqs], np.float32)
    for threads in (1, 3, 8):
        got = oracle.score_batch(table, tr3, codes, offsets, threads)
        assert got.view(np.uint32).tolist() == one.view(np.uint32).tolist()


@pytest.mark.parametrize("name", ["100.hmm", "1301.hmm", "2405.hmm"])
def test_oracle_equals_compiled_reference_on_random_inputs(oracle, reflib, name):
    """Direct check against oracle/_ref (reference object code) on seeded inputs incl. biased compositions."""
    m = reflib.model(hmm_path(name))
    rt, rtr = m.table()
    h = oracle.load_hmm(hmm_path(name))
    table, tr3 = oracle.prepare(h["match_emissions"])
    assert rt.view(np.uint32).tolist() == table.view(np.uint32).tolist()
    assert rtr.view(np.uint32).tolist() == tr3.view(np.uint32).tolist()
    rng = np.random.default_rng(int(name.split(".")[0]))
    for n in (0, 1, 5, 64, 300, 777):
        codes = rng.integers(0, 20, size=n, dtype=np.uint8)
        seq = "#" + "".join(LETTERS[c] for c in codes)
        assert bits(m.run_on_sequence(seq)) == bits(oracle.score_codes(table, tr3, codes)), (name, n)
    # a sequence emitted from the model's own consensus scores high: exercises the J (multi-hit) branch
    cons = h["match_emissions"][1:].argmax(axis=1).astype(np.uint8)
    seq = "#" + "".join(LETTERS[c] for c in np.concatenate([cons, cons[: len(cons) // 2]]))
    want = m.run_on_sequence(seq)
    assert want > 0
    assert bits(want) == bits(oracle.score_string(table, tr3, seq))
    # batch helper of the reference shim agrees with per-sequence calls
    seqs = [rng.integers(0, 20, size=int(k), dtype=np.uint8) for k in (3, 0, 40, 17)]
    codes, offsets = pack(seqs)
    got = m.run_batch(codes, offsets, threads=2)
    want = [bits(oracle.score_codes(table, tr3, s)) for s in seqs]
    assert [bits(v) for v in got] == want
