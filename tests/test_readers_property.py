"""Property tests of the FASTA readers (row A2 of SURVEY.md section 8): on randomly generated FASTA-like text -- blank
lines, '>' inside lines, foreign letters, '#', CRs, missing final newline, text before the first header -- this
implementation's string reader, its packed reader, the oracle restatement and (when built) the reference's own reader
must agree on which records survive and on their content."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import hmm_fasta_viterbi_b200 as msv
from oracle_lib import LETTERS, RefLib

ALPHABET = LETTERS + "XBZ#acd \r>-*"
line = st.text(alphabet=ALPHABET, min_size=0, max_size=30)
header = st.text(alphabet=LETTERS + " |_>", min_size=0, max_size=12).map(lambda s: ">" + s)
record = st.tuples(header, st.lists(line.filter(lambda s: not s.startswith(">")), min_size=0, max_size=4))


@st.composite
def fasta_text(draw):
    records = draw(st.lists(record, min_size=1, max_size=6))
    body = "\n".join("\n".join([h] + ls) for h, ls in records)
    if draw(st.booleans()):
        body += "\n"
    return body


def expected_records(text: str):
    """Straight restatement of the record rules (FASTA_protein_sequences.cpp:18-41) in Python."""
    lines = text.split("\n")
    if text.endswith("\n"):
        lines = lines[:-1]
    out = []
    for ln in lines:
        if ln[:1] == ">":
            out.append("#")
        else:
            out[-1] += ln  # the strategy always starts with a header
    allowed = set("#" + LETTERS)
    return [r for r in out if set(r) <= allowed]


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(text=fasta_text())
def test_fasta_readers_agree(tmp_path, oracle, text):
    path = tmp_path / "case.fsa"
    path.write_bytes(text.encode("latin-1"))
    want = expected_records(text)
    assert msv.FASTA_protein_sequences(str(path)).sequences == want
    assert oracle.load_fasta(str(path)) == want
    if RefLib.available():
        assert RefLib().load_fasta(str(path)) == want
    # the packed reader drops, in addition, records with a '#' inside (they cannot be scored; the reference would throw)
    scorable = [r for r in want if "#" not in r[1:]]
    packed = msv.Packed_sequences.from_fasta_file(str(path))
    off = packed.offsets
    got = ["#" + "".join(LETTERS[c] for c in packed.residues[int(off[q]):int(off[q + 1])]) for q in range(len(packed))]
    assert got == scorable
    assert packed.rejected == text.count("\n>") + (1 if text.startswith(">") else 0) - len(scorable)


@settings(max_examples=60, deadline=None)
@given(lengths=st.lists(st.integers(min_value=0, max_value=400), min_size=0, max_size=60), parts=st.integers(min_value=1, max_value=9))
def test_partition_properties(lengths, parts):
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    bounds = msv._cabi.partition_by_cells(offsets, parts)
    assert bounds[0] == 0 and bounds[-1] == len(lengths) and (np.diff(bounds) >= 0).all()
    total = int(offsets[-1])
    longest = max(lengths, default=0)
    for r in range(parts):  # no slice is further from the ideal share than one sequence
        share = int(offsets[bounds[r + 1]] - offsets[bounds[r]])
        assert abs(share - total / parts) <= 2 * longest + 1


# ---- .hmm reader (row A1) ---------------------------------------------------------------------------------------------
number = st.one_of(st.floats(min_value=0.0, max_value=12.0, allow_nan=False).map(lambda v: f"{v:.5f}"), st.just("*"))


@st.composite
def hmm_text(draw):
    leng = draw(st.integers(min_value=1, max_value=6))
    name = draw(st.text(alphabet="ABCdef_-0123456789", min_size=1, max_size=10))
    rows = lambda n: "  ".join(draw(number) for _ in range(n))
    out = ["HMMER3/b [3.1dev | April 2012; reverse compatibility mode]", f"NAME  {name}", "ACC   PB000001", f"LENG  {leng}",
           "ALPH  amino", "RF    no", "CS    yes", "MAP   yes", "DATE  Mon Dec 17 23:47:13 2012", "NSEQ  5", "EFFN  1.3", "CKSUM 1"]
    stats = [("MSV", draw(st.floats(-12, -5)), draw(st.floats(0.5, 0.9))), ("VITERBI", draw(st.floats(-12, -5)), draw(st.floats(0.5, 0.9))),
             ("FORWARD", draw(st.floats(-6, -2)), draw(st.floats(0.5, 0.9)))]
    for kind, a, b in draw(st.permutations(stats)):
        out.append(f"STATS LOCAL {kind:<8} {a:9.4f}  {b:.5f}")
    out.append("HMM          A        C        D        E        F        G        H        I        K        L        M        N"
               "        P        Q        R        S        T        V        W        Y   ")
    out.append("            m->m     m->i     m->d     i->m     i->i     d->m     d->d")
    out.append("  COMPO   " + rows(20))
    out.append("          " + rows(20))
    out.append("          " + rows(7))
    for node in range(1, leng + 1):
        out.append(f"{node:7d}   " + rows(20) + f"  {node:5d} - -")
        out.append("          " + rows(20))
        out.append("          " + rows(7))
    out.append("//")
    return "\n".join(out) + "\n"


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(text=hmm_text())
def test_hmm_readers_agree(tmp_path, oracle, text):
    path = tmp_path / "case.hmm"
    path.write_text(text)
    mine = msv.Profile_HMM(str(path))
    want = oracle.load_hmm(str(path))
    refs = [want] + ([RefLib().load_hmm(str(path))] if RefLib.available() else [])
    for ref in refs:
        assert mine.name == ref["name"] and mine.model_length == ref["model_length"]
        for key in ("match_emissions", "insert_emissions", "transitions"):
            assert np.asarray(getattr(mine, key)).view(np.uint32).tolist() == ref[key].view(np.uint32).tolist(), key
        stats = np.array([mine.stats_local_msv_mu, mine.stats_local_msv_lambda, mine.stats_local_viterbi_mu,
                          mine.stats_local_viterbi_lambda, mine.stats_local_forward_theta, mine.stats_local_forward_lambda], np.float32)
        assert stats.view(np.uint32).tolist() == ref["stats"].view(np.uint32).tolist()
    # model preparation on top of it: same table bits as the oracle
    table, tr3 = oracle.prepare(want["match_emissions"])
    assert msv._cabi.emission_table(mine.match_emissions).view(np.uint32).tolist() == table.view(np.uint32).tolist()
