"""The device-side FASTA parser (csrc/fasta_cuda.cu: msv_cuda_db_create_from_fasta / msv_cuda_score_fasta) against the
host reader and a straight Python restatement of the reference's record rules (data_readers/FASTA_protein_sequences.cpp:9-44):
which records survive, their content, their order -- byte for byte -- on hand-made corner cases, on random text whose lines,
headers and records straddle the parser's 16-byte spans and 4096-byte tiles, and on the bench-sized synthetic database."""
import os

import numpy as np
import pytest

import hmm_fasta_viterbi_b200 as msv
from conftest import fasta_path, hmm_path
from hmm_fasta_viterbi_b200 import _cabi
from oracle_lib import LETTERS, pack

pytestmark = pytest.mark.gpu
CODE = {ch: i for i, ch in enumerate(LETTERS)}


def expected(text: bytes):
    """(list of code arrays, rejected) by the record rules: '>' at a line start opens a record (line dropped), other lines are
    appended verbatim, a record with any byte outside the alphabet is dropped whole, text before the first header is ignored."""
    records, current, rejected = [], None, 0
    lines = text.split(b"\n")
    if text.endswith(b"\n"):
        lines = lines[:-1]

    def close(rec):
        nonlocal rejected
        if rec is None:
            return
        if all(chr(c) in CODE for c in rec):
            records.append(np.array([CODE[chr(c)] for c in rec], np.uint8))
        else:
            rejected += 1

    for ln in lines:
        if ln[:1] == b">":
            close(current)
            current = bytearray()
        elif current is not None:
            current += ln
    close(current)
    return records, rejected


def check(text: bytes):
    want, want_rejected = expected(text)
    db = msv.Database.from_fasta(text)
    residues, offsets = db.download()
    assert db.rejected == want_rejected
    assert len(offsets) - 1 == len(want) and offsets[0] == 0
    got = [residues[int(offsets[q]):int(offsets[q + 1])] for q in range(len(want))]
    for q, (a, b) in enumerate(zip(got, want)):
        assert a.tolist() == b.tolist(), q
    assert db.info()["longest"] == max([len(w) for w in want], default=0)
    db.close()
    return want


def test_corner_cases():
    for text in [b"", b"\n", b">", b">\n", b">only a header", b"ACDE\n", b"ACDE\n>h\nAC\n", b">h\nACDE", b">h\nACDE\n", b">h\n\n\nAC\n\nDE\n\n",
                 b">a\n>b\n>c\n", b">a\nAC\r\n>b\nDE\n", b">a\nACX\n>b\nDE\n>c\nF#G\n>d\nHIK\n", b">a\nAC>DE\n", b">a\nAC\n >b\nDE\n",
                 b">a\nacd\n>b\nWY\n", b"junk\nmore junk\n>a\nAC\n", b"\n\n>a\nAC\n", b">a\n" + b"A" * 5000 + b"\n", b">" + b"h" * 9000 + b"\nACD\n",
                 b">a\n" + b"\n".join([b"ACDEFGHIKL"] * 1000) + b"\n>b\nY"]:
        check(text)


def test_fixture_files_match_the_host_reader():
    for name in ("fasta_like_example.fsa", "random_FASTA.fsa"):
        with open(fasta_path(name), "rb") as f:
            text = f.read()
        want = check(text)
        host = msv.Packed_sequences.from_fasta_file(fasta_path(name))
        assert len(host) == len(want) and host.residues.tolist() == np.concatenate(want).tolist()


@pytest.mark.parametrize("seed", range(12))
def test_random_text_across_span_and_tile_boundaries(seed):
    """Line lengths, header lengths and record sizes drawn so that line starts, headers and foreign bytes fall on every
    position relative to the 16-byte spans and 4096-byte tiles; a few percent of the records carry a foreign byte."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(LETTERS.encode(), np.uint8)
    parts = []
    if seed % 3 == 0:
        parts.append(b"text before the first header\nACDEF\n")
    for _ in range(int(rng.integers(1, 400))):
        parts.append(b">" + bytes(rng.choice(np.frombuffer(b"abcXYZ |>_09", np.uint8), size=int(rng.integers(0, 70 if seed % 4 else 5000)))) + b"\n")
        n = int(rng.integers(0, 3000)) if rng.random() < 0.9 else int(rng.integers(3000, 20000))
        body = rng.choice(letters, size=n)
        if rng.random() < 0.06 and n:
            body = body.copy()
            body[int(rng.integers(0, n))] = rng.choice(np.frombuffer(b"XBZ#a \r-*>", np.uint8))
        width = int(rng.choice([1, 7, 15, 16, 17, 60, 80, 4095, 4096, 4097, 100000]))
        for at in range(0, n, width):
            parts.append(bytes(body[at:at + width]) + b"\n")
        if rng.random() < 0.1:
            parts.append(b"\n")
    text = b"".join(parts)
    if seed % 2:
        text = text[:-1]  # no newline at the end of the file
    check(text)


def test_scores_from_text_equal_scores_from_packed(oracle, tmp_path):
    """msv_cuda_score_fasta-style route (text -> GPU parser -> scan) gives the bits of the packed route, from bytes and from
    an mmap'ed file (the pageable text is staged through the pinned ring by several threads); rejected records are skipped."""
    h = oracle.load_hmm(hmm_path("400.hmm"))
    model = msv.Model(_cabi.emission_table(h["match_emissions"]), *_cabi.model_transitions(h["model_length"]))
    packed = msv.Packed_sequences.synthetic_swissprot_like(40_000, 3)
    codes, offsets = np.ascontiguousarray(packed.residues), np.ascontiguousarray(packed.offsets)
    letters = np.frombuffer(LETTERS.encode(), np.uint8)[codes]
    lines = []
    for q in range(len(offsets) - 1):
        lines.append(f">seq{q} synthetic\n".encode())
        body = letters[int(offsets[q]):int(offsets[q + 1])]
        for at in range(0, body.size, 60):
            lines.append(bytes(body[at:at + 60]) + b"\n")
    text = b"".join(lines)
    want = model.score_batch(codes, offsets)
    got, rejected = model.score_fasta(text)
    assert rejected == 0 and (got.view(np.uint32) == want.view(np.uint32)).all()
    path = tmp_path / "db.fasta"
    path.write_bytes(text)
    got, rejected = model.score_fasta_file(str(path))
    assert rejected == 0 and (got.view(np.uint32) == want.view(np.uint32)).all()
    # spoil three records: they disappear, the others keep their scores and their order
    spoiled = bytearray(text)
    bad = [5, 17_000, 39_999]
    starts = np.cumsum([0] + [len(x) for x in lines])
    header_lines = np.flatnonzero([ln.startswith(b">") for ln in lines])
    for q in bad:
        spoiled[int(starts[header_lines[q] + 1])] = ord("X")
    got, rejected = model.score_fasta(bytes(spoiled))
    keep = np.setdiff1d(np.arange(len(offsets) - 1), bad)
    assert rejected == 3 and (got.view(np.uint32) == want[keep].view(np.uint32)).all()
    # raw C entry point with a capacity that is too small: an error, not an overrun
    import ctypes as C
    out = np.zeros(10, np.float32)
    n, rej = C.c_size_t(), C.c_size_t()
    rc = _cabi.lib.msv_cuda_score_fasta(model.handle, text, len(text), out.ctypes.data, 10, C.byref(n), C.byref(rej))
    assert rc == _cabi.MSV_ERR_INVALID_ARGUMENT and n.value == len(offsets) - 1 and not out.any()
    big = np.zeros(len(offsets) - 1, np.float32)
    rc = _cabi.lib.msv_cuda_score_fasta(model.handle, text, len(text), big.ctypes.data, big.size, C.byref(n), C.byref(rej))
    assert rc == 0 and (big.view(np.uint32) == want.view(np.uint32)).all()
