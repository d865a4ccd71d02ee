"""ctypes bindings for the CPU checker libraries (TEST INFRASTRUCTURE).

* ``Oracle``  -> oracle/libmsv_oracle.so  (oracle/msv_oracle.c, the plain-C restatement)
* ``RefLib``  -> oracle/_ref/libmsv_ref.so (the reference's unmodified sources, oracle/ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmsv_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libmsv_ref.so")
LETTERS = "ACDEFGHIKLMNPQRSTVWY"

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build_oracle() -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference is reachable)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle", "ref"], check=True)


def encode(text: str) -> np.ndarray:
    """'#ACD...' or 'ACD...' -> uint8 codes 0..19 (order of MSV_HMM.cpp:29-31)."""
    if text.startswith("#"):
        text = text[1:]
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(LETTERS):
        lut[ord(ch)] = i
    codes = lut[np.frombuffer(text.encode("ascii"), dtype=np.uint8)]
    if (codes == 255).any():
        raise KeyError("foreign residue letter")
    return np.ascontiguousarray(codes)


def pack(seqs) -> tuple[np.ndarray, np.ndarray]:
    """list of uint8 code arrays -> (concatenated codes, uint64 offsets[n+1])."""
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    codes = np.concatenate(seqs).astype(np.uint8) if len(seqs) and int(offsets[-1]) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(codes), offsets


class Oracle:
    def __init__(self) -> None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.oracle_hmm_load.restype = C.c_void_p
        L.oracle_hmm_load.argtypes = [C.c_char_p]
        L.oracle_hmm_free.argtypes = [C.c_void_p]
        L.oracle_hmm_model_length.restype = C.c_size_t
        L.oracle_hmm_model_length.argtypes = [C.c_void_p]
        L.oracle_hmm_name.restype = C.c_char_p
        L.oracle_hmm_name.argtypes = [C.c_void_p]
        for fn in ("oracle_hmm_match", "oracle_hmm_insert", "oracle_hmm_transitions", "oracle_hmm_stats"):
            getattr(L, fn).restype = C.POINTER(C.c_float)
            getattr(L, fn).argtypes = [C.c_void_p]
        L.oracle_fasta_load.restype = C.c_void_p
        L.oracle_fasta_load.argtypes = [C.c_char_p]
        L.oracle_fasta_free.argtypes = [C.c_void_p]
        L.oracle_fasta_count.restype = C.c_size_t
        L.oracle_fasta_count.argtypes = [C.c_void_p]
        L.oracle_fasta_record.restype = C.c_char_p
        L.oracle_fasta_record.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_msv_prepare.argtypes = [_f32p, C.c_size_t, _f32p, _f32p]
        L.oracle_msv_length_transitions.argtypes = [C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_msv_score_codes.restype = C.c_float
        L.oracle_msv_score_codes.argtypes = [_f32p, C.c_size_t, _f32p, _u8p, C.c_size_t, C.c_void_p]
        L.oracle_msv_score_string.restype = C.c_int
        L.oracle_msv_score_string.argtypes = [_f32p, C.c_size_t, _f32p, C.c_char_p, C.POINTER(C.c_float)]
        L.oracle_msv_score_batch.argtypes = [_f32p, C.c_size_t, _f32p, _u8p, _u64p, C.c_size_t, _f32p, C.c_int]
        L.oracle_viterbi_prepare.argtypes = [_f32p, C.c_size_t, _f32p]
        L.oracle_viterbi_score_codes.restype = C.c_float
        L.oracle_viterbi_score_codes.argtypes = [_f32p, _f32p, C.c_size_t, _f32p, _u8p, C.c_size_t, C.c_void_p]
        L.oracle_viterbi_score_batch.argtypes = [_f32p, _f32p, C.c_size_t, _f32p, _u8p, _u64p, C.c_size_t, _f32p, C.c_int]

    # ---- readers ----
    def load_hmm(self, path: str) -> dict:
        h = self.lib.oracle_hmm_load(path.encode())
        if not h:
            raise OSError(f"oracle could not parse {path}")
        try:
            m = self.lib.oracle_hmm_model_length(h)
            grab = lambda fn, cols: np.ctypeslib.as_array(getattr(self.lib, fn)(h), shape=(m, cols)).copy()
            return {
                "model_length": int(m),
                "name": self.lib.oracle_hmm_name(h).decode(),
                "match_emissions": grab("oracle_hmm_match", 20),
                "insert_emissions": grab("oracle_hmm_insert", 20),
                "transitions": grab("oracle_hmm_transitions", 7),
                "stats": np.ctypeslib.as_array(self.lib.oracle_hmm_stats(h), shape=(6,)).copy(),
            }
        finally:
            self.lib.oracle_hmm_free(h)

    def load_fasta(self, path: str) -> list[str]:
        fa = self.lib.oracle_fasta_load(path.encode())
        if not fa:
            raise OSError(f"oracle could not open {path}")
        try:
            return [self.lib.oracle_fasta_record(fa, i).decode() for i in range(self.lib.oracle_fasta_count(fa))]
        finally:
            self.lib.oracle_fasta_free(fa)

    # ---- model + recurrence ----
    def prepare(self, match_emissions: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        m = match_emissions.shape[0]
        table = np.empty((20, m), dtype=np.float32)
        tr3 = np.empty(3, dtype=np.float32)
        self.lib.oracle_msv_prepare(np.ascontiguousarray(match_emissions, np.float32), m, table, tr3)
        return table, tr3

    def length_transitions(self, residues: int) -> tuple[np.float32, np.float32]:
        a, b = C.c_float(), C.c_float()
        self.lib.oracle_msv_length_transitions(residues, C.byref(a), C.byref(b))
        return np.float32(a.value), np.float32(b.value)

    def score_codes(self, table: np.ndarray, tr3: np.ndarray, codes: np.ndarray) -> np.float32:
        codes = np.ascontiguousarray(codes, np.uint8)
        if codes.size == 0:
            codes = np.zeros(1, np.uint8)
            n = 0
        else:
            n = codes.size
        return np.float32(self.lib.oracle_msv_score_codes(table, table.shape[1], tr3, codes, n, None))

    def score_string(self, table: np.ndarray, tr3: np.ndarray, seq: str) -> np.float32:
        out = C.c_float()
        rc = self.lib.oracle_msv_score_string(table, table.shape[1], tr3, seq.encode(), C.byref(out))
        if rc != 0:
            raise KeyError("foreign residue letter")
        return np.float32(out.value)

    def score_batch(self, table, tr3, codes, offsets, threads: int = 1) -> np.ndarray:
        n = len(offsets) - 1
        out = np.empty(n, dtype=np.float32)
        codes = np.ascontiguousarray(codes, np.uint8)
        if codes.size == 0:
            codes = np.zeros(1, np.uint8)
        self.lib.oracle_msv_score_batch(table, table.shape[1], tr3, codes, np.ascontiguousarray(offsets, np.uint64), n, out,
                                        threads)
        return out

    # ---- Plan-7 local Viterbi (oracle/viterbi_oracle.c; parity unpinned, see its header) ----
    def viterbi_prepare(self, transitions: np.ndarray) -> np.ndarray:
        t = np.ascontiguousarray(transitions, np.float32)
        out = np.empty_like(t)
        self.lib.oracle_viterbi_prepare(t, t.shape[0], out)
        return out

    def viterbi_score_codes(self, table, logtr, tr3, codes) -> np.float32:
        codes = np.ascontiguousarray(codes, np.uint8)
        n = codes.size
        if n == 0:
            codes = np.zeros(1, np.uint8)
        return np.float32(self.lib.oracle_viterbi_score_codes(table, np.ascontiguousarray(logtr, np.float32), table.shape[1], tr3,
                                                              codes, n, None))

    def viterbi_score_batch(self, table, logtr, tr3, codes, offsets, threads: int = 1) -> np.ndarray:
        n = len(offsets) - 1
        out = np.empty(n, dtype=np.float32)
        codes = np.ascontiguousarray(codes, np.uint8)
        if codes.size == 0:
            codes = np.zeros(1, np.uint8)
        self.lib.oracle_viterbi_score_batch(table, np.ascontiguousarray(logtr, np.float32), table.shape[1], tr3, codes,
                                            np.ascontiguousarray(offsets, np.uint64), n, out, threads)
        return out


def synthetic_database(kind: str, count: int, seed: int, shortest: int = 10_000, longest: int = 35_000):
    """BASELINE.json's synthetic workloads from oracle/synthetic_db.cpp (no product library involved):
    kind 'swissprot_like' (configs 3/4) or 'long_uniform' (config 5).  Returns (codes uint8, offsets uint64[n+1])."""
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    L = C.CDLL(ORACLE_SO)
    for fn in ("oracle_synthetic_swissprot_like", "oracle_synthetic_long_uniform"):
        getattr(L, fn).restype = C.c_void_p
    L.oracle_synthetic_swissprot_like.argtypes = [C.c_size_t, C.c_uint64]
    L.oracle_synthetic_long_uniform.argtypes = [C.c_size_t, C.c_uint64, C.c_size_t, C.c_size_t]
    L.oracle_synthetic_count.restype = C.c_size_t
    L.oracle_synthetic_count.argtypes = [C.c_void_p]
    L.oracle_synthetic_residues.restype = C.POINTER(C.c_uint8)
    L.oracle_synthetic_residues.argtypes = [C.c_void_p]
    L.oracle_synthetic_offsets.restype = C.POINTER(C.c_uint64)
    L.oracle_synthetic_offsets.argtypes = [C.c_void_p]
    L.oracle_synthetic_free.argtypes = [C.c_void_p]
    h = (L.oracle_synthetic_swissprot_like(count, seed) if kind == "swissprot_like"
         else L.oracle_synthetic_long_uniform(count, seed, shortest, longest))
    try:
        n = L.oracle_synthetic_count(h)
        offsets = np.ctypeslib.as_array(L.oracle_synthetic_offsets(h), shape=(n + 1,)).copy()
        total = int(offsets[-1])
        codes = np.ctypeslib.as_array(L.oracle_synthetic_residues(h), shape=(max(total, 1),))[:total].copy()
        return codes, offsets
    finally:
        L.oracle_synthetic_free(h)


class RefLib:
    """The reference's own code (only where oracle/_ref/libmsv_ref.so exists)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self) -> None:
        L = self.lib = C.CDLL(REF_SO)
        vp = C.c_void_p
        L.ref_profile_load.restype = vp
        L.ref_profile_load.argtypes = [C.c_char_p]
        L.ref_profile_free.argtypes = [vp]
        L.ref_profile_model_length.restype = C.c_size_t
        L.ref_profile_model_length.argtypes = [vp]
        L.ref_profile_name.restype = C.c_char_p
        L.ref_profile_name.argtypes = [vp]
        L.ref_profile_rows.restype = C.c_size_t
        L.ref_profile_rows.argtypes = [vp, C.c_int]
        L.ref_profile_copy.argtypes = [vp, C.c_int, _f32p]
        L.ref_profile_stats.argtypes = [vp, _f32p]
        L.ref_fasta_load.restype = vp
        L.ref_fasta_load.argtypes = [C.c_char_p]
        L.ref_fasta_free.argtypes = [vp]
        L.ref_fasta_count.restype = C.c_size_t
        L.ref_fasta_count.argtypes = [vp]
        L.ref_fasta_record.restype = C.c_char_p
        L.ref_fasta_record.argtypes = [vp, C.c_size_t]
        L.ref_msv_create.restype = vp
        L.ref_msv_create.argtypes = [vp]
        L.ref_msv_free.argtypes = [vp]
        L.ref_msv_model_length.restype = C.c_size_t
        L.ref_msv_model_length.argtypes = [vp]
        L.ref_msv_copy_table.argtypes = [vp, _f32p]
        L.ref_msv_transitions.argtypes = [vp, _f32p]
        L.ref_msv_run_on_sequence.restype = C.c_float
        L.ref_msv_run_on_sequence.argtypes = [vp, C.c_char_p]
        L.ref_msv_run_batch.argtypes = [vp, _u8p, _u64p, C.c_size_t, _f32p, C.c_int]

    def load_hmm(self, path: str) -> dict:
        h = self.lib.ref_profile_load(path.encode())
        try:
            m = self.lib.ref_profile_model_length(h)
            out = {"model_length": int(m), "name": self.lib.ref_profile_name(h).decode()}
            for which, key, cols in ((0, "match_emissions", 20), (1, "insert_emissions", 20), (2, "transitions", 7)):
                rows = self.lib.ref_profile_rows(h, which)
                buf = np.empty((rows, cols), dtype=np.float32)
                self.lib.ref_profile_copy(h, which, buf)
                out[key] = buf
            st = np.empty(6, np.float32)
            self.lib.ref_profile_stats(h, st)
            out["stats"] = st
            return out
        finally:
            self.lib.ref_profile_free(h)

    def load_fasta(self, path: str) -> list[str]:
        fa = self.lib.ref_fasta_load(path.encode())
        try:
            return [self.lib.ref_fasta_record(fa, i).decode() for i in range(self.lib.ref_fasta_count(fa))]
        finally:
            self.lib.ref_fasta_free(fa)

    class Model:
        def __init__(self, lib, hmm_path: str) -> None:
            self.lib = lib
            prof = lib.ref_profile_load(hmm_path.encode())
            self.handle = lib.ref_msv_create(prof)
            lib.ref_profile_free(prof)
            self.model_length = int(lib.ref_msv_model_length(self.handle))

        def table(self) -> tuple[np.ndarray, np.ndarray]:
            t = np.empty((20, self.model_length), np.float32)
            self.lib.ref_msv_copy_table(self.handle, t)
            tr3 = np.empty(3, np.float32)
            self.lib.ref_msv_transitions(self.handle, tr3)
            return t, tr3

        def run_on_sequence(self, seq: str) -> np.float32:
            return np.float32(self.lib.ref_msv_run_on_sequence(self.handle, seq.encode()))

        def run_batch(self, codes, offsets, threads: int = 1) -> np.ndarray:
            n = len(offsets) - 1
            out = np.empty(n, np.float32)
            codes = np.ascontiguousarray(codes, np.uint8)
            if codes.size == 0:
                codes = np.zeros(1, np.uint8)
            self.lib.ref_msv_run_batch(self.handle, codes, np.ascontiguousarray(offsets, np.uint64), n, out, threads)
            return out

        def __del__(self) -> None:
            try:
                self.lib.ref_msv_free(self.handle)
            except Exception:
                pass

    def model(self, hmm_path: str) -> "RefLib.Model":
        return RefLib.Model(self.lib, hmm_path)
