"""Python mirror of the reference's host interface, driving the C++ host layer (libmsv_host.so) through ctypes.

Class and method names follow the reference (data_readers/Profile_HMM.hpp:21-49,
data_readers/FASTA_protein_sequences.hpp:9-14, algorithms/MSV_HMM.hpp:17-24) so that parity tests read like the
reference's own tests.  All parsing and all arithmetic happen in the C++/CUDA libraries; this file only marshals.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi  # noqa: F401  (loads libmsv_cuda.so first so that libmsv_host.so resolves against it)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmsv_host.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
lib = C.CDLL(LIB_PATH)

_vp = C.c_void_p
_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
lib.msvh_last_error.restype = C.c_char_p
lib.msvh_profile_load.restype = _vp
lib.msvh_profile_load.argtypes = [C.c_char_p]
lib.msvh_profile_free.argtypes = [_vp]
lib.msvh_profile_model_length.restype = C.c_size_t
lib.msvh_profile_model_length.argtypes = [_vp]
lib.msvh_profile_name.restype = C.c_char_p
lib.msvh_profile_name.argtypes = [_vp]
lib.msvh_profile_rows.restype = C.c_size_t
lib.msvh_profile_rows.argtypes = [_vp, C.c_int]
lib.msvh_profile_copy.argtypes = [_vp, C.c_int, _f32]
lib.msvh_profile_stats.argtypes = [_vp, _f32]
lib.msvh_fasta_load.restype = _vp
lib.msvh_fasta_load.argtypes = [C.c_char_p]
lib.msvh_fasta_free.argtypes = [_vp]
lib.msvh_fasta_count.restype = C.c_size_t
lib.msvh_fasta_count.argtypes = [_vp]
lib.msvh_fasta_record.restype = C.c_char_p
lib.msvh_fasta_record.argtypes = [_vp, C.c_size_t]
lib.msvh_packed_from_fasta_file.restype = _vp
lib.msvh_packed_from_fasta_file.argtypes = [C.c_char_p, C.POINTER(C.c_size_t)]
lib.msvh_packed_from_fasta.restype = _vp
lib.msvh_packed_from_fasta.argtypes = [_vp]
lib.msvh_packed_synthetic_swissprot_like.restype = _vp
lib.msvh_packed_synthetic_swissprot_like.argtypes = [C.c_size_t, C.c_uint64]
lib.msvh_packed_synthetic_long_uniform.restype = _vp
lib.msvh_packed_synthetic_long_uniform.argtypes = [C.c_size_t, C.c_uint64, C.c_size_t, C.c_size_t]
lib.msvh_packed_from_arrays.restype = _vp
lib.msvh_packed_from_arrays.argtypes = [_vp, _vp, C.c_size_t]
lib.msvh_packed_free.argtypes = [_vp]
lib.msvh_packed_count.restype = C.c_size_t
lib.msvh_packed_count.argtypes = [_vp]
lib.msvh_packed_total.restype = C.c_uint64
lib.msvh_packed_total.argtypes = [_vp]
lib.msvh_packed_residues.restype = C.POINTER(C.c_uint8)
lib.msvh_packed_residues.argtypes = [_vp]
lib.msvh_packed_offsets.restype = C.POINTER(C.c_uint64)
lib.msvh_packed_offsets.argtypes = [_vp]
lib.msvh_msv_create.restype = _vp
lib.msvh_msv_create.argtypes = [_vp, C.c_int]
lib.msvh_msv_clone.restype = _vp
lib.msvh_msv_clone.argtypes = [_vp]
lib.msvh_msv_free.argtypes = [_vp]
lib.msvh_msv_length.restype = C.c_size_t
lib.msvh_msv_length.argtypes = [_vp]
lib.msvh_msv_run_on_sequence.argtypes = [_vp, C.c_char_p, C.POINTER(C.c_float)]
lib.msvh_msv_parallel_run_on_sequence.argtypes = [_vp, C.c_char_p, C.c_int, C.POINTER(C.c_float)]
lib.msvh_msv_parallel_run_on_packed.argtypes = [_vp, _vp, _f32]
lib.msvh_device_database_create.restype = _vp
lib.msvh_device_database_create.argtypes = [_vp, C.c_int]
lib.msvh_device_database_free.argtypes = [_vp]
lib.msvh_msv_parallel_run_on_device_database.argtypes = [_vp, _vp, _f32]
lib.msvh_viterbi_create.restype = _vp
lib.msvh_viterbi_create.argtypes = [_vp, C.c_int]
lib.msvh_viterbi_free.argtypes = [_vp]
lib.msvh_viterbi_parallel_run_on_sequence.argtypes = [_vp, C.c_char_p, C.POINTER(C.c_float)]
lib.msvh_viterbi_parallel_run_on_packed.argtypes = [_vp, _vp, _f32]
lib.msvh_viterbi_parallel_run_on_device_database.argtypes = [_vp, _vp, _f32]
lib.msvh_viterbi_filter.restype = C.c_long
lib.msvh_viterbi_filter.argtypes = [_vp, _vp, C.c_float, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.msvh_viterbi_filter_survivors.restype = C.c_long
lib.msvh_viterbi_filter_survivors.argtypes = [_vp, _vp, C.c_float, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.msvh_msv_filter.restype = C.c_long
lib.msvh_msv_filter.argtypes = [_vp, _vp, C.c_float, C.c_size_t, _vp, _vp, _vp, _vp]
lib.msvh_msv_parallel_run_on_packed_devices.argtypes = [_vp, _vp, C.POINTER(C.c_int), C.c_int, C.c_int, _f32]


def _raise(status: int) -> None:
    message = lib.msvh_last_error().decode(errors="replace")
    if status == -2:
        raise KeyError(message)  # std::out_of_range in C++ (foreign residue letter)
    raise RuntimeError(message)


class Profile_HMM:
    """data_readers/Profile_HMM.hpp:21-49."""

    def __init__(self, file_path: str) -> None:
        self._h = lib.msvh_profile_load(os.fspath(file_path).encode())
        if not self._h:
            _raise(-1)
        self.model_length = int(lib.msvh_profile_model_length(self._h))
        self.name = lib.msvh_profile_name(self._h).decode()
        mats = []
        for which, cols in ((0, 20), (1, 20), (2, 7)):
            rows = lib.msvh_profile_rows(self._h, which)
            buf = np.empty((rows, cols), np.float32)
            lib.msvh_profile_copy(self._h, which, buf)
            mats.append(buf)
        self.match_emissions, self.insert_emissions, self.transitions = mats
        st = np.empty(6, np.float32)
        lib.msvh_profile_stats(self._h, st)
        (self.stats_local_msv_mu, self.stats_local_msv_lambda, self.stats_local_viterbi_mu, self.stats_local_viterbi_lambda,
         self.stats_local_forward_theta, self.stats_local_forward_lambda) = (np.float32(v) for v in st)

    def __del__(self) -> None:
        if getattr(self, "_h", None):
            lib.msvh_profile_free(self._h)
            self._h = None


class FASTA_protein_sequences:
    """data_readers/FASTA_protein_sequences.hpp:9-14: `.sequences` is a list of '#'-prefixed strings."""

    def __init__(self, file_path: str) -> None:
        self._h = lib.msvh_fasta_load(os.fspath(file_path).encode())
        if not self._h:
            _raise(-1)
        self.sequences = [lib.msvh_fasta_record(self._h, i).decode("latin-1") for i in range(lib.msvh_fasta_count(self._h))]

    def __del__(self) -> None:
        if getattr(self, "_h", None):
            lib.msvh_fasta_free(self._h)
            self._h = None


class Packed_sequences:
    """Packed device-facing layout (host/data_readers/Packed_sequences.hpp): uint8 codes + uint64 offsets."""

    def __init__(self, handle) -> None:
        if not handle:
            _raise(-1)
        self._h = handle

    @classmethod
    def from_fasta_file(cls, path: str) -> "Packed_sequences":
        rejected = C.c_size_t(0)
        self = cls(lib.msvh_packed_from_fasta_file(os.fspath(path).encode(), C.byref(rejected)))
        self.rejected = rejected.value
        return self

    @classmethod
    def from_fasta(cls, fasta: FASTA_protein_sequences) -> "Packed_sequences":
        return cls(lib.msvh_packed_from_fasta(fasta._h))

    @classmethod
    def from_arrays(cls, residues: np.ndarray, offsets: np.ndarray) -> "Packed_sequences":
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        return cls(lib.msvh_packed_from_arrays(residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1))

    @classmethod
    def synthetic_swissprot_like(cls, count: int, seed: int) -> "Packed_sequences":
        return cls(lib.msvh_packed_synthetic_swissprot_like(count, seed))

    @classmethod
    def synthetic_long_uniform(cls, count: int, seed: int, shortest: int, longest: int) -> "Packed_sequences":
        return cls(lib.msvh_packed_synthetic_long_uniform(count, seed, shortest, longest))

    def __len__(self) -> int:
        return int(lib.msvh_packed_count(self._h))

    @property
    def total_residues(self) -> int:
        return int(lib.msvh_packed_total(self._h))

    @property
    def residues(self) -> np.ndarray:
        """Zero-copy view (valid while this object lives)."""
        n = self.total_residues
        if n == 0:
            return np.zeros(0, np.uint8)
        return np.ctypeslib.as_array(lib.msvh_packed_residues(self._h), shape=(n,))

    @property
    def offsets(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib.msvh_packed_offsets(self._h), shape=(len(self) + 1,))

    def __del__(self) -> None:
        if getattr(self, "_h", None):
            lib.msvh_packed_free(self._h)
            self._h = None


class Device_database:
    """A Packed_sequences uploaded once and kept in HBM (host/algorithms/MSV_HMM.hpp), scanned by any number of models."""

    def __init__(self, packed: Packed_sequences, device: int = 0) -> None:
        self._h = lib.msvh_device_database_create(packed._h, device)
        if not self._h:
            _raise(-1)
        self._n = len(packed)

    def __len__(self) -> int:
        return self._n

    def __del__(self) -> None:
        if getattr(self, "_h", None):
            lib.msvh_device_database_free(self._h)
            self._h = None


class MSV_HMM:
    """algorithms/MSV_HMM.hpp:17-24 plus the batch entry point this implementation adds."""

    def __init__(self, base_hmm: Profile_HMM, device: int = 0) -> None:
        self._h = lib.msvh_msv_create(base_hmm._h, device)
        if not self._h:
            _raise(-1)
        self.model_length = int(lib.msvh_msv_length(self._h))

    def run_on_sequence(self, seq: str) -> np.float32:
        out = C.c_float()
        status = lib.msvh_msv_run_on_sequence(self._h, seq.encode("latin-1"), C.byref(out))
        if status:
            _raise(status)
        return np.float32(out.value)

    def parallel_run_on_sequence(self, seq: str, should_specialize: bool = False) -> np.float32:
        out = C.c_float()
        status = lib.msvh_msv_parallel_run_on_sequence(self._h, seq.encode("latin-1"), int(should_specialize), C.byref(out))
        if status:
            _raise(status)
        return np.float32(out.value)

    def parallel_run_on_sequences(self, database, devices=None, gather: str = "host") -> np.ndarray:
        """Whole database in one call; `devices` (list of GPU indices) spreads it over several GPUs from this process,
        `gather` = "host" | "peer" | "nccl" picks how the scores come together (MSV_HMM::Score_gather)."""
        if isinstance(database, FASTA_protein_sequences):
            database = Packed_sequences.from_fasta(database)
        out = np.empty(max(len(database), 1), np.float32)
        if isinstance(database, Device_database):
            status = lib.msvh_msv_parallel_run_on_device_database(self._h, database._h, out)
            if status:
                _raise(status)
            return out[: len(database)]
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            status = lib.msvh_msv_parallel_run_on_packed_devices(self._h, database._h, arr, len(devices), {"host": 0, "peer": 1, "nccl": 2}[gather], out)
        else:
            status = lib.msvh_msv_parallel_run_on_packed(self._h, database._h, out)
        if status:
            _raise(status)
        return out[: len(database)]

    def msv_filter(self, database: Device_database, threshold: float = 0.02) -> dict:
        """HMMER3-style MSV filter: sequences with Gumbel P-value <= threshold, as arrays index/score/bits/p_value."""
        cap = len(database)
        index = np.empty(max(cap, 1), np.uint64)
        score, bits, p = (np.empty(max(cap, 1), np.float32) for _ in range(3))
        found = lib.msvh_msv_filter(self._h, database._h, float(threshold), cap, index.ctypes.data, score.ctypes.data,
                                    bits.ctypes.data, p.ctypes.data)
        if found < 0:
            _raise(int(found))
        return {"index": index[:found], "score": score[:found], "bits": bits[:found], "p_value": p[:found]}

    def __del__(self) -> None:
        if getattr(self, "_h", None):
            lib.msvh_msv_free(self._h)
            self._h = None


class Viterbi_HMM:
    """algorithms/Viterbi_HMM.hpp: Plan-7 local Viterbi (match/insert/delete) with MSV_HMM's calling conventions."""

    def __init__(self, base_hmm: Profile_HMM, device: int = 0) -> None:
        self._h = lib.msvh_viterbi_create(base_hmm._h, device)
        if not self._h:
            _raise(-1)

    def parallel_run_on_sequence(self, seq: str) -> np.float32:
        out = C.c_float()
        status = lib.msvh_viterbi_parallel_run_on_sequence(self._h, seq.encode("latin-1"), C.byref(out))
        if status:
            _raise(status)
        return np.float32(out.value)

    def parallel_run_on_sequences(self, database) -> np.ndarray:
        if isinstance(database, FASTA_protein_sequences):
            database = Packed_sequences.from_fasta(database)
        out = np.empty(max(len(database), 1), np.float32)
        if isinstance(database, Device_database):
            status = lib.msvh_viterbi_parallel_run_on_device_database(self._h, database._h, out)
        else:
            status = lib.msvh_viterbi_parallel_run_on_packed(self._h, database._h, out)
        if status:
            _raise(status)
        return out[: len(database)]

    def viterbi_filter(self, database: Device_database, threshold: float = 1e-3) -> dict:
        """HMMER3's second filter stage: sequences with Viterbi Gumbel P-value <= threshold (index/score/bits/p_value)."""
        cap = len(database)
        index = np.empty(max(cap, 1), np.uint64)
        score, bits, p = (np.empty(max(cap, 1), np.float32) for _ in range(3))
        found = lib.msvh_viterbi_filter(self._h, database._h, float(threshold), cap, index.ctypes.data, score.ctypes.data,
                                        bits.ctypes.data, p.ctypes.data)
        if found < 0:
            _raise(int(found))
        return {"index": index[:found], "score": score[:found], "bits": bits[:found], "p_value": p[:found]}

    def viterbi_filter_survivors(self, database: Device_database, threshold: float = 1e-3) -> dict:
        """The Viterbi filter over the survivors of the last MSV_HMM.msv_filter on this database (their index list stayed on
        the GPU): index/score/bits/p_value of the sequences that also pass this stage."""
        cap = len(database)
        index = np.empty(max(cap, 1), np.uint64)
        score, bits, p = (np.empty(max(cap, 1), np.float32) for _ in range(3))
        found = lib.msvh_viterbi_filter_survivors(self._h, database._h, float(threshold), cap, index.ctypes.data, score.ctypes.data,
                                                  bits.ctypes.data, p.ctypes.data)
        if found < 0:
            _raise(int(found))
        return {"index": index[:found], "score": score[:found], "bits": bits[:found], "p_value": p[:found]}

    def __del__(self) -> None:
        if getattr(self, "_h", None):
            lib.msvh_viterbi_free(self._h)
            self._h = None
