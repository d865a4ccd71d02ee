"""ctypes binding of the C ABI in include/msv_cuda.h (libmsv_cuda.so, sm_100a CUDA).

There is deliberately no fallback: if the library is missing the import fails, and if no B200 is present every
``msv_cuda_*`` call raises ``MsvCudaError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSV_CUDA_LIBRARY") or os.path.join(_PKG, "libmsv_cuda.so")  # (the override is a kernel-development aid)

MSV_OK = 0
MSV_ERR_INVALID_ARGUMENT = -1
MSV_ERR_NO_DEVICE = -2
MSV_ERR_CUDA = -3
MSV_ERR_BAD_RESIDUE = -4
MSV_ERR_MODEL_TOO_LONG = -5
MSV_ERR_OUT_OF_MEMORY = -6

# every symbol include/msv_cuda.h declares (tests check that the .so exports all of them)
DECLARED_SYMBOLS = (
    "msv_cuda_abi_version", "msv_cuda_last_error", "msv_cuda_device_count",
    "msv_host_emission_table", "msv_host_model_transitions", "msv_host_length_transitions", "msv_host_encode",
    "msv_host_partition_by_cells",
    "msv_cuda_model_create", "msv_cuda_model_destroy", "msv_cuda_model_geometry", "msv_cuda_model_plan", "msv_cuda_model_plan_long_sequences", "msv_cuda_model_speculation",
    "msv_cuda_db_create", "msv_cuda_db_destroy", "msv_cuda_db_info",
    "msv_cuda_db_score_device", "msv_cuda_db_score_gather", "msv_cuda_db_score", "msv_cuda_score_batch", "msv_cuda_score_sequence",
    "msv_cuda_db_filter_device", "msv_cuda_db_score_filter", "msv_cuda_host_register", "msv_cuda_host_unregister",
    "msv_cuda_launch_count",
    "msv_cuda_score_batch_gather", "msv_cuda_model_device", "msv_cuda_model_wave_geometry",
    "msv_cuda_db_create_from_fasta", "msv_cuda_db_refill_from_fasta", "msv_cuda_db_download", "msv_cuda_score_fasta",
    "msv_cuda_db_msv_filter", "msv_cuda_db_viterbi_subset_device", "msv_cuda_db_viterbi_filter_survivors",
    "msv_cuda_multi_create", "msv_cuda_multi_destroy", "msv_cuda_multi_score_batch", "msv_cuda_multi_gathered",
    "msv_host_viterbi_transitions", "msv_cuda_viterbi_model_create", "msv_cuda_viterbi_model_destroy",
    "msv_cuda_viterbi_model_geometry", "msv_cuda_db_viterbi_device", "msv_cuda_db_viterbi", "msv_cuda_db_viterbi_filter", "msv_cuda_viterbi_batch",
)


class MsvCudaError(RuntimeError):
    def __init__(self, status: int, message: str) -> None:
        super().__init__(f"[msv_cuda status {status}] {message}")
        self.status = status


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the MSV scan.")
    return C.CDLL(LIB_PATH)


lib = _load()

_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64 = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_fp = C.POINTER(C.c_float)

lib.msv_cuda_abi_version.restype = C.c_int
lib.msv_cuda_last_error.restype = C.c_char_p
lib.msv_cuda_device_count.argtypes = [C.POINTER(C.c_int)]
lib.msv_host_emission_table.argtypes = [_f32, C.c_size_t, _f32]
lib.msv_host_model_transitions.argtypes = [C.c_size_t, _fp, _fp, _fp]
lib.msv_host_length_transitions.argtypes = [C.c_size_t, _fp, _fp]
lib.msv_host_encode.argtypes = [C.c_char_p, C.c_size_t, _u8, C.POINTER(C.c_size_t)]
lib.msv_host_partition_by_cells.argtypes = [_u64, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]
lib.msv_cuda_model_create.argtypes = [_f32, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_void_p)]
lib.msv_cuda_model_destroy.argtypes = [C.c_void_p]
lib.msv_cuda_model_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
lib.msv_cuda_model_plan.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.msv_cuda_model_plan_long_sequences.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.msv_cuda_model_speculation.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_int)]
lib.msv_cuda_db_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
lib.msv_cuda_db_destroy.argtypes = [C.c_void_p]
lib.msv_cuda_db_info.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
lib.msv_cuda_db_score_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.msv_cuda_db_score_gather.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_void_p]
lib.msv_cuda_db_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.msv_cuda_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
lib.msv_cuda_score_sequence.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, _fp]
lib.msv_cuda_db_filter_device.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
lib.msv_cuda_db_score_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
lib.msv_cuda_host_register.argtypes = [C.c_void_p, C.c_size_t]
lib.msv_cuda_host_unregister.argtypes = [C.c_void_p]
lib.msv_host_viterbi_transitions.argtypes = [_f32, C.c_size_t, _f32]
lib.msv_cuda_viterbi_model_create.argtypes = [_f32, _f32, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_int,
                                              C.POINTER(C.c_void_p)]
lib.msv_cuda_viterbi_model_destroy.argtypes = [C.c_void_p]
lib.msv_cuda_viterbi_model_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
lib.msv_cuda_db_viterbi_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.msv_cuda_db_viterbi.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.msv_cuda_db_viterbi_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
lib.msv_cuda_viterbi_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
lib.msv_cuda_score_batch_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_int, C.c_size_t]
lib.msv_cuda_model_device.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
lib.msv_cuda_db_msv_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.POINTER(C.c_size_t)]
lib.msv_cuda_db_viterbi_subset_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
lib.msv_cuda_db_viterbi_filter_survivors.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
lib.msv_cuda_db_create_from_fasta.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
lib.msv_cuda_db_refill_from_fasta.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
lib.msv_cuda_db_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.msv_cuda_score_fasta.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
lib.msv_cuda_model_wave_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.msv_cuda_multi_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
lib.msv_cuda_multi_destroy.argtypes = [C.c_void_p]
lib.msv_cuda_multi_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
lib.msv_cuda_multi_gathered.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
lib.msv_cuda_launch_count.restype = C.c_uint64
lib.msv_cuda_launch_count.argtypes = [C.c_int]
for _name in DECLARED_SYMBOLS:
    if _name not in ("msv_cuda_last_error", "msv_cuda_launch_count", "msv_cuda_abi_version"):
        getattr(lib, _name).restype = C.c_int


def check(status: int) -> None:
    if status != MSV_OK:
        raise MsvCudaError(status, lib.msv_cuda_last_error().decode(errors="replace"))


def device_count() -> int:
    n = C.c_int(0)
    status = lib.msv_cuda_device_count(C.byref(n))
    return n.value if status == MSV_OK else 0


def launch_count(reset: bool = False) -> int:
    return int(lib.msv_cuda_launch_count(1 if reset else 0))


# ---- host-side model arithmetic -----------------------------------------------------------------------------------
def emission_table(match_emissions: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(match_emissions, np.float32)
    table = np.empty((20, m.shape[0]), np.float32)
    check(lib.msv_host_emission_table(m, m.shape[0], table))
    return table


def model_transitions(model_length: int) -> tuple[np.float32, np.float32, np.float32]:
    a, b, c = C.c_float(), C.c_float(), C.c_float()
    check(lib.msv_host_model_transitions(model_length, C.byref(a), C.byref(b), C.byref(c)))
    return np.float32(a.value), np.float32(b.value), np.float32(c.value)


def length_transitions(residues: int) -> tuple[np.float32, np.float32]:
    a, b = C.c_float(), C.c_float()
    check(lib.msv_host_length_transitions(residues, C.byref(a), C.byref(b)))
    return np.float32(a.value), np.float32(b.value)


def viterbi_transitions(transitions: np.ndarray) -> np.ndarray:
    """logf of the node transition probabilities Profile_HMM parses ([model_length][7])."""
    t = np.ascontiguousarray(transitions, np.float32)
    out = np.empty_like(t)
    check(lib.msv_host_viterbi_transitions(t, t.shape[0], out))
    return out


def encode(letters: str) -> np.ndarray:
    raw = letters.encode("latin-1")
    codes = np.empty(max(len(raw), 1), np.uint8)
    bad = C.c_size_t(0)
    status = lib.msv_host_encode(raw, len(raw), codes, C.byref(bad))
    if status == MSV_ERR_BAD_RESIDUE:
        raise KeyError(lib.msv_cuda_last_error().decode())
    check(status)
    return codes[: len(raw)]


def partition_by_cells(offsets: np.ndarray, parts: int) -> np.ndarray:
    offsets = np.ascontiguousarray(offsets, np.uint64)
    bounds = (C.c_size_t * (parts + 1))()
    check(lib.msv_host_partition_by_cells(offsets, len(offsets) - 1, parts, bounds))
    return np.array(list(bounds), dtype=np.int64)


# ---- device objects -------------------------------------------------------------------------------------------------
def _ptr(a) -> int:
    """Address of a numpy array / torch tensor / raw int."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())  # torch tensor


def _host_batch(residues, offsets, out):
    """Normalise the host buffers of an end-to-end call: uint8 codes, uint64 offsets (n+1), float32 result -- all
    C-contiguous.  numpy inputs of another dtype/stride are converted; torch tensors (pinned buffers) must already fit."""
    def conform(a, np_dtype, what):
        if isinstance(a, np.ndarray):
            return np.ascontiguousarray(a, np_dtype)
        if hasattr(a, "data_ptr"):  # torch tensor: check instead of converting (a copy would lose the pinning)
            itemsize = np.dtype(np_dtype).itemsize
            if a.element_size() != itemsize or not a.is_contiguous() or a.is_floating_point() != (np_dtype == np.float32):
                raise TypeError(f"{what}: expected a contiguous {np.dtype(np_dtype).name} buffer, got {a.dtype} (contiguous={a.is_contiguous()})")
            return a
        return np.ascontiguousarray(a, np_dtype)
    offsets = conform(offsets, np.uint64, "offsets")
    n = len(offsets) - 1
    if n < 0:
        raise ValueError("offsets needs n + 1 entries")
    residues = conform(residues, np.uint8, "residues")
    if out is None:
        out = np.empty(n, np.float32)
    elif isinstance(out, np.ndarray):
        if out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or out.size < n:
            raise TypeError("out: expected a C-contiguous float32 array of at least n entries")
    else:
        conform(out, np.float32, "out")
        if out.numel() < n:
            raise TypeError("out: fewer than n entries")
    return residues, offsets, n, out


class Model:
    """Device-resident model (msv_model*)."""

    def __init__(self, emission_scores: np.ndarray, tr_B_Mk, tr_E_C, tr_E_J, device: int = 0) -> None:
        table = np.ascontiguousarray(emission_scores, np.float32)
        assert table.ndim == 2 and table.shape[0] == 20
        self.model_length = int(table.shape[1])
        self.device = device
        h = C.c_void_p()
        check(lib.msv_cuda_model_create(table, self.model_length, float(tr_B_Mk), float(tr_E_C), float(tr_E_J), device,
                                        C.byref(h)))
        self.handle = h

    @property
    def geometry(self) -> dict:
        g, k, kt, t, s = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        check(lib.msv_cuda_model_geometry(self.handle, C.byref(g), C.byref(k), C.byref(kt), C.byref(t), C.byref(s)))
        return {"lanes_per_sequence": g.value, "columns_per_lane": k.value, "tensor_columns_per_lane": kt.value,
                "threads_per_cta": t.value, "shared_bytes": s.value}

    @property
    def wave_geometry(self) -> dict:
        """Single-sequence latency kernels: CTAs of the diagonal-worker kernel (0 = model too long for it), and the chain
        kernel's columns per lane (0 = none), warps in the chain, CTAs in the cluster."""
        k, w, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib.msv_cuda_model_wave_geometry(self.handle, C.byref(k), C.byref(w), C.byref(c), C.byref(d)))
        return {"columns_per_lane": k.value, "warps": w.value, "ctas": c.value, "diagonal_ctas": d.value}

    @property
    def speculation(self) -> dict:
        """Running totals behind the choice of the speculating kernel (see msv_cuda.h) and what the next scan would use."""
        f, o, b = C.c_uint(), C.c_uint(), C.c_int()
        check(lib.msv_cuda_model_speculation(self.handle, C.byref(f), C.byref(o), C.byref(b)))
        return {"failed": f.value, "scanned": o.value, "rows_next": ("exact", "whole", "blocks")[b.value]}

    def plan(self, database: "Database") -> dict:
        """Launch plan a scan of `database` would use: kernel family (lanes per sequence) and sequences per CTA."""
        lanes, per_cta = C.c_int(), C.c_int()
        check(lib.msv_cuda_model_plan(self.handle, database.handle, C.byref(lanes), C.byref(per_cta)))
        n_long, fast_ctas, fast_slots = C.c_uint(), C.c_int(), C.c_int()
        check(lib.msv_cuda_model_plan_long_sequences(self.handle, database.handle, C.byref(n_long), C.byref(fast_ctas), C.byref(fast_slots)))
        plan = {"lanes_per_sequence": lanes.value, "sequences_per_cta": per_cta.value}
        if fast_ctas.value:  # lane-group plans: the longest sequences on a few CTAs with fewer, faster slots
            plan |= {"long_sequences": n_long.value, "fast_ctas": fast_ctas.value, "fast_sequences_per_cta": fast_slots.value}
        return plan

    def score_batch(self, residues, offsets, out=None) -> np.ndarray:
        """End-to-end call with host buffers (numpy arrays or pinned torch tensors)."""
        residues, offsets, n, out = _host_batch(residues, offsets, out)
        check(lib.msv_cuda_score_batch(self.handle, _ptr(residues), _ptr(offsets), n, _ptr(out)))
        return out

    def score_batch_gather(self, residues, offsets, gathered, first_index: int) -> None:
        """End-to-end call of a sharded run: host buffers of this rank's slice in, scores out through the fused gather
        (``gathered``: this GPU's copy of the whole score array first, then the peers' copies; device pointers / CUDA tensors)."""
        residues, offsets, n, _ = _host_batch(residues, offsets, None)
        ptrs = (C.c_void_p * len(gathered))(*[_ptr(g) for g in gathered])
        check(lib.msv_cuda_score_batch_gather(self.handle, _ptr(residues), _ptr(offsets), n, ptrs, len(gathered), first_index))

    def score_fasta(self, text) -> tuple[np.ndarray, int]:
        """FASTA text (bytes, or a uint8 numpy array / np.memmap of a file) -> (scores, rejected records).  The raw text is
        uploaded and parsed on the GPU (msv_cuda_db_refill_from_fasta into a database handle this model keeps), then scanned."""
        buf = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else text
        if not hasattr(self, "_fasta_db") or self._fasta_db is None:
            self._fasta_db = Database(np.zeros(0, np.uint8), np.zeros(1, np.uint64), device=self.device)
        rejected = self._fasta_db.refill_from_fasta(buf)
        return self._fasta_db.score(self), rejected

    def score_fasta_file(self, path: str) -> tuple[np.ndarray, int]:
        """The same for a file: mapped, not read -- the pages go from the page cache straight into the pinned staging ring."""
        if os.path.getsize(path) == 0:
            return np.zeros(0, np.float32), 0
        return self.score_fasta(np.memmap(path, dtype=np.uint8, mode="r"))

    def score_sequence(self, codes: np.ndarray) -> np.float32:
        codes = np.ascontiguousarray(codes, np.uint8)
        out = C.c_float()
        check(lib.msv_cuda_score_sequence(self.handle, codes.ctypes.data if codes.size else 0, codes.size, C.byref(out)))
        return np.float32(out.value)

    def close(self) -> None:
        if getattr(self, "_fasta_db", None) is not None:
            self._fasta_db.close()
            self._fasta_db = None
        if getattr(self, "handle", None):
            lib.msv_cuda_model_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass


GATHER_HOST, GATHER_PEER, GATHER_NCCL = 0, 1, 2


class MultiGpu:
    """Several GPUs of one box from ONE process (msv_multi*): one ``Model`` per GPU, the host database cut by cell count."""

    def __init__(self, models) -> None:
        self.models = list(models)  # keeps them alive
        handles = (C.c_void_p * len(self.models))(*[m.handle for m in self.models])
        h = C.c_void_p()
        check(lib.msv_cuda_multi_create(handles, len(self.models), C.byref(h)))
        self.handle = h

    def score_batch(self, residues, offsets, out=None, gather: int = GATHER_HOST) -> np.ndarray:
        residues, offsets, n, out = _host_batch(residues, offsets, out)
        check(lib.msv_cuda_multi_score_batch(self.handle, _ptr(residues), _ptr(offsets), n, _ptr(out), gather))
        return out

    def gathered(self) -> tuple[int, int, int]:
        """(device pointer, n, device) of the whole job's scores left on the first GPU by a PEER / NCCL call."""
        p, n, d = C.c_void_p(), C.c_size_t(), C.c_int()
        check(lib.msv_cuda_multi_gathered(self.handle, C.byref(p), C.byref(n), C.byref(d)))
        return int(p.value or 0), int(n.value), int(d.value)

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib.msv_cuda_multi_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass


class ViterbiModel:
    """Device-resident model of the Plan-7 local Viterbi scan (msv_viterbi_model*)."""

    def __init__(self, emission_scores: np.ndarray, log_transitions: np.ndarray, tr_B_Mk, tr_E_C, tr_E_J, device: int = 0) -> None:
        table = np.ascontiguousarray(emission_scores, np.float32)
        logtr = np.ascontiguousarray(log_transitions, np.float32)
        assert table.ndim == 2 and table.shape[0] == 20 and logtr.shape == (table.shape[1], 7)
        self.model_length = int(table.shape[1])
        self.device = device
        h = C.c_void_p()
        check(lib.msv_cuda_viterbi_model_create(table, logtr, self.model_length, float(tr_B_Mk), float(tr_E_C), float(tr_E_J),
                                                device, C.byref(h)))
        self.handle = h

    @property
    def geometry(self) -> dict:
        k, t, s = C.c_int(), C.c_int(), C.c_size_t()
        check(lib.msv_cuda_viterbi_model_geometry(self.handle, C.byref(k), C.byref(t), C.byref(s)))
        return {"lanes_per_sequence": 32, "columns_per_lane": k.value, "threads_per_cta": t.value, "shared_bytes": s.value}

    def score_batch(self, residues, offsets, out=None) -> np.ndarray:
        residues, offsets, n, out = _host_batch(residues, offsets, out)
        check(lib.msv_cuda_viterbi_batch(self.handle, _ptr(residues), _ptr(offsets), n, _ptr(out)))
        return out

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib.msv_cuda_viterbi_model_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass


class Database:
    """Device-resident packed database (msv_db*)."""

    def __init__(self, residues, offsets, device: int = 0) -> None:
        self.n = len(offsets) - 1
        h = C.c_void_p()
        if isinstance(residues, np.ndarray):
            residues = np.ascontiguousarray(residues, np.uint8)
        if isinstance(offsets, np.ndarray):
            offsets = np.ascontiguousarray(offsets, np.uint64)
        check(lib.msv_cuda_db_create(device, _ptr(residues), _ptr(offsets), self.n, C.byref(h)))
        self.handle = h
        self.device = device

    @classmethod
    def from_fasta(cls, text, device: int = 0) -> "Database":
        """FASTA text (bytes / uint8 array / np.memmap) parsed on the GPU; `.rejected` = records dropped for a foreign character."""
        self = cls(np.zeros(0, np.uint8), np.zeros(1, np.uint64), device=device)
        self.rejected = self.refill_from_fasta(text)
        return self

    def refill_from_fasta(self, text) -> int:
        buf = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else text
        rejected = C.c_size_t(0)
        check(lib.msv_cuda_db_refill_from_fasta(self.handle, buf.ctypes.data if buf.size else 0, buf.size, C.byref(rejected)))
        self.n = self.info()["n"]
        return int(rejected.value)

    def download(self) -> tuple[np.ndarray, np.ndarray]:
        """(residues uint8, offsets uint64[n+1]) as they sit on the device."""
        meta = self.info()
        residues = np.empty(max(meta["total_residues"], 1), np.uint8)
        offsets = np.empty(meta["n"] + 1, np.uint64)
        check(lib.msv_cuda_db_download(self.handle, residues.ctypes.data, offsets.ctypes.data))
        return residues[: meta["total_residues"]], offsets

    def info(self) -> dict:
        n, total, longest = C.c_size_t(), C.c_uint64(), C.c_uint64()
        check(lib.msv_cuda_db_info(self.handle, C.byref(n), C.byref(total), C.byref(longest)))
        return {"n": n.value, "total_residues": total.value, "longest": longest.value}

    def score(self, model: Model) -> np.ndarray:
        out = np.empty(self.n, np.float32)
        check(lib.msv_cuda_db_score(model.handle, self.handle, out.ctypes.data))
        return out

    def score_device(self, model: Model, scores_device, stream: int = 0) -> None:
        """Asynchronous scan into a device buffer (torch CUDA tensor or raw pointer) on `stream`."""
        check(lib.msv_cuda_db_score_device(model.handle, self.handle, _ptr(scores_device), stream))

    def viterbi(self, model: "ViterbiModel") -> np.ndarray:
        out = np.empty(self.n, np.float32)
        check(lib.msv_cuda_db_viterbi(model.handle, self.handle, out.ctypes.data))
        return out

    def viterbi_filter(self, model: "ViterbiModel", mu: float, lam: float):
        """Raw scores, bit scores and Gumbel P-values of the Viterbi scan (host arrays)."""
        scores, bits, p = (np.empty(self.n, np.float32) for _ in range(3))
        check(lib.msv_cuda_db_viterbi_filter(model.handle, self.handle, float(mu), float(lam), scores.ctypes.data, bits.ctypes.data,
                                             p.ctypes.data))
        return scores, bits, p

    def viterbi_device(self, model: "ViterbiModel", scores_device, stream: int = 0) -> None:
        check(lib.msv_cuda_db_viterbi_device(model.handle, self.handle, _ptr(scores_device), stream))

    def score_gather(self, model: Model, gathered, first_index: int, stream: int = 0) -> None:
        """Scan fused with the gather: scores go to gathered[r][first_index + q] for every array of `gathered` (this GPU's and
        the peers' copies of the whole score array: device pointers or torch CUDA tensors)."""
        ptrs = (C.c_void_p * len(gathered))(*[_ptr(g) for g in gathered])
        check(lib.msv_cuda_db_score_gather(model.handle, self.handle, ptrs, len(gathered), first_index, stream))

    def filter_device(self, scores_device, mu: float, lam: float, bits_device=None, pvalues_device=None, stream: int = 0) -> None:
        """Bit scores and Gumbel P-values (HMMER3 MSV filter conventions) from raw scores resident on the device."""
        check(lib.msv_cuda_db_filter_device(self.handle, _ptr(scores_device), float(mu), float(lam), _ptr(bits_device),
                                            _ptr(pvalues_device), stream))

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib.msv_cuda_db_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass
