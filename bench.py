#!/usr/bin/env python
"""bench.py -- MSV GCUPS at M=1400 on the 1M-sequence synthetic Swiss-Prot-like database (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

A "step" is one scan of the whole database: LENG x sum(L) DP cells.  GCUPS = cells / seconds / 1e9.

  value      : device-resident path.  The packed database already sits in HBM (msv_cuda_db_create, untimed).  N = 1: a step
               is msv_cuda_db_score_device (ONE kernel launch).  N > 1: the ONE 1M-sequence database is cut into N
               contiguous slices of equal cell count (msv_host_partition_by_cells) -- `scaling: strong` -- and a step is
               every rank's scan of its slice with the score gather FUSED into the kernel (stores into every rank's copy of
               the job's score array over NVLink peer memory) + one device-side barrier.  CUDA events on the launching
               stream, barrier + synchronize on both sides, max over ranks.
  e2e        : the reference-facing call with HOST buffers.  N = 1: msv_cuda_score_batch, i.e. what
               MSV_HMM::parallel_run_on_sequences executes: H2D of residues + offsets from pinned memory, validation,
               longest-first bucketing, scan, D2H of the scores.  N > 1: every rank does the same for its slice
               (msv_cuda_score_batch_gather), and the step ends when rank 0 holds the WHOLE job's scores in one host buffer.
  roofline   : this path is bound by the fp32 ALU (3 lane-ops per cell: 1 add + 2 max), not by HBM
               or tensor cores; `achieved` is lane-op throughput of the scan kernel, `peak` = 148 SMs x 128 lanes x
               max SM clock.  The HBM side (1 residue byte per LENG cells) is reported next to it against the measured
               copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline / --impl reference : the reference's own MSV_HMM::run_on_sequence compiled from its unmodified sources
               (oracle/_ref) -- or the C port oracle/ when that library is absent -- on all host threads, on a bounded
               sample of the same database.  That arm loads nothing of the product (its database comes from
               oracle/synthetic_db.cpp, the same bytes).
  side keys  : weak_scaling (N > 1: every rank scans its own 1M-sequence database), nccl_gather (plain scan + NCCL
               all-gather), single_process_multi_gpu (msv_cuda_multi_score_batch from rank 0), other_configs (config 2
               latency, config 3 model sweep, config 5, hit-rate sweep, FASTA text -> scores).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

MODEL_FILE = "1400.hmm"
SEED = 20261018
LANE_OPS_PER_CELL = 3  # 1 FADD + 2 FMNMX (reference MSV_HMM.cpp:103-104)
SETTLE_STEPS = 7       # N > 1: uncounted steps before the counted warm-up (communicator / symmetric-memory set-up settles)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--sequences", type=int, default=1_000_000, help="sequences of the job's database (weak: per GPU)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="strong",
                    help="N > 1: strong = ONE database cut by cell count (config 4 as written); weak = one database per GPU")
    ap.add_argument("--model", default=MODEL_FILE)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the bounded baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peer-gather", action="store_true", help="N > 1: gather the scores with NCCL instead of the fused peer-memory stores")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the side measurements (configs 2, 3, 5, hit-rate sweep, FASTA)")
    ap.add_argument("--no-side-keys", action="store_true", help="N > 1: skip weak_scaling / nccl_gather / single_process_multi_gpu")
    ap.add_argument("--single-process-child", default=None, help=argparse.SUPPRESS)  # internal: see single_process_before_ranks
    return ap.parse_args()


def measured_peaks() -> dict:
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def workload_config(args, leng: int) -> dict:
    """Identical for both arms (the driver compares the dicts)."""
    per = "GPU" if (args.scaling == "weak" and args.gpus > 1) else "job"
    return {"workload": f"config4: {args.model} (LENG {leng}) x {args.sequences} synthetic Swiss-Prot-like sequences "
                        f"per {per}, mt19937_64 seed {SEED}" + ("(+rank)" if per == "GPU" else ""),
            "model": args.model, "sequences": args.sequences, "scaling": args.scaling, "gpus": args.gpus,
            "l2": "inputs larger than L2 (≈347 MB of residues per scan of the database)"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.recording = False
        self.last = None  # the most recent sample, inside the timed region or not (a very short region may see none)
        self._stop_flag = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self) -> None:
        if self.nv is None:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.last = mhz
                if self.recording:  # only what falls into the timed region counts
                    self.samples.append(mhz)
                    for name, bit in names.items():
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self) -> dict:
        self._stop_flag.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else (float(self.last) if self.last is not None else None)
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU legs (the only place bench.py touches oracle/) ------------------------------------------------------------
def oracle_module():
    tests = os.path.join(REPO, "tests")
    if tests not in sys.path:
        sys.path.insert(0, tests)
    import oracle_lib

    return oracle_lib


def cpu_reference_scorer(model_path: str):
    """Returns (kind, LENG, fn(codes, offsets, threads) -> scores): the reference's compiled code, else the C port."""
    ol = oracle_module()
    if ol.RefLib.available():
        ref_model = ol.RefLib().model(model_path)
        return "reference", ref_model.model_length - 1, lambda codes, offsets, threads: ref_model.run_batch(codes, offsets, threads)
    oracle = ol.Oracle()
    parsed = oracle.load_hmm(model_path)
    table, tr3 = oracle.prepare(parsed["match_emissions"])
    return "port", parsed["model_length"] - 1, lambda codes, offsets, threads: oracle.score_batch(table, tr3, codes, offsets, threads)


def bounded_sample(codes, offsets, leng: int, gcups_guess: float, seconds: float):
    """First sequences of the database worth about `seconds` of CPU time."""
    want_cells = gcups_guess * 1e9 * seconds
    n = int(np.searchsorted(offsets.astype(np.float64) * leng, want_cells))
    n = max(64, min(n, len(offsets) - 1))
    return codes[: int(offsets[n])], offsets[: n + 1].copy(), n


def time_cpu(scorer, codes, offsets, leng: int, threads: int):
    t0 = time.perf_counter()
    scores = scorer(codes, offsets, threads)
    dt = time.perf_counter() - t0
    return leng * float(offsets[-1]) / dt / 1e9, dt, scores


def run_reference_arm(args) -> None:
    """--impl reference: the reference's CPU implementation of the path on all host threads, same metric/config.
    Loads oracle/_ref (or the C port) and the checker-side workload generator only -- nothing of the product."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    model_path = os.path.join(REPO, "fixtures", "profile_HMMs", args.model)
    cores = os.cpu_count() or 1
    kind, leng, scorer = cpu_reference_scorer(model_path)
    all_codes, all_offsets = oracle_module().synthetic_database("swissprot_like", args.sequences, SEED)
    # size one step to ~cpu_seconds/(steps+warmup) so that the whole arm stays within a few minutes
    per_step = max(1.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    codes, offsets, n = bounded_sample(all_codes, all_offsets, leng, 0.13 * cores, per_step)
    for _ in range(args.warmup):
        time_cpu(scorer, codes, offsets, leng, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scorer(codes, offsets, cores)
    dt = time.perf_counter() - t0
    cells = leng * float(offsets[-1]) * args.steps
    value = cells / dt / 1e9
    sample = (f"per step: the first {n} sequences ({int(offsets[-1])} residues, {leng * float(offsets[-1]) / 1e9:.2f} Gcells) of the "
              f"{args.sequences}-sequence seed-{SEED} database of `config` (GCUPS is a rate; the GPU arm scans all of it)")
    print(json.dumps({
        "impl": "reference", "metric": "MSV GCUPS at M=1400", "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, leng),
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---- side measurements on one GPU (outside the headline's timed region) -------------------------------------------------
def event_ms(torch, stream, fn, steps: int, warm: int = 2) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(steps):
        fn()
    t1.record(stream)
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps


def load_model(msv, _cabi, name: str, device: int):
    prof = msv.Profile_HMM(os.path.join(REPO, "fixtures", "profile_HMMs", name))
    return prof, msv.Model(_cabi.emission_table(prof.match_emissions), *_cabi.model_transitions(prof.model_length), device=device)


def other_configs(torch, msv, _cabi, device: int) -> dict:
    """BASELINE.json configs 2, 3 and 5 plus the speculation hit-rate sweep and the FASTA-text-to-scores path, on the same
    GPU, device-resident scans timed with CUDA events unless stated otherwise."""
    stream = torch.cuda.current_stream()
    sm_max = 1965.0
    out: dict = {"unit": "GCUPS, device-resident unless stated"}

    def gcups(model, db, n, leng, residues, steps=3):
        scores = torch.empty(n, dtype=torch.float32, device="cuda")
        ms = event_ms(torch, stream, lambda: db.score_device(model, scores, stream.cuda_stream), steps)
        return leng * float(residues) / ms / 1e6

    # ---- config 3: every fixture model x 100k sequences ----
    sweep_db = msv.Packed_sequences.synthetic_swissprot_like(100_000, 1400)
    resident = msv.Database(sweep_db.residues, sweep_db.offsets, device=device)
    names = sorted((f for f in os.listdir(os.path.join(REPO, "fixtures", "profile_HMMs")) if f.endswith(".hmm")),
                   key=lambda s: int(s.split(".")[0]))
    config3, config3_frac = {}, {}
    roof = 148 * 128 * sm_max * 1e6 / LANE_OPS_PER_CELL / 1e9
    for name in names:
        prof, model = load_model(msv, _cabi, name, device)
        g = gcups(model, resident, len(sweep_db), prof.model_length - 1, sweep_db.total_residues)
        config3[name], config3_frac[name] = round(g, 1), round(g / roof, 3)
        model.close()
    out["config3_model_sweep_100k_sequences"] = config3
    out["config3_frac_of_fp32_alu_roofline"] = config3_frac

    # ---- config 5: 2405.hmm x 2048 titin-like sequences ----
    long_db = msv.Packed_sequences.synthetic_long_uniform(2048, 2405, 10_000, 35_000)
    long_resident = msv.Database(long_db.residues, long_db.offsets, device=device)
    prof, model = load_model(msv, _cabi, "2405.hmm", device)
    out["config5_2405hmm_2048_long_sequences"] = round(gcups(model, long_resident, len(long_db), prof.model_length - 1, long_db.total_residues), 1)
    model.close()
    long_resident.close()

    # ---- config 2 + hit-rate sweep + FASTA (each guarded: a failure is reported, not fatal to the headline) ----
    for key, fn in (("config2_benchmark_MSV_1400", config2_latency), ("hit_rate_sweep", hit_rate_sweep),
                    ("e2e_from_fasta", fasta_to_scores)):
        try:
            out[key] = fn(torch, msv, _cabi, device)
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
    out["viterbi"] = viterbi_side_line(torch, msv, _cabi, resident, sweep_db, device)
    return out


def config2_latency(torch, msv, _cabi, device: int) -> dict:
    """BASELINE.json config 2, the benchmark_MSV_1400 workload (reference benchmark_MSV_1400.cpp:5-16): 1400.hmm x
    random_FASTA.fsa, ONE sequence per call through MSV_HMM::parallel_run_on_sequence, wall clock per call (host string in,
    float out -- encode + H2D + scan + D2H inside), and the three scores' bit patterns against tests/golden."""
    hmm = msv.MSV_HMM(msv.Profile_HMM(os.path.join(REPO, "fixtures", "profile_HMMs", "1400.hmm")), device=device)
    fasta = msv.FASTA_protein_sequences(os.path.join(REPO, "fixtures", "FASTA_files", "random_FASTA.fsa"))
    for seq in fasta.sequences:  # warm-up (creates the device model)
        hmm.parallel_run_on_sequence(seq)
    per_call, got = [], []
    for _ in range(20):
        for seq in fasta.sequences:
            t0 = time.perf_counter()
            score = hmm.parallel_run_on_sequence(seq)
            per_call.append(time.perf_counter() - t0)
        got = [format(int(np.float32(hmm.parallel_run_on_sequence(s)).view(np.uint32)), "08x") for s in fasta.sequences]
    want = None
    try:
        with open(os.path.join(REPO, "tests", "golden", "msv_scores.json")) as f:
            golden = json.load(f)
        want = golden["scores"]["1400.hmm"]["random"]
    except Exception:  # noqa: BLE001 -- layout differences are reported as "unchecked", not hidden
        pass
    cells = 1400 * 3500
    us = float(np.median(per_call)) * 1e6
    return {"us_per_call_median": round(us, 1), "us_per_call_min": round(float(np.min(per_call)) * 1e6, 1),
            "gcups_per_call": round(cells / us / 1e3, 1), "calls": len(per_call), "score_bits": got,
            "golden_bits": want, "bit_exact_vs_golden": (got == want) if want is not None else None,
            "what": "MSV_HMM::parallel_run_on_sequence, 3 sequences x 3500 residues, wall clock per call"}


def plant_hits(codes, offsets, consensus, fraction: float, seed: int):
    """A copy of the database in which `fraction` of the sequences carry a planted consensus segment (a real hit: J
    overtakes N there, so the speculative rows fail and the kernel has to recover)."""
    rng = np.random.default_rng(seed)
    codes = codes.copy()
    n = len(offsets) - 1
    chosen = rng.choice(n, size=int(round(fraction * n)), replace=False) if fraction > 0 else np.zeros(0, np.int64)
    lens = np.diff(offsets.astype(np.int64))
    for q in chosen:
        L = int(lens[q])
        seg = min(L, 120)
        if seg < 40:
            continue
        at = int(offsets[q]) + int(rng.integers(0, L - seg + 1))
        start = int(rng.integers(0, max(1, len(consensus) - seg)))
        codes[at: at + seg] = consensus[start: start + seg]
    return codes


def hit_rate_sweep(torch, msv, _cabi, device: int) -> dict:
    """How much a family-rich database costs: 1400.hmm x 200k sequences with a consensus segment planted in 0 / 1 / 10 / 50 %
    of them.  Every score of the 10 % database is also compared with the end-to-end path's."""
    stream = torch.cuda.current_stream()
    prof, model = load_model(msv, _cabi, "1400.hmm", device)
    leng = prof.model_length - 1
    consensus = np.argmax(_cabi.emission_table(prof.match_emissions)[:, 1:], axis=0).astype(np.uint8)  # best log-odds per column
    base = msv.Packed_sequences.synthetic_swissprot_like(200_000, 14)
    offsets = np.ascontiguousarray(base.offsets)
    out = {"workload": "1400.hmm x 200000 sequences, planted 120-residue consensus segments", "gcups": {}}
    for fraction in (0.0, 0.01, 0.10, 0.50):
        codes = plant_hits(base.residues, offsets, consensus, fraction, 99)
        db = msv.Database(codes, offsets, device=device)
        scores = torch.empty(len(base), dtype=torch.float32, device="cuda")
        ms = event_ms(torch, stream, lambda: db.score_device(model, scores, stream.cuda_stream), 3)
        out["gcups"][f"{fraction:.2f}"] = round(leng * float(offsets[-1]) / ms / 1e6, 1)
        db.close()
    g = out["gcups"]
    out["loss_at_10pct_hits"] = round(1.0 - g["0.10"] / g["0.00"], 4)
    out["loss_at_50pct_hits"] = round(1.0 - g["0.50"] / g["0.00"], 4)
    model.close()
    return out


def write_fasta(path: str, codes: np.ndarray, offsets: np.ndarray, width: int = 60) -> int:
    """The packed database as FASTA text (one '>' header per record, `width` residues per line).  Returns the bytes."""
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)[codes]
    n = len(offsets) - 1
    lens = np.diff(offsets.astype(np.int64))
    headers = [f">sp|S{q:07d}|SYN_{q} synthetic protein {q}\n".encode() for q in range(n)]
    hlen = np.array([len(h) for h in headers], np.int64)
    lines = (lens + width - 1) // width
    rec = hlen + lens + lines
    out = np.empty(int(rec.sum()), np.uint8)
    starts = np.concatenate([[0], np.cumsum(rec)[:-1]])
    # residues with a newline after every `width` (and at the end of the record)
    for q in range(n):
        p = int(starts[q])
        h = headers[q]
        out[p: p + len(h)] = np.frombuffer(h, np.uint8)
        p += len(h)
        L = int(lens[q])
        body = letters[int(offsets[q]): int(offsets[q]) + L]
        full = L // width
        if full:
            block = np.empty((full, width + 1), np.uint8)
            block[:, :width] = body[: full * width].reshape(full, width)
            block[:, width] = 10
            out[p: p + full * (width + 1)] = block.ravel()
            p += full * (width + 1)
        tail = L - full * width
        if tail:
            out[p: p + tail] = body[full * width:]
            out[p + tail] = 10
    with open(path, "wb") as f:
        f.write(out.tobytes())
    return int(out.size)


def fasta_to_scores(torch, msv, _cabi, device: int) -> dict:
    """FASTA text file (in the page cache) -> scores on the host, 1400.hmm x a 200k-sequence database: the stage in front of
    `e2e` (reference data_readers/FASTA_protein_sequences.cpp:9-44 + the scan).  Two routes: the host reader
    (Packed_sequences::from_fasta_file, threads) followed by msv_cuda_score_batch, and msv_cuda_score_fasta, which uploads the
    raw text and classifies / encodes / cuts it on the GPU."""
    import tempfile

    prof, model = load_model(msv, _cabi, "1400.hmm", device)
    leng = prof.model_length - 1
    base = msv.Packed_sequences.synthetic_swissprot_like(200_000, 15)
    codes, offsets = np.ascontiguousarray(base.residues), np.ascontiguousarray(base.offsets)
    path = os.path.join(tempfile.gettempdir(), f"msv_bench_{os.getpid()}.fasta")
    text_bytes = write_fasta(path, codes, offsets)
    cells = leng * float(offsets[-1])
    out = {"workload": f"1400.hmm x 200000 sequences as FASTA text ({text_bytes} bytes, page cache)"}
    try:
        want = model.score_batch(codes, offsets)
        # packed host buffers (what `e2e` measures), same database, for the ratio
        for _ in range(2):
            model.score_batch(codes, offsets)
        t0 = time.perf_counter()
        for _ in range(3):
            model.score_batch(codes, offsets)
        packed_s = (time.perf_counter() - t0) / 3
        out["e2e_packed_gcups"] = round(cells / packed_s / 1e9, 1)
        # route 1: host reader + score_batch
        for _ in range(2):
            p = msv.Packed_sequences.from_fasta_file(path)
        t0 = time.perf_counter()
        for _ in range(3):
            p = msv.Packed_sequences.from_fasta_file(path)
            got = model.score_batch(p.residues, p.offsets)
        host_s = (time.perf_counter() - t0) / 3
        out["host_reader"] = {"gcups": round(cells / host_s / 1e9, 1), "ms": round(host_s * 1e3, 2),
                              "text_gb_per_s": round(text_bytes / host_s / 1e9, 2),
                              "same_bits": bool((got.view(np.uint32) == want.view(np.uint32)).all())}
        # route 2: the text goes to the GPU as it is
        if hasattr(model, "score_fasta"):
            with open(path, "rb") as f:
                text = np.frombuffer(f.read(), np.uint8)
            for _ in range(2):
                got, rejected = model.score_fasta_file(path)
            t0 = time.perf_counter()
            for _ in range(3):
                got, rejected = model.score_fasta_file(path)
            dev_s = (time.perf_counter() - t0) / 3
            out["device_parser"] = {"gcups": round(cells / dev_s / 1e9, 1), "ms": round(dev_s * 1e3, 2),
                                    "text_gb_per_s": round(text_bytes / dev_s / 1e9, 2), "rejected": int(rejected),
                                    "same_bits": bool(len(got) == len(want) and (got.view(np.uint32) == want.view(np.uint32)).all()),
                                    "slowdown_vs_e2e_packed": round(dev_s / packed_s, 3)}
            del text
    finally:
        try:
            os.remove(path)
        except OSError:
            pass
        model.close()
    return out


def viterbi_side_line(torch, msv, _cabi, resident, database, device: int) -> dict:
    """SURVEY section 8(f) rank 4, the Plan-7 local Viterbi scan (match/insert/delete) on 1400.hmm x the config-3 database:
    same GCUPS definition as the headline (LENG x residues / time), a sample checked bit-for-bit against
    oracle/viterbi_oracle.c (the checker; it is never on the measured path) and that oracle's speed on the host cores."""
    ol = oracle_module()
    oracle = ol.Oracle()
    h = oracle.load_hmm(os.path.join(REPO, "fixtures", "profile_HMMs", "1400.hmm"))
    leng = h["model_length"] - 1
    model = msv.ViterbiModel(_cabi.emission_table(h["match_emissions"]), _cabi.viterbi_transitions(h["transitions"]),
                             *_cabi.model_transitions(h["model_length"]), device=device)
    stream = torch.cuda.current_stream()
    scores = torch.empty(len(database), dtype=torch.float32, device="cuda")
    ms = event_ms(torch, stream, lambda: resident.viterbi_device(model, scores, stream.cuda_stream), 3)
    rng = np.random.default_rng(4)
    sample = rng.choice(len(database), size=256, replace=False)
    off = database.offsets
    sc, so = ol.pack([database.residues[int(off[q]):int(off[q + 1])] for q in sample])
    table, tr3 = oracle.prepare(h["match_emissions"])
    cores = os.cpu_count() or 1
    t_cpu = time.perf_counter()
    want = oracle.viterbi_score_batch(table, oracle.viterbi_prepare(h["transitions"]), tr3, sc, so, threads=cores)
    t_cpu = time.perf_counter() - t_cpu
    got = scores.cpu().numpy()[sample]
    gcups = leng * float(database.total_residues) / ms / 1e6
    return {"workload": "1400.hmm x 100000 synthetic sequences (config-3 database), device-resident", "gcups": round(gcups, 1),
            "ms": round(ms, 3), "cells_per_clk_per_sm": round(gcups * 1e9 / 148 / 1.965e9, 2), "geometry": model.geometry,
            "frac_of_fp32_alu_roofline_13_ops_per_cell": round(gcups * 1e9 * 13 / (148 * 128 * 1.965e9), 3),
            "mismatches_vs_oracle": int((got.view(np.uint32) != want.view(np.uint32)).sum()), "compared": int(sample.size),
            "oracle_gcups": round(leng * float(so[-1]) / t_cpu / 1e9, 3), "oracle_threads": cores}


# ---- our arm ----------------------------------------------------------------------------------------------------------
def single_process_child(args) -> None:
    """One process, all GPUs, through the C ABI (msv_cuda_multi_score_batch): the whole job's database in pinned host memory
    in, the whole job's scores in one host buffer out.  No torch, no NCCL process group; prints one JSON object."""
    import hmm_fasta_viterbi_b200 as msv
    from hmm_fasta_viterbi_b200 import _cabi

    ngpu = min(args.gpus, _cabi.device_count())
    profile = msv.Profile_HMM(os.path.join(REPO, "fixtures", "profile_HMMs", args.model))
    packed = msv.Packed_sequences.synthetic_swissprot_like(args.sequences, SEED)
    codes, offsets = np.ascontiguousarray(packed.residues), np.ascontiguousarray(packed.offsets)
    out = np.empty(len(offsets) - 1, np.float32)
    for a in (codes, offsets, out):  # page-locked, as the e2e legs of the ranks use
        _cabi.check(_cabi.lib.msv_cuda_host_register(a.ctypes.data, a.nbytes))
    cells = float(offsets[-1]) * (profile.model_length - 1)
    models = [msv.Model(_cabi.emission_table(profile.match_emissions), *_cabi.model_transitions(profile.model_length), device=g)
              for g in range(ngpu)]
    multi = _cabi.MultiGpu(models)
    report, kept = {"gpus": ngpu}, None
    for name, mode in (("host", _cabi.GATHER_HOST), ("peer", _cabi.GATHER_PEER), ("nccl", _cabi.GATHER_NCCL)):
        try:
            for _ in range(3):
                multi.score_batch(codes, offsets, out, gather=mode)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                multi.score_batch(codes, offsets, out, gather=mode)
            dt = (time.perf_counter() - t0) / args.steps
            report[name] = {"e2e_gcups": round(cells / dt / 1e9, 1), "ms_per_step": round(dt * 1e3, 3)}
            if kept is None:
                kept = out.copy()
            else:
                report[name]["same_bits_as_first_mode"] = bool((kept.view(np.uint32) == out.view(np.uint32)).all())
        except Exception as e:  # noqa: BLE001
            report[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if kept is not None:
        np.save(args.single_process_child, kept)
        report["scores_file"] = args.single_process_child
    report["api"] = ("msv_cuda_multi_score_batch: one process, one host thread per GPU, pinned host database in, host scores out; "
                     "measured in a process of its own before the ranks created their CUDA contexts (a GPU shared by two "
                     "processes time-slices, which is not what a user of the single-process API has)")
    print(json.dumps(report))


def single_process_before_ranks(args, rank: int, world: int):
    """N > 1: rank 0 runs single_process_child in a child process while the other ranks wait on the rendezvous store, before
    any rank has touched CUDA.  (Measured from rank 0 itself after the main loop, the same call took 66 ms instead of 26 ms at
    N = 2 -- profiles/r02/bench_b_n2_strong.json vs multi_probe_v1.json: every other GPU then also holds another rank's
    context.)"""
    import subprocess
    import tempfile
    from datetime import timedelta

    import torch.distributed as dist

    if os.environ.get("TORCHELASTIC_USE_AGENT_STORE") != "True":  # not under torchrun: no store to wait on before the process group
        return None
    store = dist.TCPStore(os.environ["MASTER_ADDR"], int(os.environ["MASTER_PORT"]), world, False, timedelta(seconds=900))
    result = None
    if rank == 0:
        scores_file = os.path.join(tempfile.gettempdir(), f"msv_single_process_{os.getpid()}.npy")
        try:
            done = subprocess.run([sys.executable, os.path.abspath(__file__), "--gpus", str(world), "--steps", "5", "--sequences", str(args.sequences),
                                   "--model", args.model, "--single-process-child", scores_file],
                                  capture_output=True, text=True, timeout=600,
                                  env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
            result = json.loads(done.stdout.strip().splitlines()[-1]) if done.returncode == 0 else {
                "error": f"child exited {done.returncode}: {done.stderr[-300:]}"}
        except Exception as e:  # noqa: BLE001
            result = {"error": f"{type(e).__name__}: {e}"[:300]}
        store.set("msv_single_process_done", "1")
    else:
        store.wait(["msv_single_process_done"])
    return result


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.single_process_child:
        single_process_child(args)
        return
    single_process = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.scaling == "strong" and not args.no_side_keys:
        single_process = single_process_before_ranks(args, int(os.environ.get("RANK", "0")), int(os.environ["WORLD_SIZE"]))

    import torch
    import torch.distributed as dist

    import hmm_fasta_viterbi_b200 as msv
    from hmm_fasta_viterbi_b200 import _cabi, sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or _cabi.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the MSV scan has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.gpus = world
    strong = world > 1 and args.scaling == "strong"

    model_path = os.path.join(REPO, "fixtures", "profile_HMMs", args.model)
    profile = msv.Profile_HMM(model_path)
    leng = profile.model_length - 1
    model = msv.Model(_cabi.emission_table(profile.match_emissions), *_cabi.model_transitions(profile.model_length), device=local)

    # ---- the job's database and this rank's slice of it ----
    # strong (config 4 as written): every rank builds the same 1M-sequence database and keeps its cell-balanced slice;
    # weak: every rank has its own database (seed + rank) -- N GPUs scan an N-million-sequence database
    packed = msv.Packed_sequences.synthetic_swissprot_like(args.sequences, SEED + (0 if strong else rank))
    all_codes, all_offsets = packed.residues, packed.offsets
    if strong:
        bounds = sharded.shard_bounds(all_offsets, world)
        codes, offsets, first_index, _ = sharded.local_slice(all_codes, all_offsets, rank, world)
        n_total = len(all_offsets) - 1
    else:
        bounds = None
        codes, offsets, first_index = all_codes, all_offsets, 0
        n_total = (len(all_offsets) - 1) * world
    n_local = len(offsets) - 1
    cells_local = leng * float(offsets[-1])

    # pinned host copies for the end-to-end leg
    pin_codes = torch.from_numpy(np.ascontiguousarray(codes)).pin_memory()
    pin_offsets = torch.from_numpy(np.ascontiguousarray(offsets).view(np.int64)).pin_memory()
    pin_scores = torch.empty(max(n_local, 1), dtype=torch.float32).pin_memory()

    db = msv.Database(pin_codes, pin_offsets, device=local)  # resident in HBM before the timed region
    n_max = n_local
    if world > 1:
        t = torch.tensor([n_local], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_max = int(t.item())
    if not strong and world > 1:
        first_index = rank * n_max
    scores = torch.full((max(n_max, 1),), float("nan"), dtype=torch.float32, device="cuda")
    gathered = torch.empty((world * max(n_max, 1),), dtype=torch.float32, device="cuda") if world > 1 else None
    stream = torch.cuda.current_stream()

    def nccl_step() -> None:
        db.score_device(model, scores, stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, scores)

    def fence() -> None:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # N > 1: the product path FUSES the gather into the scan -- every rank's kernel stores its scores straight into all ranks'
    # copies of the job's score array over NVLink peer memory (hmm_fasta_viterbi_b200.sharded.FusedGather), followed by one
    # device-side barrier; no collective.  The plain scan + NCCL all-gather is measured next to it (`nccl_gather`).
    step, gather_kind, peer_error, fused = nccl_step, "nccl all_gather_into_tensor", None, None
    if world > 1 and not args.no_peer_gather:
        try:
            if strong:
                fused = sharded.FusedGather(n_max, torch.device("cuda", local), total=n_total, first_index=first_index)
            else:
                fused = sharded.FusedGather(n_max, torch.device("cuda", local))

            def peer_step() -> None:
                fused.scan(model, db, stream.cuda_stream)

            peer_step()
            fence()
            ok = torch.tensor([1.0], device="cuda")
        except Exception as e:  # noqa: BLE001 -- needs P2P over NVLink; fall back to the NCCL gather and say so
            peer_error = f"{type(e).__name__}: {e}"[:300]
            ok = torch.tensor([0.0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank must take the same path
        if float(ok.item()) == 1.0:
            step, gather_kind = peer_step, "fused into the scan kernel: stores to every rank's copy over NVLink peer memory"
        else:
            fused = None

    # --warmup is honoured exactly; N > 1 runs SETTLE_STEPS uncounted steps first (the first ~0.3 s after communicator /
    # symmetric-memory set-up run about 1 % slow), and at least 3 steps precede the timed region in any case
    settle = (SETTLE_STEPS if world > 1 else 0) + max(0, 3 - args.warmup)
    # (NVML is initialised and the polling thread started BEFORE the fence: nvmlInit takes tens of milliseconds when eight
    # processes do it at once, and done between the fence and the first step it let the ranks enter the timed region that far
    # apart -- every rank then waited for the slowest one at the first barrier, inside its timed region: 6.5 and 7.5 ms per
    # step at 8 GPUs for a step that takes 6.08 ms, profiles/r02/bench_l_n8_strong.json, bench_n_n8_strong.json)
    sampler = ClockSampler(local)
    if os.environ.get("MSV_BENCH_NO_CLOCK_SAMPLER"):  # diagnostic only
        sampler.nv = None
    sampler.start()
    for _ in range(settle + args.warmup):
        step()
    fence()

    # ---- value: device-resident scan (+ gather), CUDA events ----
    _cabi.launch_count(reset=True)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.recording = True
    start.record(stream)
    for _ in range(args.steps):
        step()
    stop.record(stream)
    torch.cuda.synchronize()
    sampler.recording = False
    fence()
    launches = _cabi.launch_count()
    clocks = sampler.stop()
    ms_total = start.elapsed_time(stop)

    # ---- the scan kernel alone (for the roofline), same launches, no gather ----
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(args.steps):
        db.score_device(model, scores, stream.cuda_stream)
    k1.record(stream)
    fence()
    kernel_ms = k0.elapsed_time(k1) / args.steps

    # ---- e2e: host buffers in, the whole job's scores in ONE host buffer out ----
    job_scores = torch.empty(max(n_total, 1), dtype=torch.float32).pin_memory() if (world > 1 and rank == 0) else None
    if world == 1:
        def e2e_step() -> None:
            model.score_batch(pin_codes, pin_offsets, pin_scores)
        e2e_api = "msv_cuda_score_batch (pinned host buffers in, host scores out)"
    elif fused is not None:
        def e2e_step() -> None:
            fused.scan_host(model, pin_codes, pin_offsets)  # H2D + bucketing + scan with the fused gather, then the barrier
            if rank == 0:
                job_scores[: fused.total].copy_(fused.scores, non_blocking=True)
            torch.cuda.synchronize()
        e2e_api = ("every rank: msv_cuda_score_batch_gather (pinned host slice in, scores stored into all ranks' copies) + "
                   "device barrier; rank 0: D2H of the whole job's score array")
    else:
        def e2e_step() -> None:
            model.score_batch(pin_codes, pin_offsets, pin_scores)
            scores[:n_local].copy_(pin_scores[:n_local], non_blocking=True)
            dist.all_gather_into_tensor(gathered, scores)
            if rank == 0:
                job_scores[: gathered.numel()].copy_(gathered[: job_scores.numel()], non_blocking=True)
            torch.cuda.synchronize()
        e2e_api = "every rank: msv_cuda_score_batch; NCCL all-gather; rank 0: D2H of the gathered array"
    for _ in range(2):
        e2e_step()
    fence()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    fence()

    # ---- side keys for N > 1 ----
    nccl_gather = weak = gather_forms = None
    if world > 1 and not args.no_side_keys:
        side_steps = max(3, min(args.steps, 10))
        if fused is not None:  # plain scan + NCCL all-gather of padded slices; must be the same bits as the fused gather
            for _ in range(3):
                nccl_step()
            fence()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0.record(stream)
            for _ in range(side_steps):
                nccl_step()
            n1.record(stream)
            fence()
            step()
            fence()
            if strong:
                same = all(torch.equal(fused.scores[int(bounds[r]): int(bounds[r + 1])].view(torch.int32),
                                       gathered[r * n_max: r * n_max + int(bounds[r + 1] - bounds[r])].view(torch.int32)) for r in range(world))
            else:
                same = bool(torch.equal(fused.scores.view(torch.int32), gathered.view(torch.int32)))
            flags = torch.tensor([n0.elapsed_time(n1) / side_steps, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(flags, op=dist.ReduceOp.MAX)
            nccl_gather = {"ms_per_step": float(flags[0].item()), "same_bits_as_fused_gather_on_every_rank": float(flags[1].item()) == 0.0}
        if fused is not None:
            # the two forms of the gather (launch_scan in msv_cuda.cu) and their pieces, device-timed, max over ranks; every
            # measurement is bracketed by a fence so that all ranks start together
            def timed_piece(fn) -> float:
                for _ in range(3):
                    fn()
                fence()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(side_steps):
                    fn()
                b.record(stream)
                fence()
                v = torch.tensor([a.elapsed_time(b) / side_steps], dtype=torch.float64, device="cuda")
                dist.all_reduce(v, op=dist.ReduceOp.MAX)
                return round(float(v.item()), 4)

            gather_forms = {"unit": "ms per step, max over ranks", "default": os.environ.get("MSV_CUDA_GATHER", "library default")}
            previous = os.environ.get("MSV_CUDA_GATHER")
            for form in ("stores", "push"):
                os.environ["MSV_CUDA_GATHER"] = form
                gather_forms[form] = {
                    "scan_and_barrier": timed_piece(lambda: fused.scan(model, db, stream.cuda_stream)),
                    "scan_without_barrier": timed_piece(lambda: db.score_gather(model, fused._copies, fused.first_index, stream.cuda_stream)),
                }
                fused.scores.fill_(float("nan"))
                fence()
                fused.scan(model, db, stream.cuda_stream)
                fence()
                mine_ok = torch.equal(fused.scores[int(bounds[rank]) if strong else rank * fused.slot:][:n_local].view(torch.int32),
                                      scores[:n_local].view(torch.int32))
                whole_ok = not bool(torch.isnan(fused.scores[:n_total] if strong else fused.scores).any().item()) if strong else True
                ok = torch.tensor([1.0 if (mine_ok and whole_ok) else 0.0], device="cuda")
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                gather_forms[form]["every_rank_holds_the_whole_job"] = float(ok.item()) == 1.0
            if previous is None:
                os.environ.pop("MSV_CUDA_GATHER", None)
            else:
                os.environ["MSV_CUDA_GATHER"] = previous
            gather_forms["barrier_alone"] = timed_piece(lambda: fused._handle.barrier(channel=0))
            gather_forms["plain_scan"] = timed_piece(lambda: db.score_device(model, scores, stream.cuda_stream))
        if strong:  # weak scaling next to it: every rank scans a whole 1M-sequence database of its own
            own = msv.Packed_sequences.synthetic_swissprot_like(args.sequences, SEED + rank)
            own_db = msv.Database(own.residues, own.offsets, device=local)
            own_scores = torch.empty(len(own), dtype=torch.float32, device="cuda")
            ms = event_ms(torch, stream, lambda: own_db.score_device(model, own_scores, stream.cuda_stream), side_steps, warm=3)
            w = torch.tensor([ms, leng * float(own.total_residues)], dtype=torch.float64, device="cuda")
            wmax = w.clone()
            dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
            weak = {"value": float(w[1].item()) / (float(wmax[0].item()) / 1e3) / 1e9, "unit": "GCUPS", "ms_per_step": float(wmax[0].item()),
                    "what": f"every rank scans its own {args.sequences}-sequence database (seed + rank), device-resident, no gather"}
            own_db.close()
            del own, own_scores

    # ---- max over ranks (and every rank's own numbers, to see where a strong-scaling step loses time) ----
    agg = torch.tensor([ms_total, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
    mine = torch.tensor([kernel_ms, cells_local, float(n_local)], dtype=torch.float64, device="cuda")
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    cells = torch.tensor([cells_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    ms_total, e2e_s, kernel_ms_max = (float(v) for v in agg.tolist())
    cells_job = float(cells.item())

    # ---- parity (outside every timed region) ----
    result = scores[:n_local].cpu().numpy()
    parity = {}
    if world == 1:
        e2e_result = pin_scores[:n_local].numpy()
        parity["device_equals_e2e"] = bool((result.view(np.uint32) == e2e_result.view(np.uint32)).all())
    elif rank == 0:
        job = job_scores.numpy()
        if strong:
            parity["e2e_job_buffer_equals_own_device_scan"] = bool(
                (job[int(bounds[0]): int(bounds[1])].view(np.uint32) == result.view(np.uint32)).all())
            parity["e2e_job_buffer_has_no_gaps"] = bool(not np.isnan(job[:n_total]).any())

    # ---- single process, all GPUs, through the C ABI: measured before the ranks started (single_process_before_ranks) ----
    if single_process is not None and rank == 0 and strong and "scores_file" in single_process:
        try:
            theirs = np.load(single_process.pop("scores_file"))
            single_process["same_bits_as_multi_process_job_buffer"] = bool(
                theirs.size == n_total and (theirs.view(np.uint32) == job_scores.numpy()[:n_total].view(np.uint32)).all())
        except Exception as e:  # noqa: BLE001
            single_process["same_bits_as_multi_process_job_buffer"] = f"not compared: {type(e).__name__}: {e}"[:200]

    if rank == 0:
        value = cells_job * args.steps / (ms_total / 1e3) / 1e9
        e2e_value = cells_job * args.steps / e2e_s / 1e9
        peaks = measured_peaks()
        sm_max_mhz = clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        alu_peak = sms * 128 * sm_max_mhz * 1e6 / 1e12  # T lane-ops/s
        kernel_gcups = cells_local / (kernel_ms / 1e3) / 1e9
        achieved = kernel_gcups * LANE_OPS_PER_CELL / 1e3
        hbm_peak = peaks.get("hbm_gbs") or 6650.0
        hbm_bytes = float(offsets[-1]) + 8.0 * (n_local + 1) + 4.0 * n_local * 2  # residues + offsets + order + scores
        geometry = model.geometry
        # DRAM traffic per launch comes from the committed ncu --set full capture; it is printed only when that capture is
        # of this very kernel geometry on this very workload (it cannot be measured live without a profiler)
        traffic, traffic_source = None, None
        try:
            with open(os.path.join(REPO, "profiles", "roofline_traffic.json")) as f:
                captured = json.load(f)
            if (captured.get("workload") == f"{args.model} x {args.sequences} sequences" and world == 1
                    and captured.get("geometry") == [geometry["lanes_per_sequence"], geometry["columns_per_lane"],
                                                     geometry["tensor_columns_per_lane"], geometry["threads_per_cta"]]
                    and captured.get("kernel_source_sha1") == kernel_source_sha1()):
                traffic, traffic_source = captured["traffic_bytes_per_launch"], captured["source"]
        except OSError:
            pass
        roofline = {
            "bound": "fp32_alu", "kernel": "msv_scan_warp_kernel", "achieved": achieved, "peak": alu_peak, "unit": "Tlaneop/s",
            "frac": achieved / alu_peak, "traffic": traffic, "traffic_source": traffic_source, "algorithmic_hbm_bytes": hbm_bytes,
            "peak_source": f"{sms} SMs x 128 fp32 lanes x {sm_max_mhz:.0f} MHz (max SM clock); 3 lane-ops per cell",
            "kernel_ms": kernel_ms, "kernel_gcups": kernel_gcups, "gcups_at_alu_roofline": alu_peak * 1e3 / LANE_OPS_PER_CELL,
            "smem_ceiling_gcups": sms * 32 * sm_max_mhz * 1e6 / 1e9,
            # register-only add+max mix measured on this GPU type (tools/microbench.cu, profiles/r01/microbench_b200.json):
            # 38.5 cells/clk/SM -- the FMNMX pipe runs at half rate, so this is the practical ALU ceiling
            "measured_mix_peak_gcups": sms * 38.52 * sm_max_mhz * 1e6 / 1e9,
            "frac_of_measured_mix_peak": kernel_gcups / (sms * 38.52 * sm_max_mhz * 1e6 / 1e9),
            "hbm": {"achieved": hbm_bytes / (kernel_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_bytes / (kernel_ms / 1e3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6650"},
        }
        h2d = int(pin_codes.numel() + pin_offsets.numel() * 8)
        out = {
            "metric": "MSV GCUPS at M=1400", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, leng),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "GCUPS", "ms_per_step": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": h2d if world == 1 else int(float(all_offsets[-1]) + 8 * (n_total + world)) if strong else h2d * world,
                    "d2h_bytes_per_step": int(n_total * 4) if world > 1 else int(n_local * 4), "api": e2e_api},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "parity": parity,
            "kernel": {"geometry": geometry, "cells_per_step": cells_job, "settle_steps_before_warmup": settle},
        }
        if world > 1:
            out["gather"] = gather_kind
            out["per_rank"] = [{"rank": r, "kernel_ms": round(float(v[0]), 4), "cells": float(v[1]), "sequences": int(v[2])}
                               for r, v in enumerate(per_rank)]
            out["kernel_ms_max_over_ranks"] = kernel_ms_max
            if nccl_gather is not None:
                ms_nccl = nccl_gather.pop("ms_per_step")
                out["nccl_gather"] = {"value": cells_job / (ms_nccl / 1e3) / 1e9, "unit": "GCUPS", "ms_per_step": ms_nccl,
                                      "what": "plain scan + NCCL all_gather_into_tensor"} | nccl_gather
            if gather_forms is not None:
                out["gather_forms"] = gather_forms
            if weak is not None:
                out["weak_scaling"] = weak
            if single_process is not None:
                out["single_process_multi_gpu"] = single_process
            if peer_error is not None:
                out["peer_gather_unavailable"] = peer_error
        # the bounded CPU sample: timed at N = 1 (cpu_baseline); at N > 1 it only CHECKS a sample of the gathered job buffer
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            kind, _, scorer = cpu_reference_scorer(model_path)
            if world == 1:
                c1, o1, n1 = bounded_sample(codes, offsets, leng, 0.13, min(4.0, args.cpu_seconds / 3))
                g1, _, s1 = time_cpu(scorer, c1, o1, leng, 1)
                cN, oN, nN = bounded_sample(codes, offsets, leng, 0.13 * cores, args.cpu_seconds)
                gN, dtN, sN = time_cpu(scorer, cN, oN, leng, cores)
                out["cpu_baseline"] = {
                    "value": gN, "unit": "GCUPS", "cores": cores, "kind": kind,
                    "sample": f"first {nN} sequences ({int(oN[-1])} residues, {dtN:.1f} s) of this database; "
                              f"1 thread on the first {n1}: {g1:.3f} GCUPS",
                    "value_1thread": g1,
                }
                parity["oracle_checked"] = int(nN)
                parity["oracle_mismatches"] = int((result[:nN].view(np.uint32) != np.asarray(sN, np.float32).view(np.uint32)).sum())
            elif strong:
                rng = np.random.default_rng(8)
                pick = np.sort(rng.choice(n_total, size=min(n_total, 4096), replace=False))
                ol = oracle_module()
                sc, so = ol.pack([all_codes[int(all_offsets[q]): int(all_offsets[q + 1])] for q in pick])
                want = np.asarray(scorer(sc, so, cores), np.float32)
                got = job_scores.numpy()[pick]
                parity["gathered_job_buffer_vs_" + kind] = {"checked": int(pick.size), "from_every_shard": True,
                                                            "mismatches": int((got.view(np.uint32) != want.view(np.uint32)).sum())}
        if world == 1 and not args.no_other_configs:
            out["other_configs"] = other_configs(torch, msv, _cabi, local)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kernel_source_sha1() -> str:
    """Identity of the scan kernel's source: a committed ncu capture is only quoted while the kernel is unchanged."""
    import hashlib

    h = hashlib.sha1()
    for name in ("msv_kernels.cuh", "msv_device.cuh"):
        with open(os.path.join(REPO, "hmm_fasta_viterbi_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


if __name__ == "__main__":
    main()
