#!/usr/bin/env python
"""bench.py -- MSV GCUPS at M=1400 on a 1M-sequence synthetic Swiss-Prot-like database (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

A "step" is one scan of the whole database: LENG x sum(L) DP cells.  GCUPS = cells / seconds / 1e9.

  value      : device-resident path. The packed database already sits in HBM (msv_cuda_db_create, untimed); a step is
               msv_cuda_db_score_device (ONE kernel launch) and, for N > 1, the NCCL all-gather of the fp32 scores.
               Timed with CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
  e2e        : the reference-facing call with HOST buffers -- msv_cuda_score_batch, i.e. what
               MSV_HMM::parallel_run_on_sequences executes: H2D of residues + offsets from pinned memory, validation,
               longest-first bucketing, scan, D2H of the scores -- all inside the timed region.
  roofline   : this path is bound by the fp32 ALU (3 lane-ops per cell: 1 add + 2 max), not by HBM
               or tensor cores; `achieved` is lane-op throughput of the scan kernel, `peak` = 148 SMs x 128 lanes x
               max SM clock.  The HBM side (1 residue byte per LENG cells) is reported next to it against the measured
               copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline / --impl reference : the reference's own MSV_HMM::run_on_sequence compiled from its unmodified sources
               (oracle/_ref) -- or the C port oracle/ when that library is absent -- on all host threads, on a bounded
               sample of the same database.

Scaling is weak: every rank scans its own 1M-sequence database (seed + rank), i.e. N GPUs scan an N-million-sequence
database cut into contiguous per-GPU slices; `--scaling strong` shards ONE 1M-sequence database by cell count instead.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--sequences", type=int, default=1_000_000, help="sequences per GPU (weak) or in total (strong)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--model", default=MODEL_FILE)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the bounded baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peer-gather", action="store_true", help="N > 1: gather the scores with NCCL instead of the fused peer-memory stores")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short config-3 / config-5 side measurements")
    return ap.parse_args()


def measured_peaks() -> dict:
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self) -> dict:
        self._stop_flag.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU legs (the only place bench.py touches oracle/) ------------------------------------------------------------
def cpu_reference_scorer(model_path: str):
    """Returns (kind, fn(codes, offsets, threads) -> scores)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from oracle_lib import Oracle, RefLib

    if RefLib.available():
        ref_model = RefLib().model(model_path)
        return "reference", lambda codes, offsets, threads: ref_model.run_batch(codes, offsets, threads)
    oracle = Oracle()
    table, tr3 = oracle.prepare(oracle.load_hmm(model_path)["match_emissions"])
    return "port", lambda codes, offsets, threads: oracle.score_batch(table, tr3, codes, offsets, threads)


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
    """--impl reference: the reference's CPU implementation of the path on all host threads, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import hmm_fasta_viterbi_b200 as msv

    model_path = os.path.join(REPO, "fixtures", "profile_HMMs", args.model)
    leng = msv.Profile_HMM(model_path).model_length - 1
    cores = os.cpu_count() or 1
    kind, scorer = cpu_reference_scorer(model_path)
    db = msv.Packed_sequences.synthetic_swissprot_like(min(args.sequences, 200_000), SEED)
    # size one step to ~cpu_seconds/(steps+warmup) so that the whole arm stays within a few minutes
    per_step = max(1.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    codes, offsets, n = bounded_sample(db.residues, db.offsets, leng, 0.13 * cores, per_step)
    for _ in range(args.warmup):
        time_cpu(scorer, codes, offsets, leng, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scorer(codes, offsets, cores)
    dt = time.perf_counter() - t0
    cells = leng * float(offsets[-1]) * args.steps
    value = cells / dt / 1e9
    sample = f"first {n} sequences ({int(offsets[-1])} residues) of the seed-{SEED} database per step"
    print(json.dumps({
        "impl": "reference", "metric": "MSV GCUPS at M=1400", "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, leng),
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, leng: int) -> dict:
    return {"workload": f"config4: {args.model} (LENG {leng}) x {args.sequences} synthetic Swiss-Prot-like sequences "
                        f"per {'GPU' if args.scaling == 'weak' else 'job'}, mt19937_64 seed {SEED}(+rank)",
            "model": args.model, "sequences": args.sequences, "l2": "inputs larger than L2 (≈347 MB residues per scan)"}


def other_configs(torch, msv, _cabi, device: int) -> dict:
    """Side measurements on the same GPU, outside the headline's timed region (device-resident scans, CUDA events):
    BASELINE.json config 3 (every fixture model x 100k sequences) and config 5 (2405.hmm x 2048 titin-like sequences)."""
    stream = torch.cuda.current_stream()

    def gcups(model, db, n, leng, residues, steps=3):
        scores = torch.empty(n, dtype=torch.float32, device="cuda")
        for _ in range(2):
            db.score_device(model, scores, stream.cuda_stream)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for _ in range(steps):
            db.score_device(model, scores, stream.cuda_stream)
        t1.record(stream)
        torch.cuda.synchronize()
        return leng * float(residues) / (t0.elapsed_time(t1) / steps) / 1e6

    def load(name):
        prof = msv.Profile_HMM(os.path.join(REPO, "fixtures", "profile_HMMs", name))
        return prof.model_length - 1, msv.Model(_cabi.emission_table(prof.match_emissions), *_cabi.model_transitions(prof.model_length),
                                                 device=device)

    sweep_db = msv.Packed_sequences.synthetic_swissprot_like(100_000, 1400)
    resident = msv.Database(sweep_db.residues, sweep_db.offsets, device=device)
    names = sorted((f for f in os.listdir(os.path.join(REPO, "fixtures", "profile_HMMs")) if f.endswith(".hmm")),
                   key=lambda s: int(s.split(".")[0]))
    config3 = {}
    for name in names:
        leng, model = load(name)
        config3[name] = round(gcups(model, resident, len(sweep_db), leng, sweep_db.total_residues), 1)
        model.close()
    long_db = msv.Packed_sequences.synthetic_long_uniform(2048, 2405, 10_000, 35_000)
    long_resident = msv.Database(long_db.residues, long_db.offsets, device=device)
    leng, model = load("2405.hmm")
    config5 = round(gcups(model, long_resident, len(long_db), leng, long_db.total_residues), 1)
    return {"unit": "GCUPS, device-resident", "config3_model_sweep_100k_sequences": config3,
            "config5_2405hmm_2048_long_sequences": config5, "viterbi": viterbi_side_line(torch, msv, _cabi, resident, sweep_db, device)}


def viterbi_side_line(torch, msv, _cabi, resident, database, device: int) -> dict:
    """SURVEY section 8(f) rank 4, the Plan-7 local Viterbi scan (match/insert/delete) on 1400.hmm x the config-3 database:
    same GCUPS definition as the headline (LENG x residues / time), a sample checked bit-for-bit against
    oracle/viterbi_oracle.c (the checker; it is never on the measured path) and that oracle's speed on the host cores."""
    if os.path.join(REPO, "tests") not in sys.path:
        sys.path.insert(0, os.path.join(REPO, "tests"))
    from oracle_lib import Oracle, pack
    oracle = Oracle()
    h = oracle.load_hmm(os.path.join(REPO, "fixtures", "profile_HMMs", "1400.hmm"))
    leng = h["model_length"] - 1
    model = msv.ViterbiModel(_cabi.emission_table(h["match_emissions"]), _cabi.viterbi_transitions(h["transitions"]),
                             *_cabi.model_transitions(h["model_length"]), device=device)
    stream = torch.cuda.current_stream()
    scores = torch.empty(len(database), dtype=torch.float32, device="cuda")
    for _ in range(2):
        resident.viterbi_device(model, scores, stream.cuda_stream)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(3):
        resident.viterbi_device(model, scores, stream.cuda_stream)
    t1.record(stream)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    rng = np.random.default_rng(4)
    sample = rng.choice(len(database), size=256, replace=False)
    off = database.offsets
    sc, so = pack([database.residues[int(off[q]):int(off[q + 1])] for q in sample])
    table, tr3 = oracle.prepare(h["match_emissions"])
    cores = os.cpu_count() or 1
    t_cpu = time.perf_counter()
    want = oracle.viterbi_score_batch(table, oracle.viterbi_prepare(h["transitions"]), tr3, sc, so, threads=cores)
    t_cpu = time.perf_counter() - t_cpu
    got = scores.cpu().numpy()[sample]
    gcups = leng * float(database.total_residues) / ms / 1e6
    # per cell: 7 fp32 adds + 6 two-input maxima (4-way M, 2-way I, 2-way D, E) = 13 lane-ops; the FMNMX pipe (64 lanes/clk/SM,
    # 3 FMNMX + 1.5 FMNMX3 per cell) allows 14.2 cells/clk/SM, instruction issue (about 13.3 instructions per cell with
    # the operand loads) 9.6
    return {"workload": "1400.hmm x 100000 synthetic sequences (config-3 database), device-resident", "gcups": round(gcups, 1),
            "ms": round(ms, 3), "cells_per_clk_per_sm": round(gcups * 1e9 / 148 / 1.965e9, 2), "geometry": model.geometry,
            "frac_of_fp32_alu_roofline_13_ops_per_cell": round(gcups * 1e9 * 13 / (148 * 128 * 1.965e9), 3),
            "mismatches_vs_oracle": int((got.view(np.uint32) != want.view(np.uint32)).sum()), "compared": int(sample.size),
            "oracle_gcups": round(leng * float(so[-1]) / t_cpu / 1e9, 3), "oracle_threads": cores}


# ---- our arm ----------------------------------------------------------------------------------------------------------
def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

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

    model_path = os.path.join(REPO, "fixtures", "profile_HMMs", args.model)
    profile = msv.Profile_HMM(model_path)
    leng = profile.model_length - 1
    model = msv.Model(_cabi.emission_table(profile.match_emissions), *_cabi.model_transitions(profile.model_length), device=local)

    # ---- this rank's slice of the database ----
    if args.scaling == "weak":
        packed = msv.Packed_sequences.synthetic_swissprot_like(args.sequences, SEED + rank)
        codes, offsets = packed.residues, packed.offsets
    else:
        packed = msv.Packed_sequences.synthetic_swissprot_like(args.sequences, SEED)
        codes, offsets, _, _ = sharded.local_slice(packed.residues, packed.offsets, rank, world)
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
    # copies of the gathered array over NVLink peer memory (hmm_fasta_viterbi_b200.sharded.FusedGather), followed by one
    # device-side barrier; no collective.  The plain scan + NCCL all-gather is measured next to it (`nccl_gather`).
    step, gather_kind, peer_error, symmetric = nccl_step, "nccl all_gather_into_tensor", None, None
    if world > 1 and not args.no_peer_gather:
        try:
            fused = sharded.FusedGather(n_max, torch.device("cuda", local))
            symmetric = fused.scores

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
            symmetric = None

    # at least 3 warm-up steps; 10 when N > 1 (the first ~0.3 s after communicator / symmetric-memory set-up run ≈1 % slow)
    warmup_steps = max(args.warmup, 3 if world == 1 else 10)
    for _ in range(warmup_steps):
        step()
    fence()

    # ---- value: device-resident scan (+ gather), CUDA events ----
    sampler = ClockSampler(local)
    sampler.start()
    _cabi.launch_count(reset=True)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for _ in range(args.steps):
        step()
    stop.record(stream)
    fence()
    launches = _cabi.launch_count()
    clocks = sampler.stop()
    ms_total = start.elapsed_time(stop)

    # ---- comparison: plain scan + NCCL all-gather, and the two gathered arrays must be the same bits ----
    nccl_gather = None
    if world > 1 and symmetric is not None:
        for _ in range(3):
            nccl_step()
        fence()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        for _ in range(args.steps):
            nccl_step()
        n1.record(stream)
        fence()
        same = bool(torch.equal(symmetric.view(torch.int32), gathered.view(torch.int32)))
        flags = torch.tensor([n0.elapsed_time(n1), 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        nccl_gather = {"ms_total": float(flags[0].item()), "same_bits_as_fused_gather_on_every_rank": float(flags[1].item()) == 0.0}

    # ---- the scan kernel alone (for the roofline), same launches, no gather ----
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(args.steps):
        db.score_device(model, scores, stream.cuda_stream)
    k1.record(stream)
    fence()
    kernel_ms = k0.elapsed_time(k1) / args.steps

    # ---- e2e: host buffers through msv_cuda_score_batch (H2D + bucketing + scan + D2H) ----
    for _ in range(2):
        model.score_batch(pin_codes, pin_offsets, pin_scores)
    fence()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.score_batch(pin_codes, pin_offsets, pin_scores)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks ----
    agg = torch.tensor([ms_total, e2e_s, kernel_ms], dtype=torch.float64, device="cuda")
    cells = torch.tensor([cells_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    ms_total, e2e_s, kernel_ms_max = (float(v) for v in agg.tolist())
    cells_job = float(cells.item())

    # ---- parity spot check against the oracle (outside every timed region) ----
    result = scores[:n_local].cpu().numpy()
    e2e_result = pin_scores[:n_local].numpy()
    parity = {"device_equals_e2e": bool((result.view(np.uint32) == e2e_result.view(np.uint32)).all())}

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
        # DRAM traffic per launch comes from the committed ncu --set full capture of this same workload (it cannot be
        # measured live without a profiler); null when the workload differs from the captured one.
        traffic, traffic_source = None, None
        try:
            with open(os.path.join(REPO, "profiles", "roofline_traffic.json")) as f:
                captured = json.load(f)
            if captured.get("workload") == f"{args.model} x {args.sequences} sequences" and world == 1:
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
        out = {
            "metric": "MSV GCUPS at M=1400", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": warmup_steps, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, leng) | {"geometry": model.geometry, "cells_per_step": cells_job},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "GCUPS", "ms_per_step": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": int(pin_codes.numel() + pin_offsets.numel() * 8),
                    "d2h_bytes_per_step": int(n_local * 4), "api": "msv_cuda_score_batch (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "parity": parity,
        }
        if world > 1:
            out["config"]["gather"] = gather_kind
            if nccl_gather is not None:
                ms_nccl = nccl_gather.pop("ms_total")
                out["nccl_gather"] = {"value": cells_job * args.steps / (ms_nccl / 1e3) / 1e9, "unit": "GCUPS",
                                      "ms_per_step": ms_nccl / args.steps, "what": "plain scan + NCCL all_gather_into_tensor"} | nccl_gather
            if peer_error is not None:
                out["peer_gather_unavailable"] = peer_error
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            kind, scorer = cpu_reference_scorer(model_path)
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
        if world == 1 and not args.no_other_configs:
            out["other_configs"] = other_configs(torch, msv, _cabi, local)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
