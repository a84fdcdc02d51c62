#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: DTW GCUPS (1e9 reference cell updates / s) and
all-pairs matrix wall time for the banded weighted DTW distance matrix.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C2|C4|C5] [--seqs SEQS]
                    [--mode strict|fast] [--impl b200|reference]

One "step" = one full all-pairs matrix (every ordered pair, both orientations) of the
workload: DTW kernels over this rank's shard of the work units, one NCCL all-gather of
the packed shard results (N > 1), and the scatter kernel that writes the n x n matrix.
`value` is measured with the packed sequence arena already resident in HBM; `e2e` is
the same job through the reference-facing host interface with HOST buffers (sequence
packing + H2D, kernels, gather, scatter, D2H of the matrix) inside the timed region.

Cells are counted with the reference's own visit rule (src/alignments.rs:174-175) for
every ordered pair -- not the cells the kernel chooses to run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dtw_gcups"
UNIT = "GCUPS"
DEVICE = "cuda"   # tests/test_bench_flow.py drives the multi-rank control flow on CPU with a stub aligner
PIN = True


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def workload(name, n):
    from audio_pattern_discovery_b200 import synth
    c, seqs, _ = synth.make_config(name, n)
    return c, seqs


def reference_cells_total(seqs, pct):
    """Exact sum over all ordered pairs of the reference's visited cells (closed form per
    distinct length pair)."""
    from oracle import oracle  # checker-side arithmetic only (cpu legs); cheap
    lens = np.array([len(s) for s in seqs])
    vals, counts = np.unique(lens, return_counts=True)
    total = 0
    for a, ca in zip(vals, counts):
        for b, cb in zip(vals, counts):
            pairs = ca * cb - (ca if a == b else 0)
            if pairs and a >= 1 and b >= 1:
                total += int(pairs) * oracle.pair_cells(int(a), int(b), pct)
    return total


def _band_sum(a, b, k):
    """sum_{t=0}^{k-1} max(0, min(a, b - t))"""
    if a <= 0 or b <= 0 or k <= 0:
        return 0
    k1 = min(max(b - a + 1, 0), k)
    k2 = min(b, k)
    s = a * k1
    if k2 > k1:
        s += (k2 - k1) * b - (k1 + k2 - 1) * (k2 - k1) // 2
    return s


def cells_visited(n, m, w):
    """Closed form of SURVEY.md Appendix C (src/alignments.rs:174-175): diagonals j-i = 0..w-1 hold
    min(n, m-k) cells, diagonals j-i = -1..-w hold min(m, n-k)."""
    return _band_sum(n, m, w) + _band_sum(m, n - 1, w)


def needed_cells_total(seqs, pct):
    """The cells that can influence the score (rows <= n-1, columns <= m-1), all ordered pairs --
    printed beside the reference's own count so the two GCUPS conventions convert."""
    lens = np.array([len(s) for s in seqs])
    vals, counts = np.unique(lens, return_counts=True)
    total = 0
    for a, ca in zip(vals, counts):
        for b, cb in zip(vals, counts):
            pairs = int(ca) * int(cb) - (int(ca) if a == b else 0)
            if pairs and a >= 1 and b >= 1:
                prod = float(np.float32(pct) * np.float32(max(a, b)))
                band = 0 if (prod != prod or prod <= 0) else int(min(prod, 1e9))
                w = max(band, abs(int(a) - int(b))) + 2
                total += pairs * cells_visited(int(a) - 1, int(b) - 1, w)
    return total


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def i_cell(dim, strict):
    """FP32-pipe instructions per ORDERED reference cell for the shared-distance kernel
    (SURVEY.md section 8d): FAST D + 8; STRICT (3D + 7 + 15) / 2."""
    return (1.5 * dim + 11.0) if strict else (dim + 8.0)


def cpu_leg(seqs, c, target_s=15.0, max_s=384):
    """Times the oracle's dense restatement (reference threading scheme, all host cores) on
    the first S sequences of the workload; S grows until the run takes ~target_s."""
    from oracle import oracle
    cores = os.cpu_count() or 1
    S = min(len(seqs), 24)
    ins, dele, mat = c["weights"]
    while True:
        sub = seqs[:S]
        cells = reference_cells_total(sub, c["pct"])
        t0 = time.perf_counter()
        oracle.align_all(sub, c["pct"], ins, dele, mat, workers=cores, variant="dense")
        dt = time.perf_counter() - t0
        if dt >= target_s / 3 or S >= min(len(seqs), max_s):
            break
        grow = (target_s / max(dt, 1e-3)) ** 0.5
        S = int(min(len(seqs), max_s, max(S + 8, S * min(grow, 4.0))))
    return {"gcups": cells / dt / 1e9, "seconds": dt, "cells": cells, "S": S, "cores": cores}


def literal_leg(seqs, c, S=16):
    from oracle import oracle
    cores = os.cpu_count() or 1
    sub = seqs[:min(S, len(seqs))]
    cells = reference_cells_total(sub, c["pct"])
    ins, dele, mat = c["weights"]
    t0 = time.perf_counter()
    oracle.align_all(sub, c["pct"], ins, dele, mat, workers=cores, variant="literal")
    dt = time.perf_counter() - t0
    return {"gcups": cells / dt / 1e9, "seconds": dt, "S": len(sub)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    c, seqs = workload(args.workload, args.n)
    vals = []
    info = None
    for it in range(args.warmup + args.steps):
        info = cpu_leg(seqs, c, target_s=args.ref_seconds, max_s=args.ref_max_seqs) if info is None else info
        from oracle import oracle
        sub = seqs[:info["S"]]
        t0 = time.perf_counter()
        oracle.align_all(sub, c["pct"], *c["weights"], workers=info["cores"], variant="dense")
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            vals.append(dt)
    ms = float(np.mean(vals)) * 1e3
    v = info["cells"] / (ms / 1e3) / 1e9
    sample = "first %d of %d sequences of %s (%d ordered pairs, %.3e reference cells) per step" % (
        info["S"], len(seqs), args.workload, info["S"] * (info["S"] - 1), info["cells"])
    lit = literal_leg(seqs, c)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, c, seqs),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": sample,
                             "variant": "dense rolling-band restatement, reference threading scheme "
                                        "(static row blocks, one thread per core), gcc -O2",
                             "literal_hashmap_gcups": lit["gcups"], "literal_sample_sequences": lit["S"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def workload_config(args, c, seqs):
    lens = np.array([len(s) for s in seqs])
    return {"workload": "%s: %d sequences, len %d-%d, dim %d, band %.0f%%, weights ins/del/match %s, all n(n-1) "
                        "ordered pairs" % (args.workload, len(seqs), lens.min(), lens.max(), c["dim"],
                                            100 * c["pct"], "/".join("%g" % w for w in c["weights"])),
            "n_sequences": len(seqs), "dim": c["dim"], "band_pct": c["pct"], "mode": args.mode,
            "l2": "arena %.0f MB %s" % (sum(len(s) for s in seqs) * c["dim"] * 4 / 1e6,
                                         "larger than L2 (126 MB): no flush needed"
                                         if sum(len(s) for s in seqs) * c["dim"] * 4 > 126e6
                                         else "smaller than L2: L2 flushed (256 MB write) between timed steps"),
            "parallelism": "pair-space sharded over %d GPU(s), one all-gather" % args.gpus}


def matrix_checksum(m):
    """Position-dependent 64-bit checksum of the uint32 view of a (n, n) float32 matrix (torch
    tensor on any device, or numpy array): sum_k bits[k] * ((k * 2654435761 + 1) mod 2^32)  mod 2^64.
    Identical matrices give identical checksums on every rank count and in both launch forms."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(m)) if isinstance(m, np.ndarray) else m
    flat = t.reshape(-1).view(torch.int32)
    total = 0
    step = 1 << 26
    for o in range(0, flat.numel(), step):
        v = flat[o:o + step].to(torch.int64) & 0xffffffff
        idx = torch.arange(o, o + v.numel(), dtype=torch.int64, device=v.device)
        w = (idx * 2654435761 + 1) & 0xffffffff
        total = (total + int((v * w).sum().item())) & 0xffffffffffffffff
    return "%016x" % total


def parity_sample(seqs, c, strict, fetch, n_pairs=256, seed=77):
    """Checker leg (untimed): `n_pairs` random ordered pairs of the assembled matrix against the
    oracle -- bit for bit in STRICT mode, <= 1e-5 relative in FAST mode (BASELINE.json's tolerance).
    fetch(i, j) -> matrix entries as a float32 array."""
    from oracle import oracle
    n = len(seqs)
    if n < 2:
        return {"parity_checked": 0, "parity_ok": True}
    rng = np.random.default_rng(seed)
    i = rng.integers(0, n, size=n_pairs)
    j = (i + rng.integers(1, n, size=n_pairs)) % n
    pairs = np.stack([i, j], axis=1).astype(np.uint32)
    ins, dele, mat = c["weights"]
    want = oracle.align_pairs(seqs, pairs, c["pct"], ins, dele, mat, workers=os.cpu_count() or 1, variant="dense")
    got = np.asarray(fetch(i, j), dtype=np.float32)
    if strict:
        bad = int((got.view(np.uint32) != want.view(np.uint32)).sum())
        return {"parity_checked": int(n_pairs), "parity_rule": "bit-exact vs oracle (uint32 view)", "parity_mismatches": bad,
                "parity_ok": bad == 0}
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.where(want == got, 0.0, np.abs(got - want) / np.abs(want))
    worst = float(np.nanmax(rel)) if rel.size else 0.0
    return {"parity_checked": int(n_pairs), "parity_rule": "<= 1e-5 relative vs oracle", "parity_max_rel_err": worst,
            "parity_ok": bool(worst <= 1e-5)}


def roofline_object(args, c, st, peaks, peaks_src, clocks, kernel_ms_avg, scat_ms, arena_bytes, n, ms_per_step, n_dev=1):
    strict = args.mode == "strict"
    sm_count = st["sm_count"]
    clk_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    unitw = tuple(c["weights"]) == (1.0, 1.0, 1.0)
    icell = i_cell(c["dim"], strict)
    cells_local = st["cells_reference"] / n_dev          # per GPU (units are dealt evenly in cost order)
    achieved = cells_local * icell / (kernel_ms_avg / 1e3) / 1e12          # T lane-instr / s, one GPU
    peak_max = sm_count * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    peak_obs = sm_count * 128 * clk_mhz * 1e6 / 1e12
    fma_cyc = (1.625 * c["dim"] if strict else c["dim"]) + 1.5
    return {
        "bound": "fp32-issue", "kernel": "dtw_units_*kernel (%s%s)" % (args.mode, "" if unitw else ", weighted recurrence"),
        "achieved": achieved, "peak": peak_max, "unit": "Tlane-instr/s", "frac": achieved / peak_max,
        "frac_at_observed_clock": achieved / peak_obs, "observed_sm_mhz": clk_mhz,
        "peak_source": "%d SMs x 128 FP32 lanes x %s sm_max_mhz (%s MEASURED_PEAKS.json)" % (
            sm_count, peaks.get("sm_max_mhz"), peaks_src),
        "instr_per_cell": icell,
        "instr_per_cell_note": "SURVEY.md 8(d) shared-distance model for unit penalties: FAST D+8, STRICT 1.5D+11 per ordered "
                               "reference cell%s" % ("" if unitw else "; the weighted recurrence (penalty select + multiply per "
                                                     "cell and orientation) is NOT counted, so frac understates this run"),
        "fma_pipe_model": {"cycles_per_ordered_cell": fma_cyc,
                           "frac_of_fma_pipe_at_max_clock": cells_local * fma_cyc / (kernel_ms_avg / 1e3)
                           / (sm_count * 4 * 32 * peaks.get("sm_max_mhz", 1965.0) * 1e6),
                           "note": "f32x2 ops occupy the 32-lane FMA pipe for 2 cycles (measured: "
                                   "profiles/r1_microbench*.txt); DESIGN.md section 4"},
        "flops_view": {"algorithmic_flops_per_cell": 3 * c["dim"] + 7,
                       "achieved_tflops": cells_local * (3 * c["dim"] + 7) / (kernel_ms_avg / 1e3) / 1e12,
                       "peak_fp32_tflops": 2 * peak_max,
                       "frac": cells_local * (3 * c["dim"] + 7) / (kernel_ms_avg / 1e3) / 1e12 / (2 * peak_max)},
        "kernel_ms_per_step": kernel_ms_avg, "cells_per_step_per_gpu": int(cells_local),
        "cells_computed_over_reference": (st["cells_computed"] / st["cells_reference"]) if st["cells_reference"] else None,
        "scatter_ms_per_step": scat_ms, "traffic": None,
        "hbm": {"algorithmic_bytes": int(arena_bytes + 2 * n * n * 4),
                "achieved_gbs": (arena_bytes + 2 * n * n * 4) / (ms_per_step / 1e3) / 1e9,
                "peak_gbs": peaks.get("hbm_gbs"), "note": "not the bound: < 1% of HBM peak"}}


def threshold_object(thr, sel_ms, n, peaks):
    # one read of the matrix is the order statistic's algorithmic traffic; the 4-pass radix select reads it 4 times
    gbs = 4 * n * n / (sel_ms / 1e3) / 1e9 if sel_ms > 0 else None
    return {"clustering_percentile": 0.05, "threshold": float(thr), "ms": sel_ms, "passes": 4,
            "algorithmic_bytes": 4 * n * n, "traffic_bytes": 16 * n * n, "achieved_gbs": gbs,
            "peak_gbs": peaks.get("hbm_gbs"), "frac": (gbs / peaks.get("hbm_gbs", 6650.0)) if gbs else None}


def backtrack_leg(ctx, seqs, c, mode, frac=0.01, seed=1005, max_pairs=None):
    """C5: on-device trace-back for 1 % of the unordered pairs, timed separately (K2)."""
    from oracle import oracle
    n = len(seqs)
    rng = np.random.default_rng(seed)
    total = n * (n - 1) // 2
    k = max(1, int(total * frac))
    if max_pairs:
        k = min(k, max_pairs)
    flat = rng.choice(total, size=k, replace=False)
    # unordered pair index -> (i, j), i < j
    i = (np.floor((2 * n - 1 - np.sqrt((2 * n - 1) ** 2 - 8 * flat.astype(np.float64))) / 2)).astype(np.int64)
    base = i * (2 * n - i - 1) // 2
    fix = flat < base
    i[fix] -= 1
    base = i * (2 * n - i - 1) // 2
    j = flat - base + i + 1
    pairs = np.stack([i, j], axis=1).astype(np.uint32)
    ins, dele, mat = c["weights"]
    cap = int(max(len(s) for s in seqs)) * 2 + 2
    ctx.align_pairs(pairs[:2], c["pct"], ins, dele, mat, mode, want_paths=True, path_cap=cap)   # warm-up
    t0 = time.perf_counter()
    scores, paths, lens = ctx.align_pairs(pairs, c["pct"], ins, dele, mat, mode, want_paths=True, path_cap=cap)
    wall = time.perf_counter() - t0
    kms = ctx.stats()["path_ms"]
    cells = sum(oracle.pair_cells(len(seqs[a]), len(seqs[b]), c["pct"]) for a, b in pairs[:64]) / min(64, k) * k
    # checker: a few of them against the oracle (scores bit for bit in STRICT mode; the traced path too
    # where the literal hash-map restatement, the only one that traces, is affordable)
    chk = min(4, k)
    ok = True
    for q in range(chk):
        a, b = int(pairs[q, 0]), int(pairs[q, 1])
        small = len(seqs[a]) * len(seqs[b]) <= 600 * 600
        if small:
            s_ref, p_ref = oracle.dtw(seqs[a], seqs[b], c["pct"], ins, dele, mat, variant="literal", want_path=True)
            ok &= bool(np.array_equal(p_ref, paths[q]))
        else:
            s_ref = oracle.dtw(seqs[a], seqs[b], c["pct"], ins, dele, mat, variant="dense")
        if mode == 0:
            ok &= bool(np.float32(s_ref).view(np.uint32) == np.float32(scores[q]).view(np.uint32))
    return {"pairs": int(k), "fraction_of_unordered_pairs": frac, "wall_s": wall, "kernel_ms": kms,
            "gcups_kernel": cells / (kms / 1e3) / 1e9 if kms > 0 else None, "gcups_wall": cells / wall / 1e9,
            "reference_cells": int(cells), "path_cells_returned": int(np.sum(lens)), "scores_checked_vs_oracle": chk,
            "checked_ok": ok}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from audio_pattern_discovery_b200 import APD_MODE_FAST, APD_MODE_STRICT
    from audio_pattern_discovery_b200.distributed import ShardedAligner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks, peaks_src = load_peaks()

    c, seqs = workload(args.workload, args.n)
    ins, dele, mat = c["weights"]
    strict = args.mode == "strict"
    mode = APD_MODE_STRICT if strict else APD_MODE_FAST
    n = len(seqs)
    arena_bytes = sum(len(s) for s in seqs) * c["dim"] * 4
    need_flush = arena_bytes <= 126e6
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=DEVICE) if need_flush else None

    al = ShardedAligner(seqs, device=local, mode=mode)
    al.ctx.packed_len(c["pct"], mode)                # builds the unit plan
    cells_local_t = torch.tensor([al.stats()["cells_reference"]], dtype=torch.int64, device=DEVICE)
    if world > 1:
        dist.all_reduce(cells_local_t)
    cells_total = int(cells_local_t[0])              # reference visit rule, all ordered pairs (library-side count)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=DEVICE)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- device-resident arm ----------------------------------------------------
    for _ in range(args.warmup):
        al.align_all_device(c["pct"], ins, dele, mat)
        al.synchronize()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    step_ms, kern_ms, scat_ms, launches = [], [], [], 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1.0)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(al.stream)                         # events on the stream the kernels are launched on
        al.align_all_device(c["pct"], ins, dele, mat)
        e1.record(al.stream)
        al.synchronize()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        st = al.stats()
        kern_ms.append(st["kernel_ms"]); scat_ms.append(st["scatter_ms"]); launches += st["kernel_launches"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = maxrank(float(np.sum(step_ms)))
    ms_per_step = total_ms / args.steps
    gcups = cells_total / (ms_per_step / 1e3) / 1e9
    st = al.stats()
    kernel_ms_avg = maxrank(float(np.mean(kern_ms)))
    launch_plan = al.ctx.launch_plan() if hasattr(al.ctx, "launch_plan") else None

    # ---- end-to-end arm: host buffers in, host matrix out (warm: plan, staging and pinned output reused) ----
    host_out = torch.empty((n, n), dtype=torch.float32, pin_memory=PIN) if rank == 0 else None
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_ms = []
    for it in range(1 + e2e_steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(al.stream)
        al.set_sequences(seqs)                       # host packing + H2D of the arena
        al.align_all(c["pct"], ins, dele, mat, out=host_out, to_host=(rank == 0))
        e1.record(al.stream)
        torch.cuda.synchronize()
        dt = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        if it >= 1:
            e2e_ms.append(dt)
    e2e_ms_step = maxrank(float(np.mean(e2e_ms)))
    e2e_gcups = cells_total / (e2e_ms_step / 1e3) / 1e9
    h2d = al.stats()["h2d_bytes"]

    # ---- cold end-to-end: a FRESH context per rank, first set_sequences + align_all (unit plan built
    # inside), matrix into a PAGEABLE host buffer -- the reference calls align_all exactly once (src/main.rs:191)
    cold_out = torch.empty((n, n), dtype=torch.float32) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    al2 = ShardedAligner(seqs, device=local, mode=mode)           # apd_create + packing + H2D
    t1 = time.perf_counter()
    al2.align_all(c["pct"], ins, dele, mat, out=cold_out, to_host=(rank == 0))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    cold_create_set = maxrank(t1 - t0)
    cold_total = maxrank(t2 - t0)
    al2.close()
    cold = {"value": cells_total / cold_total / 1e9, "unit": UNIT, "total_s": cold_total,
            "create_and_set_sequences_s": cold_create_set, "align_all_s": cold_total - cold_create_set,
            "rank0_create_s": getattr(al2, "t_create", None), "rank0_set_sequences_s": getattr(al2, "t_set_sequences", None),
            "host_output": "pageable", "over_warm_e2e": cold_total / (e2e_ms_step / 1e3),
            "note": "fresh apd context per rank (CUDA itself already initialised in the process), unit plan built inside, "
                    "pageable destination"}

    # ---- evidence (untimed): the threshold step, a checksum of the assembled matrix, a sample against the oracle.
    # NOTE: rank 0 only, so NO collective may be called here -- the matrix is the one every rank
    # assembled in the last end-to-end step.
    sel = None
    evidence = None
    if rank == 0:
        m = al.matrix_device()
        thr = [al.percentile_of_matrix(0.05) for _ in range(3)][-1]
        sel = threshold_object(thr, al.stats()["select_ms"], n, peaks)
        evidence = {"matrix_checksum_u64": matrix_checksum(m), "checksum_mode": args.mode,
                    "host_copy_equals_device_matrix": bool(torch.equal(host_out.view(torch.int32), m.cpu().view(torch.int32)))
                    if n <= 12000 else None}
        if not args.no_parity:
            evidence.update(parity_sample(seqs, c, strict, lambda i, j: m[torch.as_tensor(i), torch.as_tensor(j)].cpu().numpy()))
    barrier()

    other = None
    if args.other_mode_steps > 0:
        omode = APD_MODE_FAST if strict else APD_MODE_STRICT
        other = measure_other_mode(seqs, c, local, omode, args.other_mode_steps, cells_total, world, maxrank, barrier,
                                   need_flush, flush)
        other["mode"] = "fast" if strict else "strict"

    line = None
    if rank == 0:
        st_all = dict(st)
        st_all["cells_reference"] = cells_total
        st_all["cells_computed"] = st["cells_computed"] * world
        line = {
            "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, c, seqs),
            "launch_form": "one process per GPU (torch.distributed, NCCL all-gather)" if world > 1 else "one process, one GPU",
            "matrix_wall_s": ms_per_step / 1e3, "matrix_wall_s_e2e": e2e_ms_step / 1e3,
            "reference_cells": cells_total, "needed_cells": needed_cells_total(seqs, c["pct"]),
            "wall_s_timed_region": wall,
            "e2e": {"value": e2e_gcups, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(n * n * 4), "steps": e2e_steps, "ms_per_step": e2e_ms_step},
            "e2e_cold": cold,
            "gpu_launches": int(launches),
            "launch_classes": launch_plan,
            "other_mode": other,
            "threshold_select": sel,
            "clocks": clocks,
            "roofline": roofline_object(args, c, st_all, peaks, peaks_src, clocks, kernel_ms_avg, float(np.mean(scat_ms)),
                                        arena_bytes, n, ms_per_step, n_dev=world),
        }
        line.update(evidence or {})
        if args.workload == "C5" and not args.no_backtrack:
            line["backtrack"] = backtrack_leg(al.ctx, seqs, c, mode, max_pairs=args.backtrack_pairs)
        if world == 1 and not args.no_cpu:
            info = cpu_leg(seqs, c, target_s=args.ref_seconds, max_s=args.ref_max_seqs)
            lit = literal_leg(seqs, c)
            line["cpu_baseline"] = {
                "value": info["gcups"], "unit": UNIT, "cores": info["cores"], "kind": "port",
                "sample": "first %d of %d sequences of %s (%d ordered pairs, %.3e reference cells, %.1f s)" % (
                    info["S"], n, args.workload, info["S"] * (info["S"] - 1), info["cells"], info["seconds"]),
                "variant": "oracle dense rolling-band restatement, reference threading scheme, gcc -O2",
                "literal_hashmap_gcups": lit["gcups"], "literal_sample_sequences": lit["S"],
                "full_matrix_extrapolated_s": cells_total / (info["gcups"] * 1e9)}
        emit(line)
        if evidence and evidence.get("parity_ok") is False:
            print("bench.py: PARITY CHECK FAILED against the oracle: %r" % (evidence,), file=sys.stderr)
    al.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_single_process(args):
    """`--single-process`: ONE host process drives all --gpus devices through the library's own
    device group (apd_create_multi) -- the drop-in form of the reference's one blocking
    align_all call.  Host buffers in, host matrix out; no torch.distributed, no NCCL."""
    import torch
    from audio_pattern_discovery_b200 import APD_MODE_FAST, APD_MODE_STRICT, visible_devices
    from audio_pattern_discovery_b200.distributed import GroupAligner

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    G = args.gpus
    if visible_devices() < G:
        raise SystemExit("bench.py: --gpus %d but only %d device(s) visible" % (G, visible_devices()))
    peaks, peaks_src = load_peaks()
    c, seqs = workload(args.workload, args.n)
    ins, dele, mat = c["weights"]
    strict = args.mode == "strict"
    mode = APD_MODE_STRICT if strict else APD_MODE_FAST
    n = len(seqs)
    arena_bytes = sum(len(s) for s in seqs) * c["dim"] * 4

    # cold first: fresh group, first call, pageable destination (the process has no CUDA context yet)
    cold_out = np.empty((n, n), dtype=np.float32)
    t0 = time.perf_counter()
    al = GroupAligner(seqs, devices=list(range(G)), mode=mode)
    t1 = time.perf_counter()
    al.align_all(c["pct"], ins, dele, mat, out=cold_out)
    t2 = time.perf_counter()
    cells_total = al.stats()["cells_reference"]
    first = {"total_s": t2 - t0, "create_and_set_sequences_s": t1 - t0, "align_all_s": t2 - t1,
             "note": "very first call of the process: includes CUDA initialisation of all devices and module loading"}
    al.close()
    t0 = time.perf_counter()
    al = GroupAligner(seqs, devices=list(range(G)), mode=mode)
    t1 = time.perf_counter()
    al.align_all(c["pct"], ins, dele, mat, out=cold_out)
    t2 = time.perf_counter()
    cold = {"value": cells_total / (t2 - t0) / 1e9, "unit": UNIT, "total_s": t2 - t0, "create_and_set_sequences_s": t1 - t0,
            "align_all_s": t2 - t1, "host_output": "pageable", "first_call_of_process": first,
            "note": "fresh device group (CUDA already initialised in the process), unit plan built inside, pageable destination"}

    out = np.empty((n, n), dtype=np.float32)          # pageable, like the reference's Vec<f32>
    for _ in range(max(args.warmup - 1, 0)):
        al.align_all(c["pct"], ins, dele, mat, out=out)
    sampler = ClockSampler(0)
    sampler.start()
    step_ms, kern_ms, scat_ms, launches = [], [], [], 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):                        # arena resident; kernels + gather + scatter + D2H
        t0 = time.perf_counter()
        al.align_all(c["pct"], ins, dele, mat, out=out)
        step_ms.append((time.perf_counter() - t0) * 1e3)
        st = al.stats()
        kern_ms.append(st["kernel_ms"]); scat_ms.append(st["scatter_ms"]); launches += st["kernel_launches"]
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    ms_per_step = float(np.mean(step_ms))
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_ms = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        al.set_sequences(seqs)
        al.align_all(c["pct"], ins, dele, mat, out=out)
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    e2e_ms_step = float(np.mean(e2e_ms))
    st = al.stats()
    thr = [al.percentile_of_matrix(0.05) for _ in range(3)][-1]
    sel = threshold_object(thr, al.stats()["select_ms"], n, peaks)
    evidence = {"matrix_checksum_u64": matrix_checksum(out), "checksum_mode": args.mode,
                "cold_matrix_equals_warm_matrix": bool(np.array_equal(out.view(np.uint32), cold_out.view(np.uint32)))}
    if not args.no_parity:
        evidence.update(parity_sample(seqs, c, strict, lambda i, j: out[i, j]))
    kernel_ms_avg = float(np.mean(kern_ms))
    line = {
        "metric": METRIC, "value": cells_total / (ms_per_step / 1e3) / 1e9, "unit": UNIT, "n_gpus": G, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, c, seqs),
        "launch_form": "ONE process, %d GPU(s) inside the library (apd_create_multi; peer stores: %s)" % (G, al.ctx.peer_stores),
        "value_note": "host-timed apd_align_all with the arena resident: kernels + fused peer-store gather + scatter + D2H of the "
                      "matrix into pageable host memory",
        "matrix_wall_s": ms_per_step / 1e3, "matrix_wall_s_e2e": e2e_ms_step / 1e3, "reference_cells": cells_total,
        "needed_cells": needed_cells_total(seqs, c["pct"]), "wall_s_timed_region": wall,
        "e2e": {"value": cells_total / (e2e_ms_step / 1e3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(st["h2d_bytes"]),
                "d2h_bytes_per_step": int(n * n * 4), "steps": e2e_steps, "ms_per_step": e2e_ms_step},
        "e2e_cold": cold, "gpu_launches": int(launches), "launch_classes": al.ctx.launch_plan(), "threshold_select": sel,
        "clocks": clocks,
        "roofline": roofline_object(args, c, st, peaks, peaks_src, clocks, kernel_ms_avg, float(np.mean(scat_ms)), arena_bytes,
                                    n, ms_per_step, n_dev=G),
    }
    line.update(evidence)
    if args.workload == "C5" and not args.no_backtrack:
        line["backtrack"] = backtrack_leg(al.ctx, seqs, c, mode, max_pairs=args.backtrack_pairs)
    emit(line)
    if evidence.get("parity_ok") is False:
        print("bench.py: PARITY CHECK FAILED against the oracle: %r" % (evidence,), file=sys.stderr)
    al.close()
    return 0


def measure_other_mode(seqs, c, local, mode, steps, cells_total, world, maxrank, barrier, need_flush, flush):
    """Device-resident GCUPS of the other arithmetic variant (reported beside the headline)."""
    import torch
    from audio_pattern_discovery_b200.distributed import ShardedAligner
    ins, dele, mat = c["weights"]
    al = ShardedAligner(seqs, device=local, mode=mode)
    al.align_all_device(c["pct"], ins, dele, mat)
    al.synchronize()
    barrier()
    ms = []
    for _ in range(steps):
        if need_flush:
            flush.fill_(1.0)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(al.stream)
        al.align_all_device(c["pct"], ins, dele, mat)
        e1.record(al.stream)
        al.synchronize()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    barrier()
    kernel_ms = al.stats()["kernel_ms"]
    al.close()
    t = maxrank(float(sum(ms))) / steps
    return {"gcups": cells_total / (t / 1e3) / 1e9, "ms_per_step": t, "steps": steps, "kernel_ms": kernel_ms}


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints to
    fd 1 (e.g. NCCL's version banner) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1ref", "C2", "C3", "C4", "C5"])
    ap.add_argument("--seqs", dest="n", type=int, default=None, help="override the number of sequences")
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--other-mode-steps", type=int, default=2,
                    help="timed steps of the other arithmetic variant, reported beside the headline (0 = skip)")
    ap.add_argument("--ref-seconds", type=float, default=15.0)
    ap.add_argument("--ref-max-seqs", type=int, default=384)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle sample (parity_checked)")
    ap.add_argument("--single-process", action="store_true",
                    help="ONE process drives all --gpus devices through the library's device group (apd_create_multi)")
    ap.add_argument("--no-backtrack", action="store_true", help="C5: skip the on-device trace-back leg")
    ap.add_argument("--backtrack-pairs", type=int, default=None, help="C5: cap on the traced pairs (default: 1%% of all)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.single_process:
        return run_single_process(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
