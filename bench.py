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
    dtw_launches = max(st["kernel_launches"] - 1, 1)

    # ---- end-to-end arm: host buffers in, host matrix out ---------------------
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

    # the threshold step of the handoff (src/clustering.rs:101) on the assembled device matrix
    # NOTE: rank 0 only, so NO collective may be called here -- the matrix is the one every rank
    # assembled in the last end-to-end step.
    sel = None
    if rank == 0:
        m = al.matrix_device()
        thr = [al.percentile_of_matrix(0.05) for _ in range(3)][-1]
        sel_ms = al.stats()["select_ms"]
        sel = {"clustering_percentile": 0.05, "threshold": float(thr), "ms": sel_ms, "passes": 4,
               "algorithmic_bytes": 16 * n * n, "achieved_gbs": 16 * n * n / (sel_ms / 1e3) / 1e9,
               "peak_gbs": peaks.get("hbm_gbs"), "frac": 16 * n * n / (sel_ms / 1e3) / 1e9 / peaks.get("hbm_gbs", 6650.0)}
    barrier()

    other = None
    if args.other_mode_steps > 0:
        omode = APD_MODE_FAST if strict else APD_MODE_STRICT
        other = measure_other_mode(seqs, c, local, omode, args.other_mode_steps, cells_total, world, maxrank, barrier,
                                   need_flush, flush)
        other["mode"] = "fast" if strict else "strict"

    line = None
    if rank == 0:
        sm_count = st["sm_count"]
        clk_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        icell = i_cell(c["dim"], strict)
        cells_local = st["cells_reference"]
        achieved = cells_local * icell / (kernel_ms_avg / 1e3) / 1e12          # T lane-instr / s, this GPU
        peak_max = sm_count * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        peak_obs = sm_count * 128 * clk_mhz * 1e6 / 1e12
        line = {
            "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, c, seqs),
            "matrix_wall_s": ms_per_step / 1e3, "matrix_wall_s_e2e": e2e_ms_step / 1e3,
            "reference_cells": cells_total, "needed_cells": needed_cells_total(seqs, c["pct"]),
            "wall_s_timed_region": wall,
            "e2e": {"value": e2e_gcups, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(n * n * 4), "steps": e2e_steps, "ms_per_step": e2e_ms_step},
            "gpu_launches": int(launches),
            "other_mode": other,
            "threshold_select": sel,
            "clocks": clocks,
            "roofline": {
                "bound": "fp32-issue", "kernel": "dtw_units_kernel (%s)" % args.mode,
                "achieved": achieved, "peak": peak_max, "unit": "Tlane-instr/s", "frac": achieved / peak_max,
                "frac_at_observed_clock": achieved / peak_obs, "observed_sm_mhz": clk_mhz,
                "peak_source": "%d SMs x 128 FP32 lanes x %s sm_max_mhz (%s MEASURED_PEAKS.json)" % (
                    sm_count, peaks.get("sm_max_mhz"), peaks_src),
                "instr_per_cell": icell,
                "fma_pipe_model": {"cycles_per_ordered_cell": (1.625 * c["dim"] if strict else c["dim"]) + 1.5,
                                   "frac_of_fma_pipe_at_max_clock":
                                       cells_local * ((1.625 * c["dim"] if strict else c["dim"]) + 1.5)
                                       / (kernel_ms_avg / 1e3) / (sm_count * 4 * 32 * peaks.get("sm_max_mhz", 1965.0) * 1e6),
                                   "note": "f32x2 ops occupy the 32-lane FMA pipe for 2 cycles (measured: "
                                           "profiles/r1_microbench*.txt); DESIGN.md section 4"},
                "flops_view": {"algorithmic_flops_per_cell": 3 * c["dim"] + 7,
                               "achieved_tflops": cells_local * (3 * c["dim"] + 7) / (kernel_ms_avg / 1e3) / 1e12,
                               "peak_fp32_tflops": 2 * peak_max,
                               "frac": cells_local * (3 * c["dim"] + 7) / (kernel_ms_avg / 1e3) / 1e12 / (2 * peak_max),
                               "note": "3D+7 flops per ordered reference cell (SURVEY.md 8d) against the FP32 FMA peak "
                                       "(2 flops per lane-instruction); subtractions, compares and the sqrt cannot be FMAs"},
                "kernel_ms_per_step": kernel_ms_avg, "dtw_launches_per_step": dtw_launches,
                "cells_per_step_this_gpu": int(cells_local), "scatter_ms_per_step": float(np.mean(scat_ms)),
                "traffic": None,
                "hbm": {"algorithmic_bytes": int(arena_bytes + 2 * n * n * 4),
                        "achieved_gbs": (arena_bytes + 2 * n * n * 4) / (ms_per_step / 1e3) / 1e9,
                        "peak_gbs": peaks.get("hbm_gbs"), "note": "not the bound: < 1% of HBM peak"}},
        }
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
    al.close()
    if world > 1:
        dist.destroy_process_group()
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
    ap.add_argument("--workload", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--seqs", dest="n", type=int, default=None, help="override the number of sequences")
    ap.add_argument("--mode", default="strict", choices=["strict", "fast"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--other-mode-steps", type=int, default=2,
                    help="timed steps of the other arithmetic variant, reported beside the headline (0 = skip)")
    ap.add_argument("--ref-seconds", type=float, default=15.0)
    ap.add_argument("--ref-max-seqs", type=int, default=384)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
