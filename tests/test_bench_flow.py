"""Drives bench.py's B200 arm with WORLD_SIZE = 2 on the CPU: gloo instead of NCCL and a stub in
place of ShardedAligner that issues the same collectives (one all-gather per align call).  It
checks the CONTROL FLOW only -- every rank must reach every collective (a rank-0-only collective
deadlocks, which is how a 113-GPU-minute run was once lost) and rank 0 must emit one JSON line.
Nothing here measures anything."""
import json
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeStream:
    cuda_stream = 0


class _FakeEvent:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self, stream=None):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class _StubCtx:
    def __init__(self, owner):
        self.o = owner

    def packed_len(self, pct, mode=0):
        return 64

    def synchronize(self, stream=0):
        pass


class StubAligner:
    """Same call surface and the same collective per call as ShardedAligner."""

    def __init__(self, seqs, device=None, group=None, mode=0):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.n = len(seqs)
        self.stream = _FakeStream()
        self.ctx = _StubCtx(self)
        self._matrix = torch.zeros((self.n, self.n))
        self._packed = torch.full((64,), float(self.rank))
        self._gathered = torch.empty(64 * self.world)

    def set_sequences(self, seqs):
        self.n = len(seqs)

    def align_all_device(self, pct, ins=1.0, dele=1.0, mat=1.0):
        if self.world > 1:
            dist.all_gather_into_tensor(self._gathered, self._packed)
        return self._matrix

    def align_all(self, pct, ins=1.0, dele=1.0, mat=1.0, out=None, to_host=True):
        m = self.align_all_device(pct, ins, dele, mat)
        if to_host and out is not None:
            out.copy_(m)
        return out

    def synchronize(self):
        pass

    def matrix_device(self):
        return self._matrix

    def percentile_of_matrix(self, perc):
        return np.float32(1.0)

    def stats(self):
        return {"cells_reference": 1000 * (self.rank + 1), "kernel_ms": 1.0, "scatter_ms": 0.1, "kernel_launches": 2,
                "sm_count": 148, "h2d_bytes": 10, "select_ms": 0.5, "units_local": 1, "units_total": 2,
                "cells_computed": 1100 * (self.rank + 1)}

    def close(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, workload="C3"):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import audio_pattern_discovery_b200.distributed as D
    D.ShardedAligner = StubAligner
    bench.DEVICE, bench.PIN = "cpu", False
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = _FakeEvent
    bench.emit = lambda line: q.put(json.dumps(line))
    sys.argv = ["bench.py", "--gpus", str(world), "--seqs", "24", "--steps", "2", "--warmup", "1", "--no-cpu",
                "--workload", workload]
    # bench.main() destroys the process group itself
    rc = bench.main()
    q.put("rank%d rc=%s" % (rank, rc))


@pytest.mark.parametrize("world,workload", [(2, "C3"), (3, "C3"), (2, "C2")])
def test_all_ranks_walk_the_same_collectives(world, workload):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, workload)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        if p.is_alive():
            for x in procs:
                x.terminate()
            pytest.fail("bench.py deadlocked with WORLD_SIZE=%d: a collective is not reached by every rank" % world)
        assert p.exitcode == 0
    msgs = [q.get(timeout=5) for _ in range(world + 1)]
    lines = [m for m in msgs if m.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == world and d["metric"] == "dtw_gcups" and d["scaling"] == "strong"
    assert d["reference_cells"] == 1000 * world * (world + 1) // 2   # summed over ranks
    assert d["e2e"]["value"] > 0 and d["other_mode"]["gcups"] > 0 and d["threshold_select"]["threshold"] == 1.0
    assert d["gpu_launches"] > 0
    assert len(d["matrix_checksum_u64"]) == 16 and d["parity_checked"] == 256 and d["e2e_cold"]["value"] > 0


class _StubGroupCtx:
    peer_stores = True

    def launch_plan(self):
        return [{"ring": "tmem", "ring_tiles": 31, "units": 1, "ctas": 1, "warps_per_cta": 4, "ctas_per_sm": 2, "smem_bytes": 0}]


class StubGroupAligner:
    """Call surface of distributed.GroupAligner (the one-process, many-GPU form)."""

    def __init__(self, seqs, devices="all", mode=0):
        self.n = len(seqs)
        self.ctx = _StubGroupCtx()
        self.world = len(devices)

    def set_sequences(self, seqs):
        self.n = len(seqs)

    def align_all(self, pct, ins=1.0, dele=1.0, mat=1.0, out=None):
        out[...] = 0.0
        return out

    def percentile_of_matrix(self, perc):
        return np.float32(1.0)

    def stats(self):
        return {"cells_reference": 5000, "cells_computed": 5500, "kernel_ms": 1.0, "scatter_ms": 0.1, "kernel_launches": 4,
                "sm_count": 148, "h2d_bytes": 10, "select_ms": 0.5, "units_local": 2, "units_total": 2}

    def close(self):
        pass


def test_single_process_arm_emits_the_contract_line(monkeypatch, capsys):
    """`bench.py --single-process --gpus N` (the in-library device group) on the CPU with a stub in place of the
    library: one JSON line with the contract keys plus the evidence keys."""
    import bench
    import audio_pattern_discovery_b200 as pkg
    import audio_pattern_discovery_b200.distributed as D
    lines = []
    monkeypatch.setattr(D, "GroupAligner", StubGroupAligner)
    monkeypatch.setattr(pkg, "visible_devices", lambda: 4)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(bench, "emit", lambda line: lines.append(line))
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": []})
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "4", "--single-process", "--seqs", "24", "--steps", "2",
                                      "--warmup", "1", "--no-cpu"])
    saved = os.dup(1)                       # bench.main() points fd 1 at stderr for the rest of the process
    try:
        assert bench.main() == 0
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    assert len(lines) == 1
    d = lines[0]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "e2e_cold", "gpu_launches", "roofline", "clocks",
                "matrix_checksum_u64", "parity_checked", "launch_form"):
        assert key in d, key
    assert d["n_gpus"] == 4 and d["gpu_launches"] > 0 and "apd_create_multi" in d["launch_form"]
    assert d["parity_ok"] is False          # the stub's all-zero matrix cannot match the oracle: the check is live
