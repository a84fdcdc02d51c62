"""World-size-2 coverage of the N>1 host logic on CPU (gloo): each rank computes its
shard of the work units (u % world == rank) with the schedule emulator, the shards
are exchanged with an all-gather, and the assembled matrix equals the oracle's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle

from . import emul
from .conftest import random_sequences


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(21)
        seqs = random_sequences(rng, 50, 3, 30, 4, False)
        part, info = emul.align_all(seqs, 0.2, 0.75, 0.5, 1.0, rank=rank, world=world)
        mine = torch.from_numpy(part.ravel().copy())
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        full = torch.stack(gathered).sum(0).numpy().reshape(part.shape)
        cells = torch.tensor([int(info[2])], dtype=torch.int64)
        dist.all_reduce(cells)
        # max-over-ranks reduction used for the timing in bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            want = oracle.align_all(seqs, 0.2, 0.75, 0.5, 1.0, variant="dense")
            ref_cells = sum(oracle.pair_cells(len(a), len(b), 0.2)
                            for i, a in enumerate(seqs) for j, b in enumerate(seqs) if i != j)
            ok = (np.array_equal(full.view(np.uint32), want.view(np.uint32))
                  and int(cells[0]) == ref_cells and float(t[0]) == world)
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_assembly():
    emul.build()
    oracle.build()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
