"""Run under torchrun (one rank per GPU): shards the pair space, all-gathers the packed
results over NCCL, scatters them, and rank 0 compares the assembled matrix with the CPU
oracle bit for bit.  Launched by tests/test_gpu_multi.py; exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_pattern_discovery_b200 import synth  # noqa: E402
from audio_pattern_discovery_b200.distributed import ShardedAligner  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    rng = np.random.default_rng(5)
    seqs, _ = synth.make_sequences(150, rng.integers(30, 140, size=150), 20, 8, 31)
    al = ShardedAligner(seqs, device=local)
    out = al.align_all(0.1, 0.75, 0.5, 1.0, to_host=True)
    st = al.stats()
    cells = torch.tensor([st["cells_reference"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(cells)
    ok = True
    if rank == 0:
        from oracle import oracle
        want = oracle.align_all(seqs, 0.1, 0.75, 0.5, 1.0, workers=8, variant="dense")
        got = out.numpy()
        ok = np.array_equal(got.view(np.uint32), want.view(np.uint32))
        ref_cells = sum(oracle.pair_cells(len(a), len(b), 0.1)
                        for i, a in enumerate(seqs) for j, b in enumerate(seqs) if i != j)
        ok = ok and int(cells[0]) == ref_cells and st["units_local"] < st["units_total"]
        print("multi_rank_check world=%d ok=%s units_local=%d units_total=%d" % (world, ok, st["units_local"], st["units_total"]))
    al.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
