"""The N > 1 path on real GPUs: one rank per GPU under torchrun, NCCL all-gather of the
packed shard results, bit-exact against the oracle.  Needs >= 2 GPUs (skipped otherwise;
the same data path is covered on one GPU by test_two_shards_on_one_device_* and on CPU by
tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_ranks_nccl_allgather_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "multi_rank_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok=True" in r.stdout
