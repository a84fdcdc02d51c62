"""Host-side mirror of the reference interface (no GPU needed)."""
import os

import numpy as np
import pytest

from audio_pattern_discovery_b200 import AlignmentParams, Discovery, NDSequence
from oracle import oracle


def test_alignment_params_band_is_f32_product_truncated():
    d = Discovery(warping_band_percentage=0.1)
    assert d.alignment_params(512).warping_band == 51
    assert Discovery(warping_band_percentage=0.05).alignment_params(1024).warping_band == 51
    assert Discovery(warping_band_percentage=1.0).alignment_params(4096).warping_band == 4096
    assert Discovery(warping_band_percentage=float("nan")).alignment_params(9).warping_band == 0
    rng = np.random.default_rng(0)
    for _ in range(500):
        pct = float(np.float32(rng.uniform(0, 1.2)))
        n = int(rng.integers(1, 5000))
        assert Discovery(warping_band_percentage=pct).alignment_params(n).warping_band == oracle.warping_band(pct, n)
    p = d.alignment_params(10)
    assert (p.insertion_penalty, p.deletion_penalty, p.match_penalty) == (1.0, 1.0, 1.0)
    assert AlignmentParams.default(7).warping_band == 7


def test_discovery_from_reference_style_toml(tmp_path):
    text = """
dft_win = 256
dft_step = 128
ceps_filter = 32
auto_encoder = 10
learning_rate = 0.1
epochs = 25
epoch_drop = 5.0
drop = 0.5
vat_moving = 15
vat_percentile = 0.95
vat_min_len = 150
warping_band_percentage = 0.1
insertion_penalty = 0.75
deletion_penalty = 0.5
match_penalty = 1.0
alignment_workers = 4
clustering_percentile = 0.05
"""
    f = tmp_path / "Discovery.toml"
    f.write_text(text)
    d = Discovery.from_toml(str(f))
    assert d.alignment_workers == 4 and d.insertion_penalty == 0.75 and d.clustering_percentile == 0.05
    f.write_text(text.replace("match_penalty = 1.0\n", ""))
    with pytest.raises(KeyError):
        Discovery.from_toml(str(f))


def test_ndsequence_layout_contract():
    a = np.arange(12, dtype=np.float32)
    s = NDSequence(4, a)
    assert s.len() == 3 and len(s) == 3
    assert np.array_equal(s.vec(1), [4, 5, 6, 7])
    assert NDSequence(5, a).len() == 2  # ragged tail ignored (integer division)
    assert NDSequence.from_array(a.reshape(3, 4)).n_bins == 4


def test_synthetic_workloads_are_deterministic_and_shaped():
    """bench.py and the parity tests regenerate the BASELINE.json workloads from seeds on every box."""
    from audio_pattern_discovery_b200 import synth
    for name, n, dim in (("C2", 50, 20), ("C3", 20, 20), ("C4", 40, 8), ("C5", 2, 20)):
        c1, a, la = synth.make_config(name, n)
        c2, b, lb = synth.make_config(name, n)
        assert len(a) == n and all(x.dtype == np.float32 and x.shape[1] == dim for x in a)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)) and np.array_equal(la, lb)
        assert c1["pct"] == c2["pct"]
    c, seqs, _ = synth.make_config("C2", 300)
    lens = np.array([len(s) for s in seqs])
    assert lens.min() >= 64 and lens.max() <= 256 and c["weights"] == (0.75, 0.5, 1.0)
    c, seqs, _ = synth.make_config("C4", 300)
    lens = np.array([len(s) for s in seqs])
    assert lens.min() >= 97 and lens.max() <= 1024 and c["pct"] == 0.05
    assert all(len(s) == 512 for s in synth.make_config("C3", 5)[1])
    assert all(len(s) == 4096 for s in synth.make_config("C5", 2)[1])


def test_unit_list_is_the_same_set_for_every_enumeration_block():
    """The planner enumerates units in blocks of 32 x sharers rows (L2 locality per device); the block size
    may only permute the list inside a class: same units, same classes, and row_block consecutive units of
    one cost bucket share their column block."""
    from .emul import plan_units
    rng = np.random.default_rng(12)
    lens = np.concatenate([np.full(700, 96), rng.integers(1, 300, size=200)])
    base, info0 = plan_units(lens, 20, 0.1, 32)
    key = lambda u: set(map(tuple, u.tolist()))
    assert len(key(base)) == len(base)                      # no duplicates
    for rb in (64, 256, 512):
        u, info = plan_units(lens, 20, 0.1, rb)
        assert key(u) == key(base) and np.array_equal(info, info0)
    # equal-length sequences: one bucket, so the enumeration order is visible directly
    u, _ = plan_units(np.full(2000, 128), 20, 0.1, 128)
    runs = np.flatnonzero(np.diff(u[:, 1]) != 0)
    assert np.median(np.diff(runs)) == 128                  # 128 consecutive units per column block
