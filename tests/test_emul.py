"""Runs the kernel's per-lane program (csrc/dtw_core.h) and the host planner
(csrc/host_plan.cpp) through the host-side schedule emulator and compares the
full matrix with the oracle bit for bit.  No GPU involved; the emulator is test
code, not a fallback (nothing in the package loads it)."""
import numpy as np
import pytest

from oracle import oracle

from . import emul
from .conftest import random_sequences


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


CASES = [
    # (n_seq, len_lo, len_hi, dim, pct, (ins, del, mat), integer)
    (9, 1, 12, 1, 0.0, (1.0, 1.0, 1.0), True),
    (12, 1, 30, 3, 0.1, (1.0, 1.0, 1.0), True),
    (12, 2, 40, 2, 0.25, (0.75, 0.5, 1.0), True),
    (40, 5, 60, 8, 0.1, (1.0, 1.0, 1.0), False),
    (35, 20, 90, 10, 0.05, (0.75, 0.5, 1.0), False),
    (20, 30, 70, 20, 1.0, (1.0, 1.0, 1.0), False),
    (10, 50, 130, 26, 0.1, (0.5, 1.0, 0.25), False),
    (70, 1, 20, 4, 0.3, (1.0, 1.0, 1.0), True),   # more than two 32-lane blocks
]


@pytest.mark.parametrize("case", CASES)
def test_emulated_schedule_matches_oracle_bitwise(case):
    n, lo, hi, dim, pct, (ins, dele, mat), integer = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    seqs = random_sequences(rng, n, lo, hi, dim, integer)
    want = oracle.align_all(seqs, pct, ins, dele, mat, workers=4, variant="dense")
    got, info = emul.align_all(seqs, pct, ins, dele, mat, strict=True)
    assert np.array_equal(bits(got), bits(want))
    ref_cells = sum(oracle.pair_cells(len(a), len(b), pct)
                    for i, a in enumerate(seqs) for j, b in enumerate(seqs) if i != j and len(a) and len(b))
    assert int(info[2]) == ref_cells


def test_cold_exact_path_schedule_matches_oracle():
    """strict=2 forces every unit through run_unit_exact (the generic-sqrt re-run path)."""
    rng = np.random.default_rng(31)
    for dim, pct, integer in ((3, 0.2, True), (20, 0.1, False), (8, 1.0, False)):
        seqs = random_sequences(rng, 37, 1, 45, dim, integer)
        want = oracle.align_all(seqs, pct, 0.75, 0.5, 1.0, variant="dense")
        got, _ = emul.align_all(seqs, pct, 0.75, 0.5, 1.0, strict=2)
        assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("dim", [5, 13, 16, 24, 30, 32])
def test_every_padded_frame_width(dim):
    """One kernel instantiation per padded width 4..32: the quarter-points at which the
    recurrence cells are interleaved with the distance work differ per width."""
    rng = np.random.default_rng(100 + dim)
    seqs = random_sequences(rng, 14, 6, 40, dim, integer=(dim % 2 == 1))
    want = oracle.align_all(seqs, 0.2, 0.75, 0.5, 1.0, variant="dense")
    got, _ = emul.align_all(seqs, 0.2, 0.75, 0.5, 1.0)
    assert np.array_equal(bits(got), bits(want))
    unit, _ = emul.align_all(seqs, 0.2)
    assert np.array_equal(bits(unit), bits(oracle.align_all(seqs, 0.2, variant="dense")))


def test_fast_mode_within_tolerance():
    rng = np.random.default_rng(3)
    seqs = random_sequences(rng, 20, 30, 80, 20, False)
    want = oracle.align_all(seqs, 0.1, 0.75, 0.5, 1.0, variant="dense")
    got, _ = emul.align_all(seqs, 0.1, 0.75, 0.5, 1.0, strict=False)
    off = ~np.eye(len(seqs), dtype=bool)
    rel = np.abs(got[off] - want[off]) / np.abs(want[off])
    assert rel.max() <= 1e-5  # BASELINE.json north_star tolerance


def test_empty_and_single_frame_sequences():
    rng = np.random.default_rng(4)
    seqs = random_sequences(rng, 6, 3, 9, 2, True)
    seqs[1] = np.zeros((0, 2), dtype=np.float32)
    seqs[3] = np.ones((1, 2), dtype=np.float32)
    seqs[4] = np.ones((1, 2), dtype=np.float32) * 2
    want = oracle.align_all(seqs, 0.5, variant="literal")
    got, _ = emul.align_all(seqs, 0.5)
    assert np.array_equal(bits(got), bits(want))
    assert np.isinf(got[1, 0]) and got[3, 4] == 0.0 and np.isinf(got[3, 0])


def test_wide_band_uses_global_ring_class():
    rng = np.random.default_rng(5)
    seqs = random_sequences(rng, 5, 250, 300, 4, False)
    got, info = emul.align_all(seqs, 1.0)
    want = oracle.align_all(seqs, 1.0, variant="dense")
    assert np.array_equal(bits(got), bits(want))
    assert any(int(info[8 + c]) == 1 for c in range(int(info[1])))  # a gstate class exists


@pytest.mark.parametrize("world", [2, 3, 8])
def test_shards_partition_the_matrix(world):
    rng = np.random.default_rng(6)
    seqs = random_sequences(rng, 45, 4, 25, 3, True)
    full, info = emul.align_all(seqs, 0.2)
    acc = np.zeros_like(full)
    seen = np.zeros(full.shape, dtype=np.int32)
    cells = 0
    for r in range(world):
        part, pinfo = emul.align_all(seqs, 0.2, rank=r, world=world)
        seen += (part != 0)
        acc += part
        cells += int(pinfo[2])
    assert np.array_equal(bits(acc), bits(full))
    assert seen.max() <= 1
    assert cells == int(info[2])


def test_cells_visited_closed_form():
    L = emul.lib()
    for n, m, w in [(512, 512, 53), (4096, 4096, 4098), (64, 256, 194), (128, 1024, 898), (1, 1, 2),
                    (5, 9, 6), (9, 5, 6), (7, 7, 2), (3, 20, 19), (20, 3, 19), (2, 1, 3)]:
        assert L.apd_emul_cells_visited(n, m, w) == oracle.cells_visited(n, m, w)
    for pct, n, m in [(0.1, 512, 512), (0.05, 1024, 128), (1.0, 4096, 4096), (0.0, 9, 4), (float("nan"), 5, 5)]:
        want = oracle.window(oracle.warping_band(pct, max(n, m)), n, m)
        assert L.apd_emul_window(pct, n, m) == min(want, n + m + 8)


@pytest.mark.parametrize("name", ["C2", "C3", "C4", "C5"])
def test_planner_at_full_baseline_sizes(name):
    """The host planner on the full BASELINE.json shapes (lengths only, no DP): every unordered pair
    in exactly one unit, the reference cell count equals the closed form summed in Python, the
    shards of 8 ranks add up, and the launch classes are the expected ring homes."""
    import time

    import bench
    from audio_pattern_discovery_b200 import synth
    c = synth.config(name)
    lens = np.full(c["n"], c["lens"]) if np.isscalar(c["lens"]) else np.asarray(c["lens"])
    t0 = time.perf_counter()
    info = emul.plan_info(lens, c["dim"], c["pct"])
    assert time.perf_counter() - t0 < 30
    vals, counts = np.unique(lens, return_counts=True)
    want = 0
    for a, ca in zip(vals, counts):
        for b, cb in zip(vals, counts):
            pairs = int(ca) * int(cb) - (int(ca) if a == b else 0)
            w = oracle.window(oracle.warping_band(c["pct"], int(max(a, b))), int(a), int(b))
            want += pairs * bench.cells_visited(int(a), int(b), w)
    assert int(info[2]) == want
    assert sum(int(emul.plan_info(lens, c["dim"], c["pct"], r, 8)[2]) for r in range(8)) == want
    if name == "C3":
        assert int(info[1]) == 1 and int(info[4]) <= 32 and int(info[8]) == 0   # one class, ring fits TMEM
        assert want == 10000 * 9999 * 51463
    if name == "C5":
        assert int(info[8]) == 1                                                 # unbanded 4096: global ring


@pytest.mark.parametrize("rho", [0, 1, 2, 3])
def test_every_row_grid_gives_the_same_bits(rho):
    """The row tiles may start at any of four offsets (dtw_core.h: row_geometry(n, rho)); the rows
    below n-1 in the last tile are computed from the zero frames behind the sequence and must
    never reach the score.  Every grid, hot and cold path, tie-heavy and real-valued data."""
    rng = np.random.default_rng(500 + rho)
    try:
        emul.force_rho(rho)
        for dim, pct, integer, w3 in ((2, 0.0, True, (1.0, 1.0, 1.0)), (3, 0.2, True, (0.75, 0.5, 1.0)),
                                      (20, 0.1, False, (1.0, 1.0, 1.0)), (8, 1.0, False, (0.5, 1.0, 0.25))):
            seqs = random_sequences(rng, 36, 1, 50, dim, integer)
            want = oracle.align_all(seqs, pct, *w3, variant="dense")
            for strict in (1, 2):
                got, _ = emul.align_all(seqs, pct, *w3, strict=strict)
                assert np.array_equal(bits(got), bits(want)), (rho, dim, pct, strict)
        used = emul.force_rho(-1)
        assert used[rho] > 0 and used.sum() == used[rho]
    finally:
        emul.force_rho(-1)


def test_chosen_row_grid_needs_fewer_tiles_on_the_headline_shape():
    """C3's geometry (len 512, band 10 % -> w = 53): a 4-column block needs 110 rows = 27.5 tiles.
    The end-anchored grid (rho = 0) spends 29 tiles per block, the grid the kernel picks 28."""
    rng = np.random.default_rng(77)
    seqs = [rng.normal(size=(512, 4)).astype(np.float32) for _ in range(3)]
    want = oracle.align_all(seqs, 0.1, variant="dense")
    try:
        emul.force_rho(0)
        got0, info0 = emul.align_all(seqs, 0.1)
        emul.force_rho(-1)
        got, info = emul.align_all(seqs, 0.1)
        used = emul.force_rho(-1)
    finally:
        emul.force_rho(-1)
    assert np.array_equal(bits(got0), bits(want)) and np.array_equal(bits(got), bits(want))
    assert used[0] == 0 and used.sum() > 0              # the anchored grid is not the one chosen here
    assert int(info[3]) < 0.975 * int(info0[3])         # >= 2.5 % fewer lane-tiles


@pytest.mark.parametrize("case", CASES)
def test_two_column_tiles_match_oracle_bitwise(case):
    """The 12-warps-per-SM kernel runs the same lane program on 4 x 2-column tiles."""
    n, lo, hi, dim, pct, (ins, dele, mat), integer = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    seqs = random_sequences(rng, n, lo, hi, dim, integer)
    want = oracle.align_all(seqs, pct, ins, dele, mat, workers=4, variant="dense")
    try:
        emul.tile_cols(2)
        for strict in (1, 2):
            got, info = emul.align_all(seqs, pct, ins, dele, mat, strict=strict)
            assert np.array_equal(bits(got), bits(want)), strict
        for rho in (0, 1, 2, 3):
            emul.force_rho(rho)
            got, _ = emul.align_all(seqs, pct, ins, dele, mat, strict=True)
            assert np.array_equal(bits(got), bits(want)), rho
    finally:
        emul.force_rho(-1)
        emul.tile_cols(4)


@pytest.mark.parametrize("dim", [5, 13, 16, 24, 30, 32])
def test_two_column_tiles_every_padded_frame_width(dim):
    rng = np.random.default_rng(300 + dim)
    seqs = random_sequences(rng, 14, 6, 40, dim, integer=(dim % 2 == 1))
    want = oracle.align_all(seqs, 0.2, 0.75, 0.5, 1.0, variant="dense")
    try:
        emul.tile_cols(2)
        got, _ = emul.align_all(seqs, 0.2, 0.75, 0.5, 1.0, strict=True)
        gotf, _ = emul.align_all(seqs, 0.2, 0.75, 0.5, 1.0, strict=False)
    finally:
        emul.tile_cols(4)
    assert np.array_equal(bits(got), bits(want))
    off = ~np.eye(len(seqs), dtype=bool)
    assert np.max(np.abs(gotf[off] - want[off]) / np.abs(want[off])) <= 1e-5
