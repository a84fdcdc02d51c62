"""Pins the CPU oracle: hand-derived KATs, the committed golden
vectors (tests/golden, produced by the independent Python transliteration), and
bit-for-bit agreement of the three restatements on seeded random data."""
import os

import numpy as np
import pytest

from oracle import literal, oracle

from .conftest import random_sequences

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dtw_golden.npz")
INF = np.float32(np.inf)


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def col(v):
    return np.array(v, dtype=np.float32).reshape(-1, 1)


@pytest.mark.parametrize("variant", ["literal", "dense"])
def test_known_answers(variant):
    big = 1.0  # band >= len
    kat = [([0, 1, 2], [0, 1, 2], 0.0), ([0, 0], [1, 1], 0.25), ([3], [7], 0.0), ([3], [7, 7], INF),
           ([1, 2, 3, 4], [1, 1, 2, 3, 5], 0.0), ([0, 1, 0, 0], [1, 0, 1, 0], 0.375)]
    for x, y, want in kat:
        got = oracle.dtw(col(x), col(y), big, variant=variant)
        assert bits(got) == bits(np.float32(want)), (x, y, got, want)
    # B7: band asymmetry with equal penalties at pct = 0 (band 0 => w = 2)
    x = col([2, 4, 1, 2, 3, 0, 1, 0, 2, 4, 0, 1])
    y = col([2, 4, 1, 2, 1, 0, 3, 0, 1, 2, 2, 0])
    assert oracle.dtw(x, y, 0.0, variant=variant) == np.float32(0.375)
    assert bits(oracle.dtw(y, x, 0.0, variant=variant)) == bits(np.float32(8.0) / np.float32(24.0))
    assert literal.dtw(x, y, 0.0) == np.float32(0.375)


def test_golden_pairs():
    g = np.load(GOLDEN)
    meta = g["meta"]
    for k in range(len(meta)):
        pct, ins, dele, mat, _ = meta[k]
        x, y = g["x%d" % k], g["y%d" % k]
        s_lit, path = oracle.dtw(x, y, pct, ins, dele, mat, variant="literal", want_path=True)
        s_den = oracle.dtw(x, y, pct, ins, dele, mat, variant="dense")
        want = g["score_bits%d" % k]
        assert bits(s_lit) == want[0], k
        assert bits(s_den) == want[0], k
        assert np.array_equal(path, g["path%d" % k]), k


@pytest.mark.parametrize("tag", ["A", "B"])
def test_golden_matrices(tag):
    g = np.load(GOLDEN)
    lens = g["all%s_lens" % tag]
    dim = int(g["all%s_dim" % tag][0])
    flat = g["all%s_flat" % tag]
    pct, ins, dele, mat = g["all%s_params" % tag]
    seqs, o = [], 0
    for t in lens:
        seqs.append(flat[o:o + t * dim].reshape(t, dim))
        o += t * dim
    want = g["all%s_matrix_bits" % tag]
    for variant in ("literal", "dense"):
        for workers in (1, 3, 4):
            got = oracle.align_all(seqs, pct, ins, dele, mat, workers=workers, variant=variant)
            assert np.array_equal(bits(got), want)
    assert np.all(np.diag(got) == 0.0)


def test_three_restatements_agree_bitwise():
    rng = np.random.default_rng(7)
    for trial in range(60):
        integer = trial % 2 == 0
        dim = int(rng.integers(1, 6))
        x, y = random_sequences(rng, 2, 1, 20, dim, integer)
        pct = float(rng.choice([0.0, 0.05, 0.2, 0.5, 1.0]))
        ins, dele, mat = [float(v) for v in rng.choice([1.0, 0.5, 0.75, 0.25], size=3)]
        a = oracle.dtw(x, y, pct, ins, dele, mat, variant="literal")
        b = oracle.dtw(x, y, pct, ins, dele, mat, variant="dense")
        c = literal.dtw(x, y, pct, ins, dele, mat)
        assert bits(a) == bits(b) == bits(c), (trial, a, b, c)


def test_dense_equals_literal_larger():
    rng = np.random.default_rng(8)
    for trial in range(12):
        dim = int(rng.choice([8, 10, 20, 26]))
        x, y = random_sequences(rng, 2, 60, 300, dim, integer=(trial % 3 == 0))
        pct = float(rng.choice([0.0, 0.05, 0.1, 1.0]))
        a = oracle.dtw(x, y, pct, 0.75, 0.5, 1.0, variant="literal")
        b = oracle.dtw(x, y, pct, 0.75, 0.5, 1.0, variant="dense")
        assert bits(a) == bits(b)


def test_transposed_band_symmetry():
    """D(y,x) with (ins, del) equals D(x,y) with (del, ins) evaluated on the transposed
    band only -- so in general the matrix is NOT symmetric."""
    rng = np.random.default_rng(9)
    asym = 0
    for _ in range(200):
        x, y = random_sequences(rng, 2, 4, 12, 1, integer=True)
        if oracle.dtw(x, y, 0.0) != oracle.dtw(y, x, 0.0):
            asym += 1
    assert asym > 0


def test_score_edge_cases():
    e = np.zeros((0, 3), dtype=np.float32)
    one = np.ones((1, 3), dtype=np.float32)
    two = np.ones((2, 3), dtype=np.float32)
    for v in ("literal", "dense"):
        assert oracle.dtw(e, e, 1.0, variant=v) == INF          # src/alignments.rs:117-118
        assert oracle.dtw(e, two, 1.0, variant=v) == INF
        assert oracle.dtw(two, e, 1.0, variant=v) == INF
        assert oracle.dtw(one, one * 5, 1.0, variant=v) == 0.0  # D[0,0] / 2
        assert oracle.dtw(one, two, 1.0, variant=v) == INF
        assert oracle.dtw(two, two, 1.0, variant=v) == 0.0
    x = np.array([[np.nan]], dtype=np.float32).repeat(3, 0)
    assert np.isnan(oracle.dtw(x, np.ones((3, 1), np.float32), 1.0))


def test_band_and_cells():
    assert oracle.warping_band(0.1, 512) == 51      # 0.1f32 * 512 = 51.2000008
    assert oracle.warping_band(0.05, 1024) == 51
    assert oracle.warping_band(1.0, 4096) == 4096
    assert oracle.warping_band(float("nan"), 10) == 0
    assert oracle.warping_band(-1.0, 10) == 0
    assert oracle.window(51, 512, 512) == 53
    # SURVEY.md Appendix C
    assert oracle.cells_visited(512, 512, 53) == 51463
    assert oracle.cells_visited(4096, 4096, 4098) == 16777216
    assert oracle.cells_visited(64, 256, 194) == 14431
    assert oracle.cells_visited(128, 1024, 898) == 123071
    for n, m, w in [(5, 9, 6), (9, 5, 6), (1, 1, 2), (7, 7, 2), (3, 20, 19)]:
        assert oracle.cells_visited(n, m, w) == literal.cells_visited(n, m, w)


def test_percentile_quirks():
    x = np.array([3, 1, np.nan, 2, 0], dtype=np.float32)
    # index from the unfiltered length (5 * 0.5 = 2), NaN dropped: sorted = [0,1,2,3] -> 2
    assert oracle.percentile(x, 0.5) == 2.0
    with pytest.raises(IndexError):
        oracle.percentile(np.array([1, 2], dtype=np.float32), 1.0)


def test_upgma_small():
    # two tight groups {0,1,2} and {3,4}; asymmetric entries are both read
    pts = np.array([0.0, 0.1, 0.2, 5.0, 5.1], dtype=np.float32)
    d = np.abs(pts[:, None] - pts[None, :]).astype(np.float32)
    merges, thr, assign = oracle.upgma(d, 0.5)
    assert merges[0][2] == 5
    assert {merges[0][0], merges[0][1]} in ({0, 1}, {1, 2}, {3, 4})
    assert len(set(assign[:3])) == 1 and len(set(assign[3:])) == 1
    assert len(merges) <= 4
