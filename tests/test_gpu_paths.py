"""K2 (csrc/pair_path.cu): scores and device-traced warping paths of requested ordered pairs
against the oracle's literal restatement (which records the branch taken at every cell,
SURVEY.md Appendix A.8).  Lengths sit on and around the kernel's 128-row slab and 4-column tile
boundaries; bands from 0 % to unbanded; weighted and unit penalties; every padded frame width."""
import numpy as np
import pytest

from oracle import oracle

from .conftest import random_sequences

pytestmark = pytest.mark.gpu


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def _check(ctx, seqs, pairs, pct, w3, band=None, path_cap=None):
    cap = path_cap or (max(len(s) for s in seqs) * 2 + 2)
    scores, paths, lens = ctx.align_pairs(pairs, pct, *w3, want_paths=True, path_cap=cap, warping_band=band)
    for (i, j), s, p, ln in zip(pairs, scores, paths, lens):
        ws, wp = oracle.dtw(seqs[i], seqs[j], pct, *w3, want_path=True, band=band)
        assert bits(s) == bits(ws), (i, j, len(seqs[i]), len(seqs[j]), s, ws)
        assert ln == len(wp), (i, j, ln, len(wp))
        assert np.array_equal(p, wp[:cap]), (i, j)


@pytest.mark.parametrize("dim,pct,w3,integer", [
    (1, 0.0, (1.0, 1.0, 1.0), True),
    (3, 0.25, (0.75, 0.5, 1.0), True),
    (20, 0.1, (1.0, 1.0, 1.0), False),
    (26, 1.0, (0.5, 1.0, 0.25), False),
    (8, 0.05, (1.0, 1.0, 1.0), False),
])
def test_slab_and_tile_boundaries(dim, pct, w3, integer):
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(dim * 7 + 1)
    lens = [1, 2, 3, 4, 5, 6, 9, 127, 128, 129, 130, 131, 133, 255, 256, 257, 258, 261, 300]
    seqs = []
    for t in lens:
        seqs.append(rng.integers(0, 3, size=(t, dim)).astype(np.float32) if integer else rng.normal(size=(t, dim)).astype(np.float32))
    pairs = [(a, b) for a in range(len(lens)) for b in range(len(lens)) if a != b and (a * 31 + b * 17) % 4 == 0]
    with Context(0) as ctx:
        ctx.set_sequences(seqs)
        _check(ctx, seqs, pairs, pct, w3)
        # the caller's own AlignmentParams.warping_band (src/alignments.rs:79) instead of the percentage
        _check(ctx, seqs, pairs[:25], 0.0, w3, band=7)
        # FAST arithmetic: same path rule, scores within 1e-5 of the oracle (ties aside, real-valued data only)
        if not integer:
            from audio_pattern_discovery_b200 import APD_MODE_FAST
            got = ctx.align_pairs(pairs, pct, *w3, mode=APD_MODE_FAST)
            want = np.array([oracle.dtw(seqs[i], seqs[j], pct, *w3, variant="dense") for i, j in pairs])
            fin = np.isfinite(want) & (want > 0)
            assert np.all(np.isinf(got[~np.isfinite(want)]))
            assert np.max(np.abs(got[fin] - want[fin]) / want[fin]) <= 1e-5


def test_truncated_path_buffer_reports_full_length():
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(3)
    seqs = random_sequences(rng, 6, 40, 90, 5)
    with Context(0) as ctx:
        ctx.set_sequences(seqs)
        _check(ctx, seqs, [(0, 1), (2, 5), (4, 3)], 0.3, (1.0, 1.0, 1.0), path_cap=10)


def test_pairs_match_the_matrix_kernel():
    """K1 (all pairs, both orientations per lane) and K2 (one ordered pair per warp) are different
    programs; in STRICT mode they must produce the same bits."""
    from audio_pattern_discovery_b200 import Context, synth
    c, seqs, _ = synth.make_config("C4", 40)
    with Context(0) as ctx:
        ctx.set_sequences(seqs)
        mat = ctx.align_all(c["pct"])
        pairs = [(i, j) for i in range(40) for j in range(40) if i != j]
        got = ctx.align_pairs(pairs, c["pct"])
    want = np.array([mat[i, j] for i, j in pairs], dtype=np.float32)
    assert np.array_equal(bits(got), bits(want))


def test_long_pair_beyond_the_old_shared_memory_limit():
    """Round 1's kernel kept three anti-diagonals in shared memory (<= ~19 000 frames).  30 000 x 300,
    window = |n - m| + 2 (the whole matrix), score against the dense oracle, path checked structurally."""
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(11)
    x = np.cumsum(rng.normal(size=(30000, 4)), axis=0).astype(np.float32) * 0.05
    y = x[::100] + rng.normal(size=(300, 4)).astype(np.float32) * 0.01
    with Context(0) as ctx:
        ctx.set_sequences([x, y])
        for pr in ((0, 1), (1, 0)):
            scores, paths, lens = ctx.align_pairs([pr], 0.0, want_paths=True, path_cap=31000)
            want = oracle.dtw([x, y][pr[0]], [x, y][pr[1]], 0.0, variant="dense")
            assert bits(scores[0]) == bits(want)
            p = paths[0].astype(np.int64)
            n, m = (30000, 300) if pr == (0, 1) else (300, 30000)
            assert lens[0] == len(p) and tuple(p[0]) == (n - 1, m - 1) and tuple(p[-1])[0] >= 1 and tuple(p[-1])[1] >= 1
            step = p[:-1] - p[1:]
            assert set(map(tuple, step.tolist())) <= {(1, 0), (0, 1), (1, 1)}
            assert p[-1, 0] == 1 or p[-1, 1] == 1


def test_many_pairs_in_one_call():
    """Hundreds of requests (persistent warps pulling jobs, most expensive first), trivial pairs mixed in."""
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(21)
    seqs = random_sequences(rng, 60, 1, 150, 10)
    seqs[7] = seqs[7][:1]
    seqs[9] = np.zeros((0, 10), np.float32)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 60, size=(700, 2)) if a != b]
    with Context(0) as ctx:
        ctx.set_sequences(seqs)
        scores, paths, lens = ctx.align_pairs(pairs, 0.15, 0.75, 0.5, 1.0, want_paths=True, path_cap=320)
        st = ctx.stats()
    assert st["path_ms"] > 0
    for (i, j), s, p, ln in zip(pairs, scores, paths, lens):
        if len(seqs[i]) == 0 or len(seqs[j]) == 0:
            assert np.isinf(s) and ln == 0
            continue
        ws, wp = oracle.dtw(seqs[i], seqs[j], 0.15, 0.75, 0.5, 1.0, want_path=True)
        assert bits(s) == bits(ws) and np.array_equal(p, wp), (i, j)
