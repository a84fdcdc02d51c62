"""Parity of the CUDA path (through the C ABI / the Python mirror of the reference
interface) against the CPU oracle.  STRICT mode must be bit-exact; FAST mode must
be within the 1e-5 relative tolerance BASELINE.json's north_star states."""
import os

import numpy as np
import pytest

from oracle import oracle

from .conftest import random_sequences
from .test_emul import CASES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dtw_golden.npz")
REL_TOL = 1e-5


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def gpu_matrix(seqs, pct, ins=1.0, dele=1.0, mat=1.0, fast=False):
    from audio_pattern_discovery_b200 import APD_MODE_FAST, APD_MODE_STRICT, AlignmentWorkers, Discovery
    w = AlignmentWorkers.new(seqs, mode=APD_MODE_FAST if fast else APD_MODE_STRICT)
    d = Discovery(warping_band_percentage=pct, insertion_penalty=ins, deletion_penalty=dele, match_penalty=mat)
    w.align_all(d)
    n = len(seqs)
    return w.result.lock().unwrap().reshape(n, n).copy(), w


def assert_close(got, want):
    assert np.array_equal(np.isinf(got), np.isinf(want))
    f = np.isfinite(want) & (want != 0)
    rel = np.abs(got[f] - want[f]) / np.abs(want[f])
    assert rel.size == 0 or rel.max() <= REL_TOL, rel.max()
    assert np.all(got[want == 0] == 0)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_golden_matrices_bitwise(tag):
    g = np.load(GOLDEN)
    lens = g["all%s_lens" % tag]
    dim = int(g["all%s_dim" % tag][0])
    flat = g["all%s_flat" % tag]
    pct, ins, dele, mat = g["all%s_params" % tag]
    seqs, o = [], 0
    for t in lens:
        seqs.append(flat[o:o + t * dim].reshape(t, dim))
        o += t * dim
    got, _ = gpu_matrix(seqs, pct, ins, dele, mat)
    assert np.array_equal(bits(got), g["all%s_matrix_bits" % tag])


def test_golden_pairs_scores_and_paths():
    from audio_pattern_discovery_b200 import Alignment, AlignmentParams
    g = np.load(GOLDEN)
    meta = g["meta"]
    for k in range(len(meta)):
        pct, ins, dele, mat, _ = meta[k]
        x, y = g["x%d" % k], g["y%d" % k]
        band = oracle.warping_band(pct, max(len(x), len(y)))
        a = Alignment.new()
        a.construct_alignment(x, y, AlignmentParams(band, ins, dele, mat))
        assert bits(a.score())[0] == g["score_bits%d" % k][0], k
        assert np.array_equal(a.path, g["path%d" % k]), k


@pytest.mark.parametrize("case", CASES)
def test_random_cases_bitwise(case):
    n, lo, hi, dim, pct, (ins, dele, mat), integer = case
    rng = np.random.default_rng(hash(case) & 0xffff)
    seqs = random_sequences(rng, n, lo, hi, dim, integer)
    want = oracle.align_all(seqs, pct, ins, dele, mat, workers=8, variant="dense")
    got, w = gpu_matrix(seqs, pct, ins, dele, mat)
    assert np.array_equal(bits(got), bits(want))
    st = w.stats()
    ref_cells = sum(oracle.pair_cells(len(a), len(b), pct)
                    for i, a in enumerate(seqs) for j, b in enumerate(seqs) if i != j and len(a) and len(b))
    assert st["cells_reference"] == ref_cells
    assert st["kernel_launches"] >= 2  # DTW + scatter kernels really ran
    if not integer:
        fast, _ = gpu_matrix(seqs, pct, ins, dele, mat, fast=True)
        assert_close(fast, want)


def test_empty_single_and_duplicate_sequences():
    rng = np.random.default_rng(4)
    seqs = random_sequences(rng, 8, 3, 9, 2, True)
    seqs[1] = np.zeros((0, 2), dtype=np.float32)
    seqs[3] = np.ones((1, 2), dtype=np.float32)
    seqs[4] = np.ones((1, 2), dtype=np.float32) * 2
    seqs[7] = seqs[6].copy()
    want = oracle.align_all(seqs, 0.5, variant="literal")
    got, _ = gpu_matrix(seqs, 0.5)
    assert np.array_equal(bits(got), bits(want))
    assert np.isinf(got[1, 0]) and got[3, 4] == 0.0 and np.isinf(got[3, 0]) and got[6, 7] == 0.0


def test_zero_and_one_sequence():
    from audio_pattern_discovery_b200 import AlignmentWorkers, Discovery
    for seqs in ([], [np.ones((5, 3), np.float32)]):
        w = AlignmentWorkers.new(seqs)
        w.align_all(Discovery())
        assert np.all(w.result.lock().unwrap() == 0)


def test_nan_and_inf_frames_propagate_like_the_reference():
    rng = np.random.default_rng(11)
    seqs = random_sequences(rng, 5, 6, 12, 3, False)
    seqs[2][3, 1] = np.nan
    seqs[4][0, 0] = np.inf
    want = oracle.align_all(seqs, 1.0, variant="dense")
    got, _ = gpu_matrix(seqs, 1.0)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    f = ~np.isnan(want)
    assert np.array_equal(bits(got[f]), bits(want[f]))


def test_c2_shaped_subset_both_weightings():
    from audio_pattern_discovery_b200 import synth
    c, seqs, _ = synth.make_config("C2", n=160)
    for w in (c["weights"], (1.0, 1.0, 1.0)):
        want = oracle.align_all(seqs, c["pct"], *w, workers=8, variant="dense")
        got, _ = gpu_matrix(seqs, c["pct"], *w)
        assert np.array_equal(bits(got), bits(want))
        fast, _ = gpu_matrix(seqs, c["pct"], *w, fast=True)
        assert_close(fast, want)


def test_c4_shaped_subset_dim8_variable_buckets():
    from audio_pattern_discovery_b200 import synth
    c, seqs, _ = synth.make_config("C4", n=64)
    want = oracle.align_all(seqs, c["pct"], *c["weights"], workers=8, variant="dense")
    got, _ = gpu_matrix(seqs, c["pct"], *c["weights"])
    assert np.array_equal(bits(got), bits(want))


def test_unbanded_long_pairs_use_global_ring():
    from audio_pattern_discovery_b200 import synth
    seqs, _ = synth.make_sequences(6, 700, 20, 3, 55)
    want = oracle.align_all(seqs, 1.0, workers=8, variant="dense")
    got, _ = gpu_matrix(seqs, 1.0)
    assert np.array_equal(bits(got), bits(want))


def test_forced_global_ring_matches(monkeypatch):
    rng = np.random.default_rng(12)
    seqs = random_sequences(rng, 40, 10, 60, 10, False)
    want = oracle.align_all(seqs, 0.1, workers=8, variant="dense")
    monkeypatch.setenv("APD_FORCE_GSTATE", "1")
    got, _ = gpu_matrix(seqs, 0.1)
    assert np.array_equal(bits(got), bits(want))


def test_flat_and_pointer_packing_agree():
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(13)
    seqs = random_sequences(rng, 33, 1, 40, 10, False)
    flat = np.concatenate([s.ravel() for s in seqs])
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    offs = np.concatenate([[0], np.cumsum(lens[:-1].astype(np.uint64) * 10)]).astype(np.uint64)
    with Context(0) as a, Context(0) as b:
        a.set_sequences(seqs)
        b.set_sequences_flat(flat, offs, lens, 10)
        ma = a.align_all(0.2)
        mb = b.align_all(0.2)
    assert np.array_equal(bits(ma), bits(mb))


def test_two_shards_on_one_device_assemble_the_full_matrix():
    """The N>1 data path (packed shards -> gathered buffer -> scatter) without NCCL."""
    import torch
    from audio_pattern_discovery_b200 import APD_MODE_STRICT, Context
    rng = np.random.default_rng(14)
    seqs = random_sequences(rng, 75, 5, 40, 8, False)
    want = oracle.align_all(seqs, 0.15, 0.75, 0.5, 1.0, workers=8, variant="dense")
    world = 2
    ctxs = [Context(0) for _ in range(world)]
    packed = []
    for r, c in enumerate(ctxs):
        c.set_sequences(seqs)
        c.set_shard(r, world)
        k = c.packed_len(0.15)
        t = torch.empty(k, dtype=torch.float32, device="cuda")
        c.align_packed(0.15, 0.75, 0.5, 1.0, APD_MODE_STRICT, t.data_ptr(), 0)
        c.synchronize(0)
        packed.append(t)
    gathered = torch.cat(packed)
    out = torch.empty((len(seqs), len(seqs)), dtype=torch.float32, device="cuda")
    ctxs[1].scatter_packed(gathered.data_ptr(), world, out.data_ptr(), 0)
    ctxs[1].synchronize(0)
    assert np.array_equal(bits(out.cpu().numpy()), bits(want))
    for c in ctxs:
        c.close()


def test_paths_on_device_match_oracle():
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(15)
    seqs = random_sequences(rng, 12, 20, 90, 10, False) + random_sequences(rng, 6, 5, 30, 2, True)
    seqs = [s if s.shape[1] == 10 else np.pad(s, ((0, 0), (0, 8))) for s in seqs]
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, len(seqs), size=(40, 2)) if a != b]
    with Context(0) as c:
        c.set_sequences(seqs)
        scores, paths, lens = c.align_pairs(pairs, 0.2, 0.75, 0.5, 1.0, want_paths=True, path_cap=200)
    for (i, j), s, p, ln in zip(pairs, scores, paths, lens):
        ws, wp = oracle.dtw(seqs[i], seqs[j], 0.2, 0.75, 0.5, 1.0, want_path=True)
        assert bits(s)[0] == bits(ws)[0]
        assert ln == len(wp) and np.array_equal(p, wp)


def test_full_c2_spot_check_and_properties():
    """BASELINE.json config 2 at full size (2 000 sequences, ~4e6 ordered pairs): random
    pairs against the oracle, plus size-independent properties."""
    from audio_pattern_discovery_b200 import synth
    c, seqs, _ = synth.make_config("C2")
    seqs[17] = seqs[5].copy()                       # a duplicate: score must be exactly 0
    got, w = gpu_matrix(seqs, c["pct"], *c["weights"])
    n = len(seqs)
    assert np.all(np.diag(got) == 0.0)
    assert got[5, 17] == 0.0 and got[17, 5] == 0.0
    off = ~np.eye(n, dtype=bool)
    assert np.all(np.isfinite(got[off])) and np.all(got[off] >= 0)
    rng = np.random.default_rng(16)
    pairs = rng.integers(0, n, size=(3000, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    want = oracle.align_pairs(seqs, pairs, c["pct"], *c["weights"], workers=8)
    assert np.array_equal(bits(got[pairs[:, 0], pairs[:, 1]]), bits(want))
    st = w.stats()
    assert st["ordered_pairs"] == n * (n - 1)
    # idempotence: a second call on the same context gives the same bits
    d2 = w._ctx.align_all(c["pct"], *c["weights"])
    assert np.array_equal(bits(d2), bits(got))


def test_upgma_handoff_identical_merges():
    """Matrix -> clustering(distances, n, perc) (src/main.rs:194-200): merge order and
    assignments from the GPU matrix equal those from the oracle matrix."""
    from audio_pattern_discovery_b200 import synth
    seqs, _ = synth.make_sequences(90, np.random.default_rng(3).integers(40, 100, size=90), 10, 6, 77)
    want = oracle.align_all(seqs, 0.1, workers=8, variant="dense")
    got, _ = gpu_matrix(seqs, 0.1)
    m_want, thr_want, a_want = oracle.upgma(want, 0.05)
    m_got, thr_got, a_got = oracle.upgma(got, 0.05)
    assert thr_want == thr_got
    assert [(a, b, k) for a, b, k, _, _ in m_got] == [(a, b, k) for a, b, k, _, _ in m_want]
    assert np.array_equal(a_got, a_want)
    fast, _ = gpu_matrix(seqs, 0.1, fast=True)
    m_fast, _, a_fast = oracle.upgma(fast, 0.05)
    if not any(t for *_, t in m_want):
        assert [frozenset((a, b)) for a, b, *_ in m_fast] == [frozenset((a, b)) for a, b, *_ in m_want]
        assert np.array_equal(a_fast, a_want)


@pytest.mark.parametrize("ring", ["tmem", "smem", "global"])
def test_ring_homes_agree_bitwise(ring, monkeypatch):
    """The boundary ring in tensor memory, shared memory and global scratch: same bits."""
    from audio_pattern_discovery_b200 import synth
    seqs, _ = synth.make_sequences(130, 300, 20, 5, 91)
    want = oracle.align_all(seqs, 0.1, workers=8, variant="dense")
    monkeypatch.setenv("APD_RING", ring)
    got, _ = gpu_matrix(seqs, 0.1)
    assert np.array_equal(bits(got), bits(want))
    fast, _ = gpu_matrix(seqs, 0.1, fast=True)
    assert_close(fast, want)


def test_tiny_magnitudes_take_the_exact_sqrt_path():
    """Squared distances below 2^-101 (and subnormal ones) are outside the hot path's square
    root: the unit re-runs with the generic IEEE sqrt and stays bit-exact."""
    rng = np.random.default_rng(17)
    seqs = [(rng.normal(size=(int(t), 4)) * 1e-19).astype(np.float32) for t in rng.integers(5, 40, size=20)]
    seqs += [(rng.normal(size=(int(t), 4)) * 1e-16).astype(np.float32) for t in rng.integers(5, 40, size=20)]
    want = oracle.align_all(seqs, 0.3, workers=8, variant="dense")
    got, _ = gpu_matrix(seqs, 0.3)
    assert np.array_equal(bits(got), bits(want))
    big = [(s * np.float32(1e36)).astype(np.float32) for s in seqs[20:]]  # 1e20 magnitudes: squares overflow to INF
    want = oracle.align_all(big, 0.3, workers=8, variant="dense")
    got, _ = gpu_matrix(big, 0.3)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    f = ~np.isnan(want)
    assert np.array_equal(bits(got[f]), bits(want[f]))


def test_c1_shaped_reference_defaults_and_upgma():
    """BASELINE.json config 1 at the DTW boundary: ~auto-encoder embeddings (dim 10, per-frame
    z-scored), slices of >= 150 frames, the reference's default Discovery.toml (band 100 %,
    unit penalties, clustering_percentile 0.05): matrix bits and UPGMA merges."""
    from audio_pattern_discovery_b200 import AlignmentWorkers, Discovery, NDSequence, synth
    rng = np.random.default_rng(18)
    seqs, _ = synth.make_sequences(80, rng.integers(150, 260, size=80), 10, 7, 1001, zscore=True)
    nd = [NDSequence.from_array(s) for s in seqs]
    d = Discovery()  # project/config/Discovery.toml defaults
    w = AlignmentWorkers.new(nd)
    w.align_all(d)
    got = w.result.lock().unwrap().reshape(80, 80)
    want = oracle.align_all(seqs, d.warping_band_percentage, workers=8, variant="dense")
    assert np.array_equal(bits(got), bits(want))
    mg, tg, ag = oracle.upgma(got, d.clustering_percentile)
    mw, tw, aw = oracle.upgma(want, d.clustering_percentile)
    assert tg == tw and [(a, b, k) for a, b, k, _, _ in mg] == [(a, b, k) for a, b, k, _, _ in mw]
    assert np.array_equal(ag, aw)


def test_c5_shaped_long_unbanded_pairs_and_paths():
    """BASELINE.json config 5 shape: len 4096, dim 20, unbanded; a few sequences, spot-checked
    pairs, and device-traced warping paths on shorter unbanded pairs."""
    from audio_pattern_discovery_b200 import Context, synth
    seqs, _ = synth.make_sequences(10, 4096, 20, 3, 1005)
    got, _ = gpu_matrix(seqs, 1.0)
    pairs = np.array([(0, 1), (1, 0), (3, 7), (9, 2), (4, 5), (8, 6)], dtype=np.uint32)
    want = oracle.align_pairs(seqs, pairs, 1.0, workers=6)
    assert np.array_equal(bits(got[pairs[:, 0], pairs[:, 1]]), bits(want))
    short, _ = synth.make_sequences(6, np.array([700, 650, 700, 512, 900, 700]), 20, 2, 1006)
    with Context(0) as c:
        c.set_sequences(short)
        prs = [(0, 1), (2, 5), (4, 3)]
        scores, paths, lens = c.align_pairs(prs, 1.0, want_paths=True, path_cap=2000)
    for (i, j), s, p, ln in zip(prs, scores, paths, lens):
        ws, wp = oracle.dtw(short[i], short[j], 1.0, want_path=True)
        assert bits(s)[0] == bits(ws)[0]
        assert ln == len(wp) and np.array_equal(p, wp)


def test_full_c3_matrix_spot_check_and_properties():
    """BASELINE.json config 3 at full size (10 000 x len 512 x dim 20, band 10 %): 1e8 ordered
    pairs on the GPU; random pairs against the oracle (bit-exact), FAST within 1e-5 of STRICT
    on every entry, and size-independent properties."""
    from audio_pattern_discovery_b200 import APD_MODE_FAST, Context, synth
    c, seqs, _ = synth.make_config("C3")
    seqs[4321] = seqs[77].copy()
    n = len(seqs)
    with Context(0) as ctx:
        ctx.set_sequences(seqs)
        strict = ctx.align_all(c["pct"], *c["weights"])
        st = ctx.stats()
        fast = ctx.align_all(c["pct"], *c["weights"], mode=APD_MODE_FAST)
    assert st["ordered_pairs"] == n * (n - 1)
    assert st["cells_reference"] == n * (n - 1) * 51463
    assert np.all(np.diag(strict) == 0.0)
    assert strict[77, 4321] == 0.0 and strict[4321, 77] == 0.0
    assert np.all(np.isfinite(strict)) and strict.min() >= 0.0
    off = ~np.eye(n, dtype=bool)
    nz = off & (strict > 0)
    rel = np.abs(fast[nz] - strict[nz]) / strict[nz]
    assert rel.max() <= REL_TOL, rel.max()
    # rows of the duplicate pair agree entry by entry (same sequence, same partners)
    m = np.ones(n, dtype=bool)
    m[[77, 4321]] = False
    assert np.array_equal(bits(strict[77, m]), bits(strict[4321, m]))
    assert np.array_equal(bits(strict[m, 77]), bits(strict[m, 4321]))
    rng = np.random.default_rng(19)
    pairs = rng.integers(0, n, size=(1500, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    want = oracle.align_pairs(seqs, pairs, c["pct"], *c["weights"], workers=8)
    assert np.array_equal(bits(strict[pairs[:, 0], pairs[:, 1]]), bits(want))


def test_device_percentile_matches_reference_quirks():
    """numerics::percentile (src/numerics.rs:125-133) as an exact device order statistic."""
    import torch
    from audio_pattern_discovery_b200 import ApdError, Context
    rng = np.random.default_rng(23)
    with Context(0) as c:
        for trial in range(12):
            n = int(rng.integers(1, 200000))
            x = rng.normal(size=n).astype(np.float32)
            if trial % 3 == 0:
                x[rng.integers(0, n, size=max(1, n // 50))] = np.nan     # dropped, index from unfiltered length
            if trial % 4 == 1:
                x[rng.integers(0, n, size=max(1, n // 20))] = np.inf
                x[rng.integers(0, n, size=max(1, n // 20))] = 0.0
            if trial % 5 == 2:
                x = np.abs(x)
            t = torch.from_numpy(x).cuda()
            for perc in (0.0, 0.05, 0.5, 0.9):
                try:
                    want = oracle.percentile(x, perc)
                except IndexError:
                    with pytest.raises(ApdError):
                        c.percentile_device(t.data_ptr(), n, perc)
                    continue
                got = c.percentile_device(t.data_ptr(), n, perc)
                assert bits(got)[0] == bits(want)[0] or (got == 0 and want == 0), (trial, perc, got, want)
        t = torch.arange(10, dtype=torch.float32, device="cuda")
        with pytest.raises(ApdError):                      # index == len: the reference panics
            c.percentile_device(t.data_ptr(), 10, 1.0)
        with pytest.raises(ApdError):
            c.percentile(0.5)                              # no matrix yet


def test_threshold_of_the_device_matrix_equals_host_percentile():
    from audio_pattern_discovery_b200 import Context, synth
    seqs, _ = synth.make_sequences(300, np.random.default_rng(24).integers(20, 80, size=300), 10, 9, 24)
    seqs[5] = np.ones((1, 10), np.float32)   # 1-frame sequence: +INF rows / columns take part in the sort
    with Context(0) as c:
        c.set_sequences(seqs)
        m = c.align_all(0.1)
        for perc in (0.05, 0.25, 0.999):
            assert bits(c.percentile(perc))[0] == bits(oracle.percentile(m, perc))[0]
        assert c.stats()["select_ms"] > 0


@pytest.mark.parametrize("dim", [5, 13, 16, 24, 30, 32])
def test_every_padded_frame_width_on_gpu(dim):
    rng = np.random.default_rng(100 + dim)
    seqs = random_sequences(rng, 40, 6, 60, dim, integer=(dim % 2 == 1))
    want = oracle.align_all(seqs, 0.2, 0.75, 0.5, 1.0, workers=8, variant="dense")
    got, _ = gpu_matrix(seqs, 0.2, 0.75, 0.5, 1.0)
    assert np.array_equal(bits(got), bits(want))
    got, _ = gpu_matrix(seqs, 0.2)
    assert np.array_equal(bits(got), bits(oracle.align_all(seqs, 0.2, workers=8, variant="dense")))


def test_dim_above_the_supported_maximum_is_refused():
    from audio_pattern_discovery_b200 import ApdError, Context
    with Context(0) as c:
        with pytest.raises(ApdError) as ei:
            c.set_sequences([np.zeros((4, 33), np.float32)] * 2)
        assert ei.value.status == 4  # APD_ERR_UNSUPPORTED


def test_hybrid_ring_and_launch_plan():
    """A band taller than the 32 ring tiles tensor memory holds per warp keeps the ring's tail in shared
    memory (still 2 CTAs x 4 warps per SM); the launch plan says where each class keeps its ring."""
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(41)
    seqs = [rng.normal(size=(int(t), 20)).astype(np.float32) for t in rng.integers(150, 200, size=70)]
    want = oracle.align_all(seqs, 1.0, 0.75, 0.5, 1.0, workers=8, variant="dense")       # unbanded: ring = all row tiles
    with Context(0) as c:
        c.set_sequences(seqs)
        got = c.align_all(1.0, 0.75, 0.5, 1.0)
        plan = c.launch_plan()
        gotu = c.align_all(1.0)
    assert np.array_equal(bits(got), bits(want))
    assert np.array_equal(bits(gotu), bits(oracle.align_all(seqs, 1.0, workers=8, variant="dense")))
    assert any(p["ring"] == "tmem+smem" and 32 < p["ring_tiles"] <= 57 and p["ctas_per_sm"] == 2 for p in plan), plan


def test_wide_kernel_experiment_is_bit_exact(monkeypatch):
    """APD_WIDE=1: the 12-warps-per-SM kernel on 4 x 2-column tiles (measured slower, kept as an experiment)."""
    from audio_pattern_discovery_b200 import Context
    monkeypatch.setenv("APD_WIDE", "1")
    rng = np.random.default_rng(43)
    seqs = random_sequences(rng, 90, 40, 120, 20)
    want = oracle.align_all(seqs, 0.1, workers=8, variant="dense")
    with Context(0) as c:
        c.set_sequences(seqs)
        got = c.align_all(0.1)
        plan = c.launch_plan()
    assert np.array_equal(bits(got), bits(want))
    assert any("12 warps" in p["ring"] for p in plan), plan
