"""The drop-in call on several GPUs from ONE process (apd_create_multi): the reference's single
blocking `workers.align_all(&discover)` (src/main.rs:189-195) fans out over the devices inside
the library -- pair space dealt over the group, packed results stored peer-to-peer from inside
the DTW kernels (or copied afterwards without peer access), matrix copied back in row slabs.
Bit-exact against the oracle.  Multi-device cases are skipped below 2 GPUs; the group of one
runs everywhere."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(n=150, seed=5):
    from audio_pattern_discovery_b200 import synth
    rng = np.random.default_rng(seed)
    seqs, _ = synth.make_sequences(n, rng.integers(30, 140, size=n), 20, 8, 31)
    return seqs


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_group_of_one_is_the_single_device_path():
    from audio_pattern_discovery_b200 import Context
    seqs = _case(70)
    want = oracle.align_all(seqs, 0.1, 0.75, 0.5, 1.0, workers=8, variant="dense")
    with Context(devices=[0]) as ctx:
        assert ctx.group_size == 1 and not ctx.peer_stores
        ctx.set_sequences(seqs)
        got = ctx.align_all(0.1, 0.75, 0.5, 1.0)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert ctx.percentile(0.05) == oracle.percentile(want, 0.05)


def test_group_rejects_bad_device_lists():
    from audio_pattern_discovery_b200 import ApdError, Context
    with pytest.raises(ApdError):
        Context(devices=[0, 0])
    with pytest.raises(ApdError):
        Context(devices=[_gpus() + 3])


@pytest.mark.parametrize("g", [2, 3, 4, 8])
def test_group_matrix_is_bit_exact(g):
    if _gpus() < g:
        pytest.skip("needs %d GPUs" % g)
    from audio_pattern_discovery_b200 import AlignmentWorkers, Context, Discovery
    seqs = _case()
    want = oracle.align_all(seqs, 0.1, 0.75, 0.5, 1.0, workers=8, variant="dense")
    with Context(devices=list(range(g))) as ctx:
        assert ctx.group_size == g
        ctx.set_sequences(seqs)
        out = np.full((len(seqs), len(seqs)), 7.0, np.float32)      # pageable, pre-filled: every entry must be written
        ctx.align_all(0.1, 0.75, 0.5, 1.0, out=out)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
        st = ctx.stats()
        ref_cells = sum(oracle.pair_cells(len(a), len(b), 0.1) for i, a in enumerate(seqs) for j, b in enumerate(seqs) if i != j)
        assert st["cells_reference"] == ref_cells and st["units_local"] == st["units_total"]
        assert st["kernel_launches"] >= 2 * g
        # the leader holds the whole matrix on the device: the threshold step works on the group
        assert ctx.percentile(0.05) == oracle.percentile(want, 0.05)
        # a second batch on the same group (plan rebuilt for other lengths), then the first again
        seqs2 = _case(61, seed=8)
        ctx.set_sequences(seqs2)
        got2 = ctx.align_all(1.0, 1.0, 1.0, 1.0)
        assert np.array_equal(got2.view(np.uint32), oracle.align_all(seqs2, 1.0, 1.0, 1.0, 1.0, workers=8, variant="dense").view(np.uint32))
        ctx.set_sequences(seqs)
        assert np.array_equal(ctx.align_all(0.1, 0.75, 0.5, 1.0).view(np.uint32), want.view(np.uint32))
        # requested pairs with paths are dealt over the group
        pairs = [(i, (i * 7 + 3) % len(seqs)) for i in range(0, 40) if i != (i * 7 + 3) % len(seqs)]
        scores, paths, lens = ctx.align_pairs(pairs, 0.1, 0.75, 0.5, 1.0, want_paths=True, path_cap=300)
        for (i, j), s, p in zip(pairs, scores, paths):
            s_ref, p_ref = oracle.dtw(seqs[i], seqs[j], 0.1, 0.75, 0.5, 1.0, variant="literal", want_path=True)
            assert np.float32(s).view(np.uint32) == np.float32(s_ref).view(np.uint32) and np.array_equal(p, p_ref)
    # the reference-facing mirror uses every visible GPU by default
    w = AlignmentWorkers.new(seqs)
    w.align_all(Discovery(warping_band_percentage=0.1, insertion_penalty=0.75, deletion_penalty=0.5, match_penalty=1.0))
    assert w._ctx.group_size == min(_gpus(), 8)
    assert np.array_equal(w.result.lock().unwrap().reshape(len(seqs), -1).view(np.uint32), want.view(np.uint32))


def test_group_without_peer_stores():
    """APD_NO_P2P=1: the packed shards are copied between the devices after the kernels."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "from audio_pattern_discovery_b200 import Context, synth\n"
        "from oracle import oracle\n"
        "rng = np.random.default_rng(5)\n"
        "seqs, _ = synth.make_sequences(90, rng.integers(30, 140, size=90), 20, 8, 31)\n"
        "ctx = Context(devices=[0, 1]); assert not ctx.peer_stores\n"
        "ctx.set_sequences(seqs); got = ctx.align_all(0.1, 1.0, 1.0, 1.0)\n"
        "want = oracle.align_all(seqs, 0.1, 1.0, 1.0, 1.0, workers=8, variant='dense')\n"
        "print('ok=%%s' %% np.array_equal(got.view(np.uint32), want.view(np.uint32)))\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, APD_NO_P2P="1"))
    assert r.returncode == 0 and "ok=True" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_group_c3_shape_equals_single_device():
    """C3-shaped (len 512, dim 20, band 10 %), large enough that every device gets many units."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from audio_pattern_discovery_b200 import Context, synth
    c, seqs, _ = synth.make_config("C3", 600)
    with Context(0) as one:
        one.set_sequences(seqs)
        a = one.align_all(c["pct"])
    with Context(devices="all") as grp:
        grp.set_sequences(seqs)
        b = grp.align_all(c["pct"])
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_group_takes_every_upload_form():
    """Flat upload (device-side gather) and encoded upload (auto-encoder on the device) on a device group:
    the leader builds the arena, the other members pull it over NVLink."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from audio_pattern_discovery_b200 import AlignmentWorkers, Context, Discovery
    rng = np.random.default_rng(17)
    w = ((rng.random((26, 10)) - 0.5) / 10 * 6).astype(np.float32)
    b = ((rng.random(10) - 0.5) / 10).astype(np.float32)
    ceps = [(np.cumsum(rng.normal(size=(int(t), 26)), axis=0) * 0.4).astype(np.float32) for t in rng.integers(20, 90, size=70)]
    emb = [oracle.ae_encode(x, w, b) for x in ceps]
    want = oracle.align_all(emb, 1.0, 1.0, 1.0, 1.0, workers=8, variant="dense")
    with Context(devices="all") as ctx:
        ctx.set_sequences_encoded(ceps, w, b)
        got = ctx.align_all(1.0)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        flat = np.concatenate([e.reshape(-1) for e in emb])
        offs = np.cumsum([0] + [e.size for e in emb[:-1]])
        ctx.set_sequences_flat(flat, offs, [len(e) for e in emb], 10)
        got = ctx.align_all(1.0)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    wk = AlignmentWorkers.new_encoded(ceps, w, b)
    wk.align_all(Discovery(warping_band_percentage=1.0))
    assert np.array_equal(wk.result.lock().unwrap().reshape(70, 70).view(np.uint32), want.view(np.uint32))


def test_group_tiny_inputs():
    """Fewer sequences than devices, empty and one-frame sequences: every member still takes part in the call."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(23)
    with Context(devices="all") as ctx:
        for lens in ([], [5], [5, 7], [0, 4, 1], [3, 3, 3, 1, 0, 9]):
            seqs = [rng.normal(size=(t, 6)).astype(np.float32) for t in lens]
            ctx.set_sequences(seqs, dim=6)
            got = ctx.align_all(0.5, 0.75, 0.5, 1.0)
            n = len(lens)
            assert got.shape == (n, n)
            if n >= 2:
                want = oracle.align_all(seqs, 0.5, 0.75, 0.5, 1.0, workers=2, variant="dense")
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), lens
            elif n == 1:
                assert got[0, 0] == 0.0
