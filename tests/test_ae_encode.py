"""Row f3 of SURVEY.md section 8: the embedding step in front of the path --
NDSequence::encoded / AutoEncoder::predict (src/spectrogram.rs:103-121, src/neural.rs:55-71).

CPU: the C oracle against the committed golden vectors (made by an independent pure-Python
transliteration, tests/golden/make_ae_golden.py).  GPU: apd_set_sequences_encoded writes the
embeddings into the device arena; they and the DTW matrix computed from them must equal the
oracle's bit for bit.

Tolerance: bit-exact.  Every f32 operation is restated in the reference's order; f32::exp is
the C library's expf, which the kernel re-implements with glibc's own algorithm (double
precision, one rounding to f32) -- identical wherever the double intermediate is not within
~2^-29 of an f32 rounding boundary (0 differences on 3.2e8 arguments against this image's
glibc 2.39).  If a different libm ever disagreed, the bound would be 1 ulp of exp."""
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ae_golden.npz"))


@pytest.mark.parametrize("name", ["ref", "wide", "tiny"])
def test_oracle_predict_matches_golden(name):
    w, b, x, y = (GOLD[name + s] for s in ("_w", "_b", "_x", "_y"))
    got = oracle.ae_encode(x, w, b)
    assert np.array_equal(got.view(np.uint32), y.view(np.uint32))


def test_oracle_predict_properties():
    """Per-frame z-score: mean ~ 0; the sigma floor of 1.0 keeps identical latents at ~0 instead of 0/0."""
    rng = np.random.default_rng(3)
    w = np.zeros((26, 10), np.float32)
    b = np.full(10, 0.25, np.float32)
    x = rng.normal(size=(5, 26)).astype(np.float32)
    assert np.all(np.abs(oracle.ae_encode(x, w, b)) < 1e-3)   # all latents equal -> (p - mu) / max(~0, 1) ~ 0 (sum rounding only)
    w = ((rng.random((26, 10)) - 0.5) / 10).astype(np.float32)
    y = oracle.ae_encode(x * 50, w, b)
    assert np.all(np.abs(y.mean(axis=1)) < 1e-5) and np.all(np.isfinite(y))


def _weights(rng, n_bins, n_latent, gain=6.0):
    w = ((rng.random((n_bins, n_latent)) - 0.5) / n_latent * gain).astype(np.float32)   # Mat::seeded-like, spread out
    b = ((rng.random(n_latent) - 0.5) / n_latent).astype(np.float32)
    return w, b


@pytest.mark.gpu
@pytest.mark.parametrize("n_bins,n_latent", [(26, 10), (33, 17), (3, 1), (64, 32)])
def test_device_embeddings_are_bit_exact(n_bins, n_latent):
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(100 + n_bins)
    w, b = _weights(rng, n_bins, n_latent)
    lens = [0, 1, 2, 37, 150, 64, 5, 260]
    ceps = [(rng.normal(size=(t, n_bins)) * 3.0).astype(np.float32) for t in lens]
    ceps[3][0] = 1000.0
    ceps[3][1] = -1000.0
    ceps[3][2] = 0.0
    with Context(0) as ctx:
        ctx.set_sequences_encoded(ceps, w, b)
        assert ctx.dim == n_latent
        for k, x in enumerate(ceps):
            want = oracle.ae_encode(x, w, b) if len(x) else np.zeros((0, n_latent), np.float32)
            got = ctx.get_sequence(k, len(x))
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), k


@pytest.mark.gpu
def test_golden_vectors_on_device():
    from audio_pattern_discovery_b200 import Context
    with Context(0) as ctx:
        for name in ("ref", "wide", "tiny"):
            w, b, x, y = (GOLD[name + s] for s in ("_w", "_b", "_x", "_y"))
            ctx.set_sequences_encoded([x], w, b)
            assert np.array_equal(ctx.get_sequence(0, len(x)).view(np.uint32), y.view(np.uint32)), name


@pytest.mark.gpu
def test_matrix_from_encoded_input_equals_oracle():
    """The reference's real pipeline (src/main.rs:150-161,187-195): encoded() then align_all,
    dim 10, shipped Discovery.toml (band 100 %, unit penalties) -- the embeddings never visit the host."""
    from audio_pattern_discovery_b200 import Context
    rng = np.random.default_rng(9)
    w, b = _weights(rng, 26, 10)
    ceps = [(np.cumsum(rng.normal(size=(int(t), 26)), axis=0) * 0.4).astype(np.float32) for t in rng.integers(20, 90, size=40)]
    emb = [oracle.ae_encode(x, w, b) for x in ceps]
    want = oracle.align_all(emb, 1.0, 1.0, 1.0, 1.0, workers=4, variant="dense")
    with Context(0) as ctx:
        ctx.set_sequences_encoded(ceps, w, b)
        got = ctx.align_all(1.0, 1.0, 1.0, 1.0)
        # and the plain upload of host-side embeddings gives the same matrix
        ctx.set_sequences(emb)
        got2 = ctx.align_all(1.0, 1.0, 1.0, 1.0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got2.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_encoded_argument_errors():
    from audio_pattern_discovery_b200 import ApdError, Context, _capi
    with Context(0) as ctx:
        with pytest.raises(ApdError) as ei:
            ctx.set_sequences_encoded([np.zeros((3, 65), np.float32)], np.zeros((65, 4), np.float32), np.zeros(4, np.float32))
        assert ei.value.status == _capi.APD_ERR_UNSUPPORTED
        with pytest.raises(ApdError) as ei:
            ctx.set_sequences_encoded([np.zeros((3, 8), np.float32)], np.zeros((8, 33), np.float32), np.zeros(33, np.float32))
        assert ei.value.status == _capi.APD_ERR_UNSUPPORTED
