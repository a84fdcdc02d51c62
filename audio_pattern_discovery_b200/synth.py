"""Synthetic workloads of SURVEY.md section 8(d): time-warped, noisy copies of K smooth
prototype trajectories so that the DTW matrix has real cluster structure.  Used by the
tests and by bench.py (there is no network for datasets); numpy only.
"""
import numpy as np


def _prototype(rng, length, dim):
    p = np.cumsum(rng.normal(0.0, 0.3, size=(length, dim)), axis=0)
    return p - p.mean(axis=1, keepdims=True)  # per-frame mean-centring, like src/spectrogram.rs:74-75


def _warp(rng, proto, length):
    """Monotone random resampling of proto to `length` frames (linear interpolation)."""
    steps = rng.gamma(4.0, 1.0, size=length)
    pos = np.cumsum(steps)
    pos = (pos - pos[0]) / max(pos[-1] - pos[0], 1e-9) * (proto.shape[0] - 1)
    lo = np.floor(pos).astype(np.int64)
    hi = np.minimum(lo + 1, proto.shape[0] - 1)
    f = (pos - lo)[:, None]
    return proto[lo] * (1.0 - f) + proto[hi] * f


def make_sequences(n, lens, dim, k_prototypes, seed, noise=0.1, zscore=False):
    """lens: int (fixed) or array of n lengths.  Returns (list of (T, dim) float32, labels)."""
    rng = np.random.default_rng(seed)
    lens = np.full(n, lens, dtype=np.int64) if np.isscalar(lens) else np.asarray(lens, dtype=np.int64)
    plen = int(max(lens.max() if n else 1, 8))
    protos = [_prototype(rng, plen, dim) for _ in range(k_prototypes)]
    labels = rng.integers(0, k_prototypes, size=n)
    out = []
    for s in range(n):
        x = _warp(rng, protos[labels[s]], int(lens[s])) + rng.normal(0.0, noise, size=(int(lens[s]), dim))
        if zscore:  # per-frame z-score with sigma floor 1.0, like src/neural.rs:61-62
            mu = x.mean(axis=1, keepdims=True)
            sd = np.maximum(x.std(axis=1, keepdims=True), 1.0)
            x = (x - mu) / sd
        out.append(np.ascontiguousarray(x, dtype=np.float32))
    return out, labels


# The configurations BASELINE.json names (SURVEY.md section 8 shorthand C2..C5).
def config(name, n=None):
    rng_len = np.random.default_rng({"C2": 2002, "C3": 2003, "C4": 2004, "C5": 2005, "C1ref": 2001}[name])
    if name == "C1ref":
        # The reference's own shipped configuration (project/config/Discovery.toml:5,17-20): ~200 slices
        # of >= vat_min_len = 150 frames, auto-encoder embeddings (dim 10, per-frame z-scored like
        # src/neural.rs:61-62), warping_band_percentage = 1.0 (effectively unbanded), unit penalties.
        n = n or 200
        return dict(n=n, lens=rng_len.integers(150, 261, size=n), dim=10, k=12, seed=1001, pct=1.0,
                    weights=(1.0, 1.0, 1.0), zscore=True)
    if name == "C2":
        n = n or 2000
        return dict(n=n, lens=rng_len.integers(64, 257, size=n), dim=20, k=40, seed=1002, pct=0.1,
                    weights=(0.75, 0.5, 1.0), zscore=False)
    if name == "C3":
        n = n or 10000
        return dict(n=n, lens=512, dim=20, k=100, seed=1003, pct=0.1, weights=(1.0, 1.0, 1.0), zscore=False)
    if name == "C4":
        n = n or 20000
        buckets = np.array([128, 256, 384, 512, 768, 1024])
        lens = buckets[rng_len.integers(0, len(buckets), size=n)] - rng_len.integers(0, 32, size=n)
        return dict(n=n, lens=lens, dim=8, k=200, seed=1004, pct=0.05, weights=(1.0, 1.0, 1.0), zscore=True)
    if name == "C5":
        n = n or 1000
        return dict(n=n, lens=4096, dim=20, k=20, seed=1005, pct=1.0, weights=(1.0, 1.0, 1.0), zscore=False)
    raise KeyError(name)


def make_config(name, n=None):
    c = config(name, n)
    seqs, labels = make_sequences(c["n"], c["lens"], c["dim"], c["k"], c["seed"], zscore=c["zscore"])
    return c, seqs, labels
