"""Generates tests/golden/ae_golden.npz: seeded auto-encoder weights (the value distribution of
Mat::seeded, src/numerics.rs:178-186: (U[0,1) - 0.5) / cols), cepstrum-like input frames, and the
embeddings the restatement of AutoEncoder::predict (src/neural.rs:55-71) produces for them.

The reference itself cannot run here (Rust, no cargo/rustc in the image): the vectors come from
the pure-Python transliteration below (numpy.float32 scalar arithmetic, libm expf through ctypes
-- the function a Linux build of the reference calls for f32::exp), NOT from the C oracle, so the
C oracle and the CUDA kernel are both checked against an independent restatement.

    python tests/golden/make_ae_golden.py
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libm = ctypes.CDLL("libm.so.6")
_libm.expf.restype = ctypes.c_float
_libm.expf.argtypes = [ctypes.c_float]
f32 = np.float32


def predict_frame(x, w, b):
    """AutoEncoder::predict on a 1 x n_bins Mat, operation by operation in f32."""
    n_bins, n_latent = w.shape
    p = []
    for j in range(n_latent):
        acc = f32(0.0)
        for k in range(n_bins):
            acc = f32(acc + f32(x[k] * w[k, j]))       # Mat::mul, src/numerics.rs:305-319
        acc = f32(acc + b[j])                           # add_col
        e = f32(_libm.expf(ctypes.c_float(float(f32(-acc)))))
        s = f32(f32(1.0) / f32(f32(1.0) + e))           # sigmoid
        p.append(f32(s * f32(255.0)))                   # scale(255.0)
    mu = f32(0.0)
    for v in p:
        mu = f32(mu + v)
    mu = f32(mu / f32(n_latent))                        # mean, src/numerics.rs:12-18
    sd = f32(0.0)
    for v in p:
        d = f32(v - mu)
        sd = f32(sd + f32(d * d))
    sd = f32(np.sqrt(f32(sd / f32(n_latent))))          # std, src/numerics.rs:23-29
    sigma = sd if sd > f32(1.0) else f32(1.0)           # f32::max(std, 1.0)
    return np.array([f32(f32(v - mu) / sigma) for v in p], dtype=np.float32)


def make():
    rng = np.random.default_rng(4242)
    cases = {}
    for name, n_bins, n_latent, scale in (("ref", 26, 10, 1.0), ("wide", 33, 17, 4.0), ("tiny", 3, 1, 30.0)):
        w = ((rng.random((n_bins, n_latent)) - 0.5) / n_latent * scale * 8).astype(np.float32)
        b = ((rng.random(n_latent) - 0.5) / n_latent).astype(np.float32)
        x = (rng.normal(size=(40, n_bins)) * 3.0).astype(np.float32)
        x[0] = 0.0                      # silence: all latents equal sigmoid(b)
        x[1] = 1000.0                   # saturation: exp underflow / overflow paths
        x[2] = -1000.0
        y = np.stack([predict_frame(x[t], w, b) for t in range(x.shape[0])])
        cases.update({name + "_w": w, name + "_b": b, name + "_x": x, name + "_y": y})
    np.savez_compressed(os.path.join(HERE, "ae_golden.npz"), **cases)
    print("wrote ae_golden.npz:", {k: v.shape for k, v in cases.items()})


if __name__ == "__main__":
    make()
