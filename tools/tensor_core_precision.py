"""north_star (d): would the |a|^2 + |b|^2 - 2 a.b contraction on tensor cores hold the 1e-5
bound?  Numerical experiment on the C3 workload's own frames (numpy, no GPU needed): the
frame distance is evaluated in float64 (truth), in the reference's f32 difference form, in
the f32 dot form, and in the dot form with the products rounded the way the tensor-core input
formats round them (TF32: 10 explicit mantissa bits; 3xTF32 and 3xBF16 error-compensated
splits as used by "fp32-emulating" GEMMs).  Reports the relative error of d for aligned
(same prototype, same time) frame pairs -- the cells a low-cost DTW path is made of -- and
for random frame pairs.

    python tools/tensor_core_precision.py > profiles/r1_tensor_core_precision.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_pattern_discovery_b200 import synth  # noqa: E402


def round_mantissa(x, bits):
    """Round float32 to `bits` explicit mantissa bits (round to nearest even)."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    half = np.uint64(1 << (drop - 1))
    lsb = (u >> np.uint64(drop)) & np.uint64(1)
    u = (u + half - np.uint64(1) + lsb) >> np.uint64(drop) << np.uint64(drop)
    return u.astype(np.uint32).view(np.float32)


def dot_split(a, b, bits, terms):
    """sum_k a_k b_k with operands split into `terms` pieces of `bits` mantissa bits, f32 accumulate."""
    def split(v):
        parts, r = [], v.astype(np.float32)
        for _ in range(terms):
            p = round_mantissa(r, bits)
            parts.append(p)
            r = (r - p).astype(np.float32)
        return parts
    pa, pb = split(a), split(b)
    acc = np.zeros(a.shape[0], dtype=np.float32)
    for i in range(terms):
        for j in range(terms):
            if i + j < terms:  # the usual truncation of the smallest cross terms
                acc = (acc + np.sum(pa[i] * pb[j], axis=1, dtype=np.float32)).astype(np.float32)
    return acc


def report(name, d, truth):
    rel = np.abs(d.astype(np.float64) - truth) / truth
    print("  %-34s median %.2e   p99 %.2e   max %.2e   share > 1e-5: %5.1f %%"
          % (name, np.median(rel), np.percentile(rel, 99), rel.max(), 100 * np.mean(rel > 1e-5)))


def main():
    c, seqs, labels = synth.make_config("C3", 400)
    rng = np.random.default_rng(0)
    same = []
    for k in np.unique(labels):
        idx = np.flatnonzero(labels == k)
        for _ in range(40):
            if len(idx) >= 2:
                i, j = rng.choice(idx, size=2, replace=False)
                t = int(rng.integers(0, 512))
                same.append((seqs[i][t], seqs[j][t]))
    a = np.array([p[0] for p in same], dtype=np.float32)
    b = np.array([p[1] for p in same], dtype=np.float32)
    ra = np.array([seqs[int(rng.integers(0, 400))][int(rng.integers(0, 512))] for _ in range(len(a))], dtype=np.float32)
    rb = np.array([seqs[int(rng.integers(0, 400))][int(rng.integers(0, 512))] for _ in range(len(a))], dtype=np.float32)
    for title, x, y in (("aligned frames of sequences from the same prototype (%d pairs)" % len(a), a, b),
                        ("random frame pairs (%d pairs)" % len(ra), ra, rb)):
        truth = np.sqrt(np.sum((x.astype(np.float64) - y.astype(np.float64)) ** 2, axis=1))
        nx = np.sum(x * x, axis=1, dtype=np.float32)
        ny = np.sum(y * y, axis=1, dtype=np.float32)
        print(title + ": |x|^2 median %.1f, d median %.3f" % (np.median(nx), np.median(truth)))
        diff = np.sqrt(np.sum((x - y) ** 2, axis=1, dtype=np.float32))
        report("f32 difference form (reference)", diff, truth)

        def dist_from_dot(dot):
            return np.sqrt(np.maximum((nx + ny - np.float32(2) * dot).astype(np.float32), 0))
        report("f32 dot form (CUDA cores)", dist_from_dot(np.sum(x * y, axis=1, dtype=np.float32)), truth)
        report("TF32 x1 dot (tensor core)", dist_from_dot(dot_split(x, y, 10, 1)), truth)
        report("BF16 x3 dot (tensor core)", dist_from_dot(dot_split(x, y, 7, 3)), truth)
        report("TF32 x3 dot (tensor core)", dist_from_dot(dot_split(x, y, 10, 3)), truth)
    print("Conclusion: every dot-form variant, even exact-f32 products, misses 1e-5 on aligned frames because "
          "|x|^2 + |y|^2 - 2 x.y cancels; the fused CUDA-core difference form is the only one inside the bound.")


if __name__ == "__main__":
    main()
