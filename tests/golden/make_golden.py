"""Generates tests/golden/dtw_golden.npz with the pure-Python transliteration of the
reference (oracle/literal.py).  The reference itself is Rust and cannot be built in
this image (no cargo/rustc), so these vectors pin the C oracle and the CUDA path to
an independent restatement, not to reference-produced output ("parity unpinned",
SURVEY.md section 8c).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import literal  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    cases = {}
    meta = []
    k = 0
    weights = [(1.0, 1.0, 1.0), (0.75, 0.5, 1.0), (0.5, 1.0, 0.25), (1.0, 0.3, 0.9)]
    for integer in (True, False):
        for dim in (1, 3, 8, 10, 20, 26):
            for pct in (0.0, 0.1, 0.33, 1.0):
                n = int(rng.integers(1, 28))
                m = int(rng.integers(1, 28))
                if integer:
                    x = rng.integers(0, 3, size=(n, dim)).astype(np.float32)
                    y = rng.integers(0, 3, size=(m, dim)).astype(np.float32)
                else:
                    x = rng.normal(size=(n, dim)).astype(np.float32)
                    y = rng.normal(size=(m, dim)).astype(np.float32)
                ins, dele, mat = weights[k % len(weights)]
                band = literal.warping_band(pct, max(n, m))
                sp = literal.construct_alignment(x, y, band, ins, dele, mat)
                s = literal.score(sp, n, m)
                p = np.array(literal.path(sp, n, m), dtype=np.uint32).reshape(-1, 2)
                cases["x%d" % k] = x
                cases["y%d" % k] = y
                cases["path%d" % k] = p
                meta.append((pct, ins, dele, mat, float(s)))
                cases["score_bits%d" % k] = np.array([np.float32(s)], dtype=np.float32).view(np.uint32)
                k += 1
    cases["meta"] = np.array(meta, dtype=np.float64)
    # two small all-pairs matrices (variable lengths incl. 1-frame and 2-frame sequences)
    for tag, integer, dim, pct, w in (("A", True, 2, 0.2, (1.0, 1.0, 1.0)), ("B", False, 10, 0.1, (0.75, 0.5, 1.0))):
        lens = [1, 2, 5, 9, 9, 14, 17, 23]
        seqs = []
        for t in lens:
            if integer:
                seqs.append(rng.integers(0, 3, size=(t, dim)).astype(np.float32))
            else:
                seqs.append(rng.normal(size=(t, dim)).astype(np.float32))
        mat = literal.align_all(seqs, pct, *w)
        cases["all%s_flat" % tag] = np.concatenate([s.ravel() for s in seqs])
        cases["all%s_lens" % tag] = np.array(lens, dtype=np.uint32)
        cases["all%s_dim" % tag] = np.array([dim], dtype=np.uint32)
        cases["all%s_params" % tag] = np.array([pct, *w], dtype=np.float64)
        cases["all%s_matrix_bits" % tag] = mat.view(np.uint32)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dtw_golden.npz")
    np.savez_compressed(out, **cases)
    print(out, k, "pair cases", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
