"""Times the on-device trace-back kernel (K2) on C5-shaped pairs: python tools/profile_paths.py [--n 32] [--pairs 40] [--len 4096]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_pattern_discovery_b200 import Context, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--pairs", type=int, default=40)
    ap.add_argument("--len", type=int, default=4096)
    ap.add_argument("--pct", type=float, default=1.0)
    a = ap.parse_args()
    seqs, _ = synth.make_sequences(a.n, a.len, 20, 4, 1005)
    rng = np.random.default_rng(1005)
    pairs = []
    while len(pairs) < a.pairs:
        i, j = rng.integers(0, a.n, size=2)
        if i != j:
            pairs.append((int(i), int(j)))
    with Context(0) as c:
        c.set_sequences(seqs)
        for rep in range(2):
            t0 = time.perf_counter()
            scores, paths, lens = c.align_pairs(pairs, a.pct, want_paths=True, path_cap=2 * a.len)
            dt = time.perf_counter() - t0
            cells = len(pairs) * a.len * a.len
            print("paths: %d pairs of %dx%d in %.3f s (%.1f GCUPS incl. copies), mean path length %.0f"
                  % (len(pairs), a.len, a.len, dt, cells / dt / 1e9, lens.mean()))


if __name__ == "__main__":
    main()
