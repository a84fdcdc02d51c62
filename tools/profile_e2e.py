"""Breaks the host-buffer call (AlignmentWorkers path) into its phases for the full C3 workload."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_pattern_discovery_b200 import Context, synth  # noqa: E402

c, seqs, _ = synth.make_config("C3", int(sys.argv[1]) if len(sys.argv) > 1 else None)
n = len(seqs)
out = np.empty((n, n), dtype=np.float32)
with Context(0) as ctx:
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.set_sequences(seqs)
        t1 = time.perf_counter()
        k = ctx.packed_len(c["pct"])
        t2 = time.perf_counter()
        ctx.align_all(c["pct"], *c["weights"], out=out)
        t3 = time.perf_counter()
        st = ctx.stats()
        print("set_sequences %.1f ms (h2d %.1f ms) | plan %.1f ms | align_all %.1f ms (kernel %.1f, scatter %.2f, d2h %.1f ms)"
              % ((t1 - t0) * 1e3, st["h2d_ms"], (t2 - t1) * 1e3, (t3 - t2) * 1e3, st["kernel_ms"], st["scatter_ms"], st["d2h_ms"]))
