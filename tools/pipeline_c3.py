"""The reference's stage 3 (src/main.rs:187-203) at the C3 scale on this stack: all-pairs DTW on
the GPU (STRICT), clustering threshold on the device, result-identical UPGMA on the host,
cluster sets -- with timings and the purity of the clusters against the synthetic prototypes."""
import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_pattern_discovery_b200 import AgglomerativeClustering, Context, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else None
c, seqs, labels = synth.make_config("C3", n)
n = len(seqs)
with Context(0) as ctx:
    t0 = time.perf_counter()
    ctx.set_sequences(seqs)
    d = ctx.align_all(c["pct"], *c["weights"])
    t1 = time.perf_counter()
    thr = ctx.percentile(0.05)
    t2 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    ops, clusters = AgglomerativeClustering.clustering(d.ravel(), n, 0.05, threshold=thr)
    groups = AgglomerativeClustering.cluster_sets(ops, clusters, n)
t3 = time.perf_counter()
pure = sum(np.bincount(labels[g]).max() for g in groups)
covered = sum(len(g) for g in groups)
print("n=%d: DTW matrix %.2f s | threshold (device radix select) %.4f s -> %.6f | UPGMA %.2f s, %d merges, %d clusters "
      "(%d non-singular covering %d sequences, purity %.4f, %d prototypes), %d exact ties"
      % (n, t1 - t0, t2 - t1, thr, t3 - t2, len(ops), len(clusters), len(groups), covered, pure / max(covered, 1),
         len(np.unique(labels)), sum(o.tie for o in ops)))
