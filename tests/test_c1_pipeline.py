"""BASELINE.json config 1 ("reference CPU pipeline on a synthetic 32-file chirp/whistle folder:
cepstrum extraction, ~200 slices, full pairwise banded DTW + UPGMA") at the DTW boundary: the
front-end restatement of tests/c1_frontend.py produces the slices' auto-encoder embeddings, and
from there on everything must equal the reference's arithmetic bit for bit (C1
parity is asserted from `frames` on, never from WAV bytes)."""
import numpy as np
import pytest

from oracle import oracle

from . import c1_frontend, emul


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def c1_sequences():
    seqs = c1_frontend.make_c1_sequences()
    assert 120 <= len(seqs) <= 260 and all(s.shape[1] == 10 and len(s) > 150 for s in seqs)
    return seqs


def test_kernel_schedule_on_c1_inputs_cpu(c1_sequences):
    """The kernel's lane program (host emulator) on real-pipeline-shaped inputs, defaults of
    project/config/Discovery.toml (band 100 %, unit penalties)."""
    short = sorted(c1_sequences, key=len)[:26]
    want = oracle.align_all(short, 1.0, workers=8, variant="dense")
    got, _ = emul.align_all(short, 1.0)
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.gpu
def test_c1_matrix_threshold_and_clusters_on_gpu(c1_sequences):
    from audio_pattern_discovery_b200 import AgglomerativeClustering, AlignmentWorkers, Discovery, NDSequence
    seqs = c1_sequences
    n = len(seqs)
    d = Discovery()                                   # the reference's shipped configuration
    w = AlignmentWorkers.new([NDSequence.from_array(s) for s in seqs])
    w.align_all(d)
    got = w.result.lock().unwrap().reshape(n, n)
    want = oracle.align_all(seqs, d.warping_band_percentage, workers=8, variant="dense")
    assert np.array_equal(bits(got), bits(want))
    thr = w._ctx.percentile(d.clustering_percentile)  # device threshold == host percentile
    assert bits(thr)[0] == bits(oracle.percentile(want, d.clustering_percentile))[0]
    ops, clusters = AgglomerativeClustering.clustering(got.ravel(), n, d.clustering_percentile, threshold=thr)
    mw, _, aw = oracle.upgma(want, d.clustering_percentile)
    assert [(o.merge_i, o.merge_j, o.into) for o in ops] == [(a, b, k) for a, b, k, _, _ in mw]
    assert clusters == set(int(r) for r in aw)
