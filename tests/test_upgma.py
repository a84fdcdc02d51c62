"""The fast host UPGMA (csrc/upgma.cpp via apd_upgma) takes exactly the decisions of the
literal restatement of src/clustering.rs:81-209 in the oracle: merge order, f32 merge
distances (bitwise), threshold, final assignment, tie flags."""
import numpy as np
import pytest

from audio_pattern_discovery_b200 import AgglomerativeClustering, Merge
from oracle import oracle


def run_both(d, perc):
    n = d.shape[0]
    ops, clusters = AgglomerativeClustering.clustering(d.ravel(), n, perc)
    want, thr, assign = oracle.upgma(d, perc)
    assert AgglomerativeClustering.last_threshold.view(np.uint32) == thr.view(np.uint32)
    got = [(o.merge_i, o.merge_j, o.into, np.float32(o.distance).view(np.uint32), int(o.tie)) for o in ops]
    exp = [(a, b, k, np.float32(dd).view(np.uint32), t) for a, b, k, dd, t in want]
    assert got == exp
    assert clusters == set(int(r) for r in assign)
    return ops, clusters


@pytest.mark.parametrize("seed", range(6))
def test_random_asymmetric_matrices(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 70))
    d = rng.uniform(0.1, 5.0, size=(n, n)).astype(np.float32)
    np.fill_diagonal(d, 0.0)
    for perc in (0.0, 0.05, 0.3, 0.9):
        run_both(d, perc)


def test_structured_data_with_groups():
    rng = np.random.default_rng(10)
    pts = np.concatenate([rng.normal(c, 0.2, size=(25, 3)) for c in (0.0, 4.0, 9.0)]).astype(np.float32)
    d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)).astype(np.float32)
    d = (d * rng.uniform(0.98, 1.02, size=d.shape)).astype(np.float32)  # asymmetric like the DTW matrix
    np.fill_diagonal(d, 0.0)
    ops, clusters = run_both(d, 0.25)
    groups = AgglomerativeClustering.cluster_sets(ops, clusters, len(pts))
    assert sum(len(g) for g in groups) <= len(pts)
    assert any(o.operation == Merge.Sequence2Sequence for o in ops)
    assert any(o.operation == Merge.Cluster2Cluster for o in ops) or len(ops) < 3


def test_exact_ties_are_flagged_and_resolved_in_ascending_id():
    d = np.array([[0, 1, 1, 4], [1, 0, 1, 4], [1, 1, 0, 4], [4, 4, 4, 0]], dtype=np.float32)
    ops, _ = run_both(d, 0.9)
    assert (ops[0].merge_i, ops[0].merge_j) == (0, 1) and ops[0].tie


def test_inf_and_full_merge():
    rng = np.random.default_rng(11)
    n = 30
    d = rng.uniform(1, 2, size=(n, n)).astype(np.float32)
    np.fill_diagonal(d, 0.0)
    d[7, :] = np.inf
    d[:, 7] = np.inf
    d[7, 7] = 0.0
    run_both(d, 0.5)
    run_both(d, 0.97)  # threshold = +INF: merges until one cluster or a non-finite linkage


def test_percentile_out_of_bounds_is_an_error():
    d = np.ones((3, 3), dtype=np.float32)
    with pytest.raises(IndexError):
        AgglomerativeClustering.clustering(d.ravel(), 3, 1.0)


def test_threshold_override_and_larger_case():
    rng = np.random.default_rng(12)
    n = 140
    d = rng.gamma(3.0, 1.0, size=(n, n)).astype(np.float32)
    np.fill_diagonal(d, 0.0)
    want, thr, _ = oracle.upgma(d, 0.1)
    ops, _ = AgglomerativeClustering.clustering(d.ravel(), n, 0.1, threshold=thr)
    assert [(o.merge_i, o.merge_j, o.into) for o in ops] == [(a, b, k) for a, b, k, _, _ in want]
