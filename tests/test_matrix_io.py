"""Round trips of the matrix / path files (row f4): the payload is byte-for-byte the flat
Vec<f32> the reference hands to clustering(), so re-clustering from disk is identical."""
import numpy as np
import pytest

from audio_pattern_discovery_b200 import AgglomerativeClustering, matrix_io


def test_matrix_round_trip_and_recluster(tmp_path, capsys):
    rng = np.random.default_rng(1)
    n = 40
    d = rng.gamma(2.0, 1.0, size=(n, n)).astype(np.float32)
    np.fill_diagonal(d, 0.0)
    d[3, 5] = np.inf
    stem = str(tmp_path / "run1")
    payload = matrix_io.save_matrix(stem, d, n, {"warping_band_percentage": 0.1, "mode": "strict"})
    assert open(payload, "rb").read() == d.astype("<f4").tobytes()
    for mm in (False, True):
        got, n2, params = matrix_io.load_matrix(stem, mmap=mm)
        assert n2 == n and params["mode"] == "strict"
        assert np.array_equal(np.asarray(got).view(np.uint32), d.ravel().view(np.uint32))
    ops_a, cl_a = AgglomerativeClustering.clustering(d.ravel(), n, 0.2)
    ops_b, cl_b = AgglomerativeClustering.clustering(np.asarray(matrix_io.load_matrix(stem)[0]), n, 0.2)
    assert [(o.merge_i, o.merge_j, o.into) for o in ops_a] == [(o.merge_i, o.merge_j, o.into) for o in ops_b]
    assert cl_a == cl_b


def test_corruption_and_shape_errors(tmp_path):
    d = np.arange(9, dtype=np.float32)
    stem = str(tmp_path / "m")
    matrix_io.save_matrix(stem, d, 3)
    with pytest.raises(ValueError):
        matrix_io.save_matrix(stem + "x", d, 4)
    with open(stem + ".apdm", "r+b") as f:
        f.seek(8)
        f.write(b"\x01")
    with pytest.raises(ValueError):
        matrix_io.load_matrix(stem)
    assert matrix_io.load_matrix(stem, verify=False)[1] == 3


def test_paths_round_trip(tmp_path):
    pairs = [(0, 1), (4, 2)]
    scores = [np.float32(0.375), np.float32(np.inf)]
    paths = [np.array([[3, 3], [2, 2], [1, 1]], dtype=np.uint32), np.zeros((0, 2), dtype=np.uint32)]
    stem = str(tmp_path / "p")
    matrix_io.save_paths(stem, pairs, scores, paths)
    back = matrix_io.load_paths(stem)
    assert [b[0] for b in back] == pairs
    assert back[0][1] == np.float32(0.375) and np.isinf(back[1][1])
    assert np.array_equal(back[0][2], paths[0]) and back[1][2].shape == (0, 2)
