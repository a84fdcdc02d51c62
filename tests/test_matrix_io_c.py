"""Row f4 behind the C ABI: apd_save_matrix / apd_load_matrix / apd_save_paths / apd_load_paths
(csrc/matrix_io.cpp, host code) read and write the same files as matrix_io.py, in both directions."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from audio_pattern_discovery_b200 import _capi, matrix_io

_fp = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def _matrix(n, seed=0):
    rng = np.random.default_rng(seed)
    d = rng.random((n, n)).astype(np.float32)
    np.fill_diagonal(d, 0.0)
    if n > 3:
        d[1, 2] = np.inf
        d[2, 1] = np.nan
    return d


def test_c_writer_python_reader(tmp_path, apd_lib_path):
    L = _capi.lib()
    d = _matrix(17)
    stem = str(tmp_path / "m")
    st = L.apd_save_matrix(stem.encode(), d.ctypes.data_as(_fp), 17, json.dumps({"warping_band_percentage": 0.1}).encode())
    assert st == _capi.APD_OK
    got, n, params = matrix_io.load_matrix(stem)            # verifies the sha256 the C side wrote
    assert n == 17 and params == {"warping_band_percentage": 0.1}
    assert np.array_equal(got.view(np.uint32), d.reshape(-1).view(np.uint32))


def test_python_writer_c_reader(tmp_path, apd_lib_path):
    L = _capi.lib()
    d = _matrix(23, 1)
    stem = str(tmp_path / "p")
    matrix_io.save_matrix(stem, d, 23, {"a": {"n": 99, "s": 'x"y'}, "n": [1, 2]})   # a nested "n" must not confuse the reader
    n = C.c_uint32(0)
    assert L.apd_load_matrix(stem.encode(), None, 0, C.byref(n), 1) == _capi.APD_OK and n.value == 23   # size query
    out = np.empty(23 * 23, np.float32)
    assert L.apd_load_matrix(stem.encode(), out.ctypes.data_as(_fp), out.size, C.byref(n), 1) == _capi.APD_OK
    assert np.array_equal(out.view(np.uint32), d.reshape(-1).view(np.uint32))
    # too small a buffer, a corrupted payload, a missing file: errors, never a partial result
    assert L.apd_load_matrix(stem.encode(), out.ctypes.data_as(_fp), 10, C.byref(n), 1) == _capi.APD_ERR_INVALID
    with open(stem + ".apdm", "r+b") as f:
        f.seek(40)
        f.write(b"\x01\x02\x03\x04")
    assert L.apd_load_matrix(stem.encode(), out.ctypes.data_as(_fp), out.size, C.byref(n), 1) == _capi.APD_ERR_INVALID
    assert b"checksum" in L.apd_last_error(None)
    assert L.apd_load_matrix(stem.encode(), out.ctypes.data_as(_fp), out.size, C.byref(n), 0) == _capi.APD_OK
    assert L.apd_load_matrix((stem + "_missing").encode(), None, 0, C.byref(n), 1) == _capi.APD_ERR_INVALID


def test_empty_matrix(tmp_path, apd_lib_path):
    L = _capi.lib()
    stem = str(tmp_path / "e")
    assert L.apd_save_matrix(stem.encode(), None, 0, None) == _capi.APD_OK
    got, n, params = matrix_io.load_matrix(stem)
    assert n == 0 and got.size == 0 and params == {}


def test_paths_both_directions(tmp_path, apd_lib_path):
    L = _capi.lib()
    pairs = np.array([[0, 3], [5, 1], [2, 2]], np.uint32)
    scores = np.array([0.375, np.inf, np.nan], np.float32)
    cap = 6
    paths = np.zeros((3, cap, 2), np.uint32)
    paths[0, :4] = [[4, 5], [3, 4], [2, 2], [1, 1]]
    paths[1, :2] = [[2, 1], [1, 1]]
    lens = np.array([4, 2, 0], np.uint64)
    stem = str(tmp_path / "w")
    assert L.apd_save_paths(stem.encode(), pairs.ctypes.data_as(_u32p), 3, scores.ctypes.data_as(_fp),
                            paths.ctypes.data_as(_u32p), cap, lens.ctypes.data_as(_u64p)) == _capi.APD_OK
    back = matrix_io.load_paths(stem)
    assert [b[0] for b in back] == [(0, 3), (5, 1), (2, 2)]
    assert back[0][1] == np.float32(0.375) and np.isinf(back[1][1]) and np.isnan(back[2][1])
    assert np.array_equal(back[0][2], paths[0, :4]) and np.array_equal(back[1][2], paths[1, :2]) and len(back[2][2]) == 0
    # Python writer -> C reader
    stem2 = str(tmp_path / "r")
    matrix_io.save_paths(stem2, [(7, 8), (9, 10)], [np.float32(1.0) / np.float32(3.0), np.float32(-np.inf)],
                         [np.array([[3, 3], [2, 2], [1, 1]]), np.zeros((0, 2))])
    n = C.c_uint64(0)
    pij = np.zeros((4, 2), np.uint32)
    sc = np.zeros(4, np.float32)
    pl = np.zeros(4, np.uint64)
    pp = np.zeros((4, 5, 2), np.uint32)
    assert L.apd_load_paths(stem2.encode(), pij.ctypes.data_as(_u32p), sc.ctypes.data_as(_fp), pl.ctypes.data_as(_u64p), 4,
                            pp.ctypes.data_as(_u32p), 5, C.byref(n)) == _capi.APD_OK
    assert n.value == 2 and pij[:2].tolist() == [[7, 8], [9, 10]] and pl[:2].tolist() == [3, 0]
    assert sc[0].view(np.uint32) == (np.float32(1.0) / np.float32(3.0)).view(np.uint32) and sc[1] == -np.inf
    assert pp[0, :3].tolist() == [[3, 3], [2, 2], [1, 1]]
    # count-only query
    assert L.apd_load_paths(stem2.encode(), None, None, None, 0, None, 0, C.byref(n)) == _capi.APD_OK and n.value == 2
