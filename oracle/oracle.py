"""ctypes binding of oracle/libapd_oracle.so -- TEST INFRASTRUCTURE ONLY.

Builds the library on first use with oracle/Makefile (gcc).  "PARITY UNPINNED" by
the reference's own tests (it has none); see apd_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libapd_oracle.so")
_lib = None


class Params(C.Structure):
    _fields_ = [("warping_band", C.c_uint64), ("insertion_penalty", C.c_float),
                ("deletion_penalty", C.c_float), ("match_penalty", C.c_float)]


class Merge(C.Structure):
    _fields_ = [("merge_i", C.c_uint32), ("merge_j", C.c_uint32), ("into", C.c_uint32),
                ("distance", C.c_float), ("tie", C.c_uint32)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("apd_oracle.c", "apd_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)
             or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L.apd_oracle_euclidean.restype = C.c_float
        L.apd_oracle_euclidean.argtypes = [fp, fp, C.c_size_t]
        L.apd_oracle_warping_band.restype = C.c_uint64
        L.apd_oracle_warping_band.argtypes = [C.c_float, C.c_uint64]
        L.apd_oracle_window.restype = C.c_uint64
        L.apd_oracle_window.argtypes = [C.c_uint64] * 3
        L.apd_oracle_cells_visited.restype = C.c_uint64
        L.apd_oracle_cells_visited.argtypes = [C.c_uint64] * 3
        L.apd_oracle_dtw_literal.restype = C.c_float
        L.apd_oracle_dtw_literal.argtypes = [fp, C.c_uint64, fp, C.c_uint64, C.c_uint64,
                                             C.POINTER(Params), u32p, C.c_uint64,
                                             C.POINTER(C.c_uint64)]
        L.apd_oracle_dtw_dense.restype = C.c_float
        L.apd_oracle_dtw_dense.argtypes = [fp, C.c_uint64, fp, C.c_uint64, C.c_uint64,
                                           C.POINTER(Params)]
        L.apd_oracle_align_all.restype = C.c_int
        L.apd_oracle_align_all.argtypes = [C.POINTER(fp), u32p, C.c_uint32, C.c_uint32,
                                           C.c_float, C.c_float, C.c_float, C.c_float,
                                           C.c_uint32, C.c_int, fp]
        L.apd_oracle_align_pairs.restype = C.c_int
        L.apd_oracle_align_pairs.argtypes = [C.POINTER(fp), u32p, C.c_uint32, C.c_uint32,
                                             C.c_float, C.c_float, C.c_float, C.c_float,
                                             u32p, C.c_uint64, C.c_uint32, C.c_int, fp]
        L.apd_oracle_percentile.restype = C.c_int
        L.apd_oracle_percentile.argtypes = [fp, C.c_uint64, C.c_float, fp]
        L.apd_oracle_upgma.restype = C.c_int
        L.apd_oracle_upgma.argtypes = [fp, C.c_uint32, C.c_float, C.POINTER(Merge), u32p, fp, u32p]
        L.apd_oracle_ae_encode.restype = None
        L.apd_oracle_ae_encode.argtypes = [fp, C.c_uint64, C.c_uint32, fp, fp, C.c_uint32, fp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _seq(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return a


def warping_band(pct, length):
    return int(lib().apd_oracle_warping_band(pct, length))


def window(band, n, m):
    return int(lib().apd_oracle_window(band, n, m))


def cells_visited(n, m, w):
    return int(lib().apd_oracle_cells_visited(n, m, w))


def pair_cells(n, m, pct):
    """Reference cell updates for one ordered pair."""
    return cells_visited(n, m, window(warping_band(pct, max(n, m)), n, m))


def dtw(x, y, pct, ins=1.0, dele=1.0, mat=1.0, variant="literal", band=None, want_path=False):
    """One ordered pair as the reference's pair loop computes it."""
    x, y = _seq(x), _seq(y)
    n, m, d = x.shape[0], y.shape[0], x.shape[1]
    if band is None:
        band = warping_band(pct, max(n, m))
    p = Params(band, ins, dele, mat)
    if variant == "dense":
        return np.float32(lib().apd_oracle_dtw_dense(_fp(x), n, _fp(y), m, d, C.byref(p)))
    if not want_path:
        return np.float32(lib().apd_oracle_dtw_literal(_fp(x), n, _fp(y), m, d, C.byref(p),
                                                       None, 0, None))
    cap = n + m + 2
    buf = np.zeros(2 * cap, dtype=np.uint32)
    plen = C.c_uint64(0)
    s = lib().apd_oracle_dtw_literal(_fp(x), n, _fp(y), m, d, C.byref(p),
                                     buf.ctypes.data_as(C.POINTER(C.c_uint32)), cap,
                                     C.byref(plen))
    return np.float32(s), buf[:2 * plen.value].reshape(-1, 2).copy()


def _pack(seqs):
    seqs = [_seq(s) for s in seqs]
    n = len(seqs)
    dim = seqs[0].shape[1] if n else 1
    ptrs = (C.POINTER(C.c_float) * max(n, 1))()
    for k, s in enumerate(seqs):
        ptrs[k] = _fp(s)
    lens = np.array([s.shape[0] for s in seqs], dtype=np.uint32)
    return seqs, ptrs, lens, dim


def align_all(seqs, pct, ins=1.0, dele=1.0, mat=1.0, workers=4, variant="literal"):
    """AlignmentWorkers::align_all -> (n, n) float32, diagonal 0."""
    seqs, ptrs, lens, dim = _pack(seqs)
    n = len(seqs)
    out = np.zeros((n, n), dtype=np.float32)
    rc = lib().apd_oracle_align_all(ptrs, lens.ctypes.data_as(C.POINTER(C.c_uint32)), n, dim,
                                    pct, ins, dele, mat, workers,
                                    0 if variant == "literal" else 1, _fp(out))
    if rc:
        raise RuntimeError("apd_oracle_align_all failed: %d" % rc)
    return out


def align_pairs(seqs, pairs, pct, ins=1.0, dele=1.0, mat=1.0, workers=4, variant="dense"):
    """Scores of an explicit list of ordered pairs [(i, j), ...]."""
    seqs, ptrs, lens, dim = _pack(seqs)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    out = np.zeros(len(pairs), dtype=np.float32)
    rc = lib().apd_oracle_align_pairs(ptrs, lens.ctypes.data_as(C.POINTER(C.c_uint32)),
                                      len(seqs), dim, pct, ins, dele, mat,
                                      pairs.ctypes.data_as(C.POINTER(C.c_uint32)), len(pairs),
                                      workers, 0 if variant == "literal" else 1, _fp(out))
    if rc:
        raise RuntimeError("apd_oracle_align_pairs failed: %d" % rc)
    return out


def percentile(x, perc):
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    out = C.c_float(0)
    rc = lib().apd_oracle_percentile(_fp(x), x.size, perc, C.byref(out))
    if rc:
        raise IndexError("percentile index out of bounds (the reference panics here)")
    return np.float32(out.value)


def upgma(dist, perc):
    """-> (merges [(p, q, k, distance, tie)], threshold, assignment[n])."""
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    n = dist.shape[0]
    ops = (Merge * max(n, 1))()
    n_ops = C.c_uint32(0)
    thr = C.c_float(0)
    assign = np.zeros(max(n, 1), dtype=np.uint32)
    rc = lib().apd_oracle_upgma(_fp(dist), n, perc, ops, C.byref(n_ops), C.byref(thr),
                                assign.ctypes.data_as(C.POINTER(C.c_uint32)))
    if rc:
        raise RuntimeError("apd_oracle_upgma failed: %d" % rc)
    merges = [(o.merge_i, o.merge_j, o.into, np.float32(o.distance), int(o.tie))
              for o in ops[:n_ops.value]]
    return merges, np.float32(thr.value), assign[:n].copy()


def ae_encode(frames, w_encode, b_encode):
    """NDSequence::encoded (src/spectrogram.rs:103-121) with AutoEncoder::predict per frame
    (src/neural.rs:55-71): (T, n_bins) -> (T, n_latent) float32."""
    x = _seq(frames)
    w = np.ascontiguousarray(w_encode, dtype=np.float32)
    b = np.ascontiguousarray(b_encode, dtype=np.float32).ravel()
    n_bins, n_latent = w.shape
    assert x.shape[1] == n_bins and b.size == n_latent
    out = np.zeros((x.shape[0], n_latent), dtype=np.float32)
    lib().apd_oracle_ae_encode(_fp(x), x.shape[0], n_bins, _fp(w), _fp(b), n_latent, _fp(out))
    return out
