"""Host-side mirror of the reference's src/alignments.rs on top of libapd_b200.

Same names, argument meaning and failure behaviour as the Rust surface
(file:line relative to the reference repository):

    AlignmentWorkers::new(data)            src/alignments.rs:17-26
    AlignmentWorkers::align_all(&discover) src/alignments.rs:31-67
    AlignmentWorkers.result                src/alignments.rs:13   (n*n row-major f32)
    AlignmentParams{..}, ::default(len)    src/alignments.rs:77-94
    Alignment::new / construct_alignment / score   src/alignments.rs:106-180

Everything numeric runs in the CUDA library through the C ABI of include/apd.h;
this file is plumbing only and fails loudly if the library or a GPU is missing.
"""
import ctypes as C
import threading

import numpy as np

from . import _capi
from ._capi import APD_MODE_FAST, APD_MODE_STRICT, ApdError, apd_params, apd_stats  # noqa: F401
from .spectrogram import NDSequence

_u32p = C.POINTER(C.c_uint32)
_fp = C.POINTER(C.c_float)


class AlignmentParams:
    """src/alignments.rs:77-83"""

    def __init__(self, warping_band, insertion_penalty=1.0, deletion_penalty=1.0, match_penalty=1.0):
        self.warping_band = int(warping_band)
        self.insertion_penalty = float(insertion_penalty)
        self.deletion_penalty = float(deletion_penalty)
        self.match_penalty = float(match_penalty)

    @staticmethod
    def default(len):  # noqa: A002  (the reference's argument name)
        """src/alignments.rs:86-93"""
        return AlignmentParams(len, 1.0, 1.0, 1.0)

    def __repr__(self):
        return ("AlignmentParams { warping_band: %d, insertion_penalty: %r, deletion_penalty: %r, "
                "match_penalty: %r }" % (self.warping_band, self.insertion_penalty,
                                         self.deletion_penalty, self.match_penalty))


def _c_params(pct, ins, dele, mat, mode):
    return apd_params(float(pct), float(ins), float(dele), float(mat), int(mode))


def visible_devices():
    """Number of CUDA devices the library sees (raises without one: there is no CPU path)."""
    lib = _capi.lib()
    n = C.c_int(0)
    st = lib.apd_device_count(C.byref(n))
    if st != _capi.APD_OK:
        msg = lib.apd_last_error(None)
        raise ApdError(st, msg.decode() if msg else "")
    return n.value


class Context:
    """An apd_ctx: the packed sequence arena on one GPU (device=k), or -- from this one
    process -- on a group of GPUs (devices=[...] or devices="all", apd_create_multi) that
    shares the pair space and assembles the matrix over NVLink inside the library."""

    def __init__(self, device=0, devices=None):
        self._lib = _capi.lib()
        h = C.c_void_p()
        if devices is None:
            st = self._lib.apd_create(int(device), C.byref(h))
            self.devices = [int(device)]
        elif isinstance(devices, str):
            if devices != "all":
                raise ValueError('devices must be a list of device ids or "all"')
            st = self._lib.apd_create_multi(None, 0, C.byref(h))
            self.devices = None
        else:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            st = self._lib.apd_create_multi(ids, len(devices), C.byref(h))
            self.devices = [int(d) for d in devices]
        if st != _capi.APD_OK:
            msg = self._lib.apd_last_error(None)
            raise ApdError(st, msg.decode() if msg else "")
        self._h = h
        g, ps = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.apd_group_size(self._h, C.byref(g), C.byref(ps)))
        self.group_size, self.peer_stores = g.value, bool(ps.value)
        if self.devices is None:
            self.devices = list(range(self.group_size))
        self.device = self.devices[0]
        self.n = 0
        self.dim = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.apd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st):
        _capi.check(self._h, st)

    # -- sequence packing ------------------------------------------------------
    def set_sequences(self, seqs, dim=None):
        """seqs: list of (T, D) float32 arrays (or NDSequence).  Copies them to the device."""
        arrs = []
        for s in seqs:
            if isinstance(s, NDSequence):
                a = s.as_array()
            else:
                a = np.ascontiguousarray(s, dtype=np.float32)
                if a.ndim == 1:
                    a = a.reshape(-1, 1)
            arrs.append(a)
        n = len(arrs)
        if dim is None:
            dim = arrs[0].shape[1] if n else 1
        for a in arrs:
            if a.shape[0] and a.shape[1] != dim:
                raise ValueError("all sequences must share the frame width")
        # one pointer per sequence, built without a ctypes object per element
        addr = np.fromiter((a.__array_interface__["data"][0] for a in arrs), dtype=np.uintp, count=n)
        if n == 0:
            addr = np.zeros(1, dtype=np.uintp)
        ptrs = addr.ctypes.data_as(C.POINTER(_fp))
        lens = np.fromiter((a.shape[0] for a in arrs), dtype=np.uint32, count=n)
        self._check(self._lib.apd_set_sequences(self._h, ptrs, lens.ctypes.data_as(_u32p), n, dim))
        self._keepalive = arrs  # the library copies before returning; kept only until the next call
        self.n, self.dim = n, dim

    def set_sequences_encoded(self, cepstra, w_encode, b_encode):
        """NDSequence::encoded on the device (src/spectrogram.rs:103-121, src/neural.rs:55-71):
        cepstra = list of (T, n_bins) float32 arrays, w_encode (n_bins, n_latent), b_encode
        (n_latent,).  The embeddings are written into the arena; frame width becomes n_latent."""
        arrs = [s.as_array() if isinstance(s, NDSequence) else np.ascontiguousarray(s, dtype=np.float32) for s in cepstra]
        w = np.ascontiguousarray(w_encode, dtype=np.float32)
        b = np.ascontiguousarray(b_encode, dtype=np.float32).ravel()
        if w.ndim != 2 or b.size != w.shape[1]:
            raise ValueError("w_encode must be (n_bins, n_latent) and b_encode (n_latent,)")
        n_bins, n_latent = w.shape
        for a in arrs:
            if a.ndim != 2 or (a.shape[0] and a.shape[1] != n_bins):
                raise ValueError("every cepstrum must be (T, n_bins)")
        n = len(arrs)
        addr = np.fromiter((a.__array_interface__["data"][0] for a in arrs), dtype=np.uintp, count=n)
        if n == 0:
            addr = np.zeros(1, dtype=np.uintp)
        lens = np.fromiter((a.shape[0] for a in arrs), dtype=np.uint32, count=n)
        self._check(self._lib.apd_set_sequences_encoded(
            self._h, addr.ctypes.data_as(C.POINTER(_fp)), lens.ctypes.data_as(_u32p), n, n_bins,
            w.ctypes.data_as(_fp), b.ctypes.data_as(_fp), n_latent))
        self._lens = lens
        self.n, self.dim = n, n_latent

    def get_sequence(self, index, length):
        """One sequence of the device arena, (length, dim) float32 (what the kernels read)."""
        out = np.empty((int(length), self.dim), dtype=np.float32)
        self._check(self._lib.apd_get_sequence(self._h, int(index), out.ctypes.data_as(_fp), out.size))
        return out

    def set_sequences_layout(self, lens, dim):
        """Declares a batch by its lengths and frame width only (one-process-per-GPU jobs: another rank uploads the
        frames and this one receives the packed arena over NVLink, see arena_device / arena_commit)."""
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        self._check(self._lib.apd_set_sequences_layout(self._h, lens.ctypes.data_as(_u32p), len(lens), int(dim)))
        self.n, self.dim = len(lens), int(dim)

    def arena_device(self):
        """-> (device pointer, floats) of the packed arena."""
        ptr, n = C.c_void_p(0), C.c_uint64(0)
        self._check(self._lib.apd_arena_device(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value or 0, n.value

    def arena_commit(self):
        self._check(self._lib.apd_arena_commit(self._h))

    def set_sequences_flat(self, flat, offsets, lens, dim):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        self._check(self._lib.apd_set_sequences_flat(
            self._h, flat.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.POINTER(C.c_uint64)),
            lens.ctypes.data_as(_u32p), len(lens), dim))
        self.n, self.dim = len(lens), dim

    def set_shard(self, rank, world):
        self._check(self._lib.apd_set_shard(self._h, rank, world))

    # -- the hot path ----------------------------------------------------------
    def align_all(self, pct, ins=1.0, dele=1.0, mat=1.0, mode=APD_MODE_STRICT, out=None):
        """Host-buffer call: n x n float32 (diag 0), host<->device copies included."""
        n = self.n
        if out is None:
            out = np.empty((n, n), dtype=np.float32)
        assert out.dtype == np.float32 and out.size == n * n and out.flags.c_contiguous
        p = _c_params(pct, ins, dele, mat, mode)
        self._check(self._lib.apd_align_all(self._h, C.byref(p), out.ctypes.data_as(C.c_void_p)))
        return out

    def packed_len(self, pct, mode=APD_MODE_STRICT):
        p = _c_params(pct, 1.0, 1.0, 1.0, mode)
        v = C.c_uint64(0)
        self._check(self._lib.apd_packed_len(self._h, C.byref(p), C.byref(v)))
        return v.value

    def align_packed(self, pct, ins, dele, mat, mode, d_packed_ptr, stream=0):
        """Enqueues this shard's DTW kernels on `stream` (a cudaStream_t handle).  stream=0 is NOT the legacy
        default stream: it selects the context's own private non-blocking stream, which has no implicit ordering
        with torch's or any other stream -- pass your own stream (as ShardedAligner does) or call synchronize()
        before another stream touches the buffers.  Calls on one context are ordered among themselves on the
        device whatever streams they name."""
        p = _c_params(pct, ins, dele, mat, mode)
        self._check(self._lib.apd_align_packed(self._h, C.byref(p), C.c_void_p(d_packed_ptr),
                                               C.c_void_p(stream)))

    def scatter_packed(self, d_gathered_ptr, world, d_out_ptr, stream=0):
        """Expands `world` gathered shards into the n x n device matrix on `stream` (0 = the context's own
        private stream, see align_packed)."""
        self._check(self._lib.apd_scatter_packed(self._h, C.c_void_p(d_gathered_ptr), world,
                                                 C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def synchronize(self, stream=0):
        """Waits for `stream` (0 = the context's own), collects kernel timings, checks the device error flag."""
        self._check(self._lib.apd_synchronize(self._h, C.c_void_p(stream)))

    def align_pairs(self, pairs, pct=1.0, ins=1.0, dele=1.0, mat=1.0, mode=APD_MODE_STRICT,
                    want_paths=False, path_cap=None, warping_band=None):
        """Scores (and warping paths) of explicit ordered pairs [(i, j), ...]."""
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        k = len(pairs)
        scores = np.empty(k, dtype=np.float32)
        p = _c_params(pct, ins, dele, mat, mode)
        paths = None
        lens = np.zeros(max(k, 1), dtype=np.uint64)
        paths_ptr = None
        cap = 0
        if want_paths:
            cap = int(path_cap) if path_cap is not None else 0
            paths = np.zeros((k, max(cap, 1), 2), dtype=np.uint32)
            paths_ptr = paths.ctypes.data_as(_u32p)
        args = [pairs.ctypes.data_as(_u32p), k, scores.ctypes.data_as(_fp), paths_ptr, cap,
                lens.ctypes.data_as(C.POINTER(C.c_uint64))]
        if warping_band is None:
            st = self._lib.apd_align_pairs(self._h, C.byref(p), *args)
        else:
            st = self._lib.apd_align_pairs_band(self._h, C.byref(p), int(warping_band), *args)
        self._check(st)
        if not want_paths:
            return scores
        return scores, [paths[q, :min(int(lens[q]), cap)].copy() for q in range(k)], lens[:k].copy()

    def percentile(self, perc):
        """numerics::percentile (src/numerics.rs:125-133) of the matrix the last align_all left on
        the device -- the clustering threshold of src/clustering.rs:101, without a host sort."""
        out = C.c_float(0)
        self._check(self._lib.apd_percentile_matrix(self._h, float(perc), C.byref(out)))
        return np.float32(out.value)

    def percentile_device(self, d_ptr, length, perc, stream=0):
        out = C.c_float(0)
        self._check(self._lib.apd_percentile_device(self._h, C.c_void_p(d_ptr), int(length), float(perc),
                                                    C.c_void_p(stream), C.byref(out)))
        return np.float32(out.value)

    def launch_plan(self):
        """Launch classes of the last DTW enqueue (ring home, ring height, units, grid)."""
        import json
        return json.loads(self._lib.apd_last_launch_plan(self._h).decode())

    def stats(self):
        s = apd_stats()
        self._check(self._lib.apd_get_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in apd_stats._fields_}


class _Mutex:
    """Stand-in for Arc<Mutex<Vec<f32>>> (src/alignments.rs:13): `.lock().unwrap()`
    yields the flat result vector, as at the call site src/main.rs:194-195."""

    class _Guard:
        def __init__(self, owner):
            self._owner = owner

        def unwrap(self):
            return self._owner._value

    def __init__(self, value):
        self._value = value
        self._lock = threading.Lock()

    def lock(self):
        return _Mutex._Guard(self)


class AlignmentWorkers:
    """Aligns all sequences and saves the results in a flat matrix (src/alignments.rs:11-68).

    `alignment_workers` of the Discovery config is accepted and ignored: the work is
    spread over the GPU by the library, not over host threads.
    """

    def __init__(self, data, device=None, mode=APD_MODE_STRICT, devices=None):
        self.data = list(data)
        n = len(self.data)
        self.result = _Mutex(np.zeros(n * n, dtype=np.float32))  # diag stays 0.0 (src/alignments.rs:20-23,51)
        self.mode = mode
        # Like the reference's one blocking call from one process (src/main.rs:189-195), but
        # spread over every visible GPU unless the caller names one (device=k) or some (devices=[..]).
        if device is not None:
            self._ctx = Context(device)
        else:
            self._ctx = Context(devices="all" if devices is None else devices)
        self._ctx.set_sequences(self.data)

    @staticmethod
    def new(data, device=None, mode=APD_MODE_STRICT, devices=None):
        return AlignmentWorkers(data, device, mode, devices)

    @staticmethod
    def new_encoded(cepstra, w_encode, b_encode, device=None, mode=APD_MODE_STRICT, devices=None):
        """`NDSequence::new(..).encoded(&nn)` of src/main.rs:150-161 moved onto the device: `cepstra` are the
        raw (T, n_bins) sequences, the auto-encoder (src/neural.rs:55-71) runs on the GPU and the embeddings
        go straight into the arena.  `data` keeps the cepstra."""
        w = AlignmentWorkers.__new__(AlignmentWorkers)
        w.data = list(cepstra)
        n = len(w.data)
        w.result = _Mutex(np.zeros(n * n, dtype=np.float32))
        w.mode = mode
        w._ctx = Context(device) if device is not None else Context(devices="all" if devices is None else devices)
        w._ctx.set_sequences_encoded(w.data, w_encode, b_encode)
        return w

    def align_all(self, params):
        """params: a Discovery (src/discovery.rs:7-26).  Blocking; fills self.result."""
        n = len(self.data)
        if params.alignment_workers == 0:
            # src/alignments.rs:33 divides by alignment_workers: the reference panics.
            raise ZeroDivisionError("attempt to divide by zero (alignment_workers == 0)")
        out = self.result.lock().unwrap()
        self._ctx.align_all(params.warping_band_percentage, params.insertion_penalty,
                            params.deletion_penalty, params.match_penalty, self.mode,
                            out=out.reshape(n, n))

    def stats(self):
        return self._ctx.stats()


class Alignment:
    """One ordered pair (src/alignments.rs:99-181).  `sparse` is not materialised on the
    host; `path` holds the device-traced warping path instead."""

    def __init__(self, device=0, mode=APD_MODE_STRICT):
        self.n = 0
        self.m = 0
        self._score = None
        self.path = None
        self._device = device
        self._mode = mode

    @staticmethod
    def new(device=0, mode=APD_MODE_STRICT):
        return Alignment(device, mode)

    def construct_alignment(self, x, y, params, want_path=True):
        xa = x.as_array() if isinstance(x, NDSequence) else np.ascontiguousarray(x, dtype=np.float32)
        ya = y.as_array() if isinstance(y, NDSequence) else np.ascontiguousarray(y, dtype=np.float32)
        if xa.ndim == 1:
            xa = xa.reshape(-1, 1)
        if ya.ndim == 1:
            ya = ya.reshape(-1, 1)
        self.n, self.m = xa.shape[0], ya.shape[0]
        with Context(self._device) as ctx:
            ctx.set_sequences([xa, ya], dim=xa.shape[1] if xa.shape[0] else ya.shape[1])
            cap = self.n + self.m + 2
            scores, paths, _ = ctx.align_pairs([(0, 1)], 0.0, params.insertion_penalty,
                                               params.deletion_penalty, params.match_penalty,
                                               self._mode, want_paths=True, path_cap=cap if want_path else 0,
                                               warping_band=params.warping_band)
        self._score = np.float32(scores[0])
        self.path = paths[0] if want_path else None

    def score(self):
        """src/alignments.rs:116-125"""
        if self._score is None:
            return np.float32(np.inf)  # Alignment::new() has n == m == 0
        return self._score
