"""Builds and binds the host-side schedule emulator (tests/emul/dtw_emul.cpp) -- TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "audio_pattern_discovery_b200", "csrc")
_OUT = os.path.join(_ROOT, "build", "libapd_emul.so")
_lib = None


def build():
    srcs = [os.path.join(_HERE, "dtw_emul.cpp"), os.path.join(_CSRC, "host_plan.cpp")]
    deps = srcs + [os.path.join(_CSRC, "dtw_core.h"), os.path.join(_CSRC, "host_plan.h")]
    if os.path.exists(_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(_OUT) for d in deps):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread",
           "-o", _OUT] + srcs
    subprocess.run(cmd, check=True, capture_output=True)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        L.apd_emul_align_all.restype = C.c_int
        L.apd_emul_align_all.argtypes = [C.POINTER(fp), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32,
                                         C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                         C.c_uint32, C.c_uint32, fp, C.POINTER(C.c_uint64)]
        L.apd_emul_plan_info.restype = C.c_int
        L.apd_emul_plan_info.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_float, C.c_uint32,
                                         C.c_uint32, C.POINTER(C.c_uint64)]
        L.apd_emul_cells_visited.restype = C.c_uint64
        L.apd_emul_cells_visited.argtypes = [C.c_uint64] * 3
        L.apd_emul_window.restype = C.c_int
        L.apd_emul_window.argtypes = [C.c_float, C.c_int, C.c_int]
        L.apd_emul_plan_units.restype = C.c_uint64
        L.apd_emul_plan_units.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_float, C.c_uint32,
                                          C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64)]
        L.apd_emul_tile_cols.restype = None
        L.apd_emul_tile_cols.argtypes = [C.c_int]
        L.apd_emul_force_rho.restype = None
        L.apd_emul_force_rho.argtypes = [C.c_int, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def tile_cols(tc):
    """Tile width of the emulated lane program: 4 (8-warp kernels) or 2 (12-warp kernel)."""
    lib().apd_emul_tile_cols(int(tc))


def force_rho(rho):
    """Row grid (0..3) for every unit, -1 = the kernel's own choice.  Returns the number of units
    that ran on each grid since the previous call."""
    used = np.zeros(4, dtype=np.uint64)
    lib().apd_emul_force_rho(int(rho), used.ctypes.data_as(C.POINTER(C.c_uint64)))
    return used


def align_all(seqs, pct, ins=1.0, dele=1.0, mat=1.0, strict=True, rank=0, world=1):
    seqs = [np.ascontiguousarray(s, dtype=np.float32) for s in seqs]
    seqs = [s.reshape(-1, 1) if s.ndim == 1 else s for s in seqs]
    n = len(seqs)
    dim = seqs[0].shape[1] if n else 1
    fp = C.POINTER(C.c_float)
    ptrs = (fp * max(n, 1))()
    for k, s in enumerate(seqs):
        ptrs[k] = s.ctypes.data_as(fp)
    lens = np.array([s.shape[0] for s in seqs], dtype=np.uint32)
    out = np.zeros((n, n), dtype=np.float32)
    info = np.zeros(16, dtype=np.uint64)
    rc = lib().apd_emul_align_all(ptrs, lens.ctypes.data_as(C.POINTER(C.c_uint32)), n, dim, pct, ins,
                                  dele, mat, int(strict), rank, world, out.ctypes.data_as(fp),
                                  info.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc:
        raise RuntimeError("emulator failed: %d" % rc)
    return out, info


def plan_info(lens, dim, pct, rank=0, world=1):
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    info = np.zeros(16, dtype=np.uint64)
    rc = lib().apd_emul_plan_info(lens.ctypes.data_as(C.POINTER(C.c_uint32)), len(lens), dim, pct, rank, world,
                                  info.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc:
        raise RuntimeError("planner check failed: %d" % rc)
    return info


def plan_units(lens, dim, pct, row_block=32):
    """The planner's ordered unit list as (a, B) rows, plus units per launch class."""
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    n = len(lens)
    cap = n * ((n + 31) // 32) + 1
    out = np.zeros(cap, dtype=np.uint64)
    info = np.zeros(4, dtype=np.uint64)
    k = lib().apd_emul_plan_units(lens.ctypes.data_as(C.POINTER(C.c_uint32)), n, dim, pct, row_block,
                                  out.ctypes.data_as(C.POINTER(C.c_uint64)), cap, info.ctypes.data_as(C.POINTER(C.c_uint64)))
    u = out[:k]
    return np.stack([(u >> np.uint64(32)).astype(np.int64), (u & np.uint64(0xffffffff)).astype(np.int64)], axis=1), info
