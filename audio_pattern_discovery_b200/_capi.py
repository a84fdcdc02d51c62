"""ctypes binding of libapd_b200.so (the C ABI declared in include/apd.h).

The library holds hand-written CUDA for sm_100a only.  There is no CPU path:
if the shared object is missing this module raises, and apd_create() itself
fails without a CUDA device.  Build it in-tree with

    python -m audio_pattern_discovery_b200.build
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# APD_LIB_PATH selects another build of the same library (kernel experiments: build.py --variant)
LIB_PATH = os.environ.get("APD_LIB_PATH") or os.path.join(_HERE, "libapd_b200.so")

APD_OK = 0
APD_ERR_INVALID = 1
APD_ERR_NO_DEVICE = 2
APD_ERR_CUDA = 3
APD_ERR_UNSUPPORTED = 4
APD_ERR_STATE = 5
APD_ERR_INTERNAL = 6

APD_MODE_STRICT = 0
APD_MODE_FAST = 1
APD_MAX_DIM = 32
APD_ABI_VERSION = 2
APD_MAX_DEVICES = 8
APD_AE_MAX_BINS = 64

STATUS_NAMES = {0: "APD_OK", 1: "APD_ERR_INVALID", 2: "APD_ERR_NO_DEVICE", 3: "APD_ERR_CUDA",
                4: "APD_ERR_UNSUPPORTED", 5: "APD_ERR_STATE", 6: "APD_ERR_INTERNAL"}


class ApdError(RuntimeError):
    def __init__(self, status, message):
        self.status = status
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, str(status)), message))


class apd_params(C.Structure):
    _fields_ = [("warping_band_percentage", C.c_float),
                ("insertion_penalty", C.c_float),
                ("deletion_penalty", C.c_float),
                ("match_penalty", C.c_float),
                ("mode", C.c_uint32)]


class apd_stats(C.Structure):
    _fields_ = [("n_sequences", C.c_uint64),
                ("ordered_pairs", C.c_uint64),
                ("units_total", C.c_uint64),
                ("units_local", C.c_uint64),
                ("cells_reference", C.c_uint64),
                ("cells_computed", C.c_uint64),
                ("kernel_launches", C.c_uint32),
                ("kernel_ms", C.c_float),
                ("scatter_ms", C.c_float),
                ("h2d_ms", C.c_float),
                ("d2h_ms", C.c_float),
                ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64),
                ("sm_clock_mhz", C.c_float),
                ("sm_count", C.c_uint32),
                ("select_ms", C.c_float),
                ("path_ms", C.c_float)]


class apd_merge(C.Structure):
    _fields_ = [("merge_i", C.c_uint32), ("merge_j", C.c_uint32), ("into", C.c_uint32),
                ("distance", C.c_float), ("operation", C.c_uint32), ("tie", C.c_uint32)]


# Every symbol include/apd.h declares: name -> (restype, argtypes).
_fp = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_pp = C.POINTER(apd_params)
PROTOTYPES = {
    "apd_abi_version": (C.c_uint32, []),
    "apd_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "apd_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "apd_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "apd_group_size": (C.c_int, [C.c_void_p, _u32p, _u32p]),
    "apd_destroy": (None, [C.c_void_p]),
    "apd_last_error": (C.c_char_p, [C.c_void_p]),
    "apd_set_sequences": (C.c_int, [C.c_void_p, C.POINTER(_fp), _u32p, C.c_uint32, C.c_uint32]),
    "apd_set_sequences_flat": (C.c_int, [C.c_void_p, C.c_void_p, _u64p, _u32p, C.c_uint32, C.c_uint32]),
    "apd_set_sequences_encoded": (C.c_int, [C.c_void_p, C.POINTER(_fp), _u32p, C.c_uint32, C.c_uint32, _fp, _fp,
                                           C.c_uint32]),
    "apd_get_sequence": (C.c_int, [C.c_void_p, C.c_uint32, _fp, C.c_uint64]),
    "apd_set_sequences_layout": (C.c_int, [C.c_void_p, _u32p, C.c_uint32, C.c_uint32]),
    "apd_arena_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), _u64p]),
    "apd_arena_commit": (C.c_int, [C.c_void_p]),
    "apd_set_shard": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "apd_align_all": (C.c_int, [C.c_void_p, _pp, C.c_void_p]),
    "apd_packed_len": (C.c_int, [C.c_void_p, _pp, _u64p]),
    "apd_align_packed": (C.c_int, [C.c_void_p, _pp, C.c_void_p, C.c_void_p]),
    "apd_scatter_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "apd_synchronize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "apd_align_pair": (C.c_int, [C.c_void_p, _pp, C.c_uint32, C.c_uint32, _fp, _u32p, C.c_uint64, _u64p]),
    "apd_align_pairs": (C.c_int, [C.c_void_p, _pp, _u32p, C.c_uint64, _fp, _u32p, C.c_uint64, _u64p]),
    "apd_align_pairs_band": (C.c_int, [C.c_void_p, _pp, C.c_uint64, _u32p, C.c_uint64, _fp, _u32p,
                                       C.c_uint64, _u64p]),
    "apd_percentile_matrix": (C.c_int, [C.c_void_p, C.c_float, _fp]),
    "apd_percentile_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_void_p, _fp]),
    "apd_upgma": (C.c_int, [_fp, C.c_uint32, C.c_float, _fp, C.POINTER(apd_merge), _u32p, _fp, _u32p]),
    "apd_get_stats": (C.c_int, [C.c_void_p, C.POINTER(apd_stats)]),
    "apd_last_launch_plan": (C.c_char_p, [C.c_void_p]),
    "apd_save_matrix": (C.c_int, [C.c_char_p, _fp, C.c_uint32, C.c_char_p]),
    "apd_load_matrix": (C.c_int, [C.c_char_p, _fp, C.c_uint64, _u32p, C.c_int]),
    "apd_save_paths": (C.c_int, [C.c_char_p, _u32p, C.c_uint64, _fp, _u32p, C.c_uint64, _u64p]),
    "apd_load_paths": (C.c_int, [C.c_char_p, _u32p, _fp, _u64p, C.c_uint64, _u32p, C.c_uint64, _u64p]),
}

_lib = None


def lib():
    """Loads libapd_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libapd_b200.so is missing (%s): build the CUDA library with "
                "`python -m audio_pattern_discovery_b200.build`; there is no CPU path" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        if L.apd_abi_version() != APD_ABI_VERSION:
            raise ImportError("libapd_b200.so ABI version mismatch: rebuild it")
        _lib = L
    return _lib


def check(ctx, status):
    if status != APD_OK:
        msg = lib().apd_last_error(ctx)
        raise ApdError(status, msg.decode("utf-8", "replace") if msg else "")
