"""In-tree build of libapd_b200.so (hand-written CUDA for sm_100a + the C ABI).

    python -m audio_pattern_discovery_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU, so this runs on the CPU-only
build box; the resulting .so is git-ignored but travels to the GPU box with the
repository snapshot.  The eight padded frame widths are separate translation
units and compile in parallel.
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libapd_b200.so")

DPADS = (4, 8, 12, 16, 20, 24, 28, 32)
# Translation units besides dtw_inst.cu (once per padded width).  rust/apd-sys/build.rs lists
# the same files; tests/test_rust_sources.py compares the two lists.
CUDA_UNITS = ("apd_api", "pair_path", "percentile", "ae_encode")
CXX_UNITS = ("host_plan", "upgma", "matrix_io")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-D%s=%s" % (k, os.environ[k]) for k in ("APD_X_LOOK", "APD_USE_EDGE_VARIANT", "APD_X_STAGE_TMA", "APD_COMPACT", "APD_WEIGHTED_EDGE", "APD_TWO_ROW_STEPS") if os.environ.get(k)] + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fno-fast-math,-ffp-contract=off"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wno-unknown-pragmas"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _sources():
    names = sorted(os.listdir(CSRC)) + ["../../include/apd.h"]
    return [os.path.normpath(os.path.join(CSRC, n)) for n in names]


def _fingerprint():
    h = hashlib.sha256()
    for p in _sources():
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())
    return h.hexdigest()


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed:\n%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr, flush=True)


def build(force=False, verbose=False, ptxas_info=False, variant=None, defines=()):
    """Builds libapd_b200.so if sources changed; returns its path.  `variant` (experiments only)
    builds libapd_b200.<variant>.so with extra -D defines next to it; load it with
    APD_LIB_PATH=<that file> (see _capi.py)."""
    global OBJ, LIB
    OBJ0, LIB0 = OBJ, LIB
    try:
        if variant:
            OBJ = os.path.join(ROOT, "build", "obj_" + variant)
            LIB = os.path.join(HERE, "libapd_b200.%s.so" % variant)
        return _build(force, verbose, ptxas_info, ["-D" + d for d in defines])
    finally:
        OBJ, LIB = OBJ0, LIB0


def _build(force, verbose, ptxas_info, defs):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "fingerprint")
    fp = _fingerprint() + " ".join(defs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB
    nvcc = _nvcc()
    extra = (["-Xptxas", "-v"] if ptxas_info else []) + defs
    jobs = []
    objs = []
    for d in DPADS:
        o = os.path.join(OBJ, "dtw_inst_%d.o" % d)
        objs.append(o)
        jobs.append([nvcc] + NVCC_FLAGS + extra + ["-DAPD_DPAD=%d" % d, "-c", "-o", o, os.path.join(CSRC, "dtw_inst.cu")])
    for name in CUDA_UNITS:
        o = os.path.join(OBJ, name + ".o")
        objs.append(o)
        jobs.append([nvcc] + NVCC_FLAGS + extra + ["-c", "-o", o, os.path.join(CSRC, name + ".cu")])
    for name in CXX_UNITS:
        o = os.path.join(OBJ, name + ".o")
        objs.append(o)
        jobs.append(["g++"] + CXX_FLAGS + ["-c", "-o", o, os.path.join(CSRC, name + ".cpp")])
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    _run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC", "-cudart", "static"], verbose)
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--ptxas-info", action="store_true")
    ap.add_argument("--variant", default=None, help="experiments: build libapd_b200.<variant>.so")
    ap.add_argument("-D", dest="defines", action="append", default=[], help="extra preprocessor define for a variant build")
    a = ap.parse_args()
    print(build(a.force, a.verbose, a.ptxas_info, a.variant, a.defines))
    sys.exit(0)
