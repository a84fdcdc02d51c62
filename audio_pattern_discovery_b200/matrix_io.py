"""On-disk form of the distance matrix and of alignment paths.

The reference never persists the matrix (it lives in an Arc<Mutex<Vec<f32>>> for the length
of `learn()`, src/main.rs:194-195), so re-clustering with another percentile means re-running
the whole alignment; README.md:67 promises "alignment path information" that no code writes.
Format (little endian, host code only):

    <stem>.apdm        n*n float32, row-major, result[x*n+y] -- byte-for-byte the Vec<f32> the
                       reference hands to clustering(); can be memory-mapped
    <stem>.apdm.json   {"format": "apd-matrix-1", "n": ..., "dtype": "<f4", "params": {...},
                        "sha256": hex digest of the .apdm payload}
    <stem>.apdp.json   {"format": "apd-paths-1", "paths": [{"i": .., "j": .., "score": ..,
                        "path": [[i, j], ...]}]}   (1-based cells, end to start, Appendix A.8)
"""
import hashlib
import json
import os

import numpy as np

MATRIX_FORMAT = "apd-matrix-1"
PATHS_FORMAT = "apd-paths-1"


def save_matrix(stem, distances, n_instances, params=None):
    """distances: flat n*n float32 (or (n, n)).  Returns the payload path."""
    d = np.ascontiguousarray(distances, dtype="<f4").reshape(-1)
    n = int(n_instances)
    if d.size != n * n:
        raise ValueError("distances must hold n_instances^2 entries")
    payload = stem + ".apdm"
    with open(payload, "wb") as f:
        f.write(d.tobytes())
    meta = {"format": MATRIX_FORMAT, "n": n, "dtype": "<f4", "params": dict(params or {}),
            "sha256": hashlib.sha256(d.tobytes()).hexdigest()}
    with open(payload + ".json", "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    return payload


def load_matrix(stem, mmap=False, verify=True):
    """-> (flat float32 array of n*n entries, n, params)."""
    payload = stem + ".apdm"
    with open(payload + ".json") as f:
        meta = json.load(f)
    if meta.get("format") != MATRIX_FORMAT or meta.get("dtype") != "<f4":
        raise ValueError("not an %s file" % MATRIX_FORMAT)
    n = int(meta["n"])
    if os.path.getsize(payload) != 4 * n * n:
        raise ValueError("payload size does not match n = %d" % n)
    d = np.memmap(payload, dtype="<f4", mode="r") if mmap else np.fromfile(payload, dtype="<f4")
    if verify and hashlib.sha256(np.asarray(d).tobytes()).hexdigest() != meta["sha256"]:
        raise ValueError("payload checksum mismatch")
    return d, n, meta.get("params", {})


def save_paths(stem, pairs, scores, paths):
    """pairs: [(i, j)], scores: floats, paths: list of (L, 2) arrays as returned by Context.align_pairs."""
    def num(v):
        v = float(v)
        return v if np.isfinite(v) else ("inf" if v > 0 else ("-inf" if v < 0 else "nan"))
    doc = {"format": PATHS_FORMAT,
           "paths": [{"i": int(i), "j": int(j), "score": num(s), "path": np.asarray(p, dtype=np.int64).reshape(-1, 2).tolist()}
                     for (i, j), s, p in zip(pairs, scores, paths)]}
    out = stem + ".apdp.json"
    with open(out, "w") as f:
        json.dump(doc, f)
    return out


def load_paths(stem):
    with open(stem + ".apdp.json") as f:
        doc = json.load(f)
    if doc.get("format") != PATHS_FORMAT:
        raise ValueError("not an %s file" % PATHS_FORMAT)
    out = []
    for e in doc["paths"]:
        s = e["score"]
        s = np.float32({"inf": np.inf, "-inf": -np.inf, "nan": np.nan}[s]) if isinstance(s, str) else np.float32(s)
        out.append(((e["i"], e["j"]), s, np.asarray(e["path"], dtype=np.uint32).reshape(-1, 2)))
    return out
