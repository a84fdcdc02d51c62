"""The compiled-language host mirror (include/apd_host.hpp) and the reference's call site written
against it (examples/learn_stage3.cpp): builds with g++ against libapd_b200.so; on a machine without
a GPU the alignment fails loudly with the reference's panic convention, the host-only clustering path
is checked against the oracle, and with a GPU the whole stage equals the oracle bit for bit."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio_pattern_discovery_b200")
EXE = os.path.join(ROOT, "build", "learn_stage3")

TOML = """
dft_win = 256            # DFT window
dft_step = 128
ceps_filter = 32
auto_encoder = 10
learning_rate = 0.1
epochs = 25
epoch_drop = 5.0
drop = 0.5
vat_moving = 15
vat_percentile = 0.95
vat_min_len = 150
warping_band_percentage = %s   # sakoe shiba band
insertion_penalty = %s
deletion_penalty = %s
match_penalty = %s
alignment_workers = 4
clustering_percentile = %s
"""


@pytest.fixture(scope="module")
def exe(apd_lib_path):
    src = os.path.join(ROOT, "examples", "learn_stage3.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("apd.h", "apd_host.hpp")]
    if not (os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(p) for p in [src, apd_lib_path] + hdrs)):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", EXE, src,
                        "-L", PKG, "-lapd_b200", "-Wl,-rpath," + PKG, "-pthread"], check=True, capture_output=True)
    return EXE


def write_apds(path, seqs):
    dim = seqs[0].shape[1]
    with open(path, "wb") as f:
        f.write(b"APDS" + struct.pack("<II", len(seqs), dim))
        f.write(np.array([len(s) for s in seqs], dtype="<u4").tobytes())
        for s in seqs:
            f.write(np.ascontiguousarray(s, dtype="<f4").tobytes())


def read_merges(path):
    out = []
    for ln in open(path):
        if not ln.startswith("#"):
            a, b, k, bits, op, tie = ln.split()
            out.append((int(a), int(b), int(k), int(bits), int(tie)))
    return out


def test_cluster_only_path_matches_oracle_on_cpu(exe, tmp_path):
    rng = np.random.default_rng(3)
    n = 60
    d = rng.gamma(2.0, 1.0, size=(n, n)).astype(np.float32)
    np.fill_diagonal(d, 0.0)
    d.tofile(str(tmp_path / "m.apdm"))
    r = subprocess.run([exe, "--cluster-only", str(tmp_path / "m.apdm"), str(n), "0.1", str(tmp_path / "out")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Clustering with" in r.stdout          # the reference's progress lines
    want, _, _ = oracle.upgma(d, 0.1)
    got = read_merges(str(tmp_path / "out.merges.txt"))
    assert got == [(a, b, k, int(np.float32(dd).view(np.uint32)), t) for a, b, k, dd, t in want]
    r = subprocess.run([exe, "--cluster-only", str(tmp_path / "m.apdm"), str(n), "1.0", str(tmp_path / "out")],
                       capture_output=True, text=True)
    assert r.returncode == 101 and "panicked" in r.stderr  # percentile index out of bounds


def test_alignment_fails_loudly_without_a_gpu(exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rng = np.random.default_rng(4)
    write_apds(str(tmp_path / "s.bin"), [rng.normal(size=(9, 3)).astype(np.float32) for _ in range(4)])
    (tmp_path / "Discovery.toml").write_text(TOML % ("0.2", "1.0", "1.0", "1.0", "0.05"))
    r = subprocess.run([exe, str(tmp_path / "s.bin"), str(tmp_path / "Discovery.toml"), str(tmp_path / "out")],
                       capture_output=True, text=True)
    assert r.returncode == 101 and "no CUDA device" in r.stderr
    (tmp_path / "bad.toml").write_text("dft_win = 256\n")
    r = subprocess.run([exe, str(tmp_path / "s.bin"), str(tmp_path / "bad.toml"), str(tmp_path / "out")],
                       capture_output=True, text=True)
    assert r.returncode == 101 and "missing field" in r.stderr


@pytest.mark.gpu
def test_stage3_through_the_cpp_mirror_equals_the_oracle(exe, tmp_path):
    from audio_pattern_discovery_b200 import synth
    rng = np.random.default_rng(5)
    seqs, _ = synth.make_sequences(70, rng.integers(30, 120, size=70), 10, 5, 41)
    write_apds(str(tmp_path / "s.bin"), seqs)
    (tmp_path / "Discovery.toml").write_text(TOML % ("0.1", "0.75", "0.5", "1.0", "0.05"))
    r = subprocess.run([exe, str(tmp_path / "s.bin"), str(tmp_path / "Discovery.toml"), str(tmp_path / "out")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(str(tmp_path / "out.apdm"), dtype="<f4").reshape(70, 70)
    want = oracle.align_all(seqs, 0.1, 0.75, 0.5, 1.0, workers=8, variant="dense")
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    wm, _, _ = oracle.upgma(want, 0.05)
    assert read_merges(str(tmp_path / "out.merges.txt")) == [(a, b, k, int(np.float32(dd).view(np.uint32)), t)
                                                             for a, b, k, dd, t in wm]


@pytest.mark.gpu
def test_stage3_with_the_encoder_on_the_device(exe, tmp_path):
    """src/main.rs:150-161 + 187-203: cepstra in, AutoEncoder::predict on the GPU, matrix and merges out --
    and the matrix file the C ABI wrote reads back through the Python side of the same format."""
    from audio_pattern_discovery_b200 import matrix_io
    rng = np.random.default_rng(6)
    w = ((rng.random((26, 10)) - 0.5) / 10 * 6).astype(np.float32)
    b = ((rng.random(10) - 0.5) / 10).astype(np.float32)
    ceps = [(np.cumsum(rng.normal(size=(int(t), 26)), axis=0) * 0.4).astype(np.float32) for t in rng.integers(20, 80, size=30)]
    write_apds(str(tmp_path / "c.bin"), ceps)
    with open(tmp_path / "nn.bin", "wb") as f:
        f.write(b"APDE" + struct.pack("<II", 26, 10) + w.astype("<f4").tobytes() + b.astype("<f4").tobytes())
    (tmp_path / "Discovery.toml").write_text(TOML % ("1.0", "1.0", "1.0", "1.0", "0.05"))
    r = subprocess.run([exe, str(tmp_path / "c.bin"), str(tmp_path / "Discovery.toml"), str(tmp_path / "out"), str(tmp_path / "nn.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    emb = [oracle.ae_encode(x, w, b) for x in ceps]
    want = oracle.align_all(emb, 1.0, 1.0, 1.0, 1.0, workers=8, variant="dense")
    got, n, _ = matrix_io.load_matrix(str(tmp_path / "out"))       # checks the sha256 of the header too
    assert n == 30 and np.array_equal(got.reshape(30, 30).view(np.uint32), want.view(np.uint32))
