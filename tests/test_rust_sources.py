"""The Rust side (rust/apd-sys, rust/alignments.rs) cannot be compiled in this image (no cargo /
rustc), so what can drift silently is checked mechanically:
  * build.rs compiles exactly the translation units and flags build.py compiles;
  * lib.rs declares exactly the symbols include/apd.h declares;
  * the #[repr(C)] structs of lib.rs have the field order, sizes and offsets of the C structs
    (a generated C program with static_asserts against include/apd.h, compiled with gcc);
  * the constants agree."""
import os
import re
import subprocess

from audio_pattern_discovery_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD_RS = open(os.path.join(ROOT, "rust", "apd-sys", "build.rs")).read()
LIB_RS = open(os.path.join(ROOT, "rust", "apd-sys", "src", "lib.rs")).read()
HEADER = open(os.path.join(ROOT, "include", "apd.h")).read()


def _rs_list(name):
    m = re.search(r"const %s: \[[^\]]*\] = \[(.*?)\];" % name, BUILD_RS, re.S)
    assert m, name
    return re.findall(r'"([^"]*)"|(\d+)', m.group(1))


def test_build_rs_compiles_what_build_py_compiles():
    assert tuple(int(d) for _, d in _rs_list("DPADS")) == B.DPADS
    assert tuple(s for s, _ in _rs_list("CUDA_UNITS")) == B.CUDA_UNITS
    assert tuple(s for s, _ in _rs_list("CXX_UNITS")) == B.CXX_UNITS
    nvcc_flags = [s for s, _ in _rs_list("NVCC_FLAGS")]
    assert nvcc_flags == [f for f in B.NVCC_FLAGS if f not in B.ARCH and not f.startswith("-DAPD_")]
    assert [s for s, _ in _rs_list("CXX_FLAGS")] == B.CXX_FLAGS
    assert "arch=compute_100a,code=sm_100a" in BUILD_RS
    for name in B.CUDA_UNITS:
        assert os.path.exists(os.path.join(B.CSRC, name + ".cu")), name
    for name in B.CXX_UNITS:
        assert os.path.exists(os.path.join(B.CSRC, name + ".cpp")), name


def _header_symbols():
    src = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(apd_[a-z_]+)\s*\(", src)))


def test_lib_rs_declares_every_header_symbol():
    rs = sorted(set(re.findall(r"pub fn (apd_[a-z_]+)\s*\(", LIB_RS)))
    assert rs == _header_symbols()


RS_TYPES = {"u64": ("uint64_t", 8), "u32": ("uint32_t", 4), "f32": ("float", 4), "i32": ("int32_t", 4)}


def _rs_structs():
    out = {}
    for name, body in re.findall(r"#\[repr\(C\)\][^{]*?pub struct (apd_[a-z_]+) \{(.*?)\n\}", LIB_RS, re.S):
        fields = re.findall(r"pub ([a-z_0-9]+): ([a-z0-9]+),", body)
        if fields:
            out[name] = fields
    return out


def test_lib_rs_struct_layouts_match_the_c_header(tmp_path):
    structs = _rs_structs()
    assert set(structs) == {"apd_params", "apd_stats", "apd_merge"}
    lines = ['#include <stddef.h>', '#include "apd.h"']
    for name, fields in structs.items():
        off = 0
        align = 1
        for fname, ftype in fields:
            ctype, size = RS_TYPES[ftype]
            off = (off + size - 1) // size * size            # repr(C): natural alignment, declaration order
            align = max(align, size)
            lines.append('_Static_assert(offsetof(%s, %s) == %d, "%s.%s offset");' % (name, fname, off, name, fname))
            lines.append('_Static_assert(sizeof(((%s *)0)->%s) == %d, "%s.%s size");' % (name, fname, size, name, fname))
            off += size
        total = (off + align - 1) // align * align
        lines.append('_Static_assert(sizeof(%s) == %d, "%s size");' % (name, total, name))
    lines.append("int main(void) { return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    r = subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "layout.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_constants_agree():
    for name in ("APD_ABI_VERSION", "APD_MAX_DIM", "APD_MAX_DEVICES", "APD_AE_MAX_BINS"):
        h = re.search(r"#define %s (\d+)" % name, HEADER).group(1)
        r = re.search(r"pub const %s: u32 = (\d+);" % name, LIB_RS).group(1)
        assert h == r, name
    for name, val in re.findall(r"(APD_ERR_[A-Z_]+|APD_OK) = (\d+)", HEADER):
        assert re.search(r"pub const %s: c_int = %s;" % (name, val), LIB_RS), name


def test_drop_in_uses_the_group_context_and_reuses_the_pair_context():
    src = open(os.path.join(ROOT, "rust", "alignments.rs")).read()
    assert "apd_create_multi" in src and "Context::all_devices()" in src
    body = src[src.index("pub fn construct_alignment"):]
    assert "pair_context()" in body and "Context::one_device()" not in body and "apd_create(" not in body
