"""The C-ABI library loads on a machine without a GPU, exports every symbol
include/apd.h declares, and refuses to compute without a device (no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from audio_pattern_discovery_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "apd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apd_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported(apd_lib_path):
    L = C.CDLL(apd_lib_path)
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), name
    assert set(names) == set(_capi.PROTOTYPES), "ctypes prototypes and apd.h disagree"


def test_abi_version(apd_lib_path):
    assert _capi.lib().apd_abi_version() == _capi.APD_ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(_capi.apd_params) == 20
    assert C.sizeof(_capi.apd_stats) == 8 * 6 + 4 * 5 + 4 + 8 * 2 + 4 + 4 + 4 + 4  # incl. padding before h2d_bytes


def test_no_cpu_path_without_device(apd_lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from audio_pattern_discovery_b200 import AlignmentWorkers, ApdError
    with pytest.raises(ApdError) as ei:
        AlignmentWorkers.new([np.zeros((4, 3), np.float32)] * 3)
    assert ei.value.status == _capi.APD_ERR_NO_DEVICE


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "audio_pattern_discovery_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)
