"""The C-ABI shared library loads and exports every symbol include/regex_fpga_b200.h declares."""
import ctypes
import os
import re

import pytest

import regex_fpga_b200 as R
from regex_fpga_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "regex_fpga_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"
    assert _lib.load().rfb_abi_version() == 4


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.rfb_match) == 12
    assert ctypes.sizeof(_lib.rfb_nfa_info) == 48
    assert ctypes.sizeof(_lib.rfb_batch) == 88
    assert ctypes.sizeof(_lib.rfb_result) == 72


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(R.RfbError) as e:
        R.Context(0)
    assert e.value.code == -6 and "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under regex_fpga_b200/ may reference it."""
    pkg = os.path.join(ROOT, "regex_fpga_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="replace").read()
                assert "oracle" not in txt.lower(), os.path.join(d, f)
