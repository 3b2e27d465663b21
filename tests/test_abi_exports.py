"""CPU-side checks of the C ABI: every symbol include/coldrec_b200.h declares is exported, the ctypes
table covers the header, and compute entry points refuse to run without an sm_100 device."""
import ctypes
import os
import re

import pytest
import torch

from coldrec_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "coldrec_b200.h")).read()
    return sorted(set(re.findall(r"CR_API[^;(]*?\b(cr_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_error_strings():
    lib = _lib.load()
    assert lib.cr_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.cr_strerror(-6)
    assert lib.cr_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_device_no_fallback():
    lib = _lib.load()
    assert lib.cr_device_check() == -6
    from coldrec_b200 import ops
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.spmm(torch.zeros(2, dtype=torch.int64), torch.zeros(0, dtype=torch.int32), None, torch.zeros(1, 64))
    with pytest.raises(ValueError, match="CUDA tensor"):
        ops.score_topk(torch.zeros(4, 64), torch.zeros(30, 64), 20)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "coldrec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f"{f} imports the oracle"
