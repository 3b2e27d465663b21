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


def _header_prototypes():
    """name -> list of parameter type strings, parsed from the header."""
    src = open(os.path.join(ROOT, "include", "coldrec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"CR_API\s+([^;(]*?)\b(cr_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        params = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        protos[m.group(2)] = [] if params in ([""], ["void"]) else params
    return protos


def test_ctypes_argtypes_match_the_header_prototypes():
    """Same number of parameters, and pointer / integer / float classes agree position by position — a parameter added to
    the header but not to the binding (or in another place) would otherwise shift every later argument silently."""
    protos = _header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)

    def klass_of_c(t):
        t = re.sub(r"\b\w+$", "", t).strip() if not t.endswith("*") else t      # drop the parameter name
        if "*" in t:
            return "ptr"
        if "float" in t or "double" in t:
            return "float"
        return "int"

    def klass_of_ctypes(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(t, type(ctypes.POINTER(ctypes.c_int))) and issubclass(t, ctypes._Pointer):
            return "ptr"
        if t in (ctypes.c_float, ctypes.c_double):
            return "float"
        return "int"

    for name, params in protos.items():
        argtypes = _lib.SIGNATURES[name][1]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, ctypes table {len(argtypes)}"
        for k, (c_t, py_t) in enumerate(zip(params, argtypes)):
            assert klass_of_c(c_t) == klass_of_ctypes(py_t), f"{name} parameter {k} ({c_t!r}) bound as {py_t}"


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
