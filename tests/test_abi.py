"""The C-ABI shared library loads and exports every symbol include/cwm_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from counterfactualworldmodels_b200 import _lib

HEADER = os.path.join(ROOT, "include", "cwm_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cwm_[a-z0-9_]+)\s*\(", src)))


def test_header_lists_functions():
    fns = declared_functions()
    assert "cwm_vmae_forward" in fns and "cwm_compact_mask" in fns and len(fns) >= 12


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} missing: run __graft_entry__.build()")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in cwm_b200.h but not exported"
    assert set(declared_functions()) == set(_lib.EXPORTED_SYMBOLS)


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.cwm_abi_version() == 3
    assert isinstance(lib.cwm_last_error(), bytes)


def test_argument_validation_without_gpu():
    """Bad arguments are rejected on the host before any CUDA call (safe without a device)."""
    lib = _lib.load()
    rc = lib.cwm_compact_mask(None, 1, 8, None, None, None, None)
    assert rc == -1 and b"null pointer" in lib.cwm_last_error()
    rc = lib.cwm_attention_f16(16, 1, 8, 2, 40, 16, None)
    assert rc == -3 and b"head_dim" in lib.cwm_last_error()
    e = _lib.GemmEpilogue()
    e.mode, e.out = 0, 1
    rc = lib.cwm_gemm_f16(1, 1, 4, 8, 24, ctypes.byref(e), None)
    assert rc == -1 and b"K" in lib.cwm_last_error()


def test_struct_layout_matches_header():
    # sizes follow from the C declaration order: 16 int32 + 3 float + 14 pointers (8-byte aligned)
    assert ctypes.sizeof(_lib.BlockWeights) == 12 * 8
    assert ctypes.sizeof(_lib.VmaeModel) == 16 * 4 + 3 * 4 + 4 + 14 * 8
    assert ctypes.sizeof(_lib.GemmEpilogue) == 80
