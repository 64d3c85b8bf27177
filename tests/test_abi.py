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
    assert lib.cwm_abi_version() == 7
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
    assert ctypes.sizeof(_lib.BlockWeights) == 18 * 8
    assert ctypes.sizeof(_lib.VmaeModel) == 16 * 4 + 3 * 4 + 4 + 14 * 8
    assert ctypes.sizeof(_lib.GemmEpilogue) == 80 + 56


def test_argument_validation_of_the_counterfactual_and_flow_entry_points():
    """SURVEY 8(f) entry points reject bad arguments on the host, before any CUDA call."""
    lib = _lib.load()
    rc = lib.cwm_cf_shift_masks(None, None, None, 4, 2, 8, 8, 1, None, None, None)
    assert rc == -1 and b"null pointer" in lib.cwm_last_error()
    rc = lib.cwm_cf_shift_masks(1, 1, 1, 4, 2, 8, 8, 2, 1, 1, None)      # frame out of range
    assert rc == -1 and b"frame" in lib.cwm_last_error()
    src = _lib.CfSource()
    rc = lib.cwm_cf_build_videos(ctypes.byref(src), 1, 2, 3, 32, 32, 4, 4, 16, None)
    assert rc == -1 and b"cwm_cf_source" in lib.cwm_last_error()
    src.x, src.shift_px, src.shifted_active, src.frame, src.static_frame = 16, 16, 16, 1, -1
    rc = lib.cwm_cf_build_videos(ctypes.byref(src), 1, 2, 3, 30, 32, 4, 4, 16, None)   # 30 % 4 != 0
    assert rc == -1 and b"divisible" in lib.cwm_last_error()
    rc = lib.cwm_cf_build_videos(ctypes.byref(src), 1, 2, 3, 32, 36, 4, 6, 16, None)   # patch width 6
    assert rc == -3 and b"multiple of 4" in lib.cwm_last_error()
    rc = lib.cwm_patch_gather_cf(ctypes.byref(src), 1, 3, 2, 32, 32, 2, 4, 4, 16, 128, 70, None, None, 16, None)
    assert rc == -1 and b"temporal patch size" in lib.cwm_last_error()
    fs = (ctypes.c_int64 * 5)(0, 0, 0, 0, 0)
    rc = lib.cwm_flow_sample_stats(None, fs, 1, 8, 8, 2, None, None, 0, 0, 5.0, 16, 16, 1024, None)
    assert rc == -1 and b"null pointer" in lib.cwm_last_error()
    rc = lib.cwm_flow_sample_stats(16, fs, 1, 8, 8, 4, None, None, 0, 0, 5.0, 16, 16, 8, None)   # workspace too small
    assert rc == -4 and b"workspace" in lib.cwm_last_error()
    rc = lib.cwm_flow_magnitude_sum(16, fs, 1, 8, 8, 4, None, None, 1, 0.01, 0, 16, 16, 1 << 20, None)
    assert rc == -1 and b"statistics" in lib.cwm_last_error()
    rc = lib.cwm_motion_map_finalize(16, 1, 8, 8, 0.0, 1, 0.01, 16, None)
    assert rc == -1
    assert lib.cwm_flow_stats_workspace_bytes(1, 224, 224, 64) >= 32 * 224 * 224 * 4


def test_cf_source_struct_layout():
    assert ctypes.sizeof(_lib.CfSource) == 8 + 5 * 8 + 3 * 8 + 2 * 4
