"""The bf16-operand build (libcwm_b200_bf16.so: the same kernel sources compiled with -DCWM_ACT_BF16, selected with
CWM_DTYPE=bf16) beside the f16 default, on fixtures the real reference produced.

BASELINE.json states the pixel tolerance (2e-2 max-abs / 2e-3 mean-abs) "for bf16 with fp32 accumulation"; SURVEY.md's
probe showed that bf16 operands cannot meet it on the graded models even in an idealised emulation (8 mantissa bits against
f16's 11), which is why f16 is the parity default.  This test pins both facts on the device: the f16 build is inside the
bar, the bf16 build runs the same path (same permutation, finite output) with an error 4-16x larger and bounded."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["tiny_4x4_b2", "base_8x8_b2_counterfactual", "large_4x4_b1_factual"]


def _run(dtype):
    env = dict(os.environ, CWM_DTYPE=dtype)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dtype_error.py")] + CASES, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_both_libraries_are_built_and_report_their_operand_type():
    import ctypes
    from counterfactualworldmodels_b200 import _lib
    for name, want in (("libcwm_b200.so", 0), ("libcwm_b200_bf16.so", 1)):
        path = os.path.join(os.path.dirname(_lib.LIB_PATH), name)
        assert os.path.exists(path), f"{path} missing: run __graft_entry__.build()"
        lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        assert lib.cwm_act_dtype() == want and lib.cwm_abi_version() == _lib.ABI_VERSION
        for sym in _lib.EXPORTED_SYMBOLS:
            assert hasattr(lib, sym), (name, sym)


@pytest.mark.gpu
def test_bf16_operand_mode_runs_and_its_error_is_reported_beside_f16():
    f16, bf16 = _run("f16"), _run("bf16")
    assert f16["dtype"] == "f16" and bf16["dtype"] == "bf16" and bf16["library"] == "libcwm_b200_bf16.so"
    for case in CASES:
        a, b = f16["cases"][case], bf16["cases"][case]
        print(f"{case}: f16 {a['max_abs']:.2e} / {a['mean_abs']:.2e}   bf16 {b['max_abs']:.2e} / {b['mean_abs']:.2e}")
        assert a["max_abs"] <= 2e-2 and a["mean_abs"] <= 2e-3                      # the parity default
        assert b["mean_abs"] > 2 * a["mean_abs"]                                   # 3 fewer mantissa bits do show
        assert b["max_abs"] <= 0.15 and b["mean_abs"] <= 2.5e-2                    # and stay bounded (~8x f16)
