"""Parity of the CUDA path in the operand type this process runs (CWM_DTYPE=f16, the default build, or CWM_DTYPE=bf16, the
bf16-operand twin built from the same sources) against fixtures the REAL reference produced (tests/golden): prints one
JSON line {"dtype", "cases": {case: {"max_abs", "mean_abs"}}}.  bench.py and tests/test_bf16_mode_gpu.py run it once per
type in a subprocess (a process binds to one build at import time).

    CWM_DTYPE=bf16 python tools/dtype_error.py base_8x8_b2_counterfactual large_4x4_b1_factual"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import vmae_oracle as oracle  # noqa: E402  (test infrastructure: the boundary normalisation of the fixture inputs)
from conftest import golden_case_inputs, load_golden  # noqa: E402
from counterfactualworldmodels_b200 import _lib, synthetic, vmae  # noqa: E402


def main(cases):
    dev = "cuda:0"
    out = {"dtype": "bf16" if _lib.act_dtype() == torch.bfloat16 else "f16", "library": os.path.basename(_lib.lib_path()),
           "cases": {}}
    for case in cases:
        g = load_golden(case)
        cfg_name, B, style, wseed, x = golden_case_inputs(case)
        m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
        synthetic.init_weights_(m, seed=wseed, style=style)
        m = m.to(dev).eval()
        y = m(oracle.preprocess(x).to(dev), g["mask"].to(dev)).cpu()
        err = (y - g["y"]).abs()
        out["cases"][case] = {"max_abs": float(err.max()), "mean_abs": float(err.mean()), "ref_std": float(g["y"].std())}
        del m
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1:] or ["base_8x8_b2_counterfactual", "large_4x4_b1_factual"])
