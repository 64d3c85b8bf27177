import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def reference_available():
    """True when the real reference can be imported: the staged copy under baseline/_ref (travels to the GPU box) or,
    in the build container, the /root/reference mount (oracle/ref_loader.py)."""
    import ref_loader
    return ref_loader.available()


needs_reference = pytest.mark.skipif(not reference_available(), reason="real reference neither staged nor mounted")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device")
    config.addinivalue_line("markers", "slow: CPU-heavy (large oracle runs)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(case):
    """Returns dict(mask bool [B,N], y fp32, ...) of a fixture written by oracle/make_golden.py."""
    path = os.path.join(GOLDEN_DIR, case + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {case} missing")
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    B, N = d["mask_shape"]
    d["mask"] = torch.from_numpy(np.unpackbits(d["mask"], axis=1)[:, :N].astype(bool))
    d["y"] = torch.from_numpy(d["y"])
    if "video" in d:
        d["video"] = torch.from_numpy(d["video"])
    return d


def golden_case_inputs(case):
    """(cfg_name, B, init style, weight seed, x) re-derived from seeds exactly as oracle/make_golden.py does."""
    from counterfactualworldmodels_b200 import synthetic
    import make_golden
    cfg_name, B, style, wseed, dseed, _ = make_golden.CASES[case]
    x = synthetic.make_video(B, synthetic.image_hw(cfg_name), seed=dseed, T=synthetic.CONFIGS[cfg_name]["num_frames"])
    return cfg_name, B, style, wseed, x


def load_golden_conjoined(case):
    """Fixture written by oracle/make_golden_conjoined.py: masks unpacked to bool tensors, arrays to tensors."""
    path = os.path.join(GOLDEN_DIR, case + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {case} missing")
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    for key in ("mask", "mask_ctx"):
        if key in d:
            B, N = d[key + "_shape"]
            d[key] = torch.from_numpy(np.unpackbits(d[key], axis=1)[:, :N].astype(bool))
    for key in ("y", "y_ctx", "y_predict", "video"):
        if key in d:
            d[key] = torch.from_numpy(d[key])
    return d
