"""Parity under TRAINED-LIKE activation statistics and at larger batches.

Every other fixture uses freshly initialised weights.  `oracle/make_golden_stats.py` ran the REAL reference with
(a) "trained_like" weights: four massive-activation channels per stream (max|x| / median|x| of a token ~ 60) plus a
mild drift of the token mean, and (b) "mean_drift" weights: no outliers, the token mean runs away to 3-9 sigma of the
token -- the regime in which feeding f16(x) instead of f16(LN(x)) to the tensor cores (the LayerNorm-folded GEMMs,
DESIGN.md section 3) loses accuracy.  Both LayerNorm modes run against both: CWM_FUSE_LN=1 (default) and =0.
Tolerance from BASELINE.json north_star: max-abs <= 2e-2, mean-abs <= 2e-3 in normalised pixel space.
"""
import numpy as np
import pytest
import torch

import make_golden_stats as mgs
import vmae_oracle as oracle
from conftest import load_golden
from counterfactualworldmodels_b200 import prediction, synthetic, vmae

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MAX_ABS, MEAN_ABS = 2e-2, 2e-3


def _model(cfg_name, wseed, style):
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
    synthetic.init_weights_(m, seed=wseed, style=style)
    return m.to(DEV).eval()


@pytest.mark.parametrize("fuse_ln", ["1", "0"])
@pytest.mark.parametrize("case", sorted(mgs.CASES))
def test_trained_like_statistics_match_the_reference(case, fuse_ln, monkeypatch):
    monkeypatch.setenv("CWM_FUSE_LN", fuse_ln)          # read when the device-side weights are packed
    g = load_golden(case)
    cfg_name, B, wseed, x, mask, style = mgs.case_inputs(case)
    m = _model(cfg_name, wseed, style)
    assert synthetic.weights_checksum(m) == pytest.approx(float(g["weights_checksum"][0]), rel=1e-9)
    assert torch.equal(mask, g["mask"])
    y = m(oracle.preprocess(x).to(DEV), mask.to(DEV)).cpu()
    err = (y - g["y"]).abs()
    print(f"{case} CWM_FUSE_LN={fuse_ln}: max-abs {err.max():.3e} mean-abs {err.mean():.3e} | reference statistics: "
          f"encoder |mean|/sigma up to {g['enc_mean_over_sigma'].max():.1f}, max/median up to "
          f"{g['enc_max_over_median'].max():.0f}")
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS
    # the same through the wrapper: visible patches bit-identical to the input
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    video = G.predict(x.to(DEV), mask.to(DEV), frame=None).cpu()
    ps = synthetic.oracle_cfg(cfg_name)["patch_size"]
    assert torch.equal(oracle.patchify(video, ps)[~mask], oracle.patchify(x, ps)[~mask])


@pytest.mark.parametrize("cfg,B,clumps,style", [("tiny_8x8", 16, 2, "perturbed"), ("small_4x4", 8, 2, "perturbed"),
                                                ("base_8x8", 8, 1, "reference"), ("tiny_8x8", 12, 3, "trained_like")])
def test_oracle_parity_at_batch_8_and_more(cfg, B, clumps, style):
    """The CUDA path against the CPU oracle on a batch of B >= 8 different samples and masks (the fixtures stop at 3)."""
    m = _model(cfg, 40 + B, style)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    x = synthetic.make_video(B, synthetic.image_hw(cfg), seed=50 + B)
    mask = synthetic.make_mask(B, m.mask_size, num_clumps=clumps, seed=60 + B)
    want = oracle.predict(sd, x, mask, synthetic.oracle_cfg(cfg), frame=None)
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    got = G.predict(x.to(DEV), mask.to(DEV), frame=None).cpu()
    err = (got - want).abs()
    per_sample = err.flatten(1).amax(1)
    print(f"{cfg} B={B} {style}: max-abs {err.max():.3e} mean-abs {err.mean():.3e}; worst sample {int(per_sample.argmax())}")
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS
    perm, _, n_vis = m.last_aux
    perm_o, _, nvis_o = oracle.compact_mask(mask.numpy())
    assert np.array_equal(perm.cpu().numpy(), perm_o) and n_vis == int(nvis_o[0])
