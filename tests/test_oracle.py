"""Pins the CPU oracle: against the fixtures the REAL reference produced (tests/golden, oracle/make_golden.py) and,
when /root/reference is mounted (build container only), against the reference module directly."""
import os

import numpy as np
import pytest
import torch

import vmae_oracle as oracle
from conftest import golden_case_inputs, load_golden, needs_reference
from counterfactualworldmodels_b200 import synthetic, vmae

FAST_CASES = ["tiny_4x4_b2", "tiny_8x8_b3", "small_4x4_b2", "small_4x4_allvisible_frame1half", "base_8x8_b1_factual",
              "tiny_4x4_tube2_b2", "tiny_8x8_layerscale_learnpos_b2"]


def _our_state_dict(cfg_name, wseed, style):
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
    synthetic.init_weights_(m, seed=wseed, style=style)
    return m, m.state_dict()


@pytest.mark.parametrize("case", FAST_CASES)
def test_oracle_matches_reference_fixture(case):
    g = load_golden(case)
    cfg_name, B, style, wseed, x = golden_case_inputs(case)
    assert float(x.double().sum()) == pytest.approx(float(g["x_fingerprint"][0]), abs=1e-6)
    m, sd = _our_state_dict(cfg_name, wseed, style)
    assert synthetic.weights_checksum(m) == pytest.approx(float(g["weights_checksum"][0]), abs=1e-6)
    ocfg = synthetic.oracle_cfg(cfg_name)
    y = oracle.vmae_forward(sd, oracle.preprocess(x), g["mask"], ocfg)
    assert y.shape == g["y"].shape
    assert (y - g["y"]).abs().max().item() < 2e-5
    if "video" in g:
        v = oracle.pred_patches_to_video(g["y"], x, g["mask"], ocfg["patch_size"])
        assert torch.equal(v, g["video"])


def test_oracle_compaction_matches_torch_nonzero():
    rng = np.random.RandomState(0)
    mask = rng.rand(5, 97) < 0.6
    perm, inv, nvis = oracle.compact_mask(mask)
    for b in range(5):
        vis = torch.nonzero(~torch.from_numpy(mask[b])).flatten().numpy()
        msk = torch.nonzero(torch.from_numpy(mask[b])).flatten().numpy()
        assert nvis[b] == len(vis)
        assert np.array_equal(perm[b, :nvis[b]], vis) and np.array_equal(perm[b, nvis[b]:], msk)
        assert np.array_equal(inv[b, perm[b]], np.arange(97))


def test_oracle_patchify_roundtrip_and_layout():
    x = torch.arange(2 * 2 * 3 * 8 * 12, dtype=torch.float32).reshape(2, 2, 3, 8, 12)
    p = oracle.patchify(x, (1, 4, 4))
    assert p.shape == (2, 2 * 2 * 3, 48)
    # token (t=1, h=1, w=2), element (ph=3, pw=1, c=2) -> channel innermost (patches.py:72-74)
    tok = (1 * 2 + 1) * 3 + 2
    assert p[1, tok, (3 * 4 + 1) * 3 + 2] == x[1, 1, 2, 1 * 4 + 3, 2 * 4 + 1]
    assert torch.equal(oracle.unpatchify(p, (1, 4, 4), x.shape), x)


def test_fp16_operand_emulation_is_within_tolerance_small():
    """fp16 operands + fp32 accumulate stay well inside 2e-2 / 2e-3."""
    g = load_golden("small_4x4_b2")
    cfg_name, B, style, wseed, x = golden_case_inputs("small_4x4_b2")
    _, sd = _our_state_dict(cfg_name, wseed, style)
    y = oracle.vmae_forward(sd, oracle.preprocess(x), g["mask"], synthetic.oracle_cfg(cfg_name),
                            operand_dtype=torch.float16)
    err = (y - g["y"]).abs()
    assert err.max().item() < 2e-2 and err.mean().item() < 2e-3


@needs_reference
def test_oracle_matches_reference_module_live():
    import ref_loader
    ref_vmae, ref_pred = ref_loader.import_reference()
    torch.manual_seed(0)
    ref = ref_vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_8x8")).eval().requires_grad_(False)
    synthetic.init_weights_(ref, seed=11, style="perturbed")
    x = synthetic.make_video(2, synthetic.image_hw("tiny_8x8"), seed=12)
    mask = synthetic.make_mask(2, ref.mask_size, num_clumps=2, seed=13)
    G = ref_pred.PredictorBasedGenerator(predictor=ref, imagenet_normalize_inputs=True, temporal_dim=2)
    with torch.no_grad():
        v_ref = G.predict(x.clone(), mask.clone(), frame=None)
    v = oracle.predict(ref.state_dict(), x, mask, synthetic.oracle_cfg("tiny_8x8"), frame=None)
    assert (v - v_ref).abs().max().item() < 2e-5
    # our module loads the reference state_dict unchanged
    ours = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_8x8"))
    missing = ours.load_state_dict(ref.state_dict())
    assert not missing.missing_keys and not missing.unexpected_keys


@pytest.mark.parametrize("case", ["tiny_8x8_b2_trained_like", "base_8x8_b1_trained_like", "base_8x8_b1_mean_drift"])
def test_oracle_reproduces_the_trained_like_fixtures(case):
    """oracle/make_golden_stats.py: the reference under trained-like activation statistics (outlier channels, token
    means of several sigma); the fixture records the statistics the reference actually reached."""
    import make_golden_stats as mgs
    g = load_golden(case)
    cfg_name, B, wseed, x, mask, style = mgs.case_inputs(case)
    m, sd = _our_state_dict(cfg_name, wseed, style)
    assert torch.equal(mask, g["mask"])
    y = oracle.vmae_forward(sd, oracle.preprocess(x), mask, synthetic.oracle_cfg(cfg_name))
    assert (y - g["y"]).abs().max().item() <= 2e-5 * max(1.0, g["y"].abs().max().item())
    if style == "mean_drift":
        assert g["enc_mean_over_sigma"].max() >= 3.0 and g["dec_mean_over_sigma"].max() >= 3.0
    else:
        assert g["enc_max_over_median"].max() >= 30.0
