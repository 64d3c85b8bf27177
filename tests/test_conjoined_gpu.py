"""GPU parity of the padded / conjoined (IMU-conditioned) predictors (BASELINE
config 5) through the C ABI against the fixtures the REAL reference produced and against the CPU oracle.
Tolerance (BASELINE.json north_star): index work bit-exact; predicted values max-abs <= 2e-2, mean-abs <= 2e-3."""
import os
from functools import partial

import numpy as np
import pytest
import torch

import conjoined_oracle as co
import make_golden_conjoined as mgc
import vmae_oracle as oracle
from conftest import GOLDEN_DIR, load_golden_conjoined
from counterfactualworldmodels_b200 import conjoined_vmae as C
from counterfactualworldmodels_b200 import prediction, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MAX_ABS, MEAN_ABS = 2e-2, 2e-3


def _check(got, want, what):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if want.numel() == 0:
        return
    err = (got.cpu() - want).abs()
    print(f"{what}: max-abs {err.max():.3e} mean-abs {err.mean():.3e} (ref std {want.std():.3f})")
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS, (what, err.max().item(), err.mean().item())


def _build(case):
    name, B, style, wseed, x, mask, imu, mc = mgc.case_inputs(case)
    m = synthetic.build_conjoined(C, name)
    synthetic.init_weights_(m, seed=wseed, style=style)
    return m.to(DEV).eval(), name, x, mask, imu, mc


@pytest.mark.parametrize("case", ["conj_padded_small_ragged", "conj_padded_small_ctxmasked",
                                  "conj_padded_small_predict", "conj_flow2imu_small"])
def test_conjoined_forward_matches_reference_fixture(case):
    g = load_golden_conjoined(case)
    m, name, x, mask, imu, mc = _build(case)
    assert synthetic.weights_checksum(m) == pytest.approx(float(g["weights_checksum"][0]), abs=1e-6)
    xin = x.to(DEV).transpose(1, 2)                                  # [B,C,T,H,W] view, as the wrapper hands it over
    y, yc = m(xin, mask.to(DEV), x_context=imu.to(DEV), mask_context=mc.to(DEV), output_main=True,
              output_context=True)
    _check(y, g["y"], case + " main")
    _check(yc, g["y_ctx"], case + " ctx")
    # integer side: the token order both streams used is the reference's (visible ascending, then masked ascending,
    # padding positions last) -- bit-exact against the oracle's masks
    (perm_m, _, n_m), (perm_c, _, n_c) = m.last_aux
    ocfg = synthetic.conjoined_oracle_cfg(name)
    full_m = co.padding_masks(mask, ocfg["main"]["max_pad"])[1] if ocfg["main"]["max_pad"] else \
        m.get_stream_inputs(xin.cpu(), mask, None, x_context=imu, mask_context=mc)[0][1]
    perm_o, _, nvis_o = oracle.compact_mask(full_m.numpy())
    assert np.array_equal(perm_m.cpu().numpy(), perm_o) and n_m == int(nvis_o[0])
    if "null_rows_main" in g:
        # padding rows are exactly zero, like `x * ~null_mask` (conjoined_vmae.py:998-1002)
        assert int((y.abs().sum(-1) == 0).sum()) == int(g["null_rows_main"][0])
        assert int((yc.abs().sum(-1) == 0).sum()) == int(g["null_rows_ctx"][0])
        assert torch.equal(m.main_stream.null_mask.cpu(), co.padding_masks(mask, ocfg["main"]["max_pad"])[2])
        # stateful quirk the wrappers rely on
        assert hasattr(m, "padding_mask")
        m._reset_padding_mask()
        assert not hasattr(m, "padding_mask")
    # output selection is cached between calls (conjoined_vmae.py:589-593)
    y_only = m(xin, mask.to(DEV), x_context=imu.to(DEV), mask_context=mc.to(DEV), output_main=False, output_context=True)
    assert torch.is_tensor(y_only) and torch.equal(y_only, yc)


def test_padded_vmae_matches_reference_fixture():
    g = load_golden_conjoined("padded_small_ragged")
    B, style, wseed, x, mask = mgc.padded_case_inputs("padded_small_ragged")
    m = C.PaddedVisionTransformer(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **mgc.PADDED_KW)
    synthetic.init_weights_(m, seed=wseed, style=style)
    m = m.to(DEV).eval()
    y = m(x.to(DEV).transpose(1, 2), mask.to(DEV))
    _check(y, g["y"], "padded")
    assert int((y.abs().sum(-1) == 0).sum()) == int(g["null_rows"][0])
    assert torch.equal(m.null_mask.cpu(), co.padding_masks(mask, 8)[2])


def test_conjoined_predict_wrapper_small():
    """`PredictorBasedGenerator.predict(x, mask, x_context=imu, mask_context=...)` (prediction.py:406-454): padding rows
    stripped, main-stream frames reassembled; visible patches bit-identical to the input."""
    case = "conj_padded_small_predict"
    g = load_golden_conjoined(case)
    m, name, x, mask, imu, mc = _build(case)
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    video = G.predict(x.to(DEV), mask.to(DEV), frame=None, x_context=imu.to(DEV), mask_context=mc.to(DEV)).cpu()
    assert not hasattr(m, "padding_mask")                              # reset after the call (prediction.py:451-452)
    err = (video - g["video"]).abs()
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS
    ps = (1, 4, 4)
    vp, xp = oracle.patchify(video, ps), oracle.patchify(x, ps)
    assert torch.equal(vp[~mask], xp[~mask])
    # chunked execution with a per-image IMU context tiled over the samples of a chunk (prediction.py:489-538)
    xs = x[:1].expand(4, -1, -1, -1, -1).contiguous().to(DEV)
    ms = torch.cat([mask, mask], 0).to(DEV)
    out = G.batch_predict_per_sample(xs, ms, frame=None, batch_size=2, sample_dim=0, x_context=imu[:1].to(DEV),
                                     mask_context=mc[:1].to(DEV))
    whole = G.predict(xs, ms, frame=None, x_context=imu[:1].expand(4, -1, -1).to(DEV),
                      mask_context=mc[:1].expand(4, -1).to(DEV))
    assert out.shape == (4, 2, 3, 32, 32) and torch.equal(out, whole)


def test_imu400_base_4x4_matches_reference_fixture():
    """BASELINE config 5 at full size (6272 + 64 tokens, 25 IMU tokens), one sample, through the wrapper."""
    case = "conj_imu400_base_4x4_b1"
    g = load_golden_conjoined(case)
    m, name, x, mask, imu, mc = _build(case)
    assert sum(p.numel() for p in m.parameters()) == 148265040
    xin = x.to(DEV).transpose(1, 2)
    y, yc = m(xin, mask.to(DEV), x_context=imu.to(DEV), mask_context=mc.to(DEV), output_main=True, output_context=True)
    _check(y, g["y"], "imu400 main")
    assert y.shape == (1, 3168, 48) and int((y.abs().sum(-1) == 0).sum()) == 64
    assert yc.shape == (1, 25, 96) and float(yc.abs().max()) == 0.0    # IMU fully visible -> only null rows
    m._reset_padding_mask()
    m._set_decoder_outputs(output_main=True, output_context=False)
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    video = G.predict(x.to(DEV), mask.to(DEV), frame=None, x_context=imu.to(DEV), mask_context=mc.to(DEV)).cpu()
    want = oracle.pred_patches_to_video(g["y_predict"][:, :-64], x, mask, (1, 4, 4))
    err = (video - want).abs()
    print(f"imu400 video: max-abs {err.max():.3e} mean-abs {err.mean():.3e}")
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS
    assert float(video.double().sum()) == pytest.approx(float(g["video_fingerprint"][0]), rel=1e-3)


def test_conjoined_batch_invariance():
    """Size-independent property: a sample's prediction does not depend on its batch neighbours (bit-exact), here with
    ragged rows so the null-token padding differs between the batched and the single-sample run."""
    m, name, x, mask, imu, mc = _build("conj_padded_small_predict")
    B = 6
    xs = synthetic.make_video(B, (32, 32), seed=77).to(DEV).transpose(1, 2)
    masks = synthetic.make_mask(B, (2, 8, 8), num_clumps=2, seed=78).to(DEV)
    imus = synthetic.make_imu(B, 80, seed=79).to(DEV)
    mcs = torch.zeros(B, 5, dtype=torch.bool, device=DEV)
    full = m(xs, masks, x_context=imus, mask_context=mcs, output_main=True, output_context=False)
    m._reset_padding_mask()
    for i in (0, 3, 5):
        alone = m(xs[i:i + 1], masks[i:i + 1], x_context=imus[i:i + 1], mask_context=mcs[i:i + 1])
        m._reset_padding_mask()
        assert torch.equal(alone[0], full[i])


def test_flow2imu_full_size_matches_reference_fixture():
    """a17 at FULL size: the factory `imu400_8x8patch_2frames_1tube_flowbackrgb01` (conjoined_vmae.py:1218-1228; ViT-base,
    784 tokens of the 7-channel flow | flow-back | rgb input its preprocessor builds with RAFT, 25 masked IMU tokens + the
    dummy token), called like `ImuConditionedFlowGenerator.predict_imu_from_video` does (segmentation.py:834-846).
    Fixture: the REAL reference on CPU, seeded RAFT-large in the preprocessor (oracle/make_golden_flow2imu.py).
    The flow network runs in fp32 here (the parity configuration), so its 2.7e-6-of-scale error is invisible next to the
    f16-operand error of the ViT; bars: predicted IMU tokens within 2e-2 max-abs / 2e-3 mean-abs."""
    import make_golden_flow2imu as mgf
    from counterfactualworldmodels_b200 import raft
    path = os.path.join(GOLDEN_DIR, "flow2imu_full_b2.npz")
    if not os.path.exists(path):
        pytest.skip("fixture flow2imu_full_b2 missing")
    g = np.load(path)
    torch.manual_seed(0)
    rargs = raft.get_args("")
    rargs.multiframe, rargs.scale_inputs, rargs.output_dim = True, True, None
    flow_model = raft.RAFT(rargs).eval().requires_grad_(False)
    m = mgf.build(C, 'flow_model', flow_model).to(DEV)
    assert sum(p.numel() for p in m.parameters()) == mgf.PARAMS == int(g["num_params"][0])
    assert synthetic.weights_checksum(m) == pytest.approx(float(g["weights_checksum"][0]), rel=1e-9)
    x, mask, imu, mc = mgf.inputs()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y = m(oracle.preprocess(x).to(DEV), mask=mask.to(DEV), x_context=imu.to(DEV), mask_context=mc.to(DEV),
              output_main=False, output_context=True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    _check(y, torch.from_numpy(g["y_ctx"]), "flow2imu full size, IMU tokens")
