"""End-to-end parity of the CUDA path against the fixtures the reference produced (tests/golden) and against the
CPU oracle.  Tolerance from BASELINE.json north_star: masks / indices / compaction bit-exact; predicted pixels
max-abs <= 2e-2 and mean-abs <= 2e-3 in normalised pixel space (f16 operands, fp32 accumulation)."""
import numpy as np
import pytest
import torch

import vmae_oracle as oracle
from conftest import golden_case_inputs, load_golden
from counterfactualworldmodels_b200 import prediction, synthetic, vmae

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MAX_ABS, MEAN_ABS = 2e-2, 2e-3

ALL_CASES = ["tiny_4x4_b2", "tiny_8x8_b3", "small_4x4_b2", "small_4x4_allvisible_frame1half",
             "tiny_4x4_tube2_b2", "tiny_8x8_layerscale_learnpos_b2",  # tubelet 2 / layer scale + learnable pos-embed
             "base_8x8_b1_factual", "base_8x8_b2_counterfactual", "base_4x4_b1", "large_4x4_b1_factual"]


def _model(cfg_name, wseed, style):
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
    synthetic.init_weights_(m, seed=wseed, style=style)
    return m.to(DEV).eval()


@pytest.mark.parametrize("case", ALL_CASES)
def test_forward_matches_reference_fixture(case):
    g = load_golden(case)
    cfg_name, B, style, wseed, x = golden_case_inputs(case)
    m = _model(cfg_name, wseed, style)
    assert synthetic.weights_checksum(m) == pytest.approx(float(g["weights_checksum"][0]), abs=1e-6)
    mask = g["mask"].to(DEV)
    xin = oracle.preprocess(x).to(DEV)                    # the boundary tensor: normalised, transposed VIEW
    assert not xin.is_contiguous()
    y = m(xin, mask).cpu()
    assert y.shape == g["y"].shape
    err = (y - g["y"]).abs()
    print(f"{case}: max-abs {err.max():.3e} mean-abs {err.mean():.3e} (ref std {g['y'].std():.3f})")
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS
    # integer side: the permutation the forward used is bit-exact
    perm, inv, n_vis = m.last_aux
    perm_o, inv_o, nvis_o = oracle.compact_mask(g["mask"].numpy())
    assert np.array_equal(perm.cpu().numpy(), perm_o) and n_vis == int(nvis_o[0])


@pytest.mark.parametrize("case", ["tiny_4x4_b2", "small_4x4_b2", "base_8x8_b2_counterfactual", "tiny_4x4_tube2_b2",
                                  "tiny_8x8_layerscale_learnpos_b2"])
def test_predict_wrapper_matches_reference_video(case):
    """`PredictorBasedGenerator.predict(x, mask, frame=None)` with raw [0,1] input (fused normalisation)."""
    g = load_golden(case)
    cfg_name, B, style, wseed, x = golden_case_inputs(case)
    m = _model(cfg_name, wseed, style)
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    video = G.predict(x.to(DEV), g["mask"].to(DEV), frame=None).cpu()
    ps = synthetic.oracle_cfg(cfg_name)["patch_size"]
    want = oracle.pred_patches_to_video(g["y"], x, g["mask"], ps)
    if "video" in g:
        assert torch.equal(want, g["video"])
    err = (video - want).abs()
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS
    # visible patches are bit-identical to the input, only masked patches differ
    vp, xp = oracle.patchify(video, ps), oracle.patchify(x, ps)
    assert torch.equal(vp[~g["mask"]], xp[~g["mask"]])
    last = G.predict(x.to(DEV), g["mask"].to(DEV), frame=-1).cpu()
    assert last.shape[1] == 1 and torch.equal(last[:, 0], video[:, -1])


def test_fused_normalisation_equals_boundary_normalisation():
    cfg = "tiny_8x8"
    m = _model(cfg, 5, "perturbed")
    x = synthetic.make_video(2, synthetic.image_hw(cfg), seed=1)
    mask = synthetic.make_mask(2, m.mask_size, 2, seed=1).to(DEV)
    y0 = m(oracle.preprocess(x).to(DEV), mask)
    y1 = m(x.to(DEV).transpose(1, 2), mask, input_norm=(oracle.IMAGENET_DEFAULT_MEAN, oracle.IMAGENET_DEFAULT_STD))
    assert torch.equal(y0, y1)


def test_ragged_rows_raise_like_reference():
    cfg = "tiny_4x4"
    m = _model(cfg, 1, "reference")
    x = synthetic.make_video(2, synthetic.image_hw(cfg), seed=1)
    mask = synthetic.make_mask(2, m.mask_size, 1, seed=1)
    mask[1, -1] = ~mask[1, -1]
    with pytest.raises(RuntimeError, match="invalid"):
        m(oracle.preprocess(x).to(DEV), mask.to(DEV))


def test_load_state_dict_invalidates_device_weights():
    cfg = "tiny_4x4"
    m = _model(cfg, 1, "perturbed")
    x = oracle.preprocess(synthetic.make_video(1, synthetic.image_hw(cfg), seed=2)).to(DEV)
    mask = synthetic.make_mask(1, m.mask_size, 2, seed=2).to(DEV)
    y1 = m(x, mask).clone()
    other = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg))
    synthetic.init_weights_(other, seed=99, style="perturbed")
    m.load_state_dict(other.state_dict())
    y2 = m(x, mask)
    assert (y1 - y2).abs().max().item() > 1e-2
    want = oracle.vmae_forward(other.state_dict(), x.cpu(), mask.cpu(), synthetic.oracle_cfg(cfg))
    assert (y2.cpu() - want).abs().max().item() <= MAX_ABS


def test_batch_invariance_full_size_base_8x8():
    """Size-independent property at BASELINE config 2 (batch 64): every sample's prediction is independent of its
    batch neighbours, so sample i of the batch equals the same sample run alone -- bit for bit."""
    cfg = "base_8x8"
    m = _model(cfg, 0, "reference")
    B = 64
    x = synthetic.make_video(B, synthetic.image_hw(cfg), seed=21).to(DEV)
    mask = synthetic.make_mask(B, m.mask_size, num_clumps=1, seed=22).to(DEV)
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    full = G.predict(x, mask, frame=None)
    assert torch.isfinite(full).all()
    for i in (0, 17, 63):
        alone = G.predict(x[i:i + 1], mask[i:i + 1], frame=None)
        assert torch.equal(alone[0], full[i])
    # checksum-of-checksums: chunked execution == whole batch
    chunks = G.batch_predict_per_sample(x, mask, frame=None, batch_size=24, sample_dim=0)
    assert torch.equal(chunks, full)


def test_oracle_on_gpu_box_small():
    """The CPU oracle run on the GPU box's host agrees with the CUDA path for a fresh (non-fixture) input."""
    cfg = "small_4x4"
    m = _model(cfg, 31, "perturbed")
    x = synthetic.make_video(3, synthetic.image_hw(cfg), seed=32)
    mask = synthetic.make_mask(3, m.mask_size, num_clumps=5, seed=33)
    want = oracle.vmae_forward({k: v.cpu() for k, v in m.state_dict().items()}, oracle.preprocess(x), mask,
                               synthetic.oracle_cfg(cfg))
    got = m(oracle.preprocess(x).to(DEV), mask.to(DEV)).cpu()
    err = (got - want).abs()
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS


def test_host_pipeline_equals_predict():
    """prediction.HostPipeline (pinned host buffers, copies overlapped with compute on 3 streams) returns exactly
    what a per-batch `predict` returns, for more batches than it has device slots."""
    cfg = "tiny_8x8"
    m = _model(cfg, 7, "perturbed")
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    B, n = 4, 5
    xs = [synthetic.make_video(B, synthetic.image_hw(cfg), seed=40 + i).pin_memory() for i in range(n)]
    ms = [synthetic.make_mask(B, m.mask_size, num_clumps=2, seed=50 + i).pin_memory() for i in range(n)]
    outs = [torch.empty(xs[0].shape).pin_memory() for _ in range(n)]
    pipe = prediction.HostPipeline(G, tuple(xs[0].shape), ms[0].shape[1], device=DEV)
    for i in range(n):
        pipe.submit(xs[i], ms[i], outs[i], frame=None)
    pipe.finish()
    torch.cuda.synchronize()
    for i in range(n):
        want = G.predict(xs[i].to(DEV), ms[i].to(DEV), frame=None).cpu()
        assert torch.equal(outs[i], want)
    assert pipe.h2d_bytes == xs[0].numel() * 4 + ms[0].numel() and pipe.d2h_bytes == outs[0].numel() * 4
