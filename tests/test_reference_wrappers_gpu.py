"""The drop-in claim, executed: the REFERENCE's own wrappers (`cwm.models.prediction.PredictorBasedGenerator.predict`,
prediction.py:406-454; `cwm.models.segmentation.FlowGenerator.predict_counterfactual_videos_and_flows`,
segmentation.py:346-432) drive this repo's predictor on the GPU, unchanged, and the result is compared with the same
wrapper driving the reference's own predictor on the CPU (identical weights, inputs and masks).

The reference is imported from the staged, unmodified copy under the git-ignored ``baseline/_ref/`` (it travels to the
GPU box; ``baseline/stage_reference.py``), with in-process stubs for the absent timm / kornia / matplotlib packages.
Tolerance from BASELINE.json north_star: pixels max-abs <= 2e-2, mean-abs <= 2e-3; masks / indices bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import needs_reference
from counterfactualworldmodels_b200 import synthetic, vmae

pytestmark = [pytest.mark.gpu, needs_reference]
DEV = "cuda:0"
MAX_ABS, MEAN_ABS = 2e-2, 2e-3


def _reference():
    import ref_loader
    ref_vmae, ref_pred = ref_loader.import_reference()
    import cwm.models.segmentation as ref_seg
    return ref_vmae, ref_pred, ref_seg


def _pair(cfg, wseed, style, **overrides):
    """(reference predictor on the CPU, drop-in predictor on the GPU) holding the same state_dict."""
    ref_vmae, _, _ = _reference()
    torch.manual_seed(0)
    kw = dict(synthetic.model_kwargs(cfg), **overrides)
    ref = ref_vmae.PretrainVisionTransformer(**kw).eval().requires_grad_(False)
    synthetic.init_weights_(ref, seed=wseed, style=style)
    ours = vmae.PretrainVisionTransformer(**kw)
    missing = ours.load_state_dict(ref.state_dict())          # the reference's checkpoint loads unchanged
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref, ours.to(DEV).eval()


@pytest.mark.parametrize("cfg,B,clumps,frame", [("tiny_8x8", 3, 2, None), ("small_4x4", 2, 1, -1),
                                                ("base_8x8", 2, 1, None)])
def test_reference_generator_predict_over_the_dropin_predictor(cfg, B, clumps, frame):
    _, ref_pred, _ = _reference()
    ref, ours = _pair(cfg, 4, "perturbed" if cfg != "base_8x8" else "reference")
    x = synthetic.make_video(B, synthetic.image_hw(cfg), seed=31)
    mask = synthetic.make_mask(B, ref.mask_size, num_clumps=clumps, seed=32)
    G_ref = ref_pred.PredictorBasedGenerator(predictor=ref, imagenet_normalize_inputs=True, temporal_dim=2)
    G_our = ref_pred.PredictorBasedGenerator(predictor=ours, imagenet_normalize_inputs=True, temporal_dim=2)
    with torch.no_grad():
        want = G_ref.predict(x.clone(), mask.clone(), frame=frame)
        got = G_our.predict(x.to(DEV), mask.to(DEV), frame=frame).cpu()
    assert got.shape == want.shape
    err = (got - want).abs()
    print(f"{cfg}: reference wrapper over the drop-in: max-abs {err.max():.3e} mean-abs {err.mean():.3e}")
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS
    if frame is None:       # frame 0 is fully visible: the wrapper's scatter returns the input bit for bit
        assert torch.equal(got[:, 0], x[:, 0])


def test_reference_generator_helpers_read_the_dropin_attributes():
    """The attributes the reference wrappers read off the predictor resolve on the drop-in."""
    _, ref_pred, _ = _reference()
    ref, ours = _pair("tiny_8x8", 1, "perturbed")
    G_ref = ref_pred.PredictorBasedGenerator(predictor=ref, imagenet_normalize_inputs=True, temporal_dim=2)
    G_our = ref_pred.PredictorBasedGenerator(predictor=ours, imagenet_normalize_inputs=True, temporal_dim=2)
    x = synthetic.make_video(1, synthetic.image_hw("tiny_8x8"), seed=3)
    G_ref.set_input(x)
    G_our.set_input(x.to(DEV))
    for name in ("patch_size", "mask_shape", "inp_shape", "num_frames"):
        a, b = getattr(G_ref, name, None), getattr(G_our, name, None)
        assert tuple(np.atleast_1d(a)) == tuple(np.atleast_1d(b)), name
    assert tuple(ours.mask_size) == tuple(ref.mask_size) and ours.num_patches == ref.num_patches
    assert tuple(ours.image_size) == tuple(ref.image_size)


def _sweep_inputs(mask_size, S, seed):
    T, h, w = mask_size
    rng = np.random.RandomState(seed)
    active = torch.ones(1, T, h, w, S, dtype=torch.bool)
    passive = torch.zeros(1, T, h, w, S, dtype=torch.bool)
    passive[:, -1] = True
    for s in range(S):
        ay, ax = 2 * rng.randint(1, h // 4), 2 * rng.randint(1, w // 4)
        py, px = 2 * rng.randint(1, h // 4), 2 * rng.randint(w // 4 + 1, w // 2 - 1)
        active[0, -1, ay:ay + 2, ax:ax + 2, s] = False
        passive[0, -1, py:py + 2, px:px + 2, s] = False
    preset = [[2, 0], [0, 2], [-2, 0], [0, -2], [2, 2], [-2, -2], [2, -2], [-2, 2]]
    return active.reshape(1, -1, S), passive.reshape(1, -1, S), [preset[s % 8] for s in range(S)]


@pytest.mark.parametrize("flow_model_kind", ["reference_raft", "dropin_raft"])
def test_reference_flow_generator_sweep_over_the_dropin_predictor(flow_model_kind):
    """segmentation.py:346-432 end to end: the reference's FlowGenerator builds the counterfactual prompts itself
    (its own per-sample host loop), predicts them in chunks with the drop-in predictor, and runs a flow network."""
    _, _, ref_seg = _reference()
    import cwm.models.raft.raft_model as ref_raft
    from counterfactualworldmodels_b200 import raft as our_raft
    S, hw = 6, (128, 128)       # 128 px: RAFT's 1/8-resolution maps are 16x16, enough for its 4-level pyramid
    ref, ours = _pair("tiny_8x8", 7, "perturbed", img_size=hw[0])
    torch.manual_seed(5)
    rargs = ref_raft.get_args("")
    rargs.multiframe, rargs.scale_inputs = True, True
    raft_cpu = ref_raft.RAFT(rargs).eval().requires_grad_(False)
    if flow_model_kind == "reference_raft":
        raft_dev = ref_raft.RAFT(rargs).eval().requires_grad_(False)
    else:
        raft_dev = our_raft.RAFT(our_raft.get_args("")).eval().requires_grad_(False)
        raft_dev.args.multiframe = raft_dev.multiframe = True
        raft_dev.args.scale_inputs = raft_dev.scale_inputs = True
        # the reference's `set_raft_iters` (segmentation.py:86-90) only finds instances of ITS RAFT class, so the
        # iteration count of a drop-in flow network is set on the module itself
        raft_dev.iters = 4
    raft_dev.load_state_dict(raft_cpu.state_dict())
    raft_dev = raft_dev.to(DEV)
    x = synthetic.make_video(1, hw, seed=41)[:, 0]
    a, p, shifts = _sweep_inputs(ref.mask_size, S, seed=42)

    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        F_ref = ref_seg.FlowGenerator(predictor=ref, flow_model=raft_cpu, raft_iters=4,
                                      imagenet_normalize_inputs=True, temporal_dim=2)
        F_our = ref_seg.FlowGenerator(predictor=ours, flow_model=raft_dev, raft_iters=4,
                                      imagenet_normalize_inputs=True, temporal_dim=2)
        with torch.no_grad():
            y_ref, f_ref = F_ref.predict_counterfactual_videos_and_flows(
                x.clone(), a.clone(), passive_patches=p.clone(), shifts=shifts, sample_batch_size=4, raft_iters=4)
            y_our, f_our = F_our.predict_counterfactual_videos_and_flows(
                x.to(DEV), a.to(DEV), passive_patches=p.to(DEV), shifts=shifts, sample_batch_size=4, raft_iters=4)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    y_our, f_our = y_our.cpu(), f_our.cpu()
    assert y_our.shape == y_ref.shape and f_our.shape == f_ref.shape
    err = (y_our - y_ref).abs()
    print(f"sweep ({flow_model_kind}): videos max-abs {err.max():.3e} mean-abs {err.mean():.3e}")
    assert err.max().item() <= MAX_ABS and err.mean().item() <= MEAN_ABS
    assert torch.equal(y_our[:, 0], y_ref[:, 0])         # prompts' frame 0 (all visible): bit-identical
    # the flows see the f16-operand pixel differences through a random-init RAFT: judged at 2e-2 of the flow scale
    ferr = (f_our - f_ref).abs().max().item() / max(f_ref.abs().max().item(), 1e-6)
    print(f"sweep ({flow_model_kind}): flows differ by {ferr:.2e} of scale")
    assert ferr <= 2e-2


def test_reference_imu_conditioned_generator_over_the_dropin_conjoined_predictors():
    """BASELINE config 5's driver, the reference's OWN class (segmentation.py:756-967 `ImuConditionedFlowGenerator`, with the
    `ImuGenerator` it builds inside, :549-754): flow2imu -- this repo's RAFT inside its preprocessor -- predicts the head
    motion of the static movie once, then the S counterfactuals are predicted by the drop-in IMU-conditioned padded
    predictor.  Compared with what the same class produced over the reference's predictors on the CPU
    (tests/golden/imu_sweep_128px.npz, oracle/make_golden_imu_sweep.py)."""
    import sys
    _, _, ref_seg = _reference()
    import make_golden_imu_sweep as mg
    from counterfactualworldmodels_b200 import conjoined_vmae as conj
    from counterfactualworldmodels_b200 import raft as our_raft

    def flow_net():
        torch.manual_seed(0)
        args = our_raft.get_args("")
        args.multiframe, args.scale_inputs, args.output_dim = True, True, None
        net = our_raft.RAFT(args).eval().requires_grad_(False)
        net.iters = mg.RAFT_ITERS          # the reference's set_raft_iters only finds ITS RAFT class
        return net

    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imu_sweep_128px.npz")
    d = np.load(path)
    head_flow = flow_net()
    pred, head = mg.build(conj, ref_seg, our_raft, 'flow_model', head_flow)
    G = ref_seg.ImuConditionedFlowGenerator(predictor=pred.to(DEV), head_motion_predictor=head.to(DEV),
                                            flow_model=flow_net().to(DEV), imagenet_normalize_inputs=True, temporal_dim=2,
                                            raft_iters=mg.RAFT_ITERS, seed=0)
    # reference quirk (segmentation.py:549-575): constructing the inner ImuGenerator re-sets every RAFT *of the reference's
    # class* it can see to 24 iterations -- the one inside flow2imu's preprocessor included; a drop-in flow network is not
    # found by that isinstance walk, so the same state is set by hand
    head_flow.iters = 24
    assert type(G).__module__ == "cwm.models.segmentation" and "cwm.models.segmentation" in sys.modules
    x, active, passive = mg.sweep_inputs()
    x, active, passive = x.to(DEV), active.to(DEV), passive.to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            G.set_input(x)
            h = G.get_static_imu()
            G.reset_padding_masks()
            ys, flows = G.predict_counterfactual_videos_and_flows(x, active, passive, shifts=mg.SHIFTS, sample_batch_size=2,
                                                                  raft_iters=mg.RAFT_ITERS)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert tuple(h.shape) == (1, 5, 96) and tuple(ys.shape) == (4, 2, 3, 128, 128) and tuple(flows.shape) == (4, 1, 2, 128, 128)
    want_h, want_y = d["head_motion"], d["ys_frame1"]
    e_h = float(np.abs(h.cpu().numpy() - want_h).max() / np.abs(want_h).max())
    err = np.abs(ys[:, 1, :, ::2, ::2].cpu().numpy() - want_y)
    print(f"reference ImuConditionedFlowGenerator over the drop-ins: head motion {e_h:.2e} of scale, frame-1 pixels "
          f"max-abs {err.max():.2e} mean-abs {err.mean():.2e}")
    assert e_h <= 2e-2 and err.max() <= MAX_ABS and err.mean() <= MEAN_ABS
    assert torch.isfinite(flows).all()
