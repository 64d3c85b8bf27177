"""SURVEY.md section 8(f) rank 3, first slice -- RAFT's correlation block and convex upsampling.
CPU: the numpy oracle against the fixtures the REAL reference produced (tests/golden/raft_*.npz); host-side validation.
GPU: csrc/raftcorr.cu through the mirror (``raft.CorrBlock`` / ``raft.upsample_flow``) against the fixtures and the oracle.

Tolerance: fp32 everywhere.  The volume sums D products in a different order than BLAS and the lookup contracts
multiply-adds, so values agree to a few ulp of the largest term: 2e-5 of the output scale (max |reference|) is the
bar; the oracle itself is pinned to the reference at 2e-6 of scale by oracle/make_golden_raft.py."""
import ctypes
import os

import numpy as np
import pytest
import torch

import raft_oracle as ro
from conftest import GOLDEN_DIR

CORR_CASES = ["raft_corr_l3_r2_8x12_b2", "raft_corr_l4_r4_16x24_b1", "raft_corr_l4_r3_17x19_b1"]
UPSAMPLE_CASES = ["raft_upsample_n2_5x36", "raft_upsample_n1_c3_4x7"]
DEV = "cuda:0"
ORACLE_TOL, GPU_TOL = 2e-6, 2e-5


def load(case):
    path = os.path.join(GOLDEN_DIR, case + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {case} missing")
    z = np.load(path)
    return {k: z[k] for k in z.files}


def rel_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


def stored_level(ref_l, lvl, n_rows):
    return ref_l[::7] if lvl < 2 and n_rows > 200 else ref_l


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("case", CORR_CASES)
def test_oracle_corr_block_matches_reference_fixture(case):
    d = load(case)
    B, D, H, W, L, r, seed, sy, sx = [int(v) for v in d["shape"]]
    f1, f2 = ro.make_fmaps(B, D, H, W, seed)
    pyr = ro.corr_pyramid(f1, f2, L)
    for lvl in range(L):
        assert pyr[lvl].shape == (B * H * W, H >> lvl, W >> lvl)
        assert rel_err(stored_level(pyr[lvl], lvl, B * H * W), d[f"level{lvl}"]) <= ORACLE_TOL
    for kind in ("grid", "random"):
        out = ro.corr_lookup(pyr, ro.make_coords(B, H, W, seed, kind), r)
        assert out.shape == (B, L * (2 * r + 1) ** 2, H, W)
        assert rel_err(out[:, :, ::sy, ::sx], d[f"lookup_{kind}"]) <= ORACLE_TOL


def test_oracle_matches_the_trace_of_the_reference_raft():
    d = load("raft_trace_large_128px")
    f1, f2 = d["fmap1"].astype(np.float32), d["fmap2"].astype(np.float32)
    pyr = ro.corr_pyramid(f1, f2, 4)
    for it in range(3):
        out = ro.corr_lookup(pyr, d[f"coords{it}"], 4)
        assert rel_err(out[:, :, ::2, ::2], d[f"lookup{it}"]) <= ORACLE_TOL
    # iteration 0 looks up the exact integer grid: the centre tap of level 0 is the volume's own diagonal entry
    centre = 4 * 9 + 4
    diag = pyr[0].reshape(256, 256)[np.arange(256), np.arange(256)].reshape(1, 16, 16)
    np.testing.assert_allclose(ro.corr_lookup(pyr, d["coords0"], 4)[:, centre], diag, rtol=0, atol=2e-5 * np.abs(diag).max())


@pytest.mark.parametrize("case", UPSAMPLE_CASES)
def test_oracle_upsample_matches_reference_fixture(case):
    d = load(case)
    N, C, H, W, seed, stride = [int(v) for v in d["shape"]]
    flow, mask = ro.make_upsample_inputs(N, C, H, W, seed)
    up = ro.upsample_flow(flow, mask)
    assert up.shape == (N, C, 8 * H, 8 * W)
    assert rel_err(up[:, :, ::stride], d["up"]) <= ORACLE_TOL


def test_oracle_upsample_of_a_constant_flow_is_that_flow_times_8_inside():
    flow = np.full((1, 2, 4, 5), 1.5, np.float32)
    _, mask = ro.make_upsample_inputs(1, 2, 4, 5, 3)
    up = ro.upsample_flow(flow, mask)
    # away from the zero padding every convex combination of equal neighbours is the neighbour itself
    np.testing.assert_allclose(up[:, :, 8:-8, 8:-8], 12.0, rtol=1e-6)
    assert up[:, :, :8].max() <= 12.0 + 1e-5


def test_mirror_has_no_cpu_fallback_and_the_abi_validates_arguments():
    from counterfactualworldmodels_b200 import _lib, raft
    f = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        raft.CorrBlock(f, f)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        raft.upsample_flow(torch.zeros(1, 2, 4, 4), torch.zeros(1, 576, 4, 4))
    g = raft.coords_grid(2, 3, 5, "cpu")
    assert g.shape == (2, 2, 3, 5) and g[1, 0, 2, 4] == 4 and g[1, 1, 2, 4] == 2
    lib = _lib.load()
    table = (ctypes.c_void_p * 4)(1, 1, 1, 1)
    assert lib.cwm_raft_corr_pyramid(1, 1, 1, 8, 8, 8, 4, table, None) == -1       # level 3 of 8x8 is 1x1
    assert b"smaller than 2x2" in lib.cwm_last_error()
    assert lib.cwm_raft_corr_pyramid(1, 1, 1, 8, 16, 16, 9, table, None) == -1 and b"num_levels" in lib.cwm_last_error()
    assert lib.cwm_raft_corr_pyramid(None, None, 1, 8, 16, 16, 4, table, None) == -1
    assert lib.cwm_raft_corr_lookup(table, 4, 9, 1, 1, 16, 16, 1, None) == -1 and b"radius" in lib.cwm_last_error()
    assert lib.cwm_raft_corr_lookup(table, 4, 4, None, 1, 16, 16, None, None) == -1
    assert lib.cwm_raft_upsample_flow(None, None, 1, 2, 4, 4, None, None) == -1
    assert lib.cwm_raft_upsample_flow(1, 1, 1, 0, 4, 4, 1, None) == -1
    assert lib.cwm_raft_corr_pyramid(None, None, 0, 8, 16, 16, 4, table, None) == 0  # empty batch is a no-op


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", CORR_CASES)
def test_gpu_corr_block_matches_reference_fixture(case):
    from counterfactualworldmodels_b200 import raft
    d = load(case)
    B, D, H, W, L, r, seed, sy, sx = [int(v) for v in d["shape"]]
    f1, f2 = ro.make_fmaps(B, D, H, W, seed)
    block = raft.CorrBlock(torch.from_numpy(f1).to(DEV), torch.from_numpy(f2).to(DEV), num_levels=L, radius=r)
    assert len(block.corr_pyramid) == L
    pyr_o = ro.corr_pyramid(f1, f2, L)
    for lvl in range(L):
        got = block.corr_pyramid[lvl]
        assert tuple(got.shape) == (B * H * W, 1, H >> lvl, W >> lvl)
        got = got[:, 0].cpu().numpy()
        assert rel_err(stored_level(got, lvl, B * H * W), d[f"level{lvl}"]) <= GPU_TOL
        assert rel_err(got, pyr_o[lvl]) <= GPU_TOL
    for kind in ("grid", "random"):
        coords = ro.make_coords(B, H, W, seed, kind)
        out = block(torch.from_numpy(coords).to(DEV))
        assert out.dtype == torch.float32 and out.is_contiguous()
        out = out.cpu().numpy()
        assert rel_err(out[:, :, ::sy, ::sx], d[f"lookup_{kind}"]) <= GPU_TOL
        assert rel_err(out, ro.corr_lookup(pyr_o, coords, r)) <= GPU_TOL


@pytest.mark.gpu
def test_gpu_corr_block_on_the_trace_of_the_reference_raft():
    from counterfactualworldmodels_b200 import raft
    d = load("raft_trace_large_128px")
    f1, f2 = torch.from_numpy(d["fmap1"]).to(DEV), torch.from_numpy(d["fmap2"]).to(DEV)  # f16: the mirror casts
    block = raft.CorrBlock(f1, f2, radius=4)
    for it in range(3):
        out = block(torch.from_numpy(d[f"coords{it}"]).to(DEV)).cpu().numpy()
        assert rel_err(out[:, :, ::2, ::2], d[f"lookup{it}"]) <= GPU_TOL
    vol = raft.CorrBlock.corr(f1, f2)
    assert tuple(vol.shape) == (1, 16, 16, 1, 16, 16)
    assert torch.equal(vol.reshape(256, 16, 16), block.corr_pyramid[0][:, 0])


@pytest.mark.gpu
def test_gpu_corr_block_at_the_sweep_shape_against_the_oracle():
    """224 px input -> 28x28 maps, D = 256 (RAFT-large), 3 samples: full comparison with the oracle, integer grid
    and far-out-of-bounds centres included; a non-default stream; NaN centres give zeros, not a fault."""
    from counterfactualworldmodels_b200 import raft
    B, D, H, W = 3, 256, 28, 28
    f1, f2 = ro.make_fmaps(B, D, H, W, 5)
    pyr_o = ro.corr_pyramid(f1, f2, 4)
    s = torch.cuda.Stream(device=DEV)
    with torch.cuda.stream(s):
        block = raft.CorrBlock(torch.from_numpy(f1).to(DEV), torch.from_numpy(f2).to(DEV))
        outs = [block(torch.from_numpy(ro.make_coords(B, H, W, 5, kind)).to(DEV)) for kind in ("grid", "random")]
        bad = torch.from_numpy(ro.make_coords(B, H, W, 5, "grid")).to(DEV)
        bad[0, 0, 3, 4] = float("nan")
        bad[1, 1, 0, 0] = float("inf")
        bad[2, :, 5, 5] = 3.0e9
        out_bad = block(bad)
    s.synchronize()
    for lvl in range(4):
        assert rel_err(block.corr_pyramid[lvl][:, 0].cpu().numpy(), pyr_o[lvl]) <= GPU_TOL
    for kind, out in zip(("grid", "random"), outs):
        assert rel_err(out.cpu().numpy(), ro.corr_lookup(pyr_o, ro.make_coords(B, H, W, 5, kind), 4)) <= GPU_TOL
    assert torch.isfinite(out_bad).all()
    assert out_bad[0, :, 3, 4].abs().max() == 0 and out_bad[1, :, 0, 0].abs().max() == 0
    assert out_bad[2, :, 5, 5].abs().max() == 0
    keep = torch.ones(B, H, W, dtype=torch.bool, device=DEV)
    keep[0, 3, 4] = keep[1, 0, 0] = keep[2, 5, 5] = False
    assert torch.equal(out_bad.permute(0, 2, 3, 1)[keep], outs[0].permute(0, 2, 3, 1)[keep])


@pytest.mark.gpu
def test_gpu_corr_lookup_properties_at_sweep_batch():
    """64 samples at 28x28 (a sweep chunk; the oracle would take minutes): size-independent properties.
    (i) the centre tap of level 0 at the integer grid is the volume's diagonal-by-construction entry;
    (ii) a lookup shifted by one whole pixel is the neighbouring channel of the unshifted lookup;
    (iii) the volume is linear in fmap2."""
    from counterfactualworldmodels_b200 import raft
    B, D, H, W, r = 64, 256, 28, 28, 4
    g = torch.Generator(device=DEV).manual_seed(0)
    f1 = torch.randn(B, D, H, W, device=DEV, generator=g)
    f2 = torch.randn(B, D, H, W, device=DEV, generator=g)
    block = raft.CorrBlock(f1, f2, radius=r)
    grid = raft.coords_grid(B, H, W, DEV)
    out = block(grid)
    assert tuple(out.shape) == (B, 4 * 81, H, W)
    vol = block.corr_pyramid[0].view(B, H * W, H * W)
    diag = torch.diagonal(vol, dim1=1, dim2=2).reshape(B, H, W)
    assert (out[:, 4 * 9 + 4] - diag).abs().max() <= 1e-5 * diag.abs().max()
    shifted = block(grid + torch.tensor([1.0, 0.0], device=DEV).view(1, 2, 1, 1))   # x + 1
    # channel a*9 + b samples x offset a - r: shifting the centre by +1 in x moves tap a to a + 1
    lvl0, lvl0_s = out[:, :81].view(B, 9, 9, H, W), shifted[:, :81].view(B, 9, 9, H, W)
    assert (lvl0_s[:, :-1] - lvl0[:, 1:]).abs().max() <= 1e-5 * vol.abs().max()
    block2 = raft.CorrBlock(f1, 2.0 * f2, radius=r)
    assert torch.equal(block2.corr_pyramid[0], 2.0 * block.corr_pyramid[0])
    # pooling: every level is the 2x2 mean of the previous one
    for lvl in range(1, 4):
        want = torch.nn.functional.avg_pool2d(block.corr_pyramid[lvl - 1], 2, stride=2)
        assert (block.corr_pyramid[lvl] - want).abs().max() <= 1e-6 * vol.abs().max()


@pytest.mark.gpu
@pytest.mark.parametrize("case", UPSAMPLE_CASES)
def test_gpu_upsample_matches_reference_fixture(case):
    from counterfactualworldmodels_b200 import raft
    d = load(case)
    N, C, H, W, seed, stride = [int(v) for v in d["shape"]]
    flow, mask = ro.make_upsample_inputs(N, C, H, W, seed)
    up = raft.upsample_flow(torch.from_numpy(flow).to(DEV), torch.from_numpy(mask).to(DEV)).cpu().numpy()
    assert rel_err(up[:, :, ::stride], d["up"]) <= GPU_TOL
    assert rel_err(up, ro.upsample_flow(flow, mask)) <= GPU_TOL


@pytest.mark.gpu
def test_gpu_upsample_at_the_sweep_shape():
    from counterfactualworldmodels_b200 import raft
    flow, mask = ro.make_upsample_inputs(4, 2, 28, 28, 9)
    up = raft.upsample_flow(torch.from_numpy(flow).to(DEV), torch.from_numpy(mask).to(DEV))
    assert tuple(up.shape) == (4, 2, 224, 224)
    assert rel_err(up.cpu().numpy(), ro.upsample_flow(flow, mask)) <= GPU_TOL
    # convexity: every output lies inside the range of the 3x3 neighbourhood of 8*flow (zero padding included)
    f8 = torch.nn.functional.pad(8 * torch.from_numpy(flow), (1, 1, 1, 1))
    hi = torch.nn.functional.max_pool2d(f8, 3, stride=1).repeat_interleave(8, 2).repeat_interleave(8, 3)
    lo = -torch.nn.functional.max_pool2d(-f8, 3, stride=1).repeat_interleave(8, 2).repeat_interleave(8, 3)
    u = up.cpu()
    assert (u <= hi + 1e-4).all() and (u >= lo - 1e-4).all()


def test_torch_port_used_as_cpu_baseline_agrees_with_the_oracle():
    f1, f2 = ro.make_fmaps(2, 16, 8, 12, 1)
    coords = [ro.make_coords(2, 8, 12, 1, k) for k in ("grid", "random")]
    outs = ro.torch_corr_block(torch.from_numpy(f1), torch.from_numpy(f2), [torch.from_numpy(c) for c in coords], 3, 2)
    pyr = ro.corr_pyramid(f1, f2, 3)
    for c, o in zip(coords, outs):
        assert rel_err(o.numpy(), ro.corr_lookup(pyr, c, 2)) <= ORACLE_TOL
    flow, mask = ro.make_upsample_inputs(1, 2, 4, 7, 2)
    up = ro.torch_upsample_flow(torch.from_numpy(flow), torch.from_numpy(mask)).numpy()
    assert rel_err(up, ro.upsample_flow(flow, mask)) <= ORACLE_TOL


# ------------------------------------------------------------------------------------------------ the network
def _mirror(small):
    from counterfactualworldmodels_b200 import raft
    torch.manual_seed(0)
    args = raft.get_args("")
    args.multiframe, args.scale_inputs, args.output_dim, args.small = True, True, None, small
    return raft.RAFT(args).eval().requires_grad_(False)


@pytest.mark.parametrize("small", [False, True])
def test_raft_mirror_reproduces_the_reference_parameters(small):
    """Same parameter names / shapes / seeded init as the reference's RAFT: the fixture holds an order-sensitive
    checksum of the REAL reference's state_dict under torch.manual_seed(0) (oracle/make_golden_raft.py)."""
    import make_golden_raft as mg
    d = load("raft_e2e_small_128px" if small else "raft_e2e_large_128px")
    model = _mirror(small)
    sd = model.state_dict()
    assert len(sd) == int(d["n_tensors"])
    assert mg.state_checksum(sd) == pytest.approx(float(d["init_checksum"]), rel=1e-12)
    assert sum(p.numel() for p in model.parameters()) == (990162 if small else 5257536)   # RAFT-large: SURVEY 8(c)
    for key in ("fnet.conv1.weight", "fnet.layer2.0.downsample.0.weight", "cnet.layer3.1.conv2.bias",
                "update_block.encoder.convc1.weight", "update_block.flow_head.conv2.bias"):
        assert key in sd
    if not small:
        assert "cnet.layer2.0.norm3.running_mean" in sd and "cnet.layer2.0.downsample.1.running_mean" in sd
        assert tuple(sd["update_block.mask.2.weight"].shape) == (576, 256, 1, 1)
        assert tuple(sd["update_block.gru.convz1.weight"].shape) == (128, 384, 1, 5)


def test_raft_loader_and_flow_generator_wiring():
    from counterfactualworldmodels_b200 import raft, segmentation, synthetic, vmae
    with pytest.raises(ValueError, match="not a valid raft checkpoint"):
        raft.load_raft_model(load_path="/nonexistent/raft-large.pth")
    with pytest.raises(NotImplementedError):
        a = raft.get_args("")
        a.alternate_corr = True
        raft.RAFT(a)
    model = raft.load_raft_model(load_path=None, output_dim=3, small=True)     # reference: a new RAFT with an output head
    assert model.output_block is not None and model.args.corr_radius == 3 and model.iters is None
    model.iters = 7
    assert model.iters == 7
    kw = synthetic.model_kwargs("tiny_4x4")
    G = segmentation.FlowGenerator(predictor=vmae.PretrainVisionTransformer(**kw), flow_model=_mirror(True))
    assert isinstance(G.flow_model, raft.RAFT) and not G.flow_model.training


@pytest.mark.gpu
@pytest.mark.parametrize("small", [False, True])
def test_gpu_raft_end_to_end_matches_the_reference(small):
    """The whole flow network (seeded init == the reference's, eval) on a seeded frame pair, forward and backward flow,
    4 GRU iterations: cuDNN fp32 convolutions (TF32 off) + this repo's correlation / lookup / upsampling kernels against
    the flows the REAL reference produced on CPU.  Tolerance 1e-4 of the flow scale (measured: 3e-6): fp32 convolutions
    summed in a different order, fed back through 4 recurrent iterations."""
    import make_golden_raft as mg
    d = load("raft_e2e_small_128px" if small else "raft_e2e_large_128px")
    model = _mirror(small).to(DEV)
    model.iters = int(d["iters"])
    x = mg.e2e_frames(2, 128).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        fwd = model(x)
        bwd = model(x, backward=True)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert tuple(fwd.shape) == (2, 1, 2, 128, 128)
    e_f, e_b = rel_err(fwd.cpu().numpy(), d["flow_fwd"]), rel_err(bwd.cpu().numpy()[:, :, :, ::2, ::2], d["flow_bwd"])
    print(f"raft e2e small={small}: fwd {e_f:.2e} bwd {e_b:.2e} of scale")
    assert e_f <= 1e-4 and e_b <= 1e-4


@pytest.mark.gpu
def test_gpu_raft_shared_first_frame_is_the_same_flow():
    """A counterfactual sweep shares frame 0: encoding it once (``shared_frame=0``, picked automatically by
    ``FlowGenerator.predict_flow``) gives the flows of the plain batched call, forward and backward."""
    import make_golden_raft as mg
    from counterfactualworldmodels_b200 import segmentation, synthetic, vmae
    model = _mirror(False).to(DEV)
    model.iters = 3
    x = mg.e2e_frames(2, 128).to(DEV).repeat(3, 1, 1, 1, 1)[:5]
    x[:, 1] = torch.roll(x[:, 1], shifts=(0, 1, 2), dims=(0, 2, 3))       # different second frames
    x[:, 0] = x[:1, 0]
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for backward in (False, True):
            plain = model(x, backward=backward)
            shared = model(x, backward=backward, shared_frame=0)
            assert rel_err(shared.cpu().numpy(), plain.cpu().numpy()) <= 1e-5
        G = segmentation.FlowGenerator(predictor=vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_4x4")),
                                       flow_model=model)
        calls = []
        orig = model._forward_two_images
        model._forward_two_images = lambda a, b, *r, **k: (calls.append((a.shape[0], b.shape[0])), orig(a, b, *r, **k))[1]
        auto = G.predict_flow(x, iters=3)
        x2 = x.clone()
        x2[3, 0] += 0.01
        G.predict_flow(x2, iters=3)
        del model._forward_two_images
        assert calls == [(1, 5), (5, 5)]
        assert rel_err(auto.cpu().numpy(), model(x).cpu().numpy()) <= 1e-5
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.gpu
def test_gpu_raft_mixed_precision_stays_close_to_the_fp32_reference():
    """``mixed_precision=True`` (the reference's autocast option; here: autocast encoders + the recurrent block on f16
    channels-last tensors): f16 convolutions with fp32 accumulation against the fp32 flows of the REAL reference.
    Tolerance 1e-2 of the flow scale (measured 1.8e-3: f16 activations through 4 recurrent iterations)."""
    import make_golden_raft as mg
    d = load("raft_e2e_large_128px")
    model = _mirror(False).to(DEV)
    model.iters = int(d["iters"])
    model.args.mixed_precision = True
    x = mg.e2e_frames(2, 128).to(DEV)
    errs = {}
    for half_update in (True, False):
        model.args.half_update = half_update
        fwd = model(x)
        assert fwd.dtype == torch.float32
        errs[half_update] = rel_err(fwd.cpu().numpy(), d["flow_fwd"])
    print(f"raft mixed precision: f16 update block {errs[True]:.2e}, autocast {errs[False]:.2e} of scale")
    assert errs[True] <= 1e-2 and errs[False] <= 1e-2
    assert "_half_ub" not in model.state_dict() and len(model.state_dict()) == int(d["n_tensors"])


# ------------------------------------------------------------------------------------------------ recurrent block
def test_oracle_gru_restatement_matches_torch_modules():
    """The elementwise restatement (oracle) + plain convolutions == the mirror's SepConvGRU / BasicMotionEncoder tail."""
    from counterfactualworldmodels_b200 import raft
    torch.manual_seed(3)
    gru = raft.SepConvGRU(hidden_dim=16, input_dim=24).eval().requires_grad_(False)
    h, x = torch.randn(2, 16, 6, 7), torch.randn(2, 24, 6, 7)
    want = gru(h, x)
    rows = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).numpy()  # noqa: E731
    unrows = lambda a, c: torch.from_numpy(a).reshape(2, 6, 7, c).permute(0, 3, 1, 2)  # noqa: E731
    hh = h
    for idx, pad in ((1, (0, 2)), (2, (2, 0))):
        cz, cr, cq = (getattr(gru, f"conv{g}{idx}") for g in "zrq")
        hx = torch.cat([hh, x], 1)
        zr = torch.cat([torch.nn.functional.conv2d(hx, cz.weight, None, 1, pad),
                        torch.nn.functional.conv2d(hx, cr.weight, None, 1, pad)], 1)
        z, rh = ro.gru_gate(rows(zr), torch.cat([cz.bias, cr.bias]).detach().numpy(), rows(hh))
        q = torch.nn.functional.conv2d(torch.cat([unrows(rh, 16), x], 1), cq.weight, None, 1, pad)
        hh = unrows(ro.gru_update(rows(q), cq.bias.detach().numpy(), z, rows(hh)), 16)
    assert (hh - want).abs().max() <= 5e-3          # f16 storage of z / r*h / h between the convolutions


def _rows16(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half()


@pytest.mark.gpu
def test_gpu_recurrent_block_kernels_match_the_oracle():
    from counterfactualworldmodels_b200 import _lib
    lib = _lib.load()
    s = torch.cuda.current_stream().cuda_stream
    M, C = 3 * 28 * 28 + 5, 128
    tol = dict(rtol=2e-3, atol=2e-3)                 # one f16 ulp of O(1) values
    # bias + relu into a slot of a wider buffer, second destination, flow tail
    x, bias, tail = _rows16((M, 136), 1, 2.0), torch.randn(128, generator=torch.Generator().manual_seed(2)), _rows16((M, 8), 3)
    d1 = torch.zeros(M, 384, dtype=torch.float16, device=DEV)
    d2 = torch.zeros(M, 384, dtype=torch.float16, device=DEV)
    xd, bd, td = x.to(DEV), bias.to(DEV), tail.to(DEV)
    _lib.check(lib.cwm_raft_bias_act_f16(xd.data_ptr(), 136, bd.data_ptr(), 1, 128, M, d1[:, 256:].data_ptr(), 384,
                                         d2[:, 128:].data_ptr(), 384, td.data_ptr(), 8, 2, s))
    want = ro.bias_act(x[:, :128].float().numpy(), bias.numpy(), True, tail[:, :2].float().numpy())
    np.testing.assert_allclose(d1[:, 256:].float().cpu().numpy(), want, **tol)
    np.testing.assert_allclose(d2[:, 128:256].float().cpu().numpy(), want, **tol)
    assert d1[:, :256].abs().max() == 0 and d2[:, :128].abs().max() == 0 and d2[:, 256:].abs().max() == 0
    # GRU gates and update
    zr, b2, hx = _rows16((M, 2 * C), 4, 2.0), torch.randn(2 * C, generator=torch.Generator().manual_seed(5)), _rows16((M, 384), 6)
    z_out = torch.empty(M, C, dtype=torch.float16, device=DEV)
    rhx = torch.zeros(M, 384, dtype=torch.float16, device=DEV)
    hxd = hx.to(DEV)
    _lib.check(lib.cwm_raft_gru_gate_f16(zr.to(DEV).data_ptr(), b2.to(DEV).data_ptr(), hxd.data_ptr(), 384, C, M,
                                         z_out.data_ptr(), rhx.data_ptr(), 384, s))
    z_w, rh_w = ro.gru_gate(zr.float().numpy(), b2.numpy(), hx[:, :C].float().numpy())
    np.testing.assert_allclose(z_out.float().cpu().numpy(), z_w, **tol)
    np.testing.assert_allclose(rhx[:, :C].float().cpu().numpy(), rh_w, **tol)
    q, bq = _rows16((M, C), 7, 2.0), torch.randn(C, generator=torch.Generator().manual_seed(8))
    dense = torch.empty(M, C, dtype=torch.float16, device=DEV)
    _lib.check(lib.cwm_raft_gru_update_f16(q.to(DEV).data_ptr(), bq.to(DEV).data_ptr(), z_out.data_ptr(), hxd.data_ptr(), 384,
                                           C, M, dense.data_ptr(), s))
    h_w = ro.gru_update(q.float().numpy(), bq.numpy(), z_out.float().cpu().numpy(), hx[:, :C].float().numpy())
    np.testing.assert_allclose(hxd[:, :C].float().cpu().numpy(), h_w, **tol)
    assert torch.equal(dense, hxd[:, :C]) and torch.equal(hxd[:, C:].cpu(), hx[:, C:])
    # flow update
    B, H, W = 3, 28, 28
    delta, bf = _rows16((B * H * W, 8), 9), torch.tensor([0.25, -0.5])
    coords = torch.from_numpy(ro.make_coords(B, H, W, 4, "random"))
    cd = coords.to(DEV).clone()
    flow16 = torch.full((B * H * W, 8), 7.0, dtype=torch.float16, device=DEV)
    _lib.check(lib.cwm_raft_flow_update(delta.to(DEV).data_ptr(), 8, bf.to(DEV).data_ptr(), cd.data_ptr(), B, H, W,
                                        flow16.data_ptr(), s))
    c_w, f_w = ro.flow_update(delta.float().numpy(), bf.numpy(), coords.numpy())
    np.testing.assert_allclose(cd.cpu().numpy(), c_w, rtol=1e-6, atol=1e-5)
    ok = np.abs(f_w) < 100                          # the far-out-of-bounds centres exceed f16 resolution
    np.testing.assert_allclose(flow16[:, :2].float().cpu().numpy()[ok], f_w[ok], rtol=2e-3, atol=2e-3)
    assert flow16[:, 2:].abs().max() == 0
    # f16 pixel-major lookup == the fp32 lookup rounded
    from counterfactualworldmodels_b200 import raft
    f1, f2 = ro.make_fmaps(2, 64, 28, 28, 6)
    block = raft.CorrBlock(torch.from_numpy(f1).to(DEV), torch.from_numpy(f2).to(DEV))
    c2 = torch.from_numpy(ro.make_coords(2, 28, 28, 6, "random")).to(DEV)
    out32 = block(c2)
    out16 = torch.full((2 * 784, 328), 9.0, dtype=torch.float16, device=DEV)
    _lib.check(lib.cwm_raft_corr_lookup_f16(raft._ptr_table(block.corr_pyramid), 4, 4, c2.data_ptr(), 2, 28, 28,
                                            out16.data_ptr(), 328, s))
    want16 = out32.permute(0, 2, 3, 1).reshape(-1, 324)
    # the fast f16 lookup takes the bilinear fractions from floor() directly instead of the reference's normalise /
    # unnormalise round trip: the same value up to the f16 rounding of the output (1 ulp of the volume's scale)
    scale = want16.abs().max().item()
    assert (out16[:, :324].float() - want16).abs().max().item() <= 1.5e-3 * scale and out16[:, 324:].abs().max() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,L", [(2, 28, 28, 4), (1, 17, 19, 4), (3, 16, 24, 3), (5, 9, 30, 2), (65, 28, 28, 4)])
def test_gpu_fast_f16_lookup_matches_the_exact_lookup(B, H, W, L, monkeypatch):
    """`raft_corr_lookup_fast_kernel` (separable blend of a staged 10x10 window; the mixed-precision recurrent block's lookup)
    against the reference-arithmetic fp32 lookup: odd map sizes, fewer levels, far-out-of-bounds / half-integer / integer /
    non-finite centres, pixel counts that do not fill the last CTA."""
    from counterfactualworldmodels_b200 import _lib, raft
    lib = _lib.load()
    f1, f2 = ro.make_fmaps(B, 32, H, W, B + H)
    block = raft.CorrBlock(torch.from_numpy(f1).to(DEV), torch.from_numpy(f2).to(DEV), num_levels=L, radius=4)
    c = torch.from_numpy(ro.make_coords(B, H, W, 3, "random"))
    c[0, 0, 0, 0], c[0, 1, 0, 1], c[0, 0, 1, 0] = float("nan"), float("inf"), -3.0e9
    c = c.to(DEV)
    want = block(c).permute(0, 2, 3, 1).reshape(-1, L * 81)
    ld = (L * 81 + 7) // 8 * 8
    out16 = torch.full((B * H * W, ld), 9.0, dtype=torch.float16, device=DEV)
    _lib.check(lib.cwm_raft_corr_lookup_f16(raft._ptr_table(block.corr_pyramid), L, 4, c.data_ptr(), B, H, W, out16.data_ptr(), ld,
                                            torch.cuda.current_stream().cuda_stream))
    scale = want.abs().max().item()
    err = (out16[:, :L * 81].float() - want).abs().max().item()
    assert err <= 1.5e-3 * scale, (err, scale)
    assert out16[:, L * 81:].abs().max().item() == 0 if ld > L * 81 else True
    assert out16[:3].abs().max().item() == 0 or torch.isfinite(out16).all()          # non-finite centres give zeros
    assert torch.isfinite(out16).all()


@pytest.mark.gpu
def test_gpu_raft_fused_update_block_matches_eager_and_reference():
    """RAFT-large, mixed precision, 4 iterations: the fused recurrent block (cuDNN convolutions + cwm_raft_*_f16 kernels)
    against the same block in eager torch ops (f16-level agreement) and against the fp32 flows of the REAL reference."""
    import make_golden_raft as mg
    d = load("raft_e2e_large_128px")
    model = _mirror(False).to(DEV)
    model.iters = int(d["iters"])
    model.args.mixed_precision = True
    x = mg.e2e_frames(2, 128).to(DEV)
    model.args.fused_update = True
    fused, fused_b = model(x), model(x, backward=True)
    model.args.fused_update = False
    eager = model(x)
    e_ref = rel_err(fused.cpu().numpy(), d["flow_fwd"])
    e_ref_b = rel_err(fused_b.cpu().numpy()[:, :, :, ::2, ::2], d["flow_bwd"])
    e_eager = rel_err(fused.cpu().numpy(), eager.cpu().numpy())
    print(f"raft fused update: vs reference fp32 {e_ref:.2e} / {e_ref_b:.2e}, vs eager f16 {e_eager:.2e} of scale")
    assert e_ref <= 1e-2 and e_ref_b <= 1e-2 and e_eager <= 5e-3
    # flow_init and the shared first frame go through the same path
    init = torch.full((2, 2, 16, 16), 0.5, device=DEV)
    model.args.fused_update = True
    a = model._forward_two_images(x[:, 0] * 255, x[:, 1] * 255, flow_init=init)[1]
    model.args.fused_update = False
    b = model._forward_two_images(x[:, 0] * 255, x[:, 1] * 255, flow_init=init)[1]
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 5e-3


# ------------------------------------------------------------------------------------------------ FlowBackRGB01
class _StubFlow(torch.nn.Module):
    """RAFT-like call signature; flow = (mean colour difference, its negative), doubled for the backward direction."""

    def forward(self, x, iters=None, backward=False):
        d = (x[:, 1:] - x[:, :-1]).mean(2, keepdim=True) * (2.0 if backward else 1.0)
        return torch.cat([d, -d], 2)


def test_frame_pair_flow_pipeline_with_a_stub_flow_network():
    """preprocessor.py:208-285 restated by hand: unnormalise -> [flow, backward flow, normalised rgb of frame 1] ->
    flow / (size / 2)."""
    from counterfactualworldmodels_b200 import preprocessor
    pre = preprocessor.get_preprocessor('flowback_rgb01', temporal_dim=2, iters=5, flow_model=_StubFlow())
    assert pre.num_channels == 7 and pre.get_num_frames() == 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 3, 8, 12, generator=g)                       # [B, C, T = 3, H, W], normalised
    y = pre(x)
    assert y.shape == (2, 7, 1, 8, 12)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)
    raw = x[:, :, :2] * std + mean
    d = (raw[:, :, 1:] - raw[:, :, :1]).mean(1, keepdim=True)
    size = torch.tensor([12.0, 8.0]).view(1, 2, 1, 1, 1) / 2
    torch.testing.assert_close(y[:, 0:2], torch.cat([d, -d], 1) / size)
    torch.testing.assert_close(y[:, 2:4], 2 * torch.cat([d, -d], 1) / size)
    torch.testing.assert_close(y[:, 4:7], (raw[:, :, 1:2] - mean) / std)
    with pytest.raises(NotImplementedError, match="flow_model"):
        preprocessor.get_preprocessor('flow01')(x)
    with pytest.raises(ValueError, match="not a valid raft checkpoint"):
        preprocessor.get_preprocessor('flow01', flow_model_ckpt="/nonexistent/raft-large.pth")
    legacy = preprocessor.get_preprocessor('flow01', flow_model=lambda frames: frames[:, :2, :1] * 0 + 1.0)
    assert legacy(x).shape == (2, 2, 1, 8, 12)                          # a plain callable returns the finished input


@pytest.mark.gpu
def test_gpu_flowback_rgb01_matches_the_reference_preprocessor():
    """`get_preprocessor('flowback_rgb01')` (the flow2imu main-stream input) with raft.RAFT under the
    reference's seeded init against the output of the REAL reference preprocessor + RAFT (3 iterations, fp32)."""
    import make_golden_raft as mg
    from counterfactualworldmodels_b200 import preprocessor
    d = load("raft_flowback_rgb01_128px")
    pre = preprocessor.get_preprocessor('flowback_rgb01', temporal_dim=2, iters=int(d["iters"]),
                                        flow_model=_mirror(False).to(DEV))
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)
    x = ((mg.e2e_frames(2, 128).transpose(1, 2) - mean) / std).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = pre(x).cpu().numpy()[:, :, :, ::2, ::2]
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert y.shape == d["y"].shape
    assert rel_err(y[:, :4], d["y"][:, :4]) <= 1e-4          # forward + backward flow, in units of half the image
    assert rel_err(y[:, 4:], d["y"][:, 4:]) <= 1e-6          # the rgb channels only go through normalise / unnormalise


@pytest.mark.gpu
def test_gpu_flow_generator_video_and_flow_entry_points():
    """`predict_video_and_flow`, `predict_flow_per_sample`, `predict_video_and_flow_per_sample` (segmentation.py:170-245)
    are compositions of `predict` and `predict_flow`: shapes, the sample axis, and consistency with doing it by hand."""
    from counterfactualworldmodels_b200 import segmentation, synthetic, vmae
    cfg = "base_8x8"                                        # 224 px: RAFT's 4-level pyramid needs >= 128 px frames
    model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg))
    synthetic.init_weights_(model, seed=0)
    flow_model = _mirror(True).to(DEV)                      # RAFT-small keeps the test light
    G = segmentation.FlowGenerator(predictor=model.to(DEV).eval(), imagenet_normalize_inputs=True, temporal_dim=2,
                                   flow_model=flow_model, raft_iters=2)
    assert flow_model.iters == 2
    H, W = synthetic.image_hw(cfg)
    x = synthetic.make_video(2, (H, W), seed=1).to(DEV)
    S = 3
    masks = torch.stack([synthetic.make_mask(2, model.mask_size, num_clumps=2, seed=s) for s in range(S)], -1).to(DEV)
    G.set_input(x)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # TF32 convolutions pick batch-size dependent algorithms
    try:
        ys, flows = G.predict_video_and_flow_per_sample(x, masks)
        only = G.predict_flow_per_sample(x, masks)
        y1 = G.predict(x, masks[..., 1], frame=None)
        by_hand = G.predict_flow(y1)
        x_pred, f_pred = G.predict_video_and_flow(x, masks[..., 0])
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert tuple(ys.shape) == (2, 2, 3, H, W, S) and tuple(flows.shape) == (2, 1, 2, H, W, S)
    assert torch.equal(ys[..., 1], y1)               # the predictor is batch-invariant bit for bit
    assert rel_err(only.cpu().numpy(), flows.cpu().numpy()) <= 1e-5
    e = rel_err(flows[..., 1].cpu().numpy(), by_hand.cpu().numpy())
    print(f"per-sample flow vs by hand: {e:.2e} of scale (|flow| max {float(by_hand.abs().max()):.2f})")
    assert e <= 1e-4
    assert tuple(x_pred.shape) == (2, 2, 3, H, W) and tuple(f_pred.shape) == (2, 1, 2, H, W)
    assert torch.equal(x_pred[:, 0], x[:, 0])
    mags = G.compute_flow_samples_magnitude(flows[:, 0])
    assert float(mags.amin()) == 0.0 and float(mags.amax()) <= 1.0 + 1e-6


# ------------------------------------------------------------------------------------------------ IMU-conditioned sweep
@pytest.mark.gpu
@pytest.mark.parametrize("B,D,H,W", [(2, 256, 28, 28), (1, 128, 16, 16), (3, 32, 9, 13), (2, 64, 30, 17)])
def test_gpu_corr_volume_on_tensor_cores_keeps_fp32_accuracy(B, D, H, W, monkeypatch):
    """`cwm_raft_corr_pyramid_tc`: the all-pairs volume as 3 TF32 products per k-step (hi*hi + hi*lo + lo*hi) on
    tcgen05.mma.kind::tf32 against the float64 product of the same fp32 operands, on feature maps whose channels span four
    orders of magnitude.  Bar: 5e-6 of the volume's scale (measured 3e-7 ... 2.5e-6: the tensor core's fp32 accumulation
    over 256 channels; a single TF32 product would be at ~5e-4), the round-1 fp32 SIMT kernel 1e-6 (measured <= 6e-7)."""
    from counterfactualworldmodels_b200 import raft
    g = torch.Generator().manual_seed(B * 100 + D)
    f1 = torch.randn(B, D, H, W, generator=g) * torch.logspace(-2, 2, D).view(1, D, 1, 1)[:, torch.randperm(D, generator=g)]
    f2 = torch.randn(B, D, H, W, generator=g)
    want = torch.einsum("bdi,bdj->bij", f1.double().flatten(2), f2.double().flatten(2)) / (float(D) ** 0.5)
    scale = want.abs().max().item()
    monkeypatch.setenv("CWM_RAFT_CORR", "tc")
    tc = raft.CorrBlock(f1.to(DEV), f2.to(DEV), num_levels=1, radius=1).corr_pyramid[0].view(B, H * W, H * W).cpu()
    monkeypatch.setenv("CWM_RAFT_CORR", "simt")
    simt = raft.CorrBlock(f1.to(DEV), f2.to(DEV), num_levels=1, radius=1).corr_pyramid[0].view(B, H * W, H * W).cpu()
    e_tc, e_simt = (tc.double() - want).abs().max().item() / scale, (simt.double() - want).abs().max().item() / scale
    print(f"corr volume B={B} D={D} {H}x{W}: tensor cores {e_tc:.2e} of scale, fp32 SIMT {e_simt:.2e}")
    assert e_tc <= 5e-6 and e_simt <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("S,H,W,C", [(3, 28, 28, 128), (5, 56, 56, 96), (2, 112, 112, 64), (65, 9, 7, 8)])
def test_gpu_instance_norm_kernel_matches_torch(S, H, W, C):
    """`cwm_instnorm_f16`: nn.InstanceNorm2d (no affine, eps 1e-5) on NHWC f16 maps with the fused relu / shortcut add /
    relu, against torch's fp32 instance norm of the same f16-rounded input."""
    from counterfactualworldmodels_b200 import raft
    g = torch.Generator().manual_seed(S + C)
    x = (torch.randn(S, C, H, W, generator=g) * 2 + torch.randn(1, C, 1, 1, generator=g)).half()
    add = torch.randn(S, C, H, W, generator=g).half()
    enc = raft.FusedFeatureEncoder.__new__(raft.FusedFeatureEncoder)
    enc.eps, enc._ws = 1e-5, None
    rows = lambda t: t.to(DEV).permute(0, 2, 3, 1).reshape(-1, C).contiguous()         # noqa: E731  (pixel-major rows)
    back = lambda r: r.float().cpu().view(S, H, W, C).permute(0, 3, 1, 2)              # noqa: E731
    ref = torch.nn.functional.instance_norm(x.float(), eps=1e-5)
    got = back(enc._inorm(rows(x), S, relu_inner=False))
    assert (got - ref).abs().max().item() <= 4e-3
    got = back(enc._inorm(rows(x), S, relu_inner=True))
    assert (got - ref.relu()).abs().max().item() <= 4e-3
    got = back(enc._inorm(rows(x), S, relu_inner=True, add=rows(add), relu_outer=True))
    assert (got - (add.float() + ref.relu()).relu()).abs().max().item() <= 6e-3
    again = back(enc._inorm(rows(x), S, relu_inner=True, add=rows(add), relu_outer=True))
    assert torch.equal(got, again)                                                       # deterministic reduction order


@pytest.mark.gpu
@pytest.mark.parametrize("which,conv_impl,hw", [("fnet", "tcgen05", 224), ("fnet", "cudnn", 224), ("cnet", "tcgen05", 224),
                                               ("fnet", "tcgen05", 136), ("cnet", "tcgen05", 200)])
def test_gpu_fused_encoders_match_the_autocast_encoders(which, conv_impl, hw):
    """`FusedFeatureEncoder` (f16 pixel-major rows; every convolution an implicit GEMM on the tcgen05 kernel -- 2-D tiles,
    stride 2 through the tensor map, im2col stem --, instance norms in cwm_instnorm_f16 / batch norms folded into the
    weights) against the module it wraps: the fp32 reference, and no further from it than plain autocast is.  The batch
    norms get non-trivial running statistics and affines first (a fresh BatchNorm2d is the identity)."""
    from counterfactualworldmodels_b200 import raft
    model = _mirror(False).to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    for m in model.cnet.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            n = m.num_features
            m.running_mean.copy_(torch.randn(n, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(n, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(n, generator=g) * 0.2)
    enc = getattr(model, which)
    x = (2 * torch.rand(5, 3, hw, hw + 16, generator=g) - 1).to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        want = enc(x).float()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    with torch.autocast("cuda"):
        amp = enc(x).float()
    got = raft.FusedFeatureEncoder(enc, DEV, conv_impl=conv_impl)(x).float()
    assert got.shape == want.shape == (5, 256, hw // 8, (hw + 16) // 8)
    scale = want.abs().max().item()
    e_fused, e_amp = (got - want).abs().max().item() / scale, (amp - want).abs().max().item() / scale
    print(f"{which} ({conv_impl}, {hw} px): fused {e_fused:.2e} of scale, autocast {e_amp:.2e}")
    assert e_fused <= 1e-2 and e_fused <= 3 * e_amp + 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,n1,D,H,W", [(3, 3, 256, 28, 28), (4, 1, 256, 28, 28), (2, 2, 64, 9, 13), (5, 1, 128, 16, 24), (1, 1, 64, 30, 17)])
def test_gpu_corr_pyramid_from_f16_rows(B, n1, D, H, W):
    """`CorrBlock.from_rows` (level 0 as one tcgen05.mma.kind::f16 GEMM per sample on the fused encoder's f16 pixel-major
    rows; one shared first image or one per sample) against the float64 product of the same f16 values: f16 x f16 products
    are exact and the accumulation is fp32, so the bar is the fp32 SIMT kernel's (1e-6 of scale); levels 1.. are the pooled
    level 0."""
    from counterfactualworldmodels_b200 import raft
    g = torch.Generator().manual_seed(B * 10 + D)
    f1 = (torch.randn(n1, H * W, D, generator=g) * torch.logspace(-1, 1, D)[torch.randperm(D, generator=g)]).half()
    f2 = torch.randn(B, H * W, D, generator=g).half()
    want = torch.einsum("bid,bjd->bij", f1.double().expand(B, -1, -1), f2.double()) / (float(D) ** 0.5)
    blk = raft.CorrBlock.from_rows(f1.reshape(-1, D).to(DEV), f2.reshape(-1, D).to(DEV), B, H, W, num_levels=2, radius=4)
    got = blk.corr_pyramid[0].view(B, H * W, H * W).cpu().double()
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 1e-6 * scale
    lvl1 = torch.nn.functional.avg_pool2d(blk.corr_pyramid[0], 2, stride=2)
    assert torch.allclose(blk.corr_pyramid[1], lvl1, rtol=1e-6, atol=1e-6 * scale)


@pytest.mark.gpu
@pytest.mark.parametrize("B,n1,H,W", [(3, 1, 28, 28), (2, 2, 16, 24)])
def test_gpu_f16_pyramid_and_its_lookup(B, n1, H, W):
    """The f16 pyramid of the mixed-precision path (`from_rows(..., f16_pyramid=True)` + `cwm_raft_corr_lookup_f16_pyr16`)
    against the fp32 pyramid from the same rows and its f16 lookup: the extra rounding of the stored volume is 2^-11
    relative, the lookup output is f16 either way."""
    from counterfactualworldmodels_b200 import _lib, raft
    lib = _lib.load()
    D = 256
    g = torch.Generator().manual_seed(B + H)
    f1 = torch.randn(n1 * H * W, D, generator=g).half().to(DEV)
    f2 = torch.randn(B * H * W, D, generator=g).half().to(DEV)
    ref = raft.CorrBlock.from_rows(f1, f2, B, H, W)
    blk = raft.CorrBlock.from_rows(f1, f2, B, H, W, f16_pyramid=True)
    for a, b in zip(ref.corr_pyramid, blk.corr_pyramid):
        assert b.dtype == torch.float16 and a.shape == b.shape
        assert (a - b.float()).abs().max().item() <= 1.5e-3 * a.abs().max().item()
    c = torch.from_numpy(ro.make_coords(B, H, W, 5, "random")).to(DEV)
    s = torch.cuda.current_stream().cuda_stream
    out_ref = torch.empty(B * H * W, 328, dtype=torch.float16, device=DEV)
    out_16 = torch.empty_like(out_ref)
    _lib.check(lib.cwm_raft_corr_lookup_f16(raft._ptr_table(ref.corr_pyramid), 4, 4, c.data_ptr(), B, H, W, out_ref.data_ptr(), 328, s))
    _lib.check(lib.cwm_raft_corr_lookup_f16_pyr16(raft._ptr_table(blk.corr_pyramid), 4, 4, c.data_ptr(), B, H, W, out_16.data_ptr(), 328, s))
    scale = out_ref.float().abs().max().item()
    assert (out_ref.float() - out_16.float()).abs().max().item() <= 3e-3 * scale
    with pytest.raises(RuntimeError):
        blk(c)
