"""GPU parity for SURVEY.md section 8(f) rank 1: the batched motion-counterfactual kernels (csrc/counterfactual.cu)
through the mirror classes (`segmentation.FlowGenerator`, `perturbation.*`) against the fixtures the REAL reference
produced and against the numpy oracle.  Bar: masks AND pixels bit-exact."""
import numpy as np
import pytest
import torch

import counterfactual_oracle as cfo
import vmae_oracle
from counterfactualworldmodels_b200 import perturbation, prediction, segmentation, synthetic, vmae
from test_counterfactual_cpu import CF_CASES, load_cf

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tiny_predictor(cfg, depth=1):
    kw = synthetic.model_kwargs(cfg)
    kw.update(encoder_depth=depth, decoder_depth=depth, encoder_embed_dim=128, decoder_embed_dim=128,
              encoder_num_heads=2, decoder_num_heads=2)
    m = vmae.PretrainVisionTransformer(**kw)
    synthetic.init_weights_(m, seed=0, style="perturbed")
    return m.to(DEV).eval()


def _bits(t):
    return t.detach().cpu().contiguous().numpy().view(np.uint32)


@pytest.mark.parametrize("case", CF_CASES)
def test_create_motion_counterfactuals_matches_reference_fixture(case):
    d = load_cf(case)
    G = segmentation.FlowGenerator(predictor=_tiny_predictor(d["cfg"]), imagenet_normalize_inputs=True, temporal_dim=2)
    x = d["x"].to(DEV)
    passive, active = torch.from_numpy(d["passive"]).to(DEV), torch.from_numpy(d["active"]).to(DEV)
    S = passive.shape[-1]
    G.set_input(x)
    G.mask_rectangularizer.set_mode('none')
    xs, ms = G.create_motion_counterfactuals(x, masks=passive, active_patches=active, shifts=d["shifts"].tolist(),
                                             num_samples=S, fix_passive=bool(d["static"][0]))
    assert xs.shape == (S,) + tuple(x.shape[1:]) and ms.dtype == torch.bool
    assert np.array_equal(ms.cpu().numpy(), d["mask_shift"])
    assert [list(map(int, s)) for s in G.shifts] == d["shifts"].tolist()
    xs_or, ms_or = cfo.create_motion_counterfactuals(d["x"].numpy(), d["passive"], d["active"], d["shifts"].tolist(),
                                                     d["patch_size"], frame=1, fix_passive=bool(d["static"][0]))
    assert np.array_equal(_bits(xs), xs_or.view(np.uint32)), "pixels differ from the oracle (bitwise)"
    assert np.allclose(cfo.fingerprint(xs.cpu().numpy()), d["x_shift_fingerprint"], rtol=0, atol=1e-6)
    if "x_shift" in d:
        assert np.array_equal(_bits(xs), d["x_shift"].view(np.uint32))
    # default mode: same rectangularisation as the reference under the same global seed (masking.py:100-132)
    G.mask_rectangularizer.set_mode('min')
    torch.manual_seed(1234)
    _, ms_rect = G.create_motion_counterfactuals(x, masks=passive, active_patches=active, shifts=d["shifts"].tolist(),
                                                 num_samples=S, fix_passive=bool(d["static"][0]), reset_shifts=True)
    assert np.array_equal(ms_rect.cpu().numpy(), d["mask_shift_rect"])


@pytest.mark.parametrize("case", CF_CASES)
def test_make_static_and_shift_match_reference_fixture(case):
    d = load_cf(case)
    cfg, seed, clump = d["cfg"], d["seed"], int(d["clump"][0])
    m = _tiny_predictor(cfg)
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    x2 = synthetic.make_video(2, synthetic.image_hw(cfg), seed=seed + 100, counterfactual_like=True)
    m2 = synthetic.make_mask(2, m.mask_size, num_clumps=3, clump=clump, seed=seed)
    xst, mst = G.make_static(x2.to(DEV), m2.to(DEV))
    xst_or, _ = cfo.make_static(x2.numpy(), m2.numpy(), d["patch_size"])
    assert np.array_equal(_bits(xst), xst_or.view(np.uint32)) and torch.equal(mst.cpu(), m2)
    assert np.allclose(cfo.fingerprint(xst.cpu().numpy()), d["make_static_fingerprint"], rtol=0, atol=1e-6)
    act2 = torch.from_numpy(d["active"][:, :, :2].transpose(0, 2, 1).reshape(2, -1).copy())
    G.mask_rectangularizer.set_mode('none')
    xsh, msh = G._shift(x2.to(DEV), m2.to(DEV), active_patches=act2.to(DEV), shift=d["shifts"][0].tolist(), frame=1)
    xsh_or, msh_or = cfo.shift_one(x2.numpy(), m2.numpy(), act2.numpy(), d["patch_size"],
                                   shift=d["shifts"][0].tolist(), frame=1)
    assert np.array_equal(msh.cpu().numpy(), msh_or) and np.array_equal(_bits(xsh), xsh_or.view(np.uint32))
    want_mask = np.unpackbits(d["shift_one_mask"])[:msh_or.size].reshape(msh_or.shape).astype(bool)
    assert np.array_equal(msh.cpu().numpy(), want_mask)
    assert np.allclose(cfo.fingerprint(xsh.cpu().numpy()), d["shift_one_fingerprint"], rtol=0, atol=1e-6)
    assert list(G.shift) == d["shifts"][0].tolist()


def test_random_masks_shifts_frames_and_unaligned_strides():
    """Arbitrary masks / perturbation points, both frames, shifts beyond the image, and a source view whose rows
    are not 16-byte aligned (scalar load path)."""
    rng = np.random.RandomState(7)
    base = torch.rand(3, 2, 3, 24, 44, generator=torch.Generator().manual_seed(7))
    for trial in range(8):
        x = base[..., 1:41] if trial % 2 else base[..., :40]   # odd column offset -> unaligned rows
        sh = perturbation.ShiftPatchesAndMask(patch_size=(1, 4, 4))
        mask = torch.from_numpy(rng.rand(3, 2 * 6 * 10) < 0.6)
        points = torch.from_numpy(rng.rand(3, 2 * 6 * 10) < 0.15) if trial % 4 != 3 else None
        ms = [int(rng.randint(-7, 8)), int(rng.randint(-11, 12))]
        xg, mg = sh(x.to(DEV), mask=mask.to(DEV), perturbation_points=None if points is None else points.to(DEV),
                    mask_shift=ms, frame=trial % 2)
        xo, mo = cfo.perturbation_forward(x.contiguous().numpy(), mask.numpy(),
                                          None if points is None else points.numpy(), (1, 4, 4), mask_shift=ms,
                                          frame=trial % 2)
        assert np.array_equal(mg.cpu().numpy(), mo), trial
        assert np.array_equal(_bits(xg), xo.view(np.uint32)), trial
        assert tuple(sh.shift) == (ms[0] * 4, ms[1] * 4)


def test_negative_zero_and_blend_are_literal():
    """x_shift*(1-m) + x*m evaluated literally: -0.0 in the shifted source becomes +0.0 exactly like the reference."""
    x = torch.zeros(1, 2, 3, 8, 8)
    x[0, :, :, 0:4, 0:4] = -0.0
    x[0, :, 0, 4:8, 4:8] = -1.5
    mask = torch.ones(1, 8, dtype=torch.bool)
    mask[0, :4] = False
    points = torch.zeros(1, 8, dtype=torch.bool)
    points[0, 4] = True
    sh = perturbation.ShiftPatchesAndMask(patch_size=(1, 4, 4))
    xg, mg = sh(x.to(DEV), mask=mask.to(DEV), perturbation_points=points.to(DEV), mask_shift=[1, 1], frame=1)
    xo, mo = cfo.perturbation_forward(x.numpy(), mask.numpy(), points.numpy(), (1, 4, 4), mask_shift=[1, 1], frame=1)
    assert np.array_equal(_bits(xg), xo.view(np.uint32)) and np.array_equal(mg.cpu().numpy(), mo)


def test_fused_prediction_equals_materialised_prediction():
    """The fused path (virtual video read by the patch gather and the unpatchify) is bit-identical to building
    x_shift first and predicting from the tensor, and within tolerance of the CPU oracle end to end."""
    d = load_cf("cf_small_4x4_s8_moving_input")
    m = _tiny_predictor(d["cfg"], depth=2)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    G = segmentation.FlowGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    x = d["x"].to(DEV)
    passive, active = torch.from_numpy(d["passive"]).to(DEV), torch.from_numpy(d["active"]).to(DEV)
    shifts = d["shifts"].tolist()
    torch.manual_seed(3)
    y_fused = G.predict_counterfactual_videos(x, active, passive_patches=passive, shifts=shifts, sample_batch_size=3,
                                              fix_passive=False)
    G.set_input(x)
    torch.manual_seed(3)
    x_shift, mask_shift = G.create_motion_counterfactuals(x, masks=passive, active_patches=active, shifts=shifts,
                                                          fix_passive=False, reset_shifts=True)
    y_mat = G.batch_predict_per_sample(x_shift, masks=mask_shift, frame=None, batch_size=3, sample_dim=0)
    assert y_fused.shape == x_shift.shape
    assert torch.equal(y_fused, y_mat)
    # end to end against the oracle (construction + VMAE forward + scatter), tolerance of the hot path
    ocfg = dict(patch_size=d["patch_size"], enc_heads=2, dec_heads=2, eps=1e-6)
    want = vmae_oracle.predict(sd, x_shift.cpu(), mask_shift.cpu(), ocfg, frame=None)
    err = (y_fused.cpu() - want).abs()
    assert err.max().item() <= 2e-2 and err.mean().item() <= 2e-3
    # visible patches of the output are the counterfactual prompt's patches, bit for bit
    ps = d["patch_size"]
    up, xp = vmae_oracle.patchify(y_fused.cpu(), ps), vmae_oracle.patchify(x_shift.cpu(), ps)
    assert torch.equal(up[~mask_shift.cpu()], xp[~mask_shift.cpu()])


def test_get_counterfactual_prediction_runs_fused():
    m = _tiny_predictor("tiny_8x8")
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    x = synthetic.make_video(1, (64, 64), seed=3).to(DEV)
    active = G.get_zeros_mask(x).clone()
    n = 64
    active[0, :n] = True
    active[0, n + 27] = False
    mask = G.get_zeros_mask(x).clone()
    mask[0, n + 3] = False
    y = G.get_counterfactual_prediction(x, mask=mask, active_patches=active, shift=[1, -1], fix_passive=True)
    assert y.shape == x.shape and bool(torch.isfinite(y).all())
    # same thing step by step through materialised tensors
    xs, _ = G.make_static(x, mask)
    xp, mp = G._shift(xs, mask=mask, active_patches=active, shift=[1, -1], frame=1)
    assert torch.equal(y, G.predict(xp, mp, frame=None))
    assert [list(s) for s in G.shifts] == [[1, -1], [1, -1]]


def test_full_size_sweep_properties():
    """BASELINE size (224 px, 8x8 patches), S = 256 samples: size-independent properties instead of the oracle.
    (i) a zero shift reproduces the static movie and the un-shifted mask; (ii) shifting by s then reading the
    moved patch gives the source patch; (iii) the number of changed patches is bounded by the active patches."""
    S, h = 256, 28
    x = synthetic.make_video(1, (224, 224), seed=9).to(DEV)
    rng = np.random.RandomState(9)
    active = torch.ones(S, 2, h, h, dtype=torch.bool)
    passive = torch.zeros(S, 2, h, h, dtype=torch.bool)
    passive[:, 1] = True
    pos, shifts = [], []
    for s in range(S):
        ay, ax = rng.randint(4, h - 6), rng.randint(4, h - 6)
        active[s, 1, ay:ay + 2, ax:ax + 2] = False
        py, px = rng.randint(0, h - 1), rng.randint(0, h - 1)
        passive[s, 1, py, px] = False
        pos.append((ay, ax))
        shifts.append([0, 0] if s % 8 == 0 else [int(rng.randint(-3, 4)), int(rng.randint(-3, 4))])
    video, mask = perturbation.shift_patches_and_masks(x, passive.reshape(S, -1).to(DEV), active.reshape(S, -1).to(DEV),
                                                       shifts, (1, 8, 8), frame=1, static_frame=0)
    xs = video.materialize()
    static = x[:, 0:1].expand(S, 2, -1, -1, -1)
    m4 = mask.reshape(S, 2, h, h).cpu()
    assert not bool(m4[:, 0].any())
    for s in range(S):
        ay, ax = pos[s]
        dy, dx = shifts[s]
        if shifts[s] == [0, 0]:
            assert torch.equal(xs[s], static[s])
            assert torch.equal(m4[s, 1], passive[s, 1] & active[s, 1])
        if s < 32:
            src = x[0, 0, :, ay * 8:(ay + 2) * 8, ax * 8:(ax + 2) * 8]
            got = xs[s, 1, :, (ay + dy) * 8:(ay + dy + 2) * 8, (ax + dx) * 8:(ax + dx + 2) * 8]
            assert torch.equal(got, src)
            assert not bool(m4[s, 1, ay + dy:ay + dy + 2, ax + dx:ax + dx + 2].any())
    changed = (xs != static).reshape(S, 2, 3, h, 8, h, 8).any(-1).any(-2).any(2)  # [S, 2, h, h] patches that differ
    assert int(changed[:, 0].sum()) == 0 and int(changed[:, 1].sum(dim=(-1, -2)).max()) <= 4


def test_error_behaviour():
    m = _tiny_predictor("tiny_4x4")
    G = segmentation.FlowGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    x = synthetic.make_video(2, (32, 32), seed=0).to(DEV)
    G.set_input(x)
    masks = G.get_zeros_mask(x).unsqueeze(-1).expand(-1, -1, 3)
    # the reference indexes shifts[i] for i < B*S with S shifts: IndexError for B > 1 (segmentation.py:329)
    with pytest.raises(IndexError):
        G.create_motion_counterfactuals(x, masks=masks, shifts=[[1, 0], [0, 1], [1, 1]], num_samples=3)
    with pytest.raises(AssertionError):  # 2-D masks need num_samples (segmentation.py:293)
        G.create_motion_counterfactuals(x, masks=masks[..., 0])
    with pytest.raises(RuntimeError, match="flow_model"):
        G.predict_flow(x)
    sh = perturbation.ShiftPatchesAndMask(patch_size=(1, 4, 4))
    with pytest.raises(AssertionError):  # pixel shifts must be multiples of the patch size (perturbation.py:251-252)
        sh(x, mask=masks[..., 0].clone(), shift=(3, 0))


def test_sharded_sweep_single_process_equals_generator():
    """`dist.sharded_counterfactual_videos` without a process group is the plain fused sweep (the multi-rank path is
    checked by tools/dist_check.py under torchrun: bit-identical on 4 and 8 GPUs, profiles/r1_dist_check_*.log)."""
    from counterfactualworldmodels_b200 import dist as cdist
    d = load_cf("cf_tiny_8x8_s8_clump2")
    m = _tiny_predictor(d["cfg"])
    G = segmentation.FlowGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    x = d["x"].to(DEV)
    passive, active = torch.from_numpy(d["passive"]).to(DEV), torch.from_numpy(d["active"]).to(DEV)
    shifts = d["shifts"].tolist()
    torch.manual_seed(11)
    a = cdist.sharded_counterfactual_videos(G, x, active, passive, shifts=shifts, sample_batch_size=3)
    torch.manual_seed(11)
    b = G.predict_counterfactual_videos(x, active, passive_patches=passive, shifts=shifts, sample_batch_size=3)
    assert torch.equal(a, b)
    local = cdist.sharded_counterfactual_videos(G, x[0, 0], active, passive, shifts=shifts, gather=False)
    assert local.shape == a.shape
