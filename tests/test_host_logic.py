"""Host-side logic that needs no GPU: mask rectangularisation, sharding, the gloo multi-process gather, and the
"no CPU fallback" contract of the product path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from counterfactualworldmodels_b200 import dist as cdist
from counterfactualworldmodels_b200 import prediction, synthetic, vmae


def test_rectangularize_min_mode():
    torch.manual_seed(0)
    masks = torch.zeros(3, 20, dtype=torch.bool)
    masks[0, :10] = True
    masks[1, :12] = True
    masks[2, :15] = True
    before = masks.clone()
    out = prediction.RectangularizeMasks('min')(masks)
    assert out.sum(-1).tolist() == [10, 10, 10]
    assert (out & ~before).sum() == 0  # 'min' only clears bits
    assert out.data_ptr() == masks.data_ptr()  # in place, like masking.py:119-128


def test_rectangularize_noop_when_equal():
    masks = synthetic.make_mask(4, (2, 8, 8), num_clumps=2, seed=0)
    state = torch.get_rng_state()
    out = prediction.RectangularizeMasks('min')(masks.clone())
    assert torch.equal(out, masks)
    assert torch.equal(torch.get_rng_state(), state)  # no RNG draw for already-rectangular batches


def test_synthetic_mask_is_temporally_factored():
    m = synthetic.make_mask(5, (2, 28, 28), num_clumps=2, seed=3).view(5, 2, 28, 28)
    assert not m[:, 0].any()
    assert (~m[:, 1]).flatten(1).sum(-1).tolist() == [8] * 5


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 1024, 1025):
        for w in (1, 2, 3, 8):
            spans = [cdist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_forward_refuses_cpu_tensors():
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_4x4"))
    x = torch.zeros(1, 3, 2, 32, 32)
    mask = synthetic.make_mask(1, m.mask_size, 1)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(x, mask)
    with pytest.raises(NotImplementedError):
        m.encoder.blocks[0](torch.zeros(1, 4, 128))


def test_unsupported_constructor_options_raise():
    kw = synthetic.model_kwargs("tiny_4x4")
    with pytest.raises(NotImplementedError):
        vmae.PretrainVisionTransformer(**dict(kw, decoder_depth=0))
    with pytest.raises(NotImplementedError):
        vmae.PretrainVisionTransformer(**dict(kw, embed_per_frame=True))
    with pytest.raises(NotImplementedError):
        vmae.PretrainVisionTransformer(**dict(kw, drop_path_rate=0.1))


def test_supported_constructor_options_add_the_reference_parameters():
    """layer scale and learnable positional embeddings create the same state_dict entries as the reference
    (VideoMAE/utils.py:140-144, vmae.py:68-70)."""
    kw = synthetic.model_kwargs("tiny_4x4")
    m = vmae.PretrainVisionTransformer(**dict(kw, init_values=0.1, use_learnable_pos_emb=True))
    keys = set(m.state_dict().keys())
    assert {"encoder.blocks.0.gamma_1", "encoder.blocks.1.gamma_2", "decoder.blocks.0.gamma_1",
            "encoder.pos_embed"} <= keys
    assert m.state_dict()["encoder.pos_embed"].shape == (1, 128, 128)
    assert float(m.state_dict()["decoder.blocks.0.gamma_2"][0]) == pytest.approx(0.1)
    plain = set(vmae.PretrainVisionTransformer(**kw).state_dict().keys())
    assert not any("gamma" in k or k.endswith("pos_embed") for k in plain)


def test_engine_signature_tracks_weight_changes():
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_4x4"))
    e = vmae._Engine()
    s0 = e._sig(m, "cpu")
    assert e._sig(m, "cpu") == s0
    with torch.no_grad():
        m.mask_token.add_(1.0)
    s1 = e._sig(m, "cpu")
    assert s1 != s0
    m.load_state_dict(m.state_dict())
    assert e._sig(m, "cpu") != s1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, n_samples, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.arange(n_samples * 6, dtype=torch.float32).reshape(n_samples, 2, 3)
        mask = torch.arange(n_samples * 4).reshape(n_samples, 4) % 2 == 0
        xs, ms, (lo, hi) = cdist.shard_samples(x, mask)
        assert xs.shape[0] == hi - lo and torch.equal(xs, x[lo:hi]) and torch.equal(ms, mask[lo:hi])
        y_local = xs * 2 + rank * 0  # stand-in for the per-rank forward
        full = cdist.gather_samples(y_local, n_samples)
        ok_all = torch.equal(full, x * 2)
        root = cdist.gather_samples(y_local, n_samples, dst=0)
        ok_root = (root is None) if rank != 0 else torch.equal(root, x * 2)
        q.put((rank, bool(ok_all and ok_root)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [7, 8, 1])
def test_shard_and_gather_gloo_world2(n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_samples, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}


def test_gui_mask_helpers_match_reference_semantics():
    """`generate_mask_from_patch_idx_list` / `get_mask_image` / `get_masked_pred_patches` (the helpers cwm/interface.py
    calls): pure mask bookkeeping, checked against hand-computed values and, when mounted, the reference itself."""
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_8x8"))
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
    x = synthetic.make_video(2, (64, 64), seed=0)
    G.set_input(x)
    mask = G.generate_mask_from_patch_idx_list([[17, 40], [63, 0]], b=1)   # pixel coordinates, stride 64 // 8 = 8
    img = G.get_mask_image(mask)
    assert img.shape == (2, 2, 8, 8) and not img[:, 0].any()
    vis = (~img[1, 1]).nonzero().tolist()
    # the reference writes through the batch-expanded zeros mask: every row gets the patches, whatever `b` is
    assert vis == [[2, 5], [7, 0]] and (~img[0, 1]).nonzero().tolist() == vis
    assert G.inp_mask_shape == (2, 128)
    out = G.get_masked_pred_patches(torch.ones(2, 2, 3, 64, 64), mask, fill_value=[0.5, 0, 0])
    assert out.shape == (2, 2, 3, 64, 64)
    assert float(out[1, 1, 0, 16:24, 40:48].min()) == 0.5 and float(out[1, 1, 1, 16:24, 40:48].max()) == 0.0
    assert float(out[1, 1, 0, 0, 0]) == 1.0 and float(out[0, 0].max()) == 0.5 and float(out[0, 0, 1].max()) == 0.0
    if os.path.isdir("/root/reference"):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_loader
        ref_vmae, ref_pred = ref_loader.import_reference()
        R = ref_pred.PredictorBasedGenerator(
            predictor=ref_vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_8x8")),
            imagenet_normalize_inputs=True, temporal_dim=2)
        R.set_input(x)
        rmask = R.generate_mask_from_patch_idx_list([[17, 40], [63, 0]], b=1)
        assert torch.equal(rmask, mask)
        want = R.get_masked_pred_patches(torch.ones(2, 2, 3, 64, 64), rmask, fill_value=[0.5, 0, 0])
        assert torch.equal(want, out)


def _ref_generator(cfg="tiny_8x8", **kw):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    ref_vmae, ref_pred = ref_loader.import_reference()
    return ref_pred.PredictorBasedGenerator(predictor=ref_vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg)),
                                            imagenet_normalize_inputs=True, temporal_dim=2, **kw)


def test_wrapper_host_helpers_hand_checked():
    """The forward-free helpers of PredictorBasedGenerator (prediction.py:226-758), on values computed by hand."""
    from counterfactualworldmodels_b200 import masking
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_8x8"))
    gen = masking.RotatedTableUniformMaskingGenerator(input_size=m.mask_size, mask_ratio=0.9, seed=0)
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2, seed=3,
                                           mask_generator=gen)
    x = synthetic.make_video(2, (64, 64), seed=0)
    G.set_input(x)
    assert G.get_fully_visible_mask().shape == tuple(G.mask_shape) and not G.get_fully_visible_mask().any()
    a = torch.ones(1, 2, 8, 8, dtype=torch.bool)
    a[0, 1, 2:4, 2:4] = False                       # 4 visible patches
    b = torch.ones(1, 2, 8, 8, dtype=torch.bool)
    b[0, 1, 2, 2] = False                           # one of them also visible in b
    comp = G.mask_complement(b.view(1, -1), a.view(1, -1)).view(1, 2, 8, 8)
    assert (~comp[0, 1]).nonzero().tolist() == [[2, 3], [3, 2], [3, 3]] and bool(comp[0, 0].all())
    inv = G._invert_mask(a.view(1, -1)).view(1, 2, 8, 8)
    assert torch.equal(inv[:, 0], a[:, 0]) and torch.equal(inv[:, 1], ~a[:, 1])
    u = G.unmask_one_patch(torch.ones(2, 128, dtype=torch.bool), idx=[5, 6], mask_shape=(2, 8, 8), frame=1)
    assert (~u.view(2, 2, 8, 8)).nonzero().tolist() == [[0, 1, 5, 6], [1, 1, 5, 6]]
    u = G.unmask_one_patch(torch.ones(2, 128, dtype=torch.bool), idx=[1, 0, 7, 7], mask_shape=(2, 8, 8))
    assert (~u.view(2, 2, 8, 8)).nonzero().tolist() == [[1, 0, 7, 7]]
    assert (~G.unmask_one_patch(torch.ones(1, 128, dtype=torch.bool), idx=9)).nonzero().tolist() == [[0, 9]]
    assert G.patch_idx_list_from_mask(a) == [[0, 1, 2, 2], [0, 1, 2, 3], [0, 1, 3, 2], [0, 1, 3, 3]]
    pairs = G.get_frame_pairs(torch.arange(3.0).view(1, 3, 1, 1, 1).expand(1, 3, 3, 4, 4))
    assert len(pairs) == 2 and G.target_frame == 1 and [float(p[0, 0, 0, 0, 0]) for p in pairs] == [0.0, 2.0]
    near = G.get_nearby_patches(a.view(1, -1), radius=1)
    want = torch.zeros(8, 8, dtype=torch.bool)
    want[1:5, 1:5] = True
    want[2:4, 2:4] = False                          # the visible patches themselves carry the maximum distance
    assert torch.equal(near[0, 1], want) and bool(near[0, 0].all())   # a frame without visible patches: distance 0
    G.set_input(x[:1])                              # the cutout writes through the batch-expanded zeros mask: B = 1 only
    cut = G.generate_cutout_mask([[17, 17]], radius=1, frame=1).view(1, 2, 8, 8)    # pixel (17, 17) -> patch (2, 2)
    with pytest.raises(RuntimeError):               # the default frame=-1 slices `mask[:, -1:0]` = nothing, as in the
        G.generate_cutout_mask([[17, 17]], radius=1)   # reference (prediction.py:656)
    G.set_input(x)
    # a "cutout": the 3 x 3 neighbourhood of the patch is MASKED, everything else in that frame visible (:657-658)
    assert int(cut[0, 1].sum()) == 9 and bool(cut[0, 1, 1:4, 1:4].all()) and not bool(cut[0, 1, 0, 0])
    e = G._get_error(torch.ones(1, 2, 3, 4, 4), torch.zeros(1, 1, 3, 4, 4))
    assert tuple(e.shape) == (1, 1, 1, 4, 4) and float(e.max()) == 3.0
    dens = torch.rand(1, 1, 64, 64, generator=torch.Generator().manual_seed(1))
    pooled = G.patchify_energy_density(dens, mode='max')
    assert tuple(pooled.shape) == (1, 1, 8, 8) and float(pooled[0, 0, 0, 0]) == pytest.approx(float(masking.boltzmann(dens, beta=None)[0, 0, :8, :8].max()))
    masks = G.sample_random_masks(num_samples=3, num_visible=2)
    assert tuple(masks.shape) == (2, 128, 3) and (~masks).sum(1).tolist() == [[64 + 2] * 3] * 2
    assert G._sample_random_patches(batch_size=2) == [[0, 1, 2, 0], [1, 1, 1, 3]] or len(G._sample_random_patches(2)) == 2
    with pytest.raises(NotImplementedError):
        G.get_initial_mask(x)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference not mounted")
def test_wrapper_host_helpers_match_the_live_reference():
    from counterfactualworldmodels_b200 import masking
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    ref_loader.import_reference()
    import cwm.models.masking as ref_masking
    R = _ref_generator(seed=3, mask_generator=ref_masking.RotatedTableUniformMaskingGenerator(
        input_size=(2, 8, 8), mask_ratio=0.9, seed=0))
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_8x8"))
    G = prediction.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2, seed=3,
                                           mask_generator=masking.RotatedTableUniformMaskingGenerator(
                                               input_size=m.mask_size, mask_ratio=0.9, seed=0))
    x = synthetic.make_video(2, (64, 64), seed=0)
    R.set_input(x)
    G.set_input(x)
    g = torch.Generator().manual_seed(5)
    m1 = torch.rand(2, 128, generator=g) < 0.8
    m2 = torch.rand(2, 128, generator=g) < 0.5
    for frame in (-1, 0, None):
        assert torch.equal(G.mask_complement(m1, m2, frame=frame), R.mask_complement(m1, m2, frame=frame))
    assert torch.equal(G._invert_mask(m1, frame=0), R._invert_mask(m1, frame=0))
    assert torch.equal(G.get_fully_visible_mask(), R.get_fully_visible_mask())
    for radius in (1, 2, 0):
        assert torch.equal(G.get_nearby_patches(m1, radius=radius), R.get_nearby_patches(m1, radius=radius)), radius
    assert torch.equal(masking.patch_distance_transform(m1.view(2, 2, 8, 8)), ref_masking.patch_distance_transform(m1.view(2, 2, 8, 8)))
    assert torch.equal(G.unmask_one_patch(m1 | True, idx=[3, 4], mask_shape=(2, 8, 8), frame=1),
                       R.unmask_one_patch(m1 | True, idx=[3, 4], mask_shape=(2, 8, 8), frame=1))
    assert torch.equal(G.unmask_one_patch(m1 | True, idx=[1, 1, 3, 4], mask_shape=(2, 8, 8)),
                       R.unmask_one_patch(m1 | True, idx=[1, 1, 3, 4], mask_shape=(2, 8, 8)))
    img = m1.view(2, 2, 8, 8)
    assert [[int(v) for v in p] for p in G.patch_idx_list_from_mask(img)] == \
        [[int(v) for v in p] for p in R.patch_idx_list_from_mask(img)]
    assert G._sample_random_patches(batch_size=3) == R._sample_random_patches(batch_size=3)
    d5 = torch.rand(2, 2, 1, 64, 64, generator=g)     # rank 5 only: the reference hands its 3-tuple patch size to the
    for mode, beta in (("mean", 2.0), ("max", None), ("min", None)):   # 2-d pools for rank-4 input and raises (:297-299)
        assert torch.equal(G.patchify_energy_density(d5, mode=mode, beta=beta),
                           R.patchify_energy_density(d5, mode=mode, beta=beta)), mode
    pa, pb = G.get_frame_pairs(x.repeat(1, 2, 1, 1, 1)[:, :3], frame=0), R.get_frame_pairs(x.repeat(1, 2, 1, 1, 1)[:, :3], frame=0)
    assert len(pa) == len(pb) and all(torch.equal(u, v) for u, v in zip(pa, pb))
    pred, gt = torch.rand(2, 2, 3, 8, 8, generator=g), torch.rand(2, 1, 3, 8, 8, generator=g)
    assert torch.equal(G._get_error(pred, gt), R._get_error(pred, gt))
    torch.manual_seed(11)                               # the generators draw `randperm` from the GLOBAL torch generator
    ours = G.sample_random_masks(num_samples=4, num_visible=3)
    torch.manual_seed(11)
    assert torch.equal(ours, R.sample_random_masks(num_samples=4, num_visible=3))
    cut_r = ref_masking  # generate_cutout_mask is a @staticmethod taking `self` in the reference
    R.set_input(x[:1])
    G.set_input(x[:1])
    assert torch.equal(G.generate_cutout_mask([[17, 40]], radius=2, frame=1),
                       type(R).generate_cutout_mask(R, [[17, 40]], radius=2, frame=1))
