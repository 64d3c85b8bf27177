"""TEST INFRASTRUCTURE ONLY -- pins oracle/raft_oracle.py against the REAL reference (cwm.models.raft.corr.CorrBlock,
RAFT.upsample_flow) and writes tests/golden/raft_*.npz.  Needs /root/reference (build container only).

Three kinds of fixture:
  * raft_corr_*:   seeded synthetic feature maps / lookup centres (regenerated from the seed by the tests) -> the
                   reference's pyramid and lookup outputs;
  * raft_trace_*:  feature maps and per-iteration lookup centres recorded from the reference's own RAFT-large
                   (random init, 3 iterations on a seeded image pair), feature maps rounded to f16 to halve the file,
                   then the reference CorrBlock run on exactly those -> its lookup outputs (a pixel subset);
  * raft_upsample_*: seeded flow / mask logits -> RAFT.upsample_flow output (a row subset).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import raft_oracle as ro  # noqa: E402
import ref_loader  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
PIN_TOL = 2e-6  # of the output scale (max |reference|)

# name: (B, D, H, W, levels, radius, seed, pixel stride (y, x) of the stored lookup subset)
CORR_CASES = {
    "raft_corr_l3_r2_8x12_b2": (2, 32, 8, 12, 3, 2, 1, (1, 1)),
    "raft_corr_l4_r4_16x24_b1": (1, 64, 16, 24, 4, 4, 2, (2, 3)),
    "raft_corr_l4_r3_17x19_b1": (1, 24, 17, 19, 4, 3, 3, (2, 2)),  # odd sizes: pooling drops rows, scalar load path
}
# name: (N, C, H, W, seed, stored row stride)
UPSAMPLE_CASES = {"raft_upsample_n2_5x36": (2, 2, 5, 36, 1, 3), "raft_upsample_n1_c3_4x7": (1, 3, 4, 7, 2, 1)}


def close(a, b, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= PIN_TOL, f"{what}: oracle differs from the reference by {err:.2e} of scale"
    return err


def main():
    ref_loader.install_stubs()
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    import cwm.models.raft.corr as ref_corr
    import cwm.models.raft.raft_model as ref_raft

    torch.set_num_threads(4)
    for name, (B, D, H, W, L, r, seed, (sy, sx)) in CORR_CASES.items():
        f1, f2 = ro.make_fmaps(B, D, H, W, seed)
        block = ref_corr.CorrBlock(torch.from_numpy(f1), torch.from_numpy(f2), num_levels=L, radius=r)
        pyr = ro.corr_pyramid(f1, f2, L)
        out = dict(shape=np.array([B, D, H, W, L, r, seed, sy, sx]))
        worst = 0.0
        for lvl in range(L):
            ref_l = block.corr_pyramid[lvl][:, 0].numpy()
            worst = max(worst, close(pyr[lvl], ref_l, f"{name} level {lvl}"))
            out[f"level{lvl}"] = ref_l[::7] if lvl < 2 and B * H * W > 200 else ref_l
        for kind in ("grid", "random"):
            coords = ro.make_coords(B, H, W, seed, kind)
            ref_o = block(torch.from_numpy(coords)).numpy()
            # the lookup is checked on the reference's own pyramid so only the lookup's rounding enters
            or_o = ro.corr_lookup([p[:, 0].numpy() for p in block.corr_pyramid], coords, r)
            worst = max(worst, close(or_o, ref_o, f"{name} lookup {kind}"))
            out[f"lookup_{kind}"] = ref_o[:, :, ::sy, ::sx]
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: oracle == reference within {worst:.1e} of scale | {os.path.getsize(path) / 1e3:.0f} KB")

    for name, (N, C, H, W, seed, stride) in UPSAMPLE_CASES.items():
        flow, mask = ro.make_upsample_inputs(N, C, H, W, seed)
        ref_o = ref_raft.RAFT.upsample_flow(None, torch.from_numpy(flow), torch.from_numpy(mask)).numpy()
        err = close(ro.upsample_flow(flow, mask), ref_o, name)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, shape=np.array([N, C, H, W, seed, stride]), up=ref_o[:, :, ::stride])
        print(f"{name}: oracle == reference within {err:.1e} of scale | {os.path.getsize(path) / 1e3:.0f} KB")

    # ---- trace of the reference's own RAFT-large (random init) ------------------------------------------
    torch.manual_seed(0)
    args = ref_raft.get_args("")
    args.multiframe, args.scale_inputs, args.output_dim = True, True, None
    model = ref_raft.RAFT(args).eval()
    rec = dict(coords=[])

    class Recorder(ref_corr.CorrBlock):
        def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
            rec["fmaps"] = (fmap1.detach().clone(), fmap2.detach().clone())
            super().__init__(fmap1, fmap2, num_levels=num_levels, radius=radius)

        def __call__(self, coords):
            rec["coords"].append(coords.detach().clone())
            return super().__call__(coords)

    ref_raft.CorrBlock = Recorder
    g = torch.Generator().manual_seed(11)
    img1 = torch.rand(1, 3, 128, 128, generator=g) * 255.0
    img2 = torch.roll(img1, shifts=(3, -5), dims=(2, 3)) + torch.randn(1, 3, 128, 128, generator=g) * 4.0
    with torch.no_grad():
        model._forward_two_images(img1, img2, iters=3)
    ref_raft.CorrBlock = ref_corr.CorrBlock
    f1 = rec["fmaps"][0].half()
    f2 = rec["fmaps"][1].half()
    block = ref_corr.CorrBlock(f1.float(), f2.float(), num_levels=4, radius=4)
    pyr = ro.corr_pyramid(f1.float().numpy(), f2.float().numpy(), 4)
    out = dict(fmap1=f1.numpy(), fmap2=f2.numpy(), stride=np.array([2, 2]))
    worst = 0.0
    for lvl in range(4):
        worst = max(worst, close(pyr[lvl], block.corr_pyramid[lvl][:, 0].numpy(), f"trace level {lvl}"))
    for it, coords in enumerate(rec["coords"]):
        ref_o = block(coords).numpy()
        or_o = ro.corr_lookup([p[:, 0].numpy() for p in block.corr_pyramid], coords.numpy(), 4)
        worst = max(worst, close(or_o, ref_o, f"trace lookup iter {it}"))
        out[f"coords{it}"] = coords.numpy()
        out[f"lookup{it}"] = ref_o[:, :, ::2, ::2]
    path = os.path.join(GOLDEN_DIR, "raft_trace_large_128px.npz")
    np.savez_compressed(path, **out)
    print(f"raft_trace_large_128px: {len(rec['coords'])} iterations, fmaps {tuple(f1.shape)}, oracle == reference within "
          f"{worst:.1e} of scale | {os.path.getsize(path) / 1e3:.0f} KB")

    # ---- end to end: the reference's RAFT (seeded random init, eval) on a seeded frame pair ---------------------
    from counterfactualworldmodels_b200 import raft as mirror
    for small in (False, True):
        name = "raft_e2e_small_128px" if small else "raft_e2e_large_128px"
        torch.manual_seed(0)
        args = ref_raft.get_args("")
        args.multiframe, args.scale_inputs, args.output_dim, args.small = True, True, None, small
        model = ref_raft.RAFT(args).eval().requires_grad_(False)
        torch.manual_seed(0)
        margs = mirror.get_args("")
        margs.multiframe, margs.scale_inputs, margs.output_dim, margs.small = True, True, None, small
        twin = mirror.RAFT(margs)
        sd, sd_twin = model.state_dict(), twin.state_dict()
        assert list(sd.keys()) == list(sd_twin.keys()) and all(torch.equal(sd[k], sd_twin[k]) for k in sd), name
        x = e2e_frames(2, 128)
        model.iters = E2E_ITERS
        with torch.no_grad():
            fwd = model(x)
            bwd = model(x, backward=True)
        assert fwd.shape == (2, 1, 2, 128, 128)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, flow_fwd=fwd.numpy(), flow_bwd=bwd.numpy()[:, :, :, ::2, ::2],
                            init_checksum=np.array(state_checksum(sd)), n_tensors=np.array(len(sd)),
                            iters=np.array(E2E_ITERS))
        print(f"{name}: |flow| max {fwd.abs().max():.3f} mean {fwd.abs().mean():.3f}; mirror init == reference init "
              f"({len(sd)} tensors) | {os.path.getsize(path) / 1e3:.0f} KB")

    # ---- FlowBackRGB01 (preprocessor.py:208-285, :340-345): the flow2imu main-stream input -----------------------
    import tempfile
    import cwm.models.preprocessor as ref_pre
    torch.manual_seed(0)
    args = ref_raft.get_args("")
    args.multiframe, args.scale_inputs, args.output_dim = True, True, None
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = os.path.join(tmp, "raft-large.pth")
        torch.save(ref_raft.RAFT(args).state_dict(), ckpt)      # the reference can only load RAFT from a file
        pre = ref_pre.get_preprocessor('flowback_rgb01', temporal_dim=2, iters=3, flow_model_ckpt=ckpt)
    frames = e2e_frames(2, 128)                                  # [B, T, C, H, W] in [0, 1]
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)
    x = (frames.transpose(1, 2) - mean) / std                    # the predictor's input: [B, C, T, H, W], normalised
    with torch.no_grad():
        y = pre(x)
    assert y.shape == (2, 7, 1, 128, 128) and pre.get_num_frames() == 1 and pre.num_channels == 7
    path = os.path.join(GOLDEN_DIR, "raft_flowback_rgb01_128px.npz")
    np.savez_compressed(path, y=y.numpy()[:, :, :, ::2, ::2], iters=np.array(3))
    print(f"raft_flowback_rgb01_128px: {tuple(y.shape)} |flow ch| max {y[:, :4].abs().max():.4f} | "
          f"{os.path.getsize(path) / 1e3:.0f} KB")


E2E_ITERS = 4


def e2e_frames(B, side):
    """Seeded [B, 2, 3, side, side] frame pair in [0, 1]: smooth random texture, second frame shifted by (3, -5) px."""
    g = torch.Generator().manual_seed(21)
    base = torch.rand(B, 3, side // 4, side // 4, generator=g)
    f0 = torch.nn.functional.interpolate(base, size=(side, side), mode="bilinear", align_corners=False)
    f0 = (f0 + 0.1 * torch.rand(B, 3, side, side, generator=g)).clamp(0, 1)
    f1 = torch.roll(f0, shifts=(3, -5), dims=(2, 3))
    return torch.stack([f0, f1], dim=1)


def state_checksum(sd):
    """Order-sensitive fp64 checksum of a state_dict (a different init order or draw changes it)."""
    total = 0.0
    for i, (k, v) in enumerate(sd.items()):
        total += (i + 1) * float(v.double().abs().sum()) + float(v.double().sum())
    return total


if __name__ == "__main__":
    main()
