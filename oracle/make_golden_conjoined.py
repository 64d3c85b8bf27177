"""TEST INFRASTRUCTURE ONLY -- pins ``oracle/conjoined_oracle.py`` against the REAL reference
(``cwm.models.VideoMAE.conjoined_vmae``) and writes ``tests/golden/conj_*.npz``.

Run in the build container (needs /root/reference):  python oracle/make_golden_conjoined.py [case ...]

Every case builds the reference model from ``synthetic.build_conjoined`` / the reference's own factory, overwrites its
weights with ``synthetic.init_weights_(seed)``, runs the reference forward on the CPU in fp32 and asserts that the
oracle reproduces it; the fixture stores the masks, the reference outputs and fingerprints of inputs and weights so
that the GPU box can re-derive everything else from seeds.
"""
import os
import sys
import time
from functools import partial

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
import conjoined_oracle as co  # noqa: E402
from counterfactualworldmodels_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name: (model, batch, init style, weight seed, data seed, main-mask clumps per row, context mask rows or spec)
CASES = {
    # ragged visible counts (4 / 8 / 8 frame-1 patches): null-token padding in the main stream, IMU fully visible
    "conj_padded_small_ragged": ("conj_padded_small", 3, "perturbed", 11, 11, [1, 2, 2], "visible"),
    # ragged IMU masks (2 vs 1 masked tokens): null-token padding in the context stream
    "conj_padded_small_ctxmasked": ("conj_padded_small", 2, "perturbed", 12, 12, [2, 2],
                                    [[0, 0, 1, 0, 1], [0, 0, 0, 0, 1]]),
    # equal rows, IMU visible: also run through PredictorBasedGenerator.predict (video stored)
    "conj_padded_small_predict": ("conj_padded_small", 2, "perturbed", 16, 16, [2, 2], "visible"),
    # flow2imu topology: 7-channel one-frame main stream all visible, IMU fully masked + dummy token, context output
    "conj_flow2imu_small": ("conj_flow2imu_small", 2, "perturbed", 13, 13, None, "masked"),
    # BASELINE config 5 at full size, one sample, through PredictorBasedGenerator.predict
    "conj_imu400_base_4x4_b1": ("imu400_base_4x4", 1, "reference", 0, 14, [8], "visible"),
}
PADDED_CASES = {
    # standalone PaddedVisionTransformer (a13), ragged rows
    "padded_small_ragged": (3, "perturbed", 15, 15, [1, 3, 2]),
}
PADDED_KW = dict(min_padding_tokens=0, max_padding_tokens=8, img_size=32, patch_size=(4, 4), encoder_embed_dim=256,
                 encoder_depth=2, encoder_num_heads=4, encoder_num_classes=0, decoder_embed_dim=128,
                 decoder_num_heads=2, decoder_depth=2, mlp_ratio=4, qkv_bias=True, num_frames=2, tubelet_size=1)


def ragged_mask(msize, clumps, seed):
    """Row b: frame 0 visible, frame 1 masked except clumps[b] visible 2x2 blocks."""
    rows = [synthetic.make_mask(1, msize, num_clumps=c, seed=seed * 10 + b) for b, c in enumerate(clumps)]
    return torch.cat(rows, 0)


def case_inputs(case):
    """(model name, B, style, wseed, x, mask, imu, ctx mask) -- re-derivable on the GPU box from seeds alone."""
    name, B, style, wseed, dseed, clumps, cspec = CASES[case]
    if name == "imu400_base_4x4":
        img, patch, chans, seq, n_ctx = 224, 4, 3, 400, 25
    else:
        c = synthetic.CONJOINED[name]
        img, patch, chans, seq, n_ctx = c["img_size"], c["patch_size"][0], c["main_chans"], c["seq_len"], c["seq_len"] // 16
    x = synthetic.make_video(B, (img, img), seed=dseed, C=chans, counterfactual_like=(chans == 3))   # [B,T,C,H,W]
    msize = (2, img // patch, img // patch)
    if clumps is None:
        mask = torch.zeros(B, msize[0] * msize[1] * msize[2], dtype=torch.bool)
    else:
        mask = ragged_mask(msize, clumps, dseed)
    imu = synthetic.make_imu(B, seq, seed=dseed)
    if cspec == "visible":
        mc = torch.zeros(B, n_ctx, dtype=torch.bool)
    elif cspec == "masked":
        mc = torch.ones(B, n_ctx, dtype=torch.bool)
        imu = torch.zeros_like(imu)   # get_fake_head_motion (segmentation.py:814-832)
    else:
        mc = torch.tensor(cspec, dtype=torch.bool)
    return name, B, style, wseed, x, mask, imu, mc


def padded_case_inputs(case):
    B, style, wseed, dseed, clumps = PADDED_CASES[case]
    x = synthetic.make_video(B, (32, 32), seed=dseed)
    return B, style, wseed, x, ragged_mask((2, 8, 8), clumps, dseed)


def pack(mask):
    return np.packbits(mask.numpy().astype(np.uint8), axis=1), np.array(mask.shape)


def main(argv):
    ref_loader.install_stubs()
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    import cwm.models.VideoMAE.conjoined_vmae as rconj
    import cwm.models.prediction as rpred
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = argv or (list(CASES) + list(PADDED_CASES))
    for case in names:
        t0 = time.time()
        if case in PADDED_CASES:
            B, style, wseed, x, mask = padded_case_inputs(case)
            ref = rconj.PaddedVisionTransformer(norm_layer=partial(nn.LayerNorm, eps=1e-6), **PADDED_KW)
            ref = ref.eval().requires_grad_(False)
            synthetic.init_weights_(ref, seed=wseed, style=style)
            xin = x.transpose(1, 2)
            with torch.no_grad():
                y_ref = ref(xin, mask.clone())
            scfg = dict(enc_heads=4, dec_heads=2, max_pad=8, min_pad=0, pos="sinusoid", eps=1e-6)
            y_or = co.padded_forward(ref.state_dict(), xin, mask, scfg)
            err = (y_or - y_ref).abs().max().item()
            assert y_or.shape == y_ref.shape and err < 2e-5, (case, err)
            pm, fm, nm = co.padding_masks(mask, 8, 0)
            assert torch.equal(pm, ref.padding_mask) and torch.equal(fm, ref.full_input_mask) and \
                torch.equal(nm, ref.null_mask)
            m, ms = pack(mask)
            out = dict(mask=m, mask_shape=ms, y=y_ref.numpy(), oracle_vs_reference_maxabs=np.array([err]),
                       weights_checksum=np.array([synthetic.weights_checksum(ref)]),
                       x_fingerprint=np.array([float(x.double().sum()), float(x.double().pow(2).sum())]),
                       null_rows=np.array([int(ref.null_mask.sum())]))
            np.savez_compressed(os.path.join(GOLDEN_DIR, case + ".npz"), **out)
            print(f"{case}: y {tuple(y_ref.shape)} null rows {out['null_rows'][0]} oracle-vs-ref {err:.2e} "
                  f"({time.time() - t0:.1f}s)")
            continue

        name, B, style, wseed, x, mask, imu, mc = case_inputs(case)
        torch.manual_seed(0)
        ref = synthetic.build_conjoined(rconj, name).eval().requires_grad_(False)
        synthetic.init_weights_(ref, seed=wseed, style=style)
        xin = x.transpose(1, 2)                                  # [B,C,T,H,W], the predictor's input convention
        padded = hasattr(ref.main_stream, "max_padding_tokens")
        with torch.no_grad():
            y_ref, yc_ref = ref(xin, mask.clone(), x_context=imu, mask_context=mc.clone(), output_main=True,
                                output_context=True)
        extra = {}
        if padded:
            extra["null_rows_main"] = np.array([int(ref.main_stream.null_mask.sum())])
            extra["null_rows_ctx"] = np.array([int(ref.context_stream.null_mask.sum())])
            pm, fm, nm = co.padding_masks(mask, ref.main_stream.max_padding_tokens, 0)
            assert torch.equal(fm, ref.main_stream.full_input_mask) and torch.equal(nm, ref.main_stream.null_mask)
            pm, fm, nm = co.padding_masks(mc, ref.context_stream.max_padding_tokens, 0)
            assert torch.equal(fm, ref.context_stream.full_input_mask) and torch.equal(nm, ref.context_stream.null_mask)
            # stateful quirk: visible only after a forward, gone after a reset
            assert hasattr(ref, "padding_mask")
            ref._reset_padding_mask()
            assert not hasattr(ref, "padding_mask")
        # ---- pin the oracle ----
        (x_m, mask_m, _), (x_c, mask_c, _) = ref.get_stream_inputs(xin, mask, None, x_context=imu, mask_context=mc)
        ocfg = synthetic.conjoined_oracle_cfg(name)
        y_or, yc_or = co.conjoined_forward(ref.state_dict(), x_m, mask_m, x_c, mask_c, ocfg, True, True)
        err, errc = (y_or - y_ref).abs().max().item() if y_ref.numel() else 0.0, (yc_or - yc_ref).abs().max().item()
        assert y_or.shape == y_ref.shape and yc_or.shape == yc_ref.shape, (y_or.shape, y_ref.shape, yc_or.shape, yc_ref.shape)
        assert err < 2e-5 and errc < 2e-5, (case, err, errc)
        # ---- through the wrapper (equal visible counts only: predict() rectangularises ragged batches) ----
        if padded and len(set(CASES[case][5])) == 1 and CASES[case][6] == "visible":
            ref._reset_padding_mask()
            # the output selection is cached on the module between calls (conjoined_vmae.py:589-593); the wrapper
            # expects the main-stream tensor only
            ref._set_decoder_outputs(output_main=True, output_context=False)
            G = rpred.PredictorBasedGenerator(predictor=ref, imagenet_normalize_inputs=True, temporal_dim=2)
            captured = {}
            hook = ref.register_forward_hook(lambda m_, i_, o_: captured.__setitem__("y", o_.detach().clone()))
            with torch.no_grad():
                video = G.predict(x.clone(), mask.clone(), frame=None, x_context=imu, mask_context=mc.clone())
            hook.remove()
            extra["video_fingerprint"] = np.array([float(video.double().sum()), float(video.double().pow(2).sum())])
            extra["y_predict"] = captured["y"].numpy()       # predictor output on the imagenet-normalised input
            if video.numel() <= 200_000:
                extra["video"] = video.numpy()
        m, ms = pack(mask)
        mcp, mcs = pack(mc)
        out = dict(mask=m, mask_shape=ms, mask_ctx=mcp, mask_ctx_shape=mcs, y=y_ref.numpy(), y_ctx=yc_ref.numpy(),
                   oracle_vs_reference_maxabs=np.array([err, errc]),
                   weights_checksum=np.array([synthetic.weights_checksum(ref)]),
                   x_fingerprint=np.array([float(x.double().sum()), float(x.double().pow(2).sum())]),
                   imu_fingerprint=np.array([float(imu.double().sum()), float(imu.double().pow(2).sum())]),
                   num_params=np.array([sum(p.numel() for p in ref.parameters())]), **extra)
        path = os.path.join(GOLDEN_DIR, case + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case}: params {out['num_params'][0]} y {tuple(y_ref.shape)} std {float(y_ref.std()) if y_ref.numel() else 0:.3f} "
              f"y_ctx {tuple(yc_ref.shape)} std {float(yc_ref.std()):.3f} | oracle-vs-ref {err:.2e} / {errc:.2e} | "
              f"{time.time() - t0:.1f}s | {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main(sys.argv[1:])
