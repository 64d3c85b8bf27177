"""TEST INFRASTRUCTURE ONLY -- pins the oracle against the REAL reference and writes tests/golden/*.npz.

Run in the build container (needs /root/reference):  python oracle/make_golden.py [case ...]

For every case it
  1. builds the reference model (`cwm.models.VideoMAE.vmae.PretrainVisionTransformer`) with the case's kwargs,
     overwrites its weights with `synthetic.init_weights_(seed)` (reproducible on the GPU box from the seed),
  2. runs the reference end to end through `cwm.models.prediction.PredictorBasedGenerator.predict`
     (`_preprocess` -> predictor -> `pred_patches_to_video`) in fp32 on the CPU, capturing the predictor output,
  3. checks `oracle/vmae_oracle.py` against it (patch predictions, assembled video, compaction, stage taps),
  4. stores mask, predictor output y, video (small cases) / video checksum and fingerprints of x and the weights.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
import vmae_oracle as oracle  # noqa: E402
from counterfactualworldmodels_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name: (config, batch, init style, weight seed, data seed, mask spec)
#   mask spec: ("clumps", n) = synthetic.make_mask with n visible 2x2 clumps in frame 1
#              ("reference_generator",) = the reference's RotatedTableUniformMaskingGenerator(0.99, clumping 2, seed 0)
CASES = {
    "tiny_4x4_b2": ("tiny_4x4", 2, "perturbed", 1, 1, ("clumps", 2)),
    "tiny_8x8_b3": ("tiny_8x8", 3, "perturbed", 2, 2, ("clumps", 1)),
    "small_4x4_b2": ("small_4x4", 2, "perturbed", 3, 3, ("clumps", 2)),
    "small_4x4_allvisible_frame1half": ("small_4x4", 1, "perturbed", 3, 4, ("clumps", 56)),
    "base_8x8_b1_factual": ("base_8x8", 1, "reference", 0, 0, ("reference_generator",)),
    "base_8x8_b2_counterfactual": ("base_8x8", 2, "reference", 0, 5, ("clumps", 1)),
    "base_4x4_b1": ("base_4x4", 1, "reference", 0, 6, ("clumps", 8)),
    "large_4x4_b1_factual": ("large_4x4", 1, "reference", 0, 7, ("reference_generator",)),
    "tiny_4x4_tube2_b2": ("tiny_4x4_tube2", 2, "perturbed", 8, 8, ("clumps", 3)),
    "tiny_8x8_layerscale_learnpos_b2": ("tiny_8x8_layerscale_learnpos", 2, "perturbed", 9, 9, ("clumps", 2)),
}


def build_case_inputs(case):
    cfg_name, B, style, wseed, dseed, mspec = CASES[case]
    x = synthetic.make_video(B, synthetic.image_hw(cfg_name), seed=dseed, T=synthetic.CONFIGS[cfg_name]["num_frames"])
    return cfg_name, B, style, wseed, x, mspec


def main(argv):
    ref_vmae, ref_pred = ref_loader.import_reference()
    import cwm.models.masking as ref_masking
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = argv or list(CASES)
    for case in names:
        cfg_name, B, style, wseed, x, mspec = build_case_inputs(case)
        t0 = time.time()
        torch.manual_seed(0)
        ref = ref_vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name)).eval().requires_grad_(False)
        synthetic.init_weights_(ref, seed=wseed, style=style)
        msize = ref.mask_size
        if mspec[0] == "clumps":
            mask = synthetic.make_mask(B, msize, num_clumps=mspec[1], seed=CASES[case][4])
        else:
            gen = ref_masking.RotatedTableUniformMaskingGenerator(
                input_size=msize, mask_ratio=0.99, clumping_factor=2, seed=0)
            mask = gen(x).view(B, -1).clone()
        assert mask.shape == (B, ref.num_patches)

        captured = {}
        hook = ref.register_forward_hook(lambda m, i, o: captured.__setitem__("y", o.detach().clone()))
        G = ref_pred.PredictorBasedGenerator(predictor=ref, imagenet_normalize_inputs=True, temporal_dim=2)
        with torch.no_grad():
            video = G.predict(x.clone(), mask.clone(), frame=None)
        hook.remove()
        y_ref = captured["y"]
        t_ref = time.time() - t0

        # ---- pin the oracle against the reference ----
        sd = ref.state_dict()
        ocfg = synthetic.oracle_cfg(cfg_name)
        y_or = oracle.vmae_forward(sd, oracle.preprocess(x), mask, ocfg)
        err_y = (y_or - y_ref).abs().max().item()
        v_or = oracle.pred_patches_to_video(y_ref, x, mask, ocfg["patch_size"])
        assert torch.equal(v_or, video), "oracle pred_patches_to_video differs from the reference"
        perm, inv, nvis = oracle.compact_mask(mask.numpy())
        for b in range(B):
            vis_ref = torch.nonzero(~mask[b]).flatten().numpy()
            msk_ref = torch.nonzero(mask[b]).flatten().numpy()
            assert np.array_equal(perm[b, :nvis[b]], vis_ref) and np.array_equal(perm[b, nvis[b]:], msk_ref)
        assert torch.equal(oracle.sinusoid_table_cached(ref.num_patches, ref.pos_embed.shape[-1]), ref.pos_embed)
        if not ref.encoder._learnable_pos_embed:
            assert torch.equal(oracle.sinusoid_table_cached(ref.num_patches, ref.encoder.pos_embed.shape[-1]),
                               ref.encoder.pos_embed)
        tol = 2e-5
        assert err_y < tol, f"{case}: oracle vs reference max-abs {err_y}"
        # visible patches of the output video are bit-identical to the input (SURVEY section 4 [probe])
        up = oracle.patchify(video, ocfg["patch_size"])
        xp = oracle.patchify(x, ocfg["patch_size"])
        assert torch.equal(up[~mask], xp[~mask])

        out = dict(
            mask=np.packbits(mask.numpy().astype(np.uint8), axis=1), mask_shape=np.array(mask.shape),
            y=y_ref.numpy().astype(np.float32),
            x_fingerprint=np.array([float(x.double().sum()), float(x.double().pow(2).sum())]),
            weights_checksum=np.array([synthetic.weights_checksum(ref)]),
            video_fingerprint=np.array([float(video.double().sum()), float(video.double().pow(2).sum())]),
            oracle_vs_reference_maxabs=np.array([err_y]),
            num_params=np.array([sum(p.numel() for p in ref.parameters())]),
            n_masked=np.array([int(mask[0].sum())]),
        )
        if video.numel() <= 200_000:
            out["video"] = video.numpy().astype(np.float32)
        path = os.path.join(GOLDEN_DIR, case + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case}: params {out['num_params'][0]} Nmask {out['n_masked'][0]} y {tuple(y_ref.shape)} "
              f"std {y_ref.std():.3f} | oracle-vs-ref {err_y:.2e} | ref {t_ref:.1f}s | {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main(sys.argv[1:])
