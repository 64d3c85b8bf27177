"""TEST INFRASTRUCTURE ONLY -- pins oracle/counterfactual_oracle.py against the REAL reference and writes
tests/golden/cf_*.npz (batched motion-counterfactual construction).

Run in the build container (needs /root/reference):  python oracle/make_golden_counterfactual.py

For every case it builds the reference ``FlowGenerator`` (cwm/models/segmentation.py:23) around a tiny reference VMAE
(only its patch size matters here), runs ``create_motion_counterfactuals`` (segmentation.py:279-343) with the mask
rectangulariser switched off ('none') and again with the default 'min' mode under a fixed global torch seed, and
``PredictorBasedGenerator.make_static`` / ``_shift`` (prediction.py:51, :756-779); asserts the oracle reproduces every
video and mask bit for bit; stores the inputs that cannot be regenerated from a seed plus the expected outputs.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import counterfactual_oracle as cfo  # noqa: E402
import ref_loader  # noqa: E402
from counterfactualworldmodels_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name: (vmae config (patch size + image size), S, seed, static input?, clump side of the active patches)
CASES = {
    "cf_tiny_4x4_s6": ("tiny_4x4", 6, 11, True, 1),
    "cf_tiny_8x8_s8_clump2": ("tiny_8x8", 8, 12, True, 2),
    "cf_small_4x4_s8_moving_input": ("small_4x4", 8, 13, False, 2),
    "cf_base_8x8_s8_preset": ("base_8x8", 8, 14, True, 2),
}
PRESET_SHIFTS = [[2, 0], [0, 2], [-2, 0], [0, -2], [2, 2], [-2, -2], [2, -2], [-2, 2]]  # ipynb:1726


def make_case(name):
    """Seeded inputs of a case: video [1,2,3,H,W], passive masks / active patches bool [1,N,S], S mask shifts."""
    cfg, S, seed, static, clump = CASES[name]
    T, h, w = synthetic.mask_size(cfg)
    x = synthetic.make_video(1, synthetic.image_hw(cfg), seed=seed)
    rng = np.random.RandomState(seed)
    n = h * w
    active = np.ones((1, T, h, w, S), bool)
    passive = np.zeros((1, T, h, w, S), bool)
    passive[:, -1] = True
    shifts = []
    for s in range(S):
        # one active clump (sometimes at the border so the shift pushes it out of the frame), 0-2 passive clumps
        # (sometimes on top of the active one: perturbation.py:106 removes the common visible patches)
        ay = rng.choice([0, h - clump, rng.randint(0, h - clump + 1)])
        ax = rng.choice([0, w - clump, rng.randint(0, w - clump + 1)])
        active[0, -1, ay:ay + clump, ax:ax + clump, s] = False
        for k in range(rng.randint(0, 3)):
            py, px = (ay, ax) if (k == 1 and s % 3 == 0) else (rng.randint(0, h - clump + 1), rng.randint(0, w - clump + 1))
            passive[0, -1, py:py + clump, px:px + clump, s] = False
        if name.endswith("preset"):
            shifts.append(list(PRESET_SHIFTS[s % 8]))
        else:
            sh = [0, 0]
            while sh == [0, 0]:
                sh = [int(rng.randint(-3, 4)), int(rng.randint(-3, 4))]
            if s == S - 1:
                sh = [h + 1, -2]  # larger than the image: everything shifted out
            shifts.append(sh)
    return x, passive.reshape(1, T * n, S), active.reshape(1, T * n, S), shifts, static


def main():
    ref_vmae, ref_pred = ref_loader.import_reference()
    import cwm.models.segmentation as ref_seg
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, (cfg, S, seed, static, clump) in CASES.items():
        kw = synthetic.model_kwargs(cfg)
        kw.update(encoder_depth=1, decoder_depth=1, encoder_embed_dim=64, decoder_embed_dim=64, encoder_num_heads=1,
                  decoder_num_heads=1)
        ref = ref_vmae.PretrainVisionTransformer(**kw).eval().requires_grad_(False)
        G = ref_seg.FlowGenerator(predictor=ref, flow_model=nn.Identity(), imagenet_normalize_inputs=True,
                                  temporal_dim=2)
        patch_size = tuple(ref.patch_size)
        x, passive, active, shifts, static = make_case(name)
        xt, pt_, at_ = x.clone(), torch.from_numpy(passive), torch.from_numpy(active)

        def run(mode, seed_global):
            G.set_input(xt)
            G.reset_shifts()
            G.shifter.set_shapes(xt, mask=at_[..., 0])
            G.shifter.set_num_shifts(S)
            G.mask_rectangularizer.set_mode(mode)
            torch.manual_seed(seed_global)
            xs, ms = G.create_motion_counterfactuals(xt.clone(), masks=pt_.clone(), active_patches=at_.clone(),
                                                     shifts=[list(s) for s in shifts], num_samples=S,
                                                     fix_passive=static, reset_shifts=False)
            return xs, ms

        xs_ref, ms_ref = run('none', 0)
        _, ms_rect = run('min', 1234)
        xs_or, ms_or = cfo.create_motion_counterfactuals(x.numpy(), passive, active, shifts, patch_size, frame=1,
                                                         fix_passive=static)
        assert np.array_equal(ms_or, ms_ref.numpy()), f"{name}: oracle mask differs from the reference"
        assert np.array_equal(xs_or.view(np.uint32), xs_ref.numpy().view(np.uint32)), \
            f"{name}: oracle video differs from the reference (bitwise)"
        assert [list(map(int, s)) for s in G.shifts[-S:]] == [list(s) for s in shifts]

        # MakeStatic + the single-sample _shift used by get_counterfactual_prediction (prediction.py:781-813)
        x2 = synthetic.make_video(2, synthetic.image_hw(cfg), seed=seed + 100, counterfactual_like=True)
        m2 = synthetic.make_mask(2, ref.mask_size, num_clumps=3, clump=clump, seed=seed)
        G.set_input(x2)
        xst_ref, _ = G.make_static(x2.clone(), m2.clone())
        xst_or, _ = cfo.make_static(x2.numpy(), m2.numpy(), patch_size)
        assert np.array_equal(xst_or.view(np.uint32), xst_ref.numpy().view(np.uint32)), f"{name}: make_static differs"
        act2 = torch.from_numpy(active[:, :, :2].transpose(0, 2, 1).reshape(2, -1).copy())
        G.mask_rectangularizer.set_mode('none')
        xsh_ref, msh_ref = G._shift(x2.clone(), m2.clone(), active_patches=act2.clone(), shift=shifts[0], frame=1)
        xsh_or, msh_or = cfo.shift_one(x2.numpy(), m2.numpy(), act2.numpy(), patch_size, shift=shifts[0], frame=1)
        assert np.array_equal(msh_or, msh_ref.numpy()) and \
            np.array_equal(xsh_or.view(np.uint32), xsh_ref.numpy().view(np.uint32)), f"{name}: _shift differs"

        # host RNG parity: the first draws of `get_random_shift` (perturbation.py:218-234) of a fresh shifter
        import cwm.models.perturbation as ref_pert
        sh = ref_pert.ShiftPatchesAndMask(patch_size=patch_size, padding_mode='constant', max_shift_fraction=0.15,
                                          allow_fractional_shifts=False, seed=seed)
        sh.set_shapes(xt, mask=at_[..., 0])
        random_mask_shifts = np.array([sh.get_random_shift(True) for _ in range(6)] +
                                      [sh.get_random_shift(False) for _ in range(6)], np.int32)

        fp = cfo.fingerprint

        out = dict(
            passive=np.packbits(passive.astype(np.uint8)), active=np.packbits(active.astype(np.uint8)),
            shape=np.array(passive.shape), shifts=np.array(shifts, np.int32), static=np.array([int(static)]),
            mask_shift=np.packbits(ms_ref.numpy().astype(np.uint8)), mask_shift_shape=np.array(ms_ref.shape),
            mask_shift_rect_min_seed1234=np.packbits(ms_rect.numpy().astype(np.uint8)),
            x_fingerprint=fp(x.numpy()), x_shift_fingerprint=fp(xs_ref.numpy()),
            n_visible=(~ms_ref.numpy()).sum(-1).astype(np.int32),
            make_static_fingerprint=fp(xst_ref.numpy()), shift_one_fingerprint=fp(xsh_ref.numpy()),
            shift_one_mask=np.packbits(msh_ref.numpy().astype(np.uint8)),
            random_shifts=random_mask_shifts, cfg=np.array(cfg), seed=np.array([seed]), clump=np.array([clump]),
            patch_size=np.array(patch_size),
        )
        if xs_ref.numel() <= 120_000:
            out["x_shift"] = xs_ref.numpy()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: S={S} visible/row {out['n_visible'].tolist()} oracle == reference (bitwise) | "
              f"{os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
