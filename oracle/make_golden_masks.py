"""TEST INFRASTRUCTURE ONLY -- records masks drawn by the REAL reference mask generators (cwm/models/masking.py,
cwm/models/sampling.py, FlowGenerator.sample_patches_from_energy) under fixed seeds into tests/golden/masks_ref.npz.
The mirror in counterfactualworldmodels_b200/masking.py must reproduce them bit for bit (same RNG streams).
Needs /root/reference.   python oracle/make_golden_masks.py"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
from counterfactualworldmodels_b200 import synthetic  # noqa: E402


def energy_map(B, H, W, seed):
    g = torch.Generator().manual_seed(5000 + seed)
    e = torch.rand(B, 1, H, W, generator=g)
    e[:, :, H // 4:H // 2, W // 3:W // 2] += 3.0
    return e


def draw_all(masking, sampling, seg_cls, vmae_mod):
    """Every entry is drawn from freshly constructed generators so that the global torch RNG state is defined by the
    constructor's own `torch.manual_seed(seed)`.  Shared verbatim by the test (with the mirror modules)."""
    out = {}
    x3 = torch.zeros(3, 2, 3, 8, 8)
    gen = masking.RotatedTableUniformMaskingGenerator(input_size=(2, 28, 28), mask_ratio=0.99, clumping_factor=2, seed=0)
    out["rotated_uniform_28_r99_c2_seed0"] = torch.stack([gen(x3) for _ in range(2)], 0)
    gen = masking.RotatedTableUniformMaskingGenerator(input_size=(2, 56, 56), mask_ratio=0.99, clumping_factor=2, seed=0)
    out["rotated_uniform_56_r99_c2_seed0"] = gen(x3[:1])
    gen = masking.MaskingGenerator(input_size=(2, 7, 10), mask_ratio=0.6, clumping_factor=3, seed=3, always_batch=True,
                                   randomize_num_visible=True)
    out["uniform_7x10_c3_pad_randomvis_seed3"] = gen(x3)
    gen = masking.MaskingGenerator(input_size=(1, 6, 6), mask_ratio=0.5, seed=4)
    out["uniform_6x6_unbatched_seed4"] = gen()
    gen = masking.RotatedTableUniformMaskingGenerator(input_size=(3, 8, 8), mask_ratio=0.9, visible_frames=2,
                                                      context_mask_ratio=0.25, full_mask_prob=0.5, seed=5)
    out["rotated_uniform_ctx_fullprob_seed5"] = gen(x3)
    e = energy_map(2, 32, 32, 1)
    gen = sampling.RotatedTableEnergyMaskingGenerator(input_size=(2, 8, 8), mask_ratio=0, seed=6, always_batch=True,
                                                      energy_power=2, eps=1e-16, pool_mode='mean', resize=False)
    gen.num_visible = 3
    out["energy_8x8_vis3_seed6"] = torch.stack([gen(e) for _ in range(3)], -1)
    gen = sampling.RotatedTableEnergyMaskingGenerator(input_size=(2, 16, 16), mask_ratio=0, seed=7, always_batch=True,
                                                      clumping_factor=2, temperature=2.0, eps=1e-16, pool_mode='max',
                                                      resize=False)
    gen.num_visible = 2 * 4
    out["energy_16x16_c2_temp_seed7"] = gen(e)
    # FlowGenerator.sample_patches_from_energy (segmentation.py:118-128) through the wrapper's own rng streams
    kw = synthetic.model_kwargs("tiny_4x4")
    kw.update(encoder_depth=1, decoder_depth=1)
    G = seg_cls(predictor=vmae_mod.PretrainVisionTransformer(**kw).eval(), flow_model=nn.Identity(), seed=11)
    x = synthetic.make_video(1, (32, 32), seed=1)
    G.set_input(x)
    out["flowgen_uniform_energy_s5_seed11"] = G.sample_patches_from_energy(None, num_samples=5, num_visible=1)
    out["flowgen_energy_beta_s4_vis2"] = G.sample_patches_from_energy(energy_map(1, 32, 32, 2), num_samples=4,
                                                                      num_visible=2, beta=0.5)
    out["flowgen_zero_visible"] = G.sample_patches_from_energy(None, num_samples=2, num_visible=0)
    # IMU token masks (masking.py:402-476): missing-data tokens forced masked, full-mask / full-visible draws.  Drawn from
    # the reference only: these generators belong to its ImuGenerator driver, which runs unchanged over the drop-in
    # predictors and is not mirrored (tests/test_reference_wrappers_gpu.py)
    if not hasattr(masking, "MissingDataImuMaskGenerator"):
        return out
    miss = torch.zeros(4, 25, dtype=torch.bool)
    miss[1, :7] = True
    miss[3, 20:] = True
    gen = masking.MissingDataImuMaskGenerator(input_size=25, mask_ratio=0.4, full_mask_prob=0, full_vis_prob=0,
                                              truncation_mode='none', create_on_cpu=True, seed=8)
    out["imu_missing_none_r40_seed8"] = gen(miss)
    gen = masking.MissingDataImuMaskGenerator(input_size=25, mask_ratio=0.4, full_mask_prob=0.3, full_vis_prob=0.2,
                                              truncation_mode='max', seed=9)
    out["imu_missing_max_fullprob_seed9"] = torch.stack([gen(miss.clone()) for _ in range(4)], 0)
    gen = masking.ImuFullMaskGenerator(input_size=(5, 5), mask_ratio=0.5, clumping_factor=5, full_mask_prob=0.5,
                                       full_mask_per_example=True, seed=10)
    out["imu_full_per_example_seed10"] = gen(miss)
    return out


def main():
    ref_vmae, _ = ref_loader.import_reference()
    import cwm.models.masking as ref_masking
    import cwm.models.sampling as ref_sampling
    import cwm.models.segmentation as ref_seg
    out = draw_all(ref_masking, ref_sampling, ref_seg.FlowGenerator, ref_vmae)
    packed = {}
    for k, v in out.items():
        packed[k] = np.packbits(v.numpy().astype(np.uint8))
        packed[k + "__shape"] = np.array(v.shape)
        print(k, tuple(v.shape), "visible per row:", (~v.reshape(v.shape[0], -1)).sum(-1).tolist()[:4])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "masks_ref.npz"), **packed)


if __name__ == "__main__":
    main()
