"""SURVEY.md section 8(f) rank 4 -- the mask generators (host-side bookkeeping on the reference's RNG streams).
The mirror must reproduce, bit for bit, the masks the REAL reference drew under the same seeds
(tests/golden/masks_ref.npz, written by oracle/make_golden_masks.py with the very same `draw_all` script)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from counterfactualworldmodels_b200 import masking, sampling


def _golden():
    path = os.path.join(GOLDEN_DIR, "masks_ref.npz")
    if not os.path.exists(path):
        pytest.skip("masks_ref.npz missing")
    z = np.load(path)
    out = {}
    for k in z.files:
        if k.endswith("__shape"):
            continue
        shape = tuple(int(v) for v in z[k + "__shape"])
        out[k] = np.unpackbits(z[k])[:int(np.prod(shape))].reshape(shape).astype(bool)
    return out


class _CpuFlowGenerator:
    """`FlowGenerator` needs a predictor module; on a CPU-only box we build the mirror class around the parameter
    holder (no forward is ever called by the mask samplers)."""

    def __new__(cls, predictor=None, flow_model=None, seed=0):
        from counterfactualworldmodels_b200 import segmentation
        return segmentation.FlowGenerator(predictor=predictor, flow_model=flow_model, seed=seed)


def test_mirror_draws_the_reference_masks():
    import make_golden_masks
    from counterfactualworldmodels_b200 import vmae
    want = _golden()
    got = make_golden_masks.draw_all(masking, sampling, _CpuFlowGenerator, vmae)
    assert set(got) <= set(want) and len(got) >= 10      # (the fixture also holds the reference's IMU-driver masks)
    for k, v in got.items():
        assert tuple(v.shape) == want[k].shape, (k, v.shape, want[k].shape)
        assert np.array_equal(v.numpy(), want[k]), k


def test_known_answers_from_the_notebook():
    """ipynb cell 12: mask ratio 0.99, clumping 2 on the 4x4-patch model -> 6272 tokens, 3104 masked (32 visible)."""
    gen = masking.RotatedTableUniformMaskingGenerator(input_size=(2, 56, 56), mask_ratio=0.99, clumping_factor=2, seed=0)
    m = gen(torch.zeros(1, 2, 3, 8, 8))
    assert m.shape == (1, 6272) and int(m.sum()) == 3104 and not bool(m[:, :3136].any())
    gen8 = masking.RotatedTableUniformMaskingGenerator(input_size=(2, 28, 28), mask_ratio=0.99, clumping_factor=2, seed=0)
    m8 = gen8(torch.zeros(2, 2, 3, 8, 8))
    assert m8.shape == (2, 1568) and (~m8).sum(-1).tolist() == [792, 792]


def test_properties():
    gen = masking.MaskingGenerator(input_size=(1, 12, 12), mask_ratio=0.75, seed=1, always_batch=True)
    assert gen.num_visible == 36 and gen.num_masks_per_frame == 108
    gen.num_visible = 10
    assert gen.num_masks_per_frame == 134 and abs(gen.mask_ratio - 134 / 144) < 1e-12
    m = gen(torch.zeros(4, 1))
    assert m.shape == (4, 144) and (~m).sum(-1).tolist() == [10] * 4
    up = masking.upsample_masks(torch.eye(2, dtype=torch.bool)[None], (4, 6))
    assert up.shape == (1, 4, 6) and bool(up[0, 0, 0]) and bool(up[0, 1, 2]) and not bool(up[0, 0, 3])
    assert torch.equal(masking.upsample_masks(up, (2, 2)), torch.eye(2, dtype=torch.bool)[None])
    # energy sampling puts every visible clump where the energy is non-zero
    e = torch.zeros(1, 1, 16, 16)
    e[..., 4:8, 8:12] = 1.0
    g = masking.RotatedTableEnergyMaskingGenerator(input_size=(2, 8, 8), mask_ratio=0, seed=0, always_batch=True,
                                                   eps=1e-16, resize=False)
    g.num_visible = 2
    for _ in range(5):
        vis = ~g(e)[0, 64:].view(8, 8)
        assert int(vis.sum()) in (1, 2) and int(vis[2:4, 4:6].sum()) == int(vis.sum())


def test_batched_patch_sampling_properties():
    """The opt-in batched sampler (one multinomial call for the whole sweep): same shape / counts / support as the
    sequential reference sampler."""
    from counterfactualworldmodels_b200 import segmentation, synthetic, vmae
    kw = synthetic.model_kwargs("tiny_4x4")
    kw.update(encoder_depth=1, decoder_depth=1)
    G = segmentation.FlowGenerator(predictor=vmae.PretrainVisionTransformer(**kw), seed=3)
    x = synthetic.make_video(1, (32, 32), seed=1)
    G.set_input(x)
    e = torch.zeros(1, 1, 32, 32)
    e[..., 8:16, 20:28] = 1.0                       # patches rows 2-3, cols 5-6 of the 8 x 8 grid
    m = G.sample_patches_from_energy(e, num_samples=64, num_visible=2, batched=True)
    ref = G.sample_patches_from_energy(e, num_samples=4, num_visible=2)
    assert m.shape == (1, 128, 64) and m.dtype == torch.bool and ref.shape == (1, 128, 4)
    assert not bool(m[:, :64].any())                # frame 0 fully visible
    vis = ~m[0, 64:].view(8, 8, 64)
    assert int(vis[2:4, 5:7].sum()) == int(vis.sum())          # every draw lands where the energy is
    counts = vis.sum((0, 1))
    assert int(counts.min()) >= 1 and int(counts.max()) <= 2    # 2 draws with replacement
    assert len({tuple(vis[..., s].flatten().tolist()) for s in range(64)}) > 1   # samples differ
