"""masks on the device from a counter-based RNG (csrc/masks.cu, device_masks.py).

CPU: the generator against the published Random123 known answers (both the numpy oracle and the library's host entry
point), the oracle's selection logic against the contracts of the reference functions it stands in for.
GPU: kernels bit-exact against the oracle; a sweep's masks identical whether they are generated in 1, 2, 4 or 8 shards.
"""
import ctypes

import numpy as np
import pytest
import torch

import device_masks_oracle as dmo
from counterfactualworldmodels_b200 import _lib

DEV = "cuda:0"


def test_philox_known_answers_oracle_and_library():
    lib = _lib.load()
    for ctr, key, want in dmo.KAT:
        got = tuple(int(v) for v in dmo.philox4x32_10(*ctr, *key))
        assert got == want
        C, K, O = (ctypes.c_uint32 * 4)(*ctr), (ctypes.c_uint32 * 2)(*key), (ctypes.c_uint32 * 4)()
        assert lib.cwm_philox4x32_10(C, K, O) == 0
        assert tuple(O) == want
    # word addressing: word i of a stream is output[i % 4] of counter block i // 4
    w = dmo.philox_words(0x1234567890, 7, 1, 2, 10)
    blk = dmo.philox4x32_10(2, 7, 1, 2, 0x34567890, 0x12)
    assert int(w[9]) == int(blk[1]) and int(w[8]) == int(blk[0])


def test_oracle_uniform_masks_follow_the_reference_contract():
    """masking.py:347-376 / :478-545: frame 0 visible, exactly k clumps of c x c patches visible in frame 1."""
    m = dmo.mask_uniform(seed=3, row0=0, rows=64, visible_frames=1, mask_frames=1, h=8, w=12, clump=2, n_visible_cells=3)
    m = m.reshape(64, 2, 8, 12)
    assert not m[:, 0].any()
    assert ((m[:, 1] == 0).sum((1, 2)) == 12).all()
    cells = (m[:, 1].reshape(64, 4, 2, 6, 2) == 0)
    assert (cells.all((2, 4)) | ~cells.any((2, 4))).all()          # whole clumps only
    # uniform over cells: 64 rows x 3 of 24 cells
    hits = cells.all((2, 4)).sum(0).reshape(-1)
    assert hits.sum() == 192 and hits.max() <= 20 and hits.min() >= 1
    # a row is a function of (seed, global row) alone
    again = dmo.mask_uniform(3, 5, 2, 1, 1, 8, 12, 2, 3).reshape(2, 2, 8, 12)
    assert np.array_equal(again, m[5:7])
    assert not np.array_equal(dmo.mask_uniform(4, 5, 2, 1, 1, 8, 12, 2, 3).reshape(2, 2, 8, 12), m[5:7])


def test_oracle_energy_sampling_is_proportional_and_collapses_duplicates():
    """utils.py:152-172 with normalize=True: p = relu(e - min e + eps) / sum; draws with replacement."""
    e = np.zeros((1, 16), np.float32)
    e[0, 3], e[0, 9] = 3.0, 1.0
    t = dmo.energy_table(e, 1e-16)
    assert int(t[0, -1]) == 16777216 + 16777216 // 3 and int(t[0, 2]) == 0   # eps-weight cells are never drawn
    ms = dmo.mask_energy_sample(t, 4, 4, 1, seed=11, sample0=0, S=400, points=1, visible_frames=1).reshape(400, 2, 16)
    vis = (ms[:, 1] == 0)
    assert (vis.sum(-1) == 1).all() and not ms[:, 0].any()
    n3, n9 = vis[:, 3].sum(), vis[:, 9].sum()
    assert n3 + n9 == 400 and 270 <= n3 <= 330                      # 3 : 1
    two = dmo.mask_energy_sample(t, 4, 4, 1, 11, 0, 200, 2, 1).reshape(200, 2, 16)
    counts = (two[:, 1] == 0).sum(-1)
    assert set(counts.tolist()) == {1, 2}                            # duplicates collapse (then: rectangularise)
    flat = dmo.energy_table(np.full((1, 8), 2.5, np.float32), 0.0)   # flat energy with eps = 0 -> uniform
    assert flat[0].tolist() == list(range(1, 9))


def test_oracle_rectangularize_min_mode_contract():
    """masking.py:100-132 'min': every row ends with min(masked) masked tokens; only masked tokens are revealed."""
    rng = np.random.RandomState(0)
    m = (rng.rand(9, 40) < 0.6).astype(np.uint8)
    out = dmo.rectangularize(m, row0=100, seed=5)
    target = (m != 0).sum(-1).min()
    assert ((out != 0).sum(-1) == target).all()
    assert ((out != 0) <= (m != 0)).all()
    # shard invariance: rows 3..5 alone with the global target give the same rows
    part = dmo.rectangularize(m[3:6], row0=103, seed=5, target_masked=int(target))
    assert np.array_equal(part, out[3:6])
    assert np.array_equal(dmo.rectangularize(out, 100, 5), out)      # idempotent


def test_device_generators_refuse_cpu():
    from counterfactualworldmodels_b200 import device_masks as dm
    g = dm.DeviceUniformMaskingGenerator((2, 8, 8), 0.75, clumping_factor=2, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(batch_size=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dm.DeterministicRectangularizeMasks()(torch.zeros(2, 8, dtype=torch.bool))
    lib = _lib.load()
    assert lib.cwm_mask_uniform(0, 0, 2, 1, 1, 7, 8, 2, 1, None, None) == -1 and b"multiple" in lib.cwm_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,cf,k,vf,mf", [(8, 12, 2, 3, 1, 1), (28, 28, 2, 2, 1, 1), (56, 56, 2, 8, 1, 1),
                                            (14, 14, 1, 20, 0, 2), (56, 56, 1, 31, 1, 1), (6, 10, 2, 15, 2, 1)])
def test_gpu_uniform_masks_equal_the_oracle(h, w, cf, k, vf, mf):
    from counterfactualworldmodels_b200 import device_masks as dm
    g = dm.DeviceUniformMaskingGenerator((vf + mf, h, w), 0.0, visible_frames=vf, seed=0xC0FFEE1234, clumping_factor=cf,
                                         device=DEV)
    g.num_visible = k
    got = g(batch_size=5, sample_offset=17).cpu().numpy().astype(np.uint8)
    want = dmo.mask_uniform(0xC0FFEE1234, 17, 5, vf, mf, h, w, cf, k)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_energy_sampling_equals_the_oracle():
    from counterfactualworldmodels_b200 import device_masks as dm
    torch.manual_seed(0)
    energy = torch.rand(2, 1, 32, 48) ** 3
    g = dm.DeviceEnergyMaskingGenerator((2, 8, 12), 0, seed=99, clumping_factor=2, energy_power=2, eps=1e-16,
                                        pool_mode='mean', device=DEV)
    g.num_visible = 3
    got = g.sample(energy.to(DEV), 6, sample_offset=40)               # [B, N, S]
    assert tuple(got.shape) == (2, 2 * 8 * 12, 6)
    w = g._cell_weights(energy.to(DEV)).cpu().numpy()                 # torch pooling on the device, exact fp32 inputs
    table = dmo.energy_table(w, 1e-16)
    want = dmo.mask_energy_sample(table, 8, 12, 2, 99, 40, 6, 3, 1).reshape(2, 6, -1)
    assert np.array_equal(got.permute(0, 2, 1).cpu().numpy().astype(np.uint8), want)
    # flat energy: every clump equally likely, one visible clump per sample
    g.num_visible = 1
    flat = g.sample(torch.ones(1, 1, 8, 12, device=DEV), 512)
    vis = (~flat[0]).view(2, 8, 12, 512)
    assert vis[0].all() and (vis[1].sum((0, 1)) == 4).all()     # frame 0 visible, one 2x2 clump in frame 1
    per_cell = vis[1].view(4, 2, 6, 2, 512).all(1).all(2).sum(-1).flatten()
    assert per_cell.sum() == 512 and per_cell.max() <= 45 and per_cell.min() >= 5


@pytest.mark.gpu
def test_gpu_rectangularize_equals_the_oracle_and_is_shard_invariant():
    from counterfactualworldmodels_b200 import device_masks as dm
    rng = np.random.RandomState(1)
    m = rng.rand(37, 6272) < 0.52
    m[5] = m[4]                                                         # equal rows stay equal only if their keys differ
    rect = dm.DeterministicRectangularizeMasks(seed=77)
    got = rect(torch.from_numpy(m).to(DEV).clone(), row_offset=1000).cpu().numpy()
    want = dmo.rectangularize(m.astype(np.uint8), 1000, 77)
    assert np.array_equal(got.astype(np.uint8), want)
    target = int(m.sum(-1).min())
    assert (got.sum(-1) == target).all()
    for world in (2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = r * 37 // world, (r + 1) * 37 // world
            parts.append(rect(torch.from_numpy(m[lo:hi]).to(DEV).clone(), row_offset=1000 + lo, target_masked=target).cpu())
        assert torch.equal(torch.cat(parts, 0), torch.from_numpy(got)), world


@pytest.mark.gpu
def test_gpu_sweep_masks_are_identical_for_1_2_4_8_shards():
    """The acceptance test of SURVEY 8f rank 4: the masks of a 1024-sample sweep do not depend on the number of ranks."""
    from counterfactualworldmodels_b200 import device_masks as dm
    S = 1024
    uni = dm.DeviceUniformMaskingGenerator((2, 56, 56), 0.0, seed=5, clumping_factor=2, device=DEV)
    uni.num_visible = 2
    en = dm.DeviceEnergyMaskingGenerator((2, 56, 56), 0, seed=6, clumping_factor=2, device=DEV)
    en.num_visible = 1
    energy = torch.rand(1, 1, 224, 224, generator=torch.Generator().manual_seed(3)).to(DEV)
    full_u = uni(batch_size=S)
    full_e = en.sample(energy, S)
    assert (~full_u).view(S, 2, -1)[:, 1].sum(-1).eq(8).all() and (~full_e[0]).view(2, 56 * 56, S)[1].sum(0).eq(4).all()
    for world in (2, 4, 8):
        n = S // world
        pu = torch.cat([uni(batch_size=n, sample_offset=r * n) for r in range(world)], 0)
        pe = torch.cat([en.sample(energy, n, sample_offset=r * n) for r in range(world)], -1)
        assert torch.equal(pu, full_u) and torch.equal(pe, full_e), world


@pytest.mark.gpu
def test_gpu_flow_generator_with_device_masks_runs_a_sweep():
    from counterfactualworldmodels_b200 import segmentation, synthetic, vmae
    cfg = "tiny_8x8"
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg))
    synthetic.init_weights_(m, seed=1, style="perturbed")
    m = m.to(DEV).eval()
    x = synthetic.make_video(1, synthetic.image_hw(cfg), seed=2)[:, 0].to(DEV)
    outs = []
    for _ in range(2):
        G = segmentation.FlowGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2, device_masks=True, seed=4)
        G.set_input(x)
        active = G.sample_patches_from_energy(torch.rand(1, 1, 64, 64, generator=torch.Generator().manual_seed(5)).to(DEV),
                                              num_samples=6, num_visible=1)
        passive = G.sample_patches_from_energy(None, num_samples=6, num_visible=2)
        assert tuple(active.shape) == (1, 128, 6) and active.dtype == torch.bool
        y = G.predict_counterfactual_videos(x, active, passive_patches=passive, shifts=[[1, 0]] * 6, sample_batch_size=4)
        assert tuple(y.shape) == (6, 2, 3, 64, 64) and torch.isfinite(y).all()
        outs.append((active.clone(), passive.clone(), y.clone()))
    # same seeds -> same sweep, rectangulariser included (no global torch RNG involved)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
