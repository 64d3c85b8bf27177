"""CPU checks for SURVEY.md section 8(f) rank 1 (batched motion-counterfactual construction): the numpy oracle
against the fixtures the REAL reference produced (tests/golden/cf_*.npz, written by
oracle/make_golden_counterfactual.py), against the live reference when /root/reference is mounted, and the host
logic of the mirror classes (shift pre-processing, RNG parity, error behaviour)."""
import os

import numpy as np
import pytest
import torch

import counterfactual_oracle as cfo
from conftest import GOLDEN_DIR, needs_reference
from counterfactualworldmodels_b200 import perturbation, synthetic

CF_CASES = ["cf_tiny_4x4_s6", "cf_tiny_8x8_s8_clump2", "cf_small_4x4_s8_moving_input", "cf_base_8x8_s8_preset"]


def load_cf(case):
    path = os.path.join(GOLDEN_DIR, case + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {case} missing")
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    B, N, S = d["shape"]
    unpack = lambda a, shape: np.unpackbits(a)[:int(np.prod(shape))].reshape(shape).astype(bool)
    d["passive"] = unpack(d["passive"], (B, N, S))
    d["active"] = unpack(d["active"], (B, N, S))
    d["mask_shift"] = unpack(d["mask_shift"], tuple(d["mask_shift_shape"]))
    d["mask_shift_rect"] = unpack(d["mask_shift_rect_min_seed1234"], tuple(d["mask_shift_shape"]))
    d["cfg"] = str(d["cfg"])
    d["seed"] = int(d["seed"][0])
    d["patch_size"] = tuple(int(v) for v in d["patch_size"])
    d["x"] = synthetic.make_video(1, synthetic.image_hw(d["cfg"]), seed=d["seed"])
    assert np.allclose(cfo.fingerprint(d["x"].numpy()), d["x_fingerprint"], rtol=0, atol=1e-6)
    return d


@pytest.mark.parametrize("case", CF_CASES)
def test_oracle_matches_reference_fixture(case):
    d = load_cf(case)
    xs, ms = cfo.create_motion_counterfactuals(d["x"].numpy(), d["passive"], d["active"], d["shifts"].tolist(),
                                               d["patch_size"], frame=1, fix_passive=bool(d["static"][0]))
    assert np.array_equal(ms, d["mask_shift"])
    assert np.array_equal((~ms).sum(-1), d["n_visible"])
    assert np.allclose(cfo.fingerprint(xs), d["x_shift_fingerprint"], rtol=0, atol=1e-6)
    if "x_shift" in d:
        assert np.array_equal(xs.view(np.uint32), d["x_shift"].view(np.uint32))


def test_oracle_shift_semantics_small():
    """Hand-checked example: content moves by +shift, zeros enter at the border, the mask pads with 1."""
    a = np.arange(12, dtype=np.float32).reshape(1, 3, 4)
    out = cfo.shift_zero_fill(a, 1, -1, 0.0)
    assert np.array_equal(out[0], np.array([[0, 0, 0, 0], [1, 2, 3, 0], [5, 6, 7, 0]], np.float32))
    assert np.array_equal(cfo.shift_zero_fill(a, 5, 0, 1.0), np.ones_like(a))
    (sy, sx), (my, mx) = cfo.get_padding_shifts((8, -16), (1, 8, 8))
    assert (sy, sx, my, mx) == (8, -16, 1, -2)


def test_shifter_host_logic_matches_reference_rng():
    for case in CF_CASES:
        d = load_cf(case)
        sh = perturbation.ShiftPatchesAndMask(patch_size=d["patch_size"], padding_mode='constant',
                                              max_shift_fraction=0.15, allow_fractional_shifts=False, seed=d["seed"])
        sh.set_shapes(d["x"], mask=torch.from_numpy(d["active"][..., 0]))
        got = [sh.get_random_shift(True) for _ in range(6)] + [sh.get_random_shift(False) for _ in range(6)]
        assert np.array_equal(np.array(got, np.int32), d["random_shifts"]), case


def test_preprocess_shifts_sequence():
    sh = perturbation.ShiftPatchesAndMask(patch_size=(1, 8, 8))
    sh.set_num_shifts(3)
    assert sh._preprocess_shifts_sequence([[1, 0]], is_mask_shift=True) == [[1, 0]] * 3
    assert sh._preprocess_shifts_sequence([1, 2]) == [[1, 2]] * 3
    # tensor input: the reference turns [2, S] into a list of numpy arrays, wraps that list once more because its
    # entries are not lists (perturbation.py:206-208) and then fails `len(s) == 2` for S != 2 -- same here
    with pytest.raises(AssertionError):
        sh._preprocess_shifts_sequence(torch.tensor([[1, 2, 3], [4, 5, 6]]))
    with pytest.raises(AssertionError):
        sh._preprocess_shifts_sequence([[1, 0], [0, 1]])
    with pytest.raises(NotImplementedError):
        perturbation.ShiftPatchesAndMask(patch_size=(1, 8, 8), allow_fractional_shifts=True)


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device, never route through the oracle."""
    x = synthetic.make_video(1, (32, 32), seed=0)
    m = torch.zeros(2, 128, dtype=torch.bool)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        perturbation.shift_patches_and_masks(x, m, m, [[1, 0], [0, 1]], (1, 4, 4))
    src = open(os.path.join(os.path.dirname(perturbation.__file__), "perturbation.py")).read() + \
        open(os.path.join(os.path.dirname(perturbation.__file__), "segmentation.py")).read()
    assert "oracle" not in src.replace("SURVEY", "")


@needs_reference
def test_oracle_matches_live_reference():
    import ref_loader
    ref_vmae, _ = ref_loader.import_reference()
    import cwm.models.perturbation as ref_pert
    rng = np.random.RandomState(5)
    x = torch.rand(2, 2, 3, 24, 40, generator=torch.Generator().manual_seed(5))
    sh = ref_pert.ShiftPatchesAndMask(patch_size=(1, 4, 4), padding_mode='constant', allow_fractional_shifts=False)
    for trial in range(6):
        mask = torch.from_numpy(rng.rand(2, 2 * 6 * 10) < 0.7)
        points = torch.from_numpy(rng.rand(2, 2 * 6 * 10) < 0.1)
        ms = [int(rng.randint(-7, 8)), int(rng.randint(-11, 12))]
        xr, mr = sh(x.clone(), mask=mask.clone(), perturbation_points=points.clone(), mask_shift=ms, frame=trial % 2)
        xo, mo = cfo.perturbation_forward(x.numpy(), mask.numpy(), points.numpy(), (1, 4, 4), mask_shift=ms,
                                          frame=trial % 2)
        assert np.array_equal(mo, mr.numpy())
        assert np.array_equal(xo.view(np.uint32), xr.numpy().view(np.uint32))
