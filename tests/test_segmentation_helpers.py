"""Host-side argument normalisation of the counterfactual entry points (segmentation.py:289-304, :364-410 of the reference):
the helpers that bring images to 2-frame movies and patch tensors to a common sample axis."""
import pytest
import torch

from counterfactualworldmodels_b200.segmentation import _two_frame_movie, _with_sample_axis


@pytest.mark.parametrize("shape,still", [((3, 8, 8), True), ((2, 3, 8, 8), True), ((2, 1, 3, 8, 8), False), ((2, 4, 3, 8, 8), False)])
def test_two_frame_movie(shape, still):
    x = torch.arange(float(torch.Size(shape).numel())).reshape(shape)
    y, was_image = _two_frame_movie(x)
    B = 1 if len(shape) == 3 else shape[0]
    assert y.shape == (B, 2, 3, 8, 8) and was_image == still
    if len(shape) == 5 and shape[1] >= 2:
        assert torch.equal(y, x[:, :2])
    else:   # a still image / single frame is repeated, not copied
        assert torch.equal(y[:, 0], y[:, 1]) and y.stride(1) == 0


def test_two_frame_movie_rejects_other_ranks():
    with pytest.raises(AssertionError):
        _two_frame_movie(torch.zeros(4, 4))


def test_with_sample_axis():
    m = torch.zeros(2, 6, dtype=torch.bool)
    assert _with_sample_axis(m).shape == (2, 6, 1)
    assert _with_sample_axis(m, 5).shape == (2, 6, 5) and _with_sample_axis(m, 5).stride(-1) == 0
    m3 = torch.zeros(2, 6, 1, dtype=torch.bool)
    assert _with_sample_axis(m3, 4).shape == (2, 6, 4)
    m4 = torch.zeros(2, 6, 3, dtype=torch.bool)
    assert _with_sample_axis(m4, 4) is m4                 # a real sample axis is left alone (the caller asserts on it)
    assert _with_sample_axis(m, 1).shape == (2, 6, 1)
