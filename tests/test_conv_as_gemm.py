"""Pins the index arithmetic of DESIGN.md section 10 (1) -- RAFT's convolutions as one tcgen05 GEMM over a zero-padded
pixel-major buffer with tap-shifted TMA row coordinates -- against torch's conv2d, before the kernel exists."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import conv_gemm_oracle as cg


@pytest.mark.parametrize("kh,kw", [(1, 1), (3, 3), (1, 5), (5, 1), (7, 7)])
def test_convolution_equals_one_gemm_with_shifted_row_tiles(kh, kw):
    g = torch.Generator().manual_seed(kh * 10 + kw)
    S, C, H, W, N, P = 2, 8, 7, 9, 5, 3
    x = torch.randn(S, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(N, C, kh, kw, generator=g, dtype=torch.float64)
    b = torch.randn(N, generator=g, dtype=torch.float64)
    want = F.conv2d(x, w, b, 1, (kh // 2, kw // 2)).numpy()
    rows, geom = cg.to_rows(x.numpy(), P, Wp=16)            # Wp wider than W + 2P: a power-of-two pitch like 32 for 28
    out = cg.conv_as_gemm(rows, geom, w.numpy(), b.numpy())
    np.testing.assert_allclose(cg.from_rows(out, geom), want, rtol=1e-12, atol=1e-12)
    assert np.abs(out[cg.border_mask(geom)]).max() == 0     # the result is a valid padded buffer again


def test_two_chained_layers_keep_the_padding_invariant():
    """SepConvGRU-like chain: 1x5 then 5x1, with a relu between; the zeroed border of layer 1 is layer 2's padding."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 4, 28, 28, generator=g, dtype=torch.float64)
    w1 = torch.randn(6, 4, 1, 5, generator=g, dtype=torch.float64)
    w2 = torch.randn(3, 6, 5, 1, generator=g, dtype=torch.float64)
    want = F.conv2d(F.relu(F.conv2d(x, w1, None, 1, (0, 2))), w2, None, 1, (2, 0)).numpy()
    rows, geom = cg.to_rows(x.numpy(), 2, Wp=32)            # 28 + 2*2 = 32: the pitch the recurrent block would use
    assert rows.shape[0] == 32 * 32
    h = np.maximum(cg.conv_as_gemm(rows, geom, w1.numpy()), 0)
    out = cg.conv_as_gemm(h, geom, w2.numpy())
    np.testing.assert_allclose(cg.from_rows(out, geom), want, rtol=1e-12, atol=1e-12)
    assert cg.tap_offsets(1, 5, 32) == [-2, -1, 0, 1, 2] and cg.tap_offsets(5, 1, 32) == [-64, -32, 0, 32, 64]
