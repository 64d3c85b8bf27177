"""TEST INFRASTRUCTURE ONLY -- the formulation DESIGN.md section 10 (1) proposes for RAFT's convolutions on the tcgen05
GEMM, restated in numpy so that the index arithmetic is pinned before any kernel exists:

  * activations live as zero-padded pixel-major rows  X[(s, y, x), c],  y in [0, H + 2P), x in [0, Wp), Wp >= W + 2P
    (the valid pixels sit at y, x in [P, P + H) x [P, P + W); every other row is zero);
  * a kh x kw convolution (stride 1, 'same' padding, kh//2, kw//2 <= P) is ONE GEMM  out[m, n] = sum_k A[m, k] W[n, k]
    with K = kh*kw*C, where K-slab (tap t = ky*kw + kx, channels c0..c0+63) of the virtual A is the plain 2-D tile
    X[m + off_t, c0:c0+64],  off_t = (ky - kh//2) * Wp + (kx - kw//2)  -- a TMA load at a shifted row coordinate (rows
    before the first / after the last row read as zero, which is TMA's out-of-bounds fill);
  * the epilogue writes zeros to the border rows, so the output is again a valid zero-padded buffer for the next layer.

``tests/test_conv_as_gemm.py`` checks this against ``torch.nn.functional.conv2d`` for RAFT's kernel shapes (1x1, 3x3, 1x5,
5x1, 7x7) and for two chained layers.
"""
import numpy as np


def to_rows(x, P, Wp=None):
    """x [S, C, H, W] -> zero-padded pixel-major rows [S * (H + 2P) * Wp, C]."""
    S, C, H, W = x.shape
    Wp = Wp or (W + 2 * P)
    buf = np.zeros((S, H + 2 * P, Wp, C), x.dtype)
    buf[:, P:P + H, P:P + W] = x.transpose(0, 2, 3, 1)
    return buf.reshape(-1, C), (S, H, W, P, Wp)


def from_rows(rows, geom):
    S, H, W, P, Wp = geom
    return rows.reshape(S, H + 2 * P, Wp, -1)[:, P:P + H, P:P + W].transpose(0, 3, 1, 2)


def pack_weight(w):
    """w [N, C, kh, kw] -> [N, kh*kw*C]: K index = (ky*kw + kx) * C + c, matching the slab order of the virtual A."""
    N, C, kh, kw = w.shape
    return w.transpose(0, 2, 3, 1).reshape(N, kh * kw * C)


def tap_offsets(kh, kw, Wp):
    return [(ky - kh // 2) * Wp + (kx - kw // 2) for ky in range(kh) for kx in range(kw)]


def border_mask(geom):
    """True for the rows the epilogue must write as zero."""
    S, H, W, P, Wp = geom
    m = np.ones((S, H + 2 * P, Wp), bool)
    m[:, P:P + H, P:P + W] = False
    return m.reshape(-1)


def conv_as_gemm(rows, geom, w, bias=None):
    """One GEMM over the virtual A (tap-shifted row tiles, zero fill outside the buffer), border rows zeroed."""
    N, C, kh, kw = w.shape
    S, H, W, P, Wp = geom
    assert kh // 2 <= P and kw // 2 <= P and rows.shape[1] == C
    M = rows.shape[0]
    wp = pack_weight(w)
    out = np.zeros((M, N), np.float64)
    for t, off in enumerate(tap_offsets(kh, kw, Wp)):
        shifted = np.zeros_like(rows)                     # rows m + off, zero where that leaves the buffer (TMA OOB fill)
        lo, hi = max(0, -off), min(M, M - off)
        shifted[lo:hi] = rows[lo + off:hi + off]
        out += shifted.astype(np.float64) @ wp[:, t * C:(t + 1) * C].astype(np.float64).T
    if bias is not None:
        out += bias
    out[border_mask(geom)] = 0
    return out.astype(rows.dtype)
