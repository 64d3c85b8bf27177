"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the implicit-GEMM convolution in csrc/gemm.cu (`cwm_conv2d_f16`,
SURVEY 8f rank 3: the nn.Conv2d calls of cwm/models/raft/update.py:6-14, :16-60, :79-97, :121-137).

Formulation (what the kernel does, step for step):
  * activations are NHWC rows  X[(s, y, x), c]  with a row pitch `ldx` >= C (a convolution may read a column slice);
  * one 128-row M tile = `hb` image rows of `wb` pixel slots, wb = 16 or 32 >= W, hb = 128 / wb: tile row r is pixel
    (y0 + r // wb, r % wb) of sample s;
  * k-step (tap t = ky*kw + kx, channel slab j) loads the A tile as ONE box [hb, wb, 64] of the image at
    (y0 + ky - pad_h, kx - pad_w, 64 j): every element outside [0,H) x [0,W) x [0,C) is ZERO (the TMA unit's
    out-of-bounds fill) -- that is the convolution's zero padding, and nothing is ever im2col-ed;
  * weights are packed [Cout, kh, kw, cin_pad], cin_pad = ceil(C / 64) * 64, zero in the padding;
  * the store writes the box [rows, wb, 64] back clipped to [0,H) x [0,W) x [0,Cout).
`tests/test_conv_as_gemm.py` checks this against torch.nn.functional.conv2d for RAFT's kernel shapes on CPU and the CUDA
kernel against conv2d of the same f16-rounded operands on the GPU.
"""
import numpy as np

BK = 64


def pack_weight(w):
    """w [Cout, C, kh, kw] -> [Cout, kh*kw*cin_pad]: K index = ((ky*kw + kx) * cin_pad + c)."""
    N, C, kh, kw = w.shape
    cin_pad = (C + BK - 1) // BK * BK
    out = np.zeros((N, kh, kw, cin_pad), w.dtype)
    out[..., :C] = w.transpose(0, 2, 3, 1)
    return out.reshape(N, kh * kw * cin_pad)


def load_box(x, s, y0, x0, c0, hb, wb):
    """The 4-D box [hb, wb, 64] of NHWC tensor x at (s, y0, x0, c0) with zero fill outside the tensor."""
    S, H, W, C = x.shape
    box = np.zeros((hb, wb, BK), x.dtype)
    if not (0 <= s < S):
        return box
    ys, xs, cs = np.arange(y0, y0 + hb), np.arange(x0, x0 + wb), np.arange(c0, c0 + BK)
    vy, vx, vc = (ys >= 0) & (ys < H), (xs >= 0) & (xs < W), (cs >= 0) & (cs < C)
    if vy.any() and vx.any() and vc.any():
        box[np.ix_(vy, vx, vc)] = x[s][np.ix_(ys[vy], xs[vx], cs[vc])]
    return box


def conv_as_gemm(x, w, bias=None, relu=False):
    """x [S, H, W, C] (NHWC), w [Cout, C, kh, kw] -> [S, H, W, Cout]: the tile / k-step walk of the kernel in float64."""
    S, H, W, C = x.shape
    N, _, kh, kw = w.shape
    pad_h, pad_w = kh // 2, kw // 2
    wb = 16 if W <= 16 else 32
    assert W <= 32
    hb = 128 // wb
    slabs = (C + BK - 1) // BK
    wp = pack_weight(w).astype(np.float64)
    out = np.zeros((S, H, W, N), np.float64)
    tiles = (H + hb - 1) // hb
    for s in range(S):
        for yt in range(tiles):
            y0 = yt * hb
            acc = np.zeros((hb * wb, N), np.float64)
            kb = 0
            for tap in range(kh * kw):
                for j in range(slabs):
                    a = load_box(x, s, y0 + tap // kw - pad_h, tap % kw - pad_w, j * BK, hb, wb).reshape(hb * wb, BK)
                    acc += a.astype(np.float64) @ wp[:, kb * BK:(kb + 1) * BK].T
                    kb += 1
            if bias is not None:
                acc += bias
            if relu:
                acc = np.maximum(acc, 0)
            tile = acc.reshape(hb, wb, N)
            ye = min(H, y0 + hb)
            out[s, y0:ye, :, :] = tile[:ye - y0, :W]              # the clipped store
    return out
