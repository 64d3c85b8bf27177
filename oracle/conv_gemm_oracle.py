"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the implicit-GEMM convolution in csrc/gemm.cu (`cwm_conv2d_f16`,
SURVEY 8f rank 3: the nn.Conv2d calls of cwm/models/raft/update.py:6-14, :16-60, :79-97, :121-137).

Formulation (what the kernel does, step for step):
  * activations are NHWC rows  X[(s, y, x), c]  with a row pitch `ldx` >= C (a convolution may read a column slice);
  * one 128-row M tile = `hb` image rows of `wb` pixel slots, hb = 128 / wb: tile row r is OUTPUT pixel
    (y0 + r // wb, x0 + r % wb) of sample s; maps of <= 32 pixels have one column block (wb = 16 or 32 >= W), the
    encoders' wider maps (extractor.py:118-190: 112 / 56 pixels) are tiled in both directions;
  * stride 2 (the first convolution and the 1x1 shortcut of the encoders' stages 2 and 3): the box starts at the INPUT
    pixel (2 y0 + ky - pad_h, 2 x0 + kx - pad_w) and the tensor map's traversal stride 2 delivers every second pixel;
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


def load_box(x, s, y0, x0, c0, hb, wb, stride=1):
    """The 4-D box of NHWC tensor x at (s, y0, x0, c0) that delivers [hb, wb, 64] elements with zero fill outside the
    tensor; `stride` is the tensor map's traversal stride over the two pixel axes (element i of the box is pixel
    origin + i * stride)."""
    S, H, W, C = x.shape
    box = np.zeros((hb, wb, BK), x.dtype)
    if not (0 <= s < S):
        return box
    ys, xs, cs = y0 + stride * np.arange(hb), x0 + stride * np.arange(wb), np.arange(c0, c0 + BK)
    vy, vx, vc = (ys >= 0) & (ys < H), (xs >= 0) & (xs < W), (cs >= 0) & (cs < C)
    if vy.any() and vx.any() and vc.any():
        box[np.ix_(vy, vx, vc)] = x[s][np.ix_(ys[vy], xs[vx], cs[vc])]
    return box


def tile_shape(Ho, Wo):
    """(hb, wb) of a 128-row tile: one column block for maps of <= 32 pixels, else the wb in {32, 16, 8} with the least
    padded area (ties: the widest) -- csrc/gemm.cu conv2d_impl."""
    if Wo <= 32:
        wb = 16 if Wo <= 16 else 32
        return 128 // wb, wb
    best = None
    for wb in (32, 16, 8):
        hb = 128 // wb
        area = -(-Wo // wb) * wb * -(-Ho // hb) * hb
        if best is None or area < best[0]:
            best = (area, hb, wb)
    return best[1], best[2]


def conv_as_gemm(x, w, bias=None, relu=False, stride=1):
    """x [S, H, W, C] (NHWC), w [Cout, C, kh, kw] -> [S, Ho, Wo, Cout]: the tile / k-step walk of the kernel in float64
    ('same' padding k // 2, stride 1 or 2; Ho = (H - 1) // stride + 1)."""
    S, H, W, C = x.shape
    N, _, kh, kw = w.shape
    pad_h, pad_w = kh // 2, kw // 2
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    hb, wb = tile_shape(Ho, Wo)
    slabs = (C + BK - 1) // BK
    wp = pack_weight(w).astype(np.float64)
    out = np.zeros((S, Ho, Wo, N), np.float64)
    for s in range(S):
        for yt in range(-(-Ho // hb)):
            for xt in range(-(-Wo // wb)):
                y0, x0 = yt * hb, xt * wb
                acc = np.zeros((hb * wb, N), np.float64)
                kb = 0
                for tap in range(kh * kw):
                    for j in range(slabs):
                        a = load_box(x, s, stride * y0 + tap // kw - pad_h, stride * x0 + tap % kw - pad_w, j * BK, hb, wb,
                                     stride).reshape(hb * wb, BK)
                        acc += a.astype(np.float64) @ wp[:, kb * BK:(kb + 1) * BK].T
                        kb += 1
                if bias is not None:
                    acc += bias
                if relu:
                    acc = np.maximum(acc, 0)
                tile = acc.reshape(hb, wb, N)
                ye, xe = min(Ho, y0 + hb), min(Wo, x0 + wb)
                out[s, y0:ye, x0:xe, :] = tile[:ye - y0, :xe - x0]              # the clipped store
    return out


def im2col_nchw(img, k, stride, pad, ldo, scale=1.0, shift=0.0):
    """csrc/norm.cu im2col_nchw_kernel: img [S, C, H, W] -> [S*Ho*Wo, ldo], column (ky*k + kx)*C + c, zero outside the image
    (the normalisation scale * v + shift applies inside only)."""
    S, C, H, W = img.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    v = np.zeros((S, C, H + 2 * pad, W + 2 * pad), np.float64)
    v[:, :, pad:pad + H, pad:pad + W] = img.astype(np.float64) * scale + shift
    out = np.zeros((S, Ho, Wo, ldo), np.float64)
    for ky in range(k):
        for kx in range(k):
            patch = v[:, :, ky:ky + stride * Ho:stride, kx:kx + stride * Wo:stride]          # [S, C, Ho, Wo]
            out[..., (ky * k + kx) * C:(ky * k + kx + 1) * C] = patch.transpose(0, 2, 3, 1)
    return out.reshape(S * Ho * Wo, ldo)


def conv_as_gemm_halo2d(x, w, bias=None, relu=False):
    """The 2-D halo mode of csrc/gemm.cu (stride-1 convolutions on maps wider than 32 pixels: the encoders' 112 / 56 pixel
    maps): a 128-row tile is 4 image rows of 32 pixel slots that keep wo = 32 - 2 pad_w outputs each.  Per channel slab ONE
    box [4 + kh - 1, 32, 64] is staged at (y0 - pad_h, x0 - pad_w); tap (ky, kx) reads the SAME box as 128 consecutive
    flat rows starting ky * 32 + kx (slot c of image row r -> box row r + ky, slot c + kx; slots c >= wo wrap into the next
    box row and are discarded by the store, which writes [4, wo] clipped to the image)."""
    S, H, W, C = x.shape
    N, _, kh, kw = w.shape
    pad_h, pad_w = kh // 2, kw // 2
    hb, wb = 4, 32
    wo = wb - 2 * pad_w
    assert W > 32 and wo >= 16
    slabs = (C + BK - 1) // BK
    wp = pack_weight(w).astype(np.float64)
    out = np.zeros((S, H, W, N), np.float64)
    rows_box = hb + kh - 1
    for s in range(S):
        for yt in range(-(-H // hb)):
            for xt in range(-(-W // wo)):
                y0, x0 = yt * hb, xt * wo
                acc = np.zeros((hb * wb, N), np.float64)
                for j in range(slabs):
                    box = load_box(x, s, y0 - pad_h, x0 - pad_w, j * BK, rows_box, wb).reshape(rows_box * wb, BK)
                    box = np.concatenate([box, np.zeros((wb, BK), box.dtype)], 0)      # reads of discarded rows run past the box
                    for tap in range(kh * kw):
                        start = (tap // kw) * wb + tap % kw
                        a = box[start:start + hb * wb]
                        kb = tap * slabs + j
                        acc += a.astype(np.float64) @ wp[:, kb * BK:(kb + 1) * BK].T
                if bias is not None:
                    acc += bias
                if relu:
                    acc = np.maximum(acc, 0)
                tile = acc.reshape(hb, wb, N)
                ye, xe = min(H, y0 + hb), min(W, x0 + wo)
                out[s, y0:ye, x0:xe, :] = tile[:ye - y0, :xe - x0]
    return out
