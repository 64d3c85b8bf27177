"""RAFT's convolutions as implicit GEMMs (csrc/gemm.cu `cwm_conv2d_f16`).

CPU: the numpy restatement of the kernel's tile / k-step walk (4-D boxes with zero fill, packed weights) against torch's
conv2d for every kernel shape of the recurrent block.  GPU: the CUDA kernel against conv2d of the same f16-rounded
operands -- fp32 accumulation of f16 products, so the bar is the f16 rounding of the OUTPUT (2^-11 relative) plus the
accumulation-order noise of K up to 1920 terms."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import conv_gemm_oracle as cg

DEV = "cuda:0"


@pytest.mark.parametrize("kh,kw", [(1, 1), (3, 3), (1, 5), (5, 1), (7, 7)])
@pytest.mark.parametrize("H,W,C", [(7, 9, 8), (28, 28, 72), (5, 20, 130)])
def test_convolution_equals_the_box_walk(kh, kw, H, W, C):
    g = torch.Generator().manual_seed(kh * 10 + kw + H)
    S, N = 2, 5
    x = torch.randn(S, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(N, C, kh, kw, generator=g, dtype=torch.float64)
    b = torch.randn(N, generator=g, dtype=torch.float64)
    want = F.conv2d(x, w, b, 1, (kh // 2, kw // 2)).permute(0, 2, 3, 1).numpy()
    got = cg.conv_as_gemm(x.permute(0, 2, 3, 1).numpy(), w.numpy(), b.numpy())
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("kh,stride,H,W,C", [(3, 1, 40, 56, 24), (3, 2, 56, 56, 64), (1, 2, 56, 112, 96), (3, 2, 37, 45, 70),
                                             (3, 1, 20, 112, 8), (3, 2, 18, 36, 8)])
def test_strided_and_wide_convolutions_equal_the_box_walk(kh, stride, H, W, C):
    """The encoders' shapes: 2-D tiles for maps wider than 32 pixels, traversal stride 2 (odd sizes included)."""
    g = torch.Generator().manual_seed(kh * 10 + stride + H)
    S, N = 2, 3
    x = torch.randn(S, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(N, C, kh, kh, generator=g, dtype=torch.float64)
    b = torch.randn(N, generator=g, dtype=torch.float64)
    want = F.conv2d(x, w, b, stride, kh // 2).permute(0, 2, 3, 1).numpy()
    got = cg.conv_as_gemm(x.permute(0, 2, 3, 1).numpy(), w.numpy(), b.numpy(), stride=stride)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("kh,kw,H,W,C", [(3, 3, 10, 56, 24), (3, 3, 7, 112, 8), (1, 5, 9, 45, 70), (5, 1, 6, 33, 8), (7, 7, 9, 40, 8)])
def test_2d_halo_tiles_equal_the_convolution(kh, kw, H, W, C):
    """The 2-D halo mode (4 x 32-slot tiles keeping 32 - 2 pad outputs per row, every tap a shifted read of one staged box)."""
    g = torch.Generator().manual_seed(kh * 10 + kw + W)
    S, N = 2, 3
    x = torch.randn(S, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(N, C, kh, kw, generator=g, dtype=torch.float64)
    b = torch.randn(N, generator=g, dtype=torch.float64)
    want = F.conv2d(x, w, b, 1, (kh // 2, kw // 2)).permute(0, 2, 3, 1).numpy()
    got = cg.conv_as_gemm_halo2d(x.permute(0, 2, 3, 1).numpy(), w.numpy(), b.numpy())
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-11)


def test_stem_im2col_plus_gemm_equals_the_strided_convolution():
    g = torch.Generator().manual_seed(5)
    img = torch.rand(2, 3, 30, 38, generator=g, dtype=torch.float64) * 255
    w = torch.randn(4, 3, 7, 7, generator=g, dtype=torch.float64)
    want = F.conv2d(2 * (img / 255) - 1, w, None, 2, 3).permute(0, 2, 3, 1).reshape(-1, 4).numpy()
    cols = cg.im2col_nchw(img.numpy(), 7, 2, 3, 152, scale=2 / 255, shift=-1.0)
    wk = np.zeros((4, 152))
    wk[:, :147] = w.permute(0, 2, 3, 1).reshape(4, -1).numpy()
    np.testing.assert_allclose(cols @ wk.T, want, rtol=1e-10, atol=1e-10)


def test_packed_weight_layout_and_relu():
    w = np.arange(2 * 3 * 1 * 5, dtype=np.float64).reshape(2, 3, 1, 5)
    p = cg.pack_weight(w)
    assert p.shape == (2, 5 * 64) and p[1, 2 * 64 + 1] == w[1, 1, 0, 2] and p[0, 3:64].max() == 0
    x = -np.ones((1, 4, 4, 3))
    assert cg.conv_as_gemm(x, np.abs(w), relu=True).max() == 0


def _conv_gpu(x_rows, S, H, W, w, bias, relu, ldo=None, out=None):
    from counterfactualworldmodels_b200 import ops
    return ops.conv2d_f16(x_rows, S, H, W, w, bias=bias, relu=relu, ldo=ldo, out=out)


@pytest.mark.gpu
@pytest.mark.parametrize("name,Cin,Cout,kh,kw", [
    ("convc1", 328, 256, 1, 1), ("convc2", 256, 192, 3, 3), ("convf2", 128, 64, 3, 3), ("conv", 256, 128, 3, 3),
    ("convz1|r1", 384, 256, 1, 5), ("convq2", 384, 128, 5, 1), ("flow_head.conv1|mask.0", 128, 512, 3, 3),
    ("flow_head.conv2", 256, 8, 3, 3), ("mask.2", 256, 576, 1, 1), ("7x7", 64, 128, 7, 7)])
@pytest.mark.parametrize("S,H,W", [(3, 28, 28), (2, 16, 16), (1, 9, 30), (5, 20, 12), (67, 28, 28)])
def test_gpu_conv2d_matches_torch(name, Cin, Cout, kh, kw, S, H, W):
    g = torch.Generator().manual_seed(Cin + Cout + kh + S)
    x = (torch.randn(S, H, W, Cin, generator=g) * 0.7).half()
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5).half()
    b = torch.randn(Cout, generator=g)
    want = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, 1, (kh // 2, kw // 2)).permute(0, 2, 3, 1)
    got = _conv_gpu(x.reshape(-1, Cin).to(DEV), S, H, W, w.to(DEV), b.to(DEV), False).cpu().float().view(S, H, W, Cout)
    err = (got - want).abs().max().item()
    assert err <= 2e-3 * max(1.0, want.abs().max().item()), (name, err)
    got_relu = _conv_gpu(x.reshape(-1, Cin).to(DEV), S, H, W, w.to(DEV), None, True).cpu().float().view(S, H, W, Cout)
    want_relu = F.relu(want - b)
    assert (got_relu - want_relu).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())


@pytest.mark.gpu
def test_gpu_conv2d_reads_and_writes_column_slices():
    """The recurrent block's buffers: a convolution reads a slice of a wider row buffer and writes into a slice of the
    next convolution's input (HX = [h | inp | motion]); the neighbouring columns must stay untouched."""
    g = torch.Generator().manual_seed(5)
    S, H, W = 2, 28, 28
    buf = (torch.randn(S * H * W, 384, generator=g) * 0.5).half().to(DEV)
    w = (torch.randn(64, 128, 3, 3, generator=g) / 34).half().to(DEV)
    dst = torch.full((S * H * W, 256), 7.0, dtype=torch.float16, device=DEV)
    _conv_gpu(buf[:, 128:256], S, H, W, w, None, False, out=dst[:, 192:])
    want = F.conv2d(buf[:, 128:256].float().view(S, H, W, 128).permute(0, 3, 1, 2), w.float(), None, 1, 1).permute(0, 2, 3, 1)
    assert (dst[:, 192:].float().view(S, H, W, 64) - want).abs().max().item() <= 2e-3 * want.abs().max().item()
    assert bool((dst[:, :192] == 7.0).all())


@pytest.mark.gpu
def test_gpu_im2col_flow_and_7x7_as_one_gemm():
    from counterfactualworldmodels_b200 import ops
    g = torch.Generator().manual_seed(9)
    S, H, W = 2, 28, 28
    flow = torch.zeros(S * H * W, 8, dtype=torch.float16)
    flow[:, :2] = (torch.randn(S * H * W, 2, generator=g) * 3).half()
    w = (torch.randn(128, 2, 7, 7, generator=g) / 10).half()
    cols = ops.raft_im2col_flow(flow.to(DEV), S, H, W, 7, 128)
    want_cols = F.unfold(flow[:, :2].float().view(S, H, W, 2).permute(0, 3, 1, 2), 7, padding=3)   # [S, 2*49, HW], (c, tap)
    want_cols = want_cols.view(S, 2, 49, H * W).permute(0, 3, 2, 1).reshape(S * H * W, 98)
    assert torch.equal(cols[:, :98].cpu().float(), want_cols) and float(cols[:, 98:].abs().max()) == 0
    wk = torch.zeros(128, 128, dtype=torch.float16)
    wk[:, :98] = w.permute(0, 2, 3, 1).reshape(128, 98)
    got = _conv_gpu(cols, S, H, W, wk.view(128, 128, 1, 1).to(DEV), None, False).cpu().float().view(S, H, W, 128)
    want = F.conv2d(flow[:, :2].float().view(S, H, W, 2).permute(0, 3, 1, 2), w.float(), None, 1, 3).permute(0, 2, 3, 1)
    assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item()


@pytest.mark.gpu
@pytest.mark.parametrize("kh,kw", [(1, 5), (5, 1), (3, 3)])
@pytest.mark.parametrize("S,H,W", [(3, 28, 28), (2, 16, 16)])
def test_gpu_gru_gate_and_update_in_the_conv_epilogues(kh, kw, S, H, W):
    """One ConvGRU half-step (update.py:43-60) as two convolutions with the gate arithmetic in their epilogues, against
    the same step in torch fp32 on the f16-rounded operands."""
    from counterfactualworldmodels_b200 import _lib, ops
    lib = _lib.load()
    g = torch.Generator().manual_seed(kh * 7 + kw + S)
    C, M = 128, S * H * W
    HX = (torch.randn(M, 3 * C, generator=g) * 0.6).half()
    w_zr = (torch.randn(2 * C, 3 * C, kh, kw, generator=g) / (3 * C * kh * kw) ** 0.5 * 2).half()
    w_q = (torch.randn(C, 3 * C, kh, kw, generator=g) / (3 * C * kh * kw) ** 0.5 * 2).half()
    b_zr, b_q = torch.randn(2 * C, generator=g) * 0.3, torch.randn(C, generator=g) * 0.3
    pad = (kh // 2, kw // 2)
    nchw = lambda rows: rows.float().view(S, H, W, -1).permute(0, 3, 1, 2)          # noqa: E731
    rows = lambda t: t.permute(0, 2, 3, 1).reshape(M, -1)                            # noqa: E731
    h = HX[:, :C].float()
    zr = torch.sigmoid(rows(F.conv2d(nchw(HX), w_zr.float(), b_zr, 1, pad)))
    z, r = zr[:, :C], zr[:, C:]
    RHX = torch.cat([(r * h).half(), HX[:, C:]], 1)
    q = torch.tanh(rows(F.conv2d(nchw(RHX), w_q.float(), b_q, 1, pad)))
    h_new = (1 - z.half().float()) * h + z.half().float() * q

    d = lambda t: t.to(DEV).contiguous()                                             # noqa: E731
    HXd, RHXd = d(HX), d(torch.cat([torch.zeros(M, C).half(), HX[:, C:]], 1))
    Zd = torch.empty(M, C, dtype=torch.float16, device=DEV)
    Hd = torch.empty(M, C, dtype=torch.float16, device=DEV)
    pzr, pq, bzr, bq = ops.pack_conv_weight(d(w_zr)), ops.pack_conv_weight(d(w_q)), d(b_zr), d(b_q)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.cwm_conv2d_gru_gate_f16(HXd.data_ptr(), 3 * C, S, H, W, 3 * C, pzr.data_ptr(), C, kh, kw, pad[0], pad[1],
                                           bzr.data_ptr(), HXd.data_ptr(), 3 * C, Zd.data_ptr(), C, RHXd.data_ptr(), 3 * C, st))
    assert (Zd.cpu().float() - z).abs().max().item() <= 2e-3
    assert (RHXd[:, :C].cpu().float() - r * h).abs().max().item() <= 4e-3
    assert torch.equal(RHXd[:, C:].cpu(), HX[:, C:])                                 # the other slots are untouched
    _lib.check(lib.cwm_conv2d_gru_update_f16(RHXd.data_ptr(), 3 * C, S, H, W, 3 * C, pq.data_ptr(), C, kh, kw, pad[0], pad[1],
                                             bq.data_ptr(), Zd.data_ptr(), C, HXd.data_ptr(), 3 * C, Hd.data_ptr(), st))
    got = HXd[:, :C].cpu().float()
    assert (got - h_new).abs().max().item() <= 8e-3, (got - h_new).abs().max().item()
    assert torch.equal(Hd.cpu(), HXd[:, :C].cpu()) and torch.equal(HXd[:, C:].cpu(), HX[:, C:])


@pytest.mark.gpu
@pytest.mark.parametrize("name,Cin,Cout,k,stride,S,H,W", [
    ("layer1 3x3", 64, 64, 3, 1, 3, 112, 112), ("layer2 3x3/2", 64, 96, 3, 2, 3, 112, 112), ("layer2 1x1/2", 64, 96, 1, 2, 2, 112, 112),
    ("layer2 3x3", 96, 96, 3, 1, 5, 56, 56), ("layer3 3x3/2", 96, 128, 3, 2, 5, 56, 56), ("layer3 1x1/2", 96, 128, 1, 2, 2, 56, 56),
    ("layer3 3x3", 128, 128, 3, 1, 3, 28, 28), ("conv2 1x1", 128, 256, 1, 1, 3, 28, 28),
    ("odd sizes /2", 72, 40, 3, 2, 2, 37, 45), ("odd sizes", 24, 136, 3, 1, 2, 40, 50), ("1x1/2 odd", 16, 64, 1, 2, 1, 19, 75),
    ("layer1, many tiles per CTA", 64, 64, 3, 1, 9, 112, 112), ("resident weights, ragged", 64, 64, 3, 1, 3, 50, 61),
    ("5x5 wide", 64, 32, 5, 1, 2, 33, 70), ("two N tiles", 64, 320, 3, 1, 1, 20, 40)])
def test_gpu_encoder_convolutions_match_torch(name, Cin, Cout, k, stride, S, H, W):
    """RAFT's encoder convolutions (extractor.py:6-56, :118-190): wide maps tiled in both directions, stride 2 through the
    tensor map's traversal stride."""
    from counterfactualworldmodels_b200 import ops
    g = torch.Generator().manual_seed(Cin + Cout + k + S)
    x = (torch.randn(S, H, W, Cin, generator=g) * 0.7).half()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).half()
    b = torch.randn(Cout, generator=g)
    want = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride, k // 2).permute(0, 2, 3, 1)
    got = ops.conv2d_f16(x.reshape(-1, Cin).to(DEV), S, H, W, w.to(DEV), bias=b.to(DEV), relu=False, stride=stride)
    got = got.cpu().float().view(want.shape)
    err = (got - want).abs().max().item()
    assert err <= 2e-3 * max(1.0, want.abs().max().item()), (name, err)
    got_relu = ops.conv2d_f16(x.reshape(-1, Cin).to(DEV), S, H, W, w.to(DEV), relu=True, stride=stride).cpu().float().view(want.shape)
    assert (got_relu - F.relu(want - b)).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("S,H,W", [(2, 224, 224), (3, 30, 38), (1, 17, 65)])
def test_gpu_stem_im2col_matches_the_restatement(S, H, W):
    from counterfactualworldmodels_b200 import ops
    g = torch.Generator().manual_seed(S + H)
    img = torch.rand(S, 3, H, W, generator=g) * 255
    got = ops.im2col_nchw_f16(img.to(DEV), 7, 2, 3, 152, scale=2 / 255, shift=-1.0).cpu().float().numpy()
    want = cg.im2col_nchw(img.numpy(), 7, 2, 3, 152, scale=2 / 255, shift=-1.0)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-3)     # f16 rounding of values in [-1, 1]
    assert np.array_equal(got == 0, np.abs(want) < 3e-8) or np.abs(got[np.abs(want) < 3e-8]).max() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("S,H,W", [(2, 28, 28), (3, 16, 16), (1, 9, 30)])
def test_gpu_dual_destination_convolution_with_a_tail(S, H, W):
    """`cwm_conv2d_dual_f16`: the motion encoder's last convolution writes its 126 features + the 2 flow columns into both
    GRU input buffers; everything in front stays untouched."""
    from counterfactualworldmodels_b200 import _lib, ops
    g = torch.Generator().manual_seed(S + H)
    M = S * H * W
    x = (torch.randn(M, 256, generator=g) * 0.5).half().to(DEV)
    w = torch.zeros(128, 256, 3, 3)
    w[:126] = torch.randn(126, 256, 3, 3, generator=g) / 48
    b = torch.zeros(128)
    b[:126] = torch.randn(126, generator=g)
    flow16 = torch.zeros(M, 8, dtype=torch.float16)
    flow16[:, :2] = (torch.randn(M, 2, generator=g) * 4).half()
    HX = torch.full((M, 384), 7.0, dtype=torch.float16, device=DEV)
    RHX = torch.full((M, 384), -3.0, dtype=torch.float16, device=DEV)
    pk = ops.pack_conv_weight(w.half().to(DEV))
    bd, fd = b.to(DEV), flow16.to(DEV)
    _lib.check(_lib.load().cwm_conv2d_dual_f16(x.data_ptr(), 256, S, H, W, 256, pk.data_ptr(), 128, 3, 3, 1, 1, bd.data_ptr(), 1,
                                               fd.data_ptr(), 8, HX[:, 256:].data_ptr(), 384, RHX[:, 256:].data_ptr(), 384, None))
    want = F.relu(F.conv2d(x.float().cpu().view(S, H, W, 256).permute(0, 3, 1, 2), w.half().float(), b, 1, 1)).permute(0, 2, 3, 1)
    want = want.reshape(-1, 128)[:, :126]
    for buf, fill in ((HX, 7.0), (RHX, -3.0)):
        got = buf.cpu().float()
        assert (got[:, 256:382] - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
        assert bool((got[:, :256] == fill).all()) and torch.equal(got[:, 382:], flow16[:, :2].float())


@pytest.mark.gpu
@pytest.mark.parametrize("S,H,W", [(2, 28, 28), (3, 16, 16), (1, 9, 30)])
def test_gpu_flow_head_as_tap_products_plus_stencil(S, H, W):
    """The flow head's 3x3 / 256 -> 2 convolution as a 1x1 GEMM to 18 per-tap products + the stencil sum inside
    `cwm_raft_flow_update_taps`, against conv2d; the new flow lands in flow16 and in both GRU input rows."""
    from counterfactualworldmodels_b200 import _lib, ops
    g = torch.Generator().manual_seed(S * 7 + W)
    M = S * H * W
    x = torch.relu(torch.randn(M, 256, generator=g)).half()
    w = (torch.randn(2, 256, 3, 3, generator=g) / 48).half()
    b = torch.randn(2, generator=g) * 0.1
    coords = torch.randn(S, 2, H, W, generator=g) * 5
    want_c = coords + F.conv2d(x.float().view(S, H, W, 256).permute(0, 3, 1, 2), w.float(), b, 1, 1)
    wt = torch.zeros(24, 256, 1, 1, dtype=torch.float16)
    wt[:18, :, 0, 0] = w.permute(2, 3, 0, 1).reshape(18, 256)
    taps = ops.conv2d_f16(x.to(DEV), S, H, W, wt.to(DEV))
    c1 = coords.clone().to(DEV)
    flow16 = torch.zeros(M, 8, dtype=torch.float16, device=DEV)
    HX = torch.full((M, 384), 7.0, dtype=torch.float16, device=DEV)
    RHX = torch.full((M, 384), -3.0, dtype=torch.float16, device=DEV)
    bd = b.to(DEV)
    _lib.check(_lib.load().cwm_raft_flow_update_taps(taps.data_ptr(), taps.stride(0), bd.data_ptr(), c1.data_ptr(), S, H, W,
                                                     flow16.data_ptr(), HX[:, 382:].data_ptr(), 384, RHX[:, 382:].data_ptr(), 384,
                                                     None))   # (the two extra destinations are optional; raft.py passes NULL)
    scale = (want_c - coords).abs().max().item()
    assert (c1.cpu() - want_c).abs().max().item() <= 3e-3 * scale
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    want_flow = (want_c - torch.stack([xs, ys])[None]).permute(0, 2, 3, 1).reshape(M, 2)
    for got in (flow16[:, :2], HX[:, 382:], RHX[:, 382:]):
        assert (got.cpu().float() - want_flow).abs().max().item() <= 3e-3 * scale + 2e-3 * want_flow.abs().max().item()
    assert bool((HX[:, :382] == 7.0).all()) and bool((RHX[:, :382] == -3.0).all()) and float(flow16[:, 2:].abs().max()) == 0


@pytest.mark.gpu
def test_gpu_stem_im2col_reads_frame_slices_in_place():
    """The stem's im2col takes one frame of a [S, T, 3, H, W] movie through its sample stride (no copy) and writes into a
    row slice of a larger buffer: RAFT's fused encoders fold the input normalisation 2 x - 1 into it this way."""
    from counterfactualworldmodels_b200 import ops
    g = torch.Generator().manual_seed(3)
    movie = torch.rand(3, 2, 3, 40, 56, generator=g)
    dev = movie.to(DEV)
    rows = 20 * 28
    cols = torch.full((4 * rows, 152), 9.0, dtype=torch.float16, device=DEV)
    ops.im2col_nchw_f16(dev[:1, 0], 7, 2, 3, 152, scale=2.0, shift=-1.0, out=cols[:rows])
    ops.im2col_nchw_f16(dev[:, 1], 7, 2, 3, 152, scale=2.0, shift=-1.0, out=cols[rows:])
    want = np.concatenate([cg.im2col_nchw(movie[:1, 0].numpy(), 7, 2, 3, 152, 2.0, -1.0),
                           cg.im2col_nchw(movie[:, 1].numpy(), 7, 2, 3, 152, 2.0, -1.0)], 0)
    np.testing.assert_allclose(cols.cpu().float().numpy(), want, rtol=0, atol=1e-3)
