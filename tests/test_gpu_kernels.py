"""Stage-wise parity of every CUDA kernel, called through the C ABI, against the CPU oracle (integer stages:
bit-exact) or a plain fp32 torch restatement of the same op (floating-point stages, tolerance stated per test)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vmae_oracle as oracle
from counterfactualworldmodels_b200 import _lib, ops, prediction, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_device_is_sm100():
    assert _lib.load().cwm_device_check() == 0


# ---------------------------------------------------------------- a4 compaction (bit-exact)
@pytest.mark.parametrize("B,N,p", [(1, 1, 0.5), (3, 97, 0.6), (4, 1568, 0.5), (2, 6272, 0.495), (5, 256, 0.0),
                                   (5, 255, 1.0), (64, 1568, 0.5), (2, 6336, 0.3)])
def test_compact_mask_bit_exact(B, N, p):
    rng = np.random.RandomState(B * 1000 + N)
    mask = rng.rand(B, N) < p
    perm_o, inv_o, nvis_o = oracle.compact_mask(mask)
    perm, inv, nvis = ops.compact_mask(torch.from_numpy(mask).to(DEV))
    assert np.array_equal(perm.cpu().numpy(), perm_o)
    assert np.array_equal(inv.cpu().numpy(), inv_o)
    assert np.array_equal(nvis.cpu().numpy(), nvis_o)
    # and against torch itself, the way the reference computes it (vmae.py:167, :555-556)
    mt = torch.from_numpy(mask)
    for b in range(B):
        assert torch.equal(perm[b, :nvis_o[b]].cpu().long(), torch.nonzero(~mt[b]).flatten())
        assert torch.equal(perm[b, nvis_o[b]:].cpu().long(), torch.nonzero(mt[b]).flatten())


def test_compact_mask_empty_batch():
    perm, inv, nvis = ops.compact_mask(torch.zeros(0, 16, dtype=torch.bool, device=DEV))
    assert perm.shape == (0, 16)


# ---------------------------------------------------------------- a1+a2 gather
@pytest.mark.parametrize("cfg,normalize", [("tiny_4x4", True), ("tiny_8x8", False), ("base_8x8", True)])
def test_patch_gather_matches_unfold(cfg, normalize):
    hw = synthetic.image_hw(cfg)
    msize = synthetic.mask_size(cfg)
    ps = synthetic.oracle_cfg(cfg)["patch_size"]
    B = 2
    x_raw = synthetic.make_video(B, hw, seed=3)                      # [B,T,C,H,W]
    mask = synthetic.make_mask(B, msize, num_clumps=2, seed=4)
    perm_o, _, nvis_o = oracle.compact_mask(mask.numpy())
    xin = x_raw.to(DEV).transpose(1, 2)                               # non-contiguous view, like _preprocess
    perm, _, _ = ops.compact_mask(mask.to(DEV))
    norm = (oracle.IMAGENET_DEFAULT_MEAN, oracle.IMAGENET_DEFAULT_STD) if normalize else None
    a = ops.patch_gather(xin, perm, int(nvis_o[0]), ps, input_norm=norm)
    # oracle: preprocess, then im2col in Conv3d weight order (c, kt, kh, kw)
    xp = oracle.preprocess(x_raw, normalize)                          # [B,C,T,H,W]
    pt, ph, pw = ps
    Bc, C, T, H, W = xp.shape
    cols = xp.reshape(B, C, T // pt, pt, H // ph, ph, W // pw, pw).permute(0, 2, 4, 6, 1, 3, 5, 7)
    cols = cols.reshape(B, -1, C * pt * ph * pw)
    want = torch.stack([cols[b, perm_o[b, :nvis_o[b]]] for b in range(B)]).reshape(-1, cols.shape[-1])
    assert torch.equal(a.cpu(), want.to(torch.float16))               # same fp32 ops, then one rounding


# ---------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("M,C", [(7, 128), (1000, 384), (333, 512), (257, 768), (64, 1024)])
def test_layernorm_f16(M, C):
    g = torch.Generator().manual_seed(M + C)
    x = torch.randn(M, C, generator=g) * 3 + 0.5
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g)
    want = F.layer_norm(x, (C,), gamma, beta, 1e-6)
    got = ops.layernorm_f16(x.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-6).cpu().float()
    # tolerance: one f16 rounding of values up to ~|10| (2^-11 relative) + fp32 statistics noise
    assert (got - want).abs().max().item() <= 4e-3 + 1e-3 * want.abs().max().item()
    assert (got - want.to(torch.float16).float()).abs().mean().item() < 1e-4


def test_layernorm_row_gather():
    B, Ntot, Nvis, C = 3, 50, 18, 256
    x = torch.randn(B * Ntot, C)
    gamma, beta = torch.ones(C), torch.zeros(C)
    Nm = Ntot - Nvis
    got = ops.layernorm_f16(x.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-6, M=B * Nm, grp_rows=Nm, grp_stride=Ntot,
                            grp_offset=Nvis).cpu().float()
    want = F.layer_norm(x.view(B, Ntot, C)[:, -Nm:], (C,)).reshape(-1, C)
    assert (got - want).abs().max().item() < 5e-3


# ---------------------------------------------------------------- GEMM + epilogues
def _gemm_ref(a, w):
    return a.float() @ w.float().t()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1, 64, 16), (300, 768, 192), (1000, 2304, 768), (257, 384, 384),
                                   (513, 1152, 384), (130, 48, 512), (129, 192, 384), (1584, 1024, 48),
                                   (777, 3072, 768), (50688, 768, 3072),
                                   # CTA-pair (cta_group::2) tiles with ragged N / M: W rows beyond N are zero-filled
                                   # in the second CTA's half, the last pair tile has rows in one CTA only
                                   (300, 200, 128), (520, 136, 64), (257, 256, 64), (256, 512, 1024)])
def test_gemm_f32_bias(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    a = (torch.randn(M, K, generator=g)).to(torch.float16).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.float16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    got = ops.gemm_f16(a, w, _lib.EPI_F32, bias=bias)
    want = _gemm_ref(a, w) + bias
    # fp32 accumulation of exact f16 products: only summation-order noise
    assert (got - want).abs().max().item() < 2e-3
    assert (got - want).abs().mean().item() < 1e-4


@pytest.mark.parametrize("M,N,K", [(300, 768, 768), (129, 2304, 768), (1000, 1536, 512)])
def test_gemm_f16_scale_cols(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(torch.float16).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.float16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    sc = N // 3
    got = ops.gemm_f16(a, w, _lib.EPI_F16, bias=bias, scale=0.125, scale_cols=sc).float()
    want = _gemm_ref(a, w) + bias
    want[:, :sc] *= 0.125
    assert (got - want).abs().max().item() < 1e-2  # one f16 rounding of |values| < ~8
    assert torch.equal(got.half(), want.half()) or (got - want).abs().mean().item() < 5e-4


@pytest.mark.parametrize("M,N,K", [(300, 3072, 768), (129, 512, 128), (640, 2048, 512)])
def test_gemm_gelu(M, N, K):
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(torch.float16).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.float16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    got = ops.gemm_f16(a, w, _lib.EPI_GELU_F16, bias=bias).float()
    want = F.gelu(_gemm_ref(a, w) + bias)
    assert (got - want).abs().max().item() < 6e-3
    assert (got - want).abs().mean().item() < 3e-4


@pytest.mark.parametrize("M,N,K", [(300, 768, 3072), (257, 384, 1536), (1000, 1024, 1024)])
def test_gemm_residual_inplace(M, N, K):
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g).to(torch.float16).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.float16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    x = torch.randn(M, N, generator=g).to(DEV)
    want = x + _gemm_ref(a, w) + bias
    got = ops.gemm_f16(a, w, _lib.EPI_RES_F32, bias=bias, res=x, out=x)
    assert got.data_ptr() == x.data_ptr()
    assert (got - want).abs().max().item() < 2e-3


def test_gemm_gathered_residual_and_row_remap():
    """The two special uses: patch-embed (+pos[perm]) and encoder_to_decoder written into the decoder sequence."""
    B, Nvis, Ntot, K, N = 3, 37, 80, 192, 256
    g = torch.Generator().manual_seed(5)
    a = torch.randn(B * Nvis, K, generator=g).to(torch.float16).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.float16).to(DEV)
    pos = torch.randn(Ntot, N, generator=g).to(DEV)
    perm = torch.stack([torch.randperm(Ntot, generator=g) for _ in range(B)]).to(torch.int32).to(DEV)
    acc = _gemm_ref(a, w).view(B, Nvis, N)
    want_rows = acc + pos[perm[:, :Nvis].long()]
    # (1) compact output
    got = ops.gemm_f16(a, w, _lib.EPI_RES_F32, res=pos, res_gather=perm, gather_stride=Ntot, grp_rows=Nvis,
                       grp_out_stride=Nvis)
    assert (got.view(B, Nvis, N) - want_rows).abs().max().item() < 2e-3
    # (2) scattered into [B, Ntot, N]; untouched rows keep their content
    out = torch.full((B * Ntot, N), -7.0, device=DEV)
    ops.gemm_f16(a, w, _lib.EPI_RES_F32, res=pos, res_gather=perm, gather_stride=Ntot, grp_rows=Nvis,
                 grp_out_stride=Ntot, out=out)
    out = out.view(B, Ntot, N)
    assert (out[:, :Nvis] - want_rows).abs().max().item() < 2e-3
    assert bool((out[:, Nvis:] == -7.0).all())


# ---------------------------------------------------------------- attention
def _attn_ref(qkv, B, N, H):
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    p = (q @ k.transpose(-2, -1)).softmax(-1)
    return (p @ v).transpose(1, 2).reshape(B * N, H * 64)


@pytest.mark.parametrize("B,N,H", [(1, 128, 1), (2, 256, 2), (1, 100, 2), (2, 456, 4), (1, 792, 12), (2, 896, 2),
                                   (1, 1568, 6), (1, 3168, 2), (1, 129, 1), (3, 385, 1),
                                   # more work items than SMs: every persistent CTA pipelines across several items,
                                   # with and without a half (<= 128-row) last q block, and single-tile sequences
                                   (40, 300, 4), (30, 788, 12), (200, 100, 2), (64, 130, 6), (9, 520, 8)])
def test_attention_matches_fp32_softmax(B, N, H):
    g = torch.Generator().manual_seed(B * 100 + N + H)
    qkv = torch.randn(B * N, 3 * H * 64, generator=g)
    qkv[:, :H * 64] *= 0.125 * 3.0   # pre-scaled q, with enough spread that the running max moves
    qkv = qkv.to(torch.float16).to(DEV)
    got = ops.attention_f16(qkv, B, N, H).float()
    want = _attn_ref(qkv, B, N, H)
    # f16 P and V operands, fp32 accumulate: ~1e-3 absolute on O(1) outputs
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() < 4e-3
    assert (got - want).abs().mean().item() < 3e-4


@pytest.mark.parametrize("B,N,H", [(64, 256, 6), (64, 130, 6), (30, 788, 12), (16, 1568, 6), (64, 384, 6)])
def test_attention_stress_repeated(B, N, H):
    """Race detector: more CTAs than SMs, logits with a realistic spread (the lazy-rescale path runs), 15 repetitions
    per shape.  The first attention kernel (single mbarrier per hand-off) failed this on every repetition at
    N = 788 / 1568 while passing every parity test."""
    g = torch.Generator().manual_seed(B * 100 + N + H)
    qkv = torch.randn(B * N, 3 * H * 64, generator=g)
    qkv[:, :H * 64] *= 0.125 * 3.0
    qkv = qkv.to(torch.float16).to(DEV)
    want = _attn_ref(qkv, B, N, H)
    first = None
    for rep in range(15):
        got = ops.attention_f16(qkv, B, N, H)
        assert (got.float() - want).abs().max().item() < 4e-3, rep
        if first is None:
            first = got.clone()
        assert torch.equal(got, first), rep     # and run-to-run deterministic


def test_attention_large_logits_rescale_path():
    """Rows whose maximum keeps growing across KV tiles exercise the lazy O-rescale."""
    B, N, H = 1, 640, 1
    g = torch.Generator().manual_seed(9)
    qkv = torch.randn(B * N, 3 * 64, generator=g)
    ramp = torch.linspace(0.2, 4.0, N)[:, None]
    qkv[:, 64:128] *= ramp           # keys grow with position -> later tiles dominate
    qkv = qkv.to(torch.float16).to(DEV)
    got = ops.attention_f16(qkv, B, N, H).float()
    want = _attn_ref(qkv, B, N, H)
    assert (got - want).abs().max().item() < 6e-3


# ---------------------------------------------------------------- a10 / a12
def test_fill_mask_tokens():
    B, Ntot, Nvis, C = 2, 64, 20, 128
    g = torch.Generator().manual_seed(1)
    mt, pos = torch.randn(C, generator=g).to(DEV), torch.randn(Ntot, C, generator=g).to(DEV)
    perm = torch.stack([torch.randperm(Ntot, generator=g) for _ in range(B)]).to(torch.int32).to(DEV)
    x = torch.full((B, Ntot, C), 3.0, device=DEV)
    ops.fill_mask_tokens(mt, pos, perm, Nvis, x)
    assert bool((x[:, :Nvis] == 3.0).all())
    assert torch.equal(x[:, Nvis:], mt + pos[perm[:, Nvis:].long()])


@pytest.mark.parametrize("cfg", ["tiny_4x4", "tiny_8x8", "base_8x8"])
def test_unpatchify_scatter_bit_exact(cfg):
    hw, msize = synthetic.image_hw(cfg), synthetic.mask_size(cfg)
    ps = synthetic.oracle_cfg(cfg)["patch_size"]
    B = 2
    x = synthetic.make_video(B, hw, seed=8)
    mask = synthetic.make_mask(B, msize, num_clumps=3, seed=9)
    D = 3 * ps[0] * ps[1] * ps[2]
    y = torch.randn(B, int(mask[0].sum()), D)
    want = oracle.pred_patches_to_video(y, x, mask, ps)
    _, inv, nvis = ops.compact_mask(mask.to(DEV))
    got = prediction.unpatchify_scatter(y.to(DEV), x.to(DEV), inv, int(nvis[0]), ps)
    assert torch.equal(got.cpu(), want)   # pure data movement: bit-exact


# ---------------------------------------------------------------- config 5 kernels (a13-a16)
@pytest.mark.parametrize("M,C", [(50, 64), (33, 192), (7, 4), (129, 96)])
def test_layernorm_generic_widths(M, C):
    g = torch.Generator().manual_seed(M * 7 + C)
    x = torch.randn(M, C, generator=g) * 2 - 0.3
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g)
    want = F.layer_norm(x, (C,), gamma, beta, 1e-6)
    got = ops.layernorm_f16(x.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-6).cpu().float()
    assert (got - want).abs().max().item() <= 4e-3 + 1e-3 * want.abs().max().item()


def _attn_ref3(q, k, v):
    """fp32 softmax(q k^T) v on f16-rounded operands; q,k,v [B,H,N,d]."""
    s = q.float() @ k.float().transpose(-1, -2)
    return s.softmax(-1) @ v.float()


@pytest.mark.parametrize("B,Nq,Nk,H,d", [
    (2, 25, 25, 12, 32), (3, 50, 50, 6, 32), (2, 26, 26, 4, 32), (1, 1, 1, 12, 32),   # context self-attention
    (2, 788, 25, 4, 192), (2, 1600, 50, 4, 96), (1, 70, 5, 4, 64), (2, 100, 1, 4, 192),  # cross, trg direction
    (2, 25, 3140, 4, 192), (2, 50, 6336, 4, 96), (1, 5, 72, 4, 32), (2, 1, 784, 4, 192),  # cross, src direction
    (1, 300, 300, 2, 128),
    # the tile-walking trg kernel (Nq >= 512, one resident K / V tile) at the model sizes and at ragged / minimal ones
    (1, 3140, 25, 4, 192), (1, 6336, 50, 4, 96), (2, 512, 64, 4, 128), (3, 1000, 33, 2, 64), (1, 777, 7, 3, 32),
    (40, 513, 32, 4, 96),
])
def test_attention_generic(B, Nq, Nk, H, d):
    g = torch.Generator().manual_seed(Nq * 31 + Nk)
    q = (torch.randn(B, Nq, H, d, generator=g) * d ** -0.25).to(torch.float16)
    k = (torch.randn(B, Nk, H, d, generator=g) * d ** -0.25).to(torch.float16)
    v = torch.randn(B, Nk, H, d, generator=g).to(torch.float16)
    want = _attn_ref3(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3))   # [B,H,Nq,d]
    want = want.permute(0, 2, 1, 3).reshape(B * Nq, H * d)
    got = ops.attention_generic_f16(q.reshape(B * Nq, H * d).to(DEV), k.reshape(B * Nk, H * d).to(DEV),
                                    v.reshape(B * Nk, H * d).to(DEV), B, Nq, Nk, H, d).cpu().float()
    # P is rounded to f16 for the PV product (2^-11 relative per term), output rounded to f16
    err = (got - want).abs()
    assert err.max().item() <= 6e-3 and err.mean().item() <= 6e-4, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("N", [300, 1500])
def test_attention_generic_cross_layout(N):
    """The interleaved [qk | v] layout of BidirectionalCrossAttention (transformer.py:333-365): head h of qk owns
    columns [2*hd*h, 2*hd*(h+1)), first hd for the trg similarity, last hd for the src similarity."""
    B, M, H, hd = 2, 25, 4, 96
    D = H * hd
    g = torch.Generator().manual_seed(5)
    qkv = (torch.randn(B * N, 3 * D, generator=g) * 0.3).to(torch.float16)
    qkv_s = (torch.randn(B * M, 3 * D, generator=g) * 0.3).to(torch.float16)
    qk = qkv[:, :2 * D].reshape(B, N, H, 2 * hd).permute(0, 2, 1, 3)
    qk_s = qkv_s[:, :2 * D].reshape(B, M, H, 2 * hd).permute(0, 2, 1, 3)
    v = qkv[:, 2 * D:].reshape(B, N, H, hd).permute(0, 2, 1, 3)
    v_s = qkv_s[:, 2 * D:].reshape(B, M, H, hd).permute(0, 2, 1, 3)
    y = _attn_ref3(qk[..., :hd], qk_s[..., :hd], v_s).permute(0, 2, 1, 3).reshape(B * N, D)
    y_s = _attn_ref3(qk_s[..., hd:], qk[..., hd:], v).permute(0, 2, 1, 3).reshape(B * M, D)
    qd, sd = qkv.to(DEV), qkv_s.to(DEV)
    got = ops.attention_generic_f16(qd, sd, sd[:, 2 * D:], B, N, M, H, hd, 2 * hd, 2 * hd, hd).cpu().float()
    got_s = ops.attention_generic_f16(sd[:, hd:], qd[:, hd:], qd[:, 2 * D:], B, M, N, H, hd, 2 * hd, 2 * hd,
                                      hd).cpu().float()
    assert (got - y).abs().max().item() <= 6e-3
    assert (got_s - y_s).abs().max().item() <= 6e-3


def test_patch_gather_imu_scalar_and_pad_tokens():
    """IMU "video" [B, 6, 400, 1, 1] with a (16,1,1) tubelet (conjoined_vmae.py:1013-1038): scalar gather path; token
    ids >= 25 are padding positions and give zero rows (conjoined_vmae.py:130-133)."""
    B, C, L, pt = 3, 6, 400, 16
    g = torch.Generator().manual_seed(9)
    imu = torch.randn(B, C, L, generator=g)
    perm = torch.stack([torch.randperm(25 + 5, generator=g) for _ in range(B)]).to(torch.int32)
    rows = 12
    a = ops.patch_gather(imu.to(DEV)[..., None, None], perm.to(DEV), rows, (pt, 1, 1)).cpu()
    tok = imu.reshape(B, C, L // pt, pt).permute(0, 2, 1, 3).reshape(B, L // pt, C * pt)   # (c, kt) order
    for b in range(B):
        for j in range(rows):
            t = int(perm[b, j])
            want = tok[b, t].to(torch.float16) if t < 25 else torch.zeros(C * pt, dtype=torch.float16)
            assert torch.equal(a[b * rows + j], want)


def test_fill_pad_rows():
    B, rows, C, first_pad = 3, 10, 64, 20
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, rows, C, generator=g)
    perm = torch.stack([torch.randperm(28, generator=g) for _ in range(B)]).to(torch.int32)
    val = torch.randn(C, generator=g)
    for value in (val, None):
        for off in (0, 7):
            got = ops.fill_pad_rows(x.clone().to(DEV), perm.to(DEV), off, first_pad,
                                    None if value is None else value.to(DEV)).cpu()
            want = x.clone()
            sel = perm[:, off:off + rows] >= first_pad
            want[sel] = 0.0 if value is None else value
            assert torch.equal(got, want)


def test_gemm_cta_pair_equals_single_cta():
    """The cta_group::2 kernels (two CTAs share a 256-row tile, each stages half of W) give bit-identical results to
    the single-CTA kernels: same MMA shapes along K, same accumulation order."""
    import ctypes
    lib = _lib.load()
    lib.cwm_debug_gemm_cta2.argtypes = [ctypes.c_int]
    g = torch.Generator().manual_seed(5)
    try:
        for (M, N, K, mode) in [(1000, 2304, 768, _lib.EPI_F16), (777, 3072, 768, _lib.EPI_GELU_F16),
                                (1300, 768, 3072, _lib.EPI_RES_F32), (513, 1152, 384, _lib.EPI_F32),
                                (50432, 768, 768, _lib.EPI_RES_F32), (40000, 1536, 384, _lib.EPI_GELU_F16)]:
            a = torch.randn(M, K, generator=g).to(torch.float16).to(DEV)
            w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.float16).to(DEV)
            bias = torch.randn(N, generator=g).to(DEV)
            res = torch.randn(M, N, generator=g).to(DEV) if mode == _lib.EPI_RES_F32 else None
            def run(flag):
                lib.cwm_debug_gemm_cta2(flag)
                return ops.gemm_f16(a, w, mode, bias=bias, scale=0.125, scale_cols=N // 3 if mode == _lib.EPI_F16 else 0,
                                    res=None if res is None else res.clone()).clone()
            ref = run(0)
            for rep in range(15):  # repeated: a missing hand-off in the pair protocol would show up as a sporadic mismatch
                assert torch.equal(run(1), ref), (M, N, K, mode, rep)
    finally:
        lib.cwm_debug_gemm_cta2(1)


@pytest.mark.parametrize("M,C,N2", [(1000, 768, 2304), (777, 384, 1536), (300, 1024, 4096)])
def test_layernorm_folded_into_gemm_epilogues(M, C, N2):
    """The LayerNorm-free block path at the kernel level: a residual GEMM (producer) emits f16 rows + partial row
    statistics, the next GEMM (consumer, weights folded with gamma) normalises in its epilogue.  Checked against
    LayerNorm -> Linear in fp32; and the row-statistics kernel (first block of a stream) against the producer's planes."""
    import ctypes
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + C)
    K1 = 256
    a = torch.randn(M, K1, generator=g).to(torch.float16).to(DEV)
    w1 = (torch.randn(C, K1, generator=g) / K1 ** 0.5).to(torch.float16).to(DEV)
    b1 = torch.randn(C, generator=g).to(DEV)
    x0 = (torch.randn(M, C, generator=g) * 2 + 0.7).to(DEV)          # residual with a non-zero mean
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV)
    w2 = (torch.randn(N2, C, generator=g) / C ** 0.5).to(DEV)
    b2 = torch.randn(N2, generator=g).to(DEV)
    eps = 1e-6
    # ---- producer: x = x0 + a w1^T + b1, plus f16 copy and statistics planes ----
    parts = lib.cwm_gemm_ln_parts(C)
    x = x0.clone()
    x16 = torch.empty(M, C, dtype=torch.float16, device=DEV)
    stats = torch.full((parts, M, 2), float("nan"), device=DEV)
    e = _lib.GemmEpilogue()
    e.mode, e.bias, e.res, e.ldr, e.out, e.ldo = _lib.EPI_RES_F32, b1.data_ptr(), x.data_ptr(), C, x.data_ptr(), C
    e.ln_x16, e.ln_ldx16, e.ln_stats_out = x16.data_ptr(), C, stats.data_ptr()
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.cwm_gemm_f16(a.data_ptr(), w1.data_ptr(), M, C, K1, ctypes.byref(e), stream))
    want_x = x0 + a.float() @ w1.float().t() + b1
    assert (x - want_x).abs().max().item() < 2e-3
    assert torch.equal(x16, x.to(torch.float16))
    s = stats.sum(0)
    assert torch.allclose(s[:, 0], x.sum(1), rtol=1e-5, atol=1e-2) and torch.allclose(s[:, 1], (x * x).sum(1), rtol=1e-5, atol=1e-2)
    # ---- row-statistics kernel: same information in one plane ----
    x16b = torch.empty_like(x16)
    st1 = torch.empty(M, 2, device=DEV)
    _lib.check(lib.cwm_rowstats_f16(x.data_ptr(), M, C, x16b.data_ptr(), st1.data_ptr(), stream))
    assert torch.equal(x16b, x16) and torch.allclose(st1, s, rtol=1e-5, atol=1e-2)
    # ---- consumer: GELU(LN(x) w2^T + b2) with the LayerNorm folded into weights + epilogue ----
    w2f = (w2 * gamma[None, :]).to(torch.float16).contiguous()
    colsum = w2f.float().sum(1).contiguous()
    c = (w2 @ beta + b2).contiguous()
    for stats_in, n_parts in ((stats, parts), (st1, 1)):
        out = torch.empty(M, N2, dtype=torch.float16, device=DEV)
        e = _lib.GemmEpilogue()
        e.mode, e.bias, e.out, e.ldo = _lib.EPI_GELU_F16, c.data_ptr(), out.data_ptr(), N2
        e.ln_stats_in, e.ln_parts, e.ln_colsum, e.ln_width, e.ln_eps = stats_in.data_ptr(), n_parts, colsum.data_ptr(), C, eps
        _lib.check(lib.cwm_gemm_f16(x16.data_ptr(), w2f.data_ptr(), M, N2, C, ctypes.byref(e), stream))
        want = F.gelu(F.layer_norm(x, (C,), gamma, beta, eps) @ w2.t() + b2)
        err = (out.float() - want).abs()
        assert err.max().item() < 2e-2 and err.mean().item() < 1.5e-3, (err.max().item(), err.mean().item())
