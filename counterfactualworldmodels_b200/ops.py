"""Thin Python wrappers over the individual C-ABI entry points (one call = one kernel launch on the current
stream).  The model forward does not use these -- it goes through ``cwm_vmae_forward`` in one call -- they exist
for the stage-wise parity tests and for users who want a single fused op."""
import ctypes

import torch

from . import _lib
from .vmae import compact_mask  # noqa: F401  (re-export)


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _req_cuda(*ts):
    for t in ts:
        if t is not None and t.device.type != "cuda":
            raise RuntimeError("libcwm_b200 ops need CUDA (B200) tensors; there is no CPU fallback")


def patch_gather(x, perm, rows_per_sample, patch_size, input_norm=None):
    """x [B,C,T,H,W] fp32 (any strides), perm int32 [B,Ntot] -> f16 [B*rows_per_sample, C*pt*ph*pw]."""
    _req_cuda(x, perm)
    lib = _lib.load()
    B, C, T, H, W = x.shape
    pt, ph, pw = patch_size
    out = torch.empty(B * rows_per_sample, C * pt * ph * pw, dtype=torch.float16, device=x.device)
    mean = std = None
    if input_norm is not None:
        mean, std = _lib.float_array(input_norm[0]), _lib.float_array(input_norm[1])
    _lib.check(lib.cwm_patch_gather(x.data_ptr(), _lib.strides5(x), B, C, T, H, W, pt, ph, pw, perm.data_ptr(),
                                    perm.shape[1], rows_per_sample, mean, std, out.data_ptr(), _stream(x)))
    return out


def layernorm_f16(x, gamma, beta, eps, M=None, grp_rows=0, grp_stride=0, grp_offset=0):
    """x fp32 [rows, C] -> f16 [M, C] (M defaults to rows; optional row gather, see cwm_b200.h)."""
    _req_cuda(x, gamma, beta)
    lib = _lib.load()
    C = x.shape[-1]
    M = x.shape[0] if M is None else M
    out = torch.empty(M, C, dtype=torch.float16, device=x.device)
    _lib.check(lib.cwm_layernorm_f16(x.data_ptr(), M, C, gamma.data_ptr(), beta.data_ptr(), float(eps), grp_rows,
                                     grp_stride, grp_offset, out.data_ptr(), _stream(x)))
    return out


def gemm_f16(a, w, mode, bias=None, scale=1.0, scale_cols=0, res=None, res_gather=None, gather_stride=0,
             grp_rows=0, grp_out_stride=0, out=None, out_rows=None):
    """epilogue(a[M,K] @ w[N,K]^T); a, w f16 contiguous.  Returns the output tensor (f16 or fp32 by mode)."""
    _req_cuda(a, w, bias, res, res_gather, out)
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.is_contiguous() and w.is_contiguous()
    f16_out = mode in (_lib.EPI_F16, _lib.EPI_GELU_F16)
    if out is None:
        rows = M if out_rows is None else out_rows
        out = torch.zeros(rows, N, dtype=torch.float16 if f16_out else torch.float32, device=a.device)
    e = _lib.GemmEpilogue()
    e.mode = mode
    e.bias = bias.data_ptr() if bias is not None else None
    e.scale, e.scale_cols = float(scale), int(scale_cols)
    e.res = res.data_ptr() if res is not None else None
    e.ldr = res.shape[-1] if res is not None else 0
    e.res_gather = res_gather.data_ptr() if res_gather is not None else None
    e.gather_stride, e.grp_rows, e.grp_out_stride = int(gather_stride), int(grp_rows), int(grp_out_stride)
    e.out, e.ldo = out.data_ptr(), out.shape[-1]
    _lib.check(lib.cwm_gemm_f16(a.data_ptr(), w.data_ptr(), M, N, K, ctypes.byref(e), _stream(a)))
    return out


def attention_f16(qkv, B, N, H):
    """qkv f16 [B*N, 3*H*64] (q pre-scaled) -> f16 [B*N, H*64]."""
    _req_cuda(qkv)
    lib = _lib.load()
    out = torch.empty(B * N, H * 64, dtype=torch.float16, device=qkv.device)
    _lib.check(lib.cwm_attention_f16(qkv.data_ptr(), B, N, H, 64, out.data_ptr(), _stream(qkv)))
    return out


def fill_mask_tokens(mask_token, pos, perm, n_vis, x_full):
    """x_full [B, Ntot, C] fp32: rows n_vis.. of every sample <- mask_token + pos[perm]."""
    _req_cuda(mask_token, pos, perm, x_full)
    lib = _lib.load()
    B, Ntot, C = x_full.shape
    _lib.check(lib.cwm_fill_mask_tokens(mask_token.data_ptr(), pos.data_ptr(), perm.data_ptr(), B, Ntot, n_vis, C,
                                        x_full.data_ptr(), _stream(x_full)))
    return x_full


def attention_generic_f16(q, k, v, B, Nq, Nk, H, head_dim, q_head_stride=None, k_head_stride=None,
                          v_head_stride=None):
    """softmax(q k^T) v with independently strided operands (see cwm_b200.h).  q, k, v: 2-D f16 *views* whose row
    stride is ``.stride(0)`` and whose first element is head 0 (q pre-scaled) -> f16 [B*Nq, H*head_dim]."""
    _req_cuda(q, k, v)
    lib = _lib.load()
    out = torch.empty(B * Nq, H * head_dim, dtype=torch.float16, device=q.device)
    nbytes = lib.cwm_attention_generic_workspace_bytes(B, Nq, Nk, H, head_dim)
    ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=q.device)
    hs = [head_dim if s is None else int(s) for s in (q_head_stride, k_head_stride, v_head_stride)]
    _lib.check(lib.cwm_attention_generic_f16(q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0), k.stride(0),
                                             v.stride(0), hs[0], hs[1], hs[2], B, Nq, Nk, H, head_dim,
                                             out.data_ptr(), out.shape[1], ws.data_ptr(), ws.numel(), _stream(q)))
    return out


def fill_pad_rows(x, perm, perm_offset, first_pad_token, value=None):
    """x fp32 [B, rows, C] in place: rows whose token id perm[b, perm_offset + j] >= first_pad_token <- value / 0."""
    _req_cuda(x, perm, value)
    lib = _lib.load()
    B, rows, C = x.shape
    _lib.check(lib.cwm_fill_pad_rows(x.data_ptr(), B, rows, C, perm.data_ptr(), perm.shape[1], int(perm_offset),
                                     int(first_pad_token), value.data_ptr() if value is not None else None,
                                     _stream(x)))
    return x


def pack_conv_weight(w):
    """nn.Conv2d weight [Cout, Cin, kh, kw] -> f16 [Cout, kh*kw*cin_pad] (tap-major, channels padded to 64 with zeros):
    the layout ``cwm_conv2d_f16`` reads (include/cwm_b200.h)."""
    Cout, Cin, kh, kw = w.shape
    cin_pad = (Cin + 63) // 64 * 64
    p = torch.zeros(Cout, kh, kw, cin_pad, dtype=torch.float16, device=w.device)
    p[..., :Cin] = w.detach().permute(0, 2, 3, 1).to(torch.float16)
    return p.reshape(Cout, kh * kw * cin_pad).contiguous()


def conv2d_f16(x_rows, S, H, W, weight, bias=None, relu=False, ldo=None, out=None, packed=None, stride=1):
    """'Same'-padded convolution (stride 1 or 2) of NHWC f16 rows ``x_rows [S*H*W, Cin]`` (may be a column slice of a wider
    buffer) with an nn.Conv2d weight ``[Cout, Cin, kh, kw]`` on the tcgen05 implicit-GEMM kernel -> f16 rows
    ``[S*Ho*Wo, Cout]`` (``out``: optional destination, may be a column slice)."""
    _req_cuda(x_rows, weight, bias, out)
    lib = _lib.load()
    Cout, Cin, kh, kw = weight.shape
    assert x_rows.dtype == torch.float16 and x_rows.shape == (S * H * W, Cin) and x_rows.stride(1) == 1, x_rows.shape
    if packed is None:
        packed = pack_conv_weight(weight)
    assert packed.shape[1] == lib.cwm_conv2d_weight_k(Cin, kh, kw)
    Mo = S * ((H - 1) // stride + 1) * ((W - 1) // stride + 1)
    if out is None:
        out = torch.empty(Mo, ldo or Cout, dtype=torch.float16, device=x_rows.device)[:, :Cout]
    assert out.dtype == torch.float16 and out.shape == (Mo, Cout) and out.stride(1) == 1
    args = (None if bias is None else bias.data_ptr(), int(bool(relu)), out.data_ptr(), out.stride(0), _stream(x_rows))
    if stride == 1 and W <= 32:
        _lib.check(lib.cwm_conv2d_f16(x_rows.data_ptr(), x_rows.stride(0), S, H, W, Cin, packed.data_ptr(), Cout, kh, kw,
                                      kh // 2, kw // 2, *args))
    else:
        _lib.check(lib.cwm_conv2d_strided_f16(x_rows.data_ptr(), x_rows.stride(0), S, H, W, Cin, packed.data_ptr(), Cout, kh,
                                              kw, kh // 2, kw // 2, stride, *args))
    return out


def im2col_nchw_f16(img, k, stride, pad, ldo, scale=1.0, shift=0.0, out=None):
    """fp32 NCHW image ``[S, Cin, H, W]`` -> f16 rows ``[S*Ho*Wo, ldo]``: the k x k neighbourhood of every output pixel of a
    strided convolution in (ky, kx, c) order, ``scale * v + shift`` applied inside the image, zeros outside.  ``img`` may be
    a frame slice of a movie (any sample stride, each sample's [Cin, H, W] block contiguous); ``out``: optional destination
    rows (e.g. a slice of a larger buffer that holds several images)."""
    _req_cuda(img, out)
    assert img.dtype == torch.float32 and img.dim() == 4
    S, Cin, H, W = img.shape
    assert img.stride()[1:] == (H * W, W, 1), "each sample's [Cin, H, W] block must be contiguous"
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    if out is None:
        out = torch.empty(S * Ho * Wo, ldo, dtype=torch.float16, device=img.device)
    assert out.shape == (S * Ho * Wo, ldo) and out.is_contiguous() and out.dtype == torch.float16
    _lib.check(_lib.load().cwm_im2col_nchw_f16(img.data_ptr(), img.stride(0) if S > 1 else 0, S, Cin, H, W, k, stride, pad,
                                               float(scale), float(shift), out.data_ptr(), ldo, _stream(img)))
    return out


def raft_im2col_flow(flow16, S, H, W, k, ldo):
    """flow16 f16 [S*H*W, >= 2] -> f16 [S*H*W, ldo]: the k x k neighbourhood of the 2 flow channels, (ky, kx, c) order."""
    _req_cuda(flow16)
    out = torch.empty(S * H * W, ldo, dtype=torch.float16, device=flow16.device)
    _lib.check(_lib.load().cwm_raft_im2col_flow(flow16.data_ptr(), flow16.stride(0), S, H, W, k, out.data_ptr(), ldo,
                                                _stream(flow16)))
    return out
