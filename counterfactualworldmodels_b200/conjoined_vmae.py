"""Drop-in replacements for the padded / conjoined (IMU-conditioned) predictors of
``cwm/models/VideoMAE/conjoined_vmae.py`` (BASELINE config 5):

  PaddedVisionTransformer            (:24-210)    learnable null tokens pad rows with fewer visible patches
  ConjoinedPretrainVisionTransformer (:212-887)   two token streams joined by cross-attention blocks
  ConjoinedPaddedVisionTransformer   (:889-1011)
  ImuEncoder                         (:1013-1147) IMU sequence [B, 6, L] -> L/16 tokens
  imu400_base_4x4patch_2frames_1tube (:1230-1243), imu400_8x8patch_2frames_1tube_flowbackrgb01 (:1218-1228)

Same constructor arguments, ``state_dict`` keys, ``forward`` signatures and the stateful attributes the reference
wrappers rely on (``padding_mask`` / ``null_mask`` / ``full_input_mask``, ``_reset_padding_mask``, attribute
forwarding to ``main_stream``).  The modules only hold parameters; the arithmetic runs in
libcwm_b200 (include/cwm_b200.h): per stream the patch gather + embedding GEMM, one ``cwm_block_forward`` per
transformer block, one ``cwm_cross_block_forward`` per conjoining block, the decoder-input assembly, head GEMM and
the null-row fills.  No PyTorch or CPU fallback.  Reference citations are relative to /root/reference.
"""
import copy
import ctypes
from functools import partial

import numpy as np
import os

import torch
import torch.nn as nn

from . import _lib
from . import preprocessor as preproc
from .transformer import CrossAttentionTransformerBlock, pos_embedding
from .vmae import (PretrainVisionTransformer, PretrainVisionTransformerEncoder, _Packer, _cfg, compact_mask)


# ---------------------------------------------------------------------------------------------------------
# device-side state of one token stream
# ---------------------------------------------------------------------------------------------------------
class _StreamEngine:
    """f16 weight copies, block arrays and (null-token extended) positional tables of one PretrainVisionTransformer
    -like stream.  Rebuilt whenever a parameter's storage / version or a positional table changes."""

    def __init__(self):
        self.signature = None

    def ensure(self, m, device):
        sig = (str(device), id(m.encoder.pos_embed), id(m.pos_embed)) + \
            tuple((p.data_ptr(), p._version) for p in m.parameters())
        if sig != self.signature:
            self._build(m, device)
            self.signature = sig
        return self

    def _build(self, m, device):
        pk = self.pk = _Packer(device)
        enc, dec = m.encoder, m.decoder
        pe = enc.patch_embed
        self.Ce, self.Cd = enc.embed_dim, dec.embed_dim
        self.K = pe.proj.weight[0].numel()
        self.w_patch = pk.f16(pe.proj.weight.reshape(pe.proj.out_channels, -1))  # (c, kt, kh, kw) flattening
        self.b_patch = pk.f32(pe.proj.bias)
        fold = os.environ.get("CWM_FUSE_LN", "1") != "0"  # second LayerNorm of every block folded into the fc1 epilogue
        self.enc_blocks = pk.block_array(enc.blocks, fold_ln=fold)
        self.dec_blocks = pk.block_array(dec.blocks, fold_ln=fold)

        def dims(blocks, C):
            if len(blocks) == 0:
                return (1, C, C, 1.0)
            a = blocks[0].attn
            return (a.num_heads, a.head_dim, blocks[0].mlp.fc1.out_features, float(a.scale))
        self.enc_dims, self.dec_dims = dims(enc.blocks, self.Ce), dims(dec.blocks, self.Cd)
        self.ln_eps = float(enc.norm.eps)
        self.enc_norm = (pk.f32(enc.norm.weight), pk.f32(enc.norm.bias))
        self.w_e2d = pk.f16(m.encoder_to_decoder.weight)
        self.mask_token = pk.f32(m.mask_token.reshape(-1))
        self.dec_norm = (pk.f32(dec.norm.weight), pk.f32(dec.norm.bias))
        self.w_head, self.b_head = pk.f16(dec.head.weight), pk.f32(dec.head.bias)
        self.out_dim = dec.head.out_features
        # positional tables, extended by the null-token rows of a PaddedVisionTransformer: token ids >= n_tokens are
        # padding positions.  Encoder pads carry null_token_enc *instead of* embedding + position
        # (conjoined_vmae.py:130-133); decoder pads carry null_token_dec as their "position" (:156-160).
        P = int(getattr(m, 'max_padding_tokens', 0) or 0) if hasattr(m, 'null_token_enc') else 0
        pos_e = enc.pos_embed[0].detach().float().cpu()
        pos_d = m.pos_embed[0].detach().float().cpu()
        assert pos_e.shape[0] == pos_d.shape[0], (pos_e.shape, pos_d.shape)
        self.n_tokens, self.P = pos_e.shape[0], P
        self.null_enc = None
        if P > 0:
            pos_e = torch.cat([pos_e, torch.zeros(P, self.Ce)], 0)
            pos_d = torch.cat([pos_d, m.null_token_dec.detach().float().cpu().reshape(1, -1).expand(P, -1)], 0)
            self.null_enc = pk.f32(m.null_token_enc.reshape(-1))
        self.pos_enc, self.pos_dec = pk.f32(pos_e), pk.f32(pos_d)


def _pack_cross_block(pk, blk):
    ca = blk.cross_attention
    w = _lib.CrossBlockWeights()
    w.ln1_g, w.ln1_b = pk.f32(blk.norm1_cross.weight), pk.f32(blk.norm1_cross.bias)
    w.ln1s_g, w.ln1s_b = pk.f32(blk.norm1_src_cross.weight), pk.f32(blk.norm1_src_cross.bias)
    w.w_qkv = pk.f16(torch.cat([ca.qk.weight.detach(), ca.v.weight.detach()], 0))
    w.w_qkv_s = pk.f16(torch.cat([ca.qk_src.weight.detach(), ca.v_src.weight.detach()], 0))
    w.w_proj, w.b_proj = pk.f16(ca.projection.weight), pk.f32(ca.projection.bias)
    w.w_proj_s, w.b_proj_s = pk.f16(ca.projection_src.weight), pk.f32(ca.projection_src.bias)
    w.ln2_g, w.ln2_b = pk.f32(blk.norm2.weight), pk.f32(blk.norm2.bias)
    w.ln2s_g, w.ln2s_b = pk.f32(blk.norm2_src.weight), pk.f32(blk.norm2_src.bias)
    t, s = blk.mlp['trg'].layers, blk.mlp['src'].layers
    w.w_fc1, w.b_fc1, w.w_fc2, w.b_fc2 = pk.f16(t[0].weight), pk.f32(t[0].bias), pk.f16(t[2].weight), pk.f32(t[2].bias)
    w.w_fc1_s, w.b_fc1_s = pk.f16(s[0].weight), pk.f32(s[0].bias)
    w.w_fc2_s, w.b_fc2_s = pk.f16(s[2].weight), pk.f32(s[2].bias)
    pk.keep.append(w)
    geom = dict(C=ca.in_dim, Cs=ca.in_dim_src, heads=ca.num_heads, head_dim=ca.head_dim,
                hidden=t[0].out_features, hidden_s=s[0].out_features, eps=float(blk.norm2.eps), scale=float(ca.scale))
    return w, geom


class _Workspace:
    """One growable scratch buffer shared by the block-level calls of a forward (they run back to back on one stream)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _gemm(lib, a_ptr, w_ptr, M, N, K, stream, **epi):
    e = _lib.GemmEpilogue()
    for k, v in epi.items():
        setattr(e, k, v)
    _lib.check(lib.cwm_gemm_f16(a_ptr, w_ptr, M, N, K, ctypes.byref(e), stream))


def _tokenize_visible(se, x5, patch_size, perm, n_vis, B, stream, input_norm=None):
    """a1-a4 for one stream: gather the visible patches (padding positions give zero rows), embed, add the
    positional embedding of each gathered token, then overwrite padding rows with null_token_enc
    (vmae.py:140-150; conjoined_vmae.py:125-134).  -> fp32 [B * n_vis, Ce]."""
    lib = _lib.load()
    dev = x5.device
    Bx, C, T, H, W = x5.shape
    pt, ph, pw = patch_size
    Next = perm.shape[1]
    xe = torch.empty(B * n_vis, se.Ce, dtype=torch.float32, device=dev)
    if n_vis == 0 or B == 0:
        return xe
    a16 = torch.empty(B * n_vis, se.K, dtype=torch.float16, device=dev)
    if x5.dtype != torch.float32:
        x5 = x5.float()
    mean = std = None
    if input_norm is not None:  # imagenet_normalize of the raw frames (prediction.py:309-310) fused into the gather
        mean, std = _lib.float_array(input_norm[0]), _lib.float_array(input_norm[1])
    _lib.check(lib.cwm_patch_gather(x5.data_ptr(), _lib.strides5(x5), B, C, T, H, W, pt, ph, pw, perm.data_ptr(), Next,
                                    n_vis, mean, std, a16.data_ptr(), stream))
    _gemm(lib, a16.data_ptr(), se.w_patch, B * n_vis, se.Ce, se.K, stream, mode=_lib.EPI_RES_F32, bias=se.b_patch,
          res=se.pos_enc, ldr=se.Ce, res_gather=perm.data_ptr(), gather_stride=Next, grp_rows=n_vis,
          grp_out_stride=n_vis, out=xe.data_ptr(), ldo=se.Ce)
    if se.P > 0:
        _lib.check(lib.cwm_fill_pad_rows(xe.data_ptr(), B, n_vis, se.Ce, perm.data_ptr(), Next, 0, se.n_tokens,
                                         se.null_enc, stream))
    return xe


def _run_block(ws, arr, i, x, B, N, C, dims, eps, stream):
    lib = _lib.load()
    heads, head_dim, hidden, scale = dims
    nbytes = lib.cwm_block_workspace_bytes(B, N, C, heads, head_dim, hidden)
    buf = ws.get(nbytes, x.device)
    _lib.check(lib.cwm_block_forward(ctypes.byref(arr[i]), x.data_ptr(), B, N, C, heads, head_dim, hidden, eps, scale,
                                     buf.data_ptr(), buf.numel(), stream))


def _run_cross(ws, w, g, x, src, B, N, M, stream):
    lib = _lib.load()
    nbytes = lib.cwm_cross_block_workspace_bytes(B, N, M, g['C'], g['Cs'], g['heads'], g['head_dim'], g['hidden'],
                                                 g['hidden_s'])
    buf = ws.get(nbytes, x.device)
    _lib.check(lib.cwm_cross_block_forward(ctypes.byref(w), x.data_ptr(), src.data_ptr(), B, N, M, g['C'], g['Cs'],
                                           g['heads'], g['head_dim'], g['hidden'], g['hidden_s'], g['eps'], g['scale'],
                                           buf.data_ptr(), buf.numel(), stream))


def _to_decoder(se, xe, perm, n_vis, B, stream):
    """a8-a10: encoder norm, encoder_to_decoder (no bias) written at the visible rows of the decoder sequence with the
    (null-extended) positional embedding of each token added; mask_token + position on the masked rows
    (vmae.py:547-557; conjoined_vmae.py:154-165, :969-975).  -> fp32 [B * Next, Cd]."""
    lib = _lib.load()
    dev = xe.device
    Next = perm.shape[1]
    xd = torch.empty(B * Next, se.Cd, dtype=torch.float32, device=dev)
    if B == 0:
        return xd
    if n_vis > 0:
        a16 = torch.empty(B * n_vis, se.Ce, dtype=torch.float16, device=dev)
        _lib.check(lib.cwm_layernorm_f16(xe.data_ptr(), B * n_vis, se.Ce, se.enc_norm[0], se.enc_norm[1], se.ln_eps, 0, 0,
                                         0, a16.data_ptr(), stream))
        _gemm(lib, a16.data_ptr(), se.w_e2d, B * n_vis, se.Cd, se.Ce, stream, mode=_lib.EPI_RES_F32, res=se.pos_dec,
              ldr=se.Cd, res_gather=perm.data_ptr(), gather_stride=Next, grp_rows=n_vis, grp_out_stride=Next,
              out=xd.data_ptr(), ldo=se.Cd)
    _lib.check(lib.cwm_fill_mask_tokens(se.mask_token, se.pos_dec, perm.data_ptr(), B, Next, n_vis, se.Cd, xd.data_ptr(),
                                        stream))
    return xd


def _last_tokens(se, xd, perm, n_vis, B, n_ret, zero_pads, stream):
    """`get_last_tokens` (vmae.py:238-244): head(norm(x[:, -n_ret:])); with ``zero_pads`` the rows of masked padding
    positions are zeroed (conjoined_vmae.py:207-208, :998-1002).  -> fp32 [B, n_ret, D]."""
    lib = _lib.load()
    dev = xd.device
    Next = perm.shape[1]
    y = torch.empty(B, n_ret, se.out_dim, dtype=torch.float32, device=dev)
    if B == 0 or n_ret == 0:
        return y
    a16 = torch.empty(B * n_ret, se.Cd, dtype=torch.float16, device=dev)
    _lib.check(lib.cwm_layernorm_f16(xd.data_ptr(), B * n_ret, se.Cd, se.dec_norm[0], se.dec_norm[1], se.ln_eps, n_ret,
                                     Next, Next - n_ret, a16.data_ptr(), stream))
    _gemm(lib, a16.data_ptr(), se.w_head, B * n_ret, se.out_dim, se.Cd, stream, mode=_lib.EPI_F32, bias=se.b_head,
          out=y.data_ptr(), ldo=se.out_dim)
    if zero_pads and se.P > 0:
        _lib.check(lib.cwm_fill_pad_rows(y.data_ptr(), B, n_ret, se.out_dim, perm.data_ptr(), Next, Next - n_ret,
                                         se.n_tokens, None, stream))
    return y


def _visible_count(nvis, B, what):
    """Host read of the per-row visible count (the reference syncs at the same place in `x[~mask]`)."""
    counts = nvis.cpu()
    n = int(counts[0]) if B > 0 else 0
    if B > 0 and not bool((counts == n).all()):
        raise RuntimeError(f"shape '[{B}, -1, C]' is invalid: rows of the {what} mask have different numbers of "
                           f"visible tokens {counts.tolist()} (rectangularize the masks or use a padded model)")
    return n


def _require_cuda(x):
    if x.device.type != "cuda":
        raise RuntimeError("counterfactualworldmodels_b200: the conjoined VMAE forward only runs on a CUDA sm_100 "
                           "(B200) device; there is no CPU or PyTorch fallback")


# ---------------------------------------------------------------------------------------------------------
# PaddedVisionTransformer
# ---------------------------------------------------------------------------------------------------------
class PaddedVisionTransformer(PretrainVisionTransformer):
    """Allow batches with a mixed number of visible patches by padding encoder tokens (conjoined_vmae.py:24-210)."""
    PRINT_PADDING = False

    def __init__(self, min_padding_tokens=0, max_padding_tokens=16, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.min_padding_tokens = min_padding_tokens
        self.max_padding_tokens = max_padding_tokens
        self.null_token_enc = nn.Parameter(torch.zeros(1, 1, self.encoder.embed_dim), requires_grad=True)
        self.null_token_dec = nn.Parameter(torch.zeros(1, 1, self.decoder.embed_dim), requires_grad=True)
        nn.init.trunc_normal_(self.null_token_enc, mean=0., std=0.02, a=-0.02, b=0.02)
        nn.init.trunc_normal_(self.null_token_dec, mean=0., std=0.02, a=-0.02, b=0.02)
        self._stream_engine = _StreamEngine()
        self._ws = _Workspace()
        self._reset_padding_mask()

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'mask_token', 'null_token_enc', 'null_token_dec'}

    def _set_padding_mask(self, mask, device=None):
        """conjoined_vmae.py:49-116, computed on the mask's device (the reference goes through CPU tensors and
        copies back).  Integer / boolean bookkeeping only."""
        if device is not None:
            self.device = device
        else:
            device = getattr(self, 'device', mask.device)
        with torch.no_grad():
            mask = mask.bool()
            B, P = mask.size(0), self.max_padding_tokens
            self._num_visible = torch.sum((~mask).flatten(1).int(), -1, keepdim=True)
            _min_num_vis = torch.min(self._num_visible)
            _max_num_vis = torch.max(self._num_visible)
            num_padding_per_ex = _max_num_vis - self._num_visible + self.min_padding_tokens
            padding_mask = torch.arange(P, device=mask.device)[None].expand(B, -1) < num_padding_per_ex
            null_padding_mask = torch.zeros((B, P), dtype=torch.bool, device=mask.device)
            null_padding_mask[:, :1] = True
            any_visible = (torch.sum(self._num_visible) > 0).reshape(1, 1).expand(B, P)
            padding_mask = torch.where(any_visible, padding_mask, null_padding_mask)
            _min_num_vis = torch.maximum(_min_num_vis, torch.ones_like(_min_num_vis))
            _max_num_vis = torch.maximum(_max_num_vis, torch.ones_like(_max_num_vis))
            padding_mask = ~padding_mask
            min_masked = mask.size(1) - int(_max_num_vis) - self.min_padding_tokens
            self.padding_mask = padding_mask.to(device)
            self._num_visible = self._num_visible.to(device)
            self._min_num_vis, self._max_num_vis = _min_num_vis.to(device), _max_num_vis.to(device)
            self.full_input_mask = torch.cat([mask.to(device), self.padding_mask], -1).contiguous()
            self.null_mask = torch.cat([torch.zeros_like(mask[:, :min_masked]).to(device), self.padding_mask],
                                       -1).contiguous()

    def _reset_padding_mask(self):
        self.padding_mask = None
        self.full_input_mask = None
        self.null_mask = None
        self._num_visible = None
        self._min_num_vis = self._max_num_vis = None

    def get_masked_targets(self, x, mask, patch_size=None, preproc_func=None, postproc_func=None):
        """conjoined_vmae.py:167-187 (host-side indexing of training targets; kept for the IMU wrappers)."""
        if preproc_func is not None:
            x = preproc_func(x)
        pt, ph, pw = patch_size or self.patch_size
        B, C, T, H, W = x.shape
        x = x.reshape(B, C, T // pt, pt, H // ph, ph, W // pw, pw).permute(0, 2, 4, 6, 3, 5, 7, 1)
        x = x.reshape(B, (T // pt) * (H // ph) * (W // pw), pt * ph * pw, C)
        if postproc_func is not None:
            x = postproc_func(x)
        x_padded = torch.cat([x, torch.zeros(B, self.max_padding_tokens, x.size(2), x.size(3)).to(x)], 1)
        return x_padded[self.full_input_mask].reshape(mask.size(0), -1, int(np.prod(list(x.shape[2:]))))

    @torch.no_grad()
    def forward(self, x, mask, *args, reset_padding_mask=True, input_norm=None, **kwargs):
        """conjoined_vmae.py:189-210: [B, Ntot + P - Nvis_max - min_pad, D] with the rows of padding positions zeroed.
        ``input_norm=(mean, std)`` (extension): fuse imagenet_normalize of a raw input into the patch gather."""
        _require_cuda(x)
        if reset_padding_mask:
            self._reset_padding_mask()
        self.device = x.device
        B = x.shape[0]
        mask = mask.reshape(B, -1).to(x.device)
        with torch.cuda.device(x.device):
            _lib.load().cwm_launch_count_reset()
            se = self._stream_engine.ensure(self, x.device)
            stream = _stream_ptr(x.device)
            self._set_padding_mask(mask)
            perm, inv, nvis = compact_mask(self.full_input_mask)
            n_vis = _visible_count(nvis, B, "padded")
            Next = perm.shape[1]
            xe = _tokenize_visible(se, x, self.patch_size, perm, n_vis, B, stream, input_norm)
            for i in range(len(self.encoder.blocks)):
                _run_block(self._ws, se.enc_blocks, i, xe, B, n_vis, se.Ce, se.enc_dims, se.ln_eps, stream)
            xd = _to_decoder(se, xe, perm, n_vis, B, stream)
            for i in range(len(self.decoder.blocks)):
                _run_block(self._ws, se.dec_blocks, i, xd, B, Next, se.Cd, se.dec_dims, se.ln_eps, stream)
            n_ret = Next - n_vis
            y = _last_tokens(se, xd, perm, n_vis, B, n_ret if n_ret > 0 else Next, True, stream)
            self.last_forward_launches = _lib.load().cwm_last_forward_launches()
            self.last_aux = (perm, inv, n_vis)
        return y


# ---------------------------------------------------------------------------------------------------------
# IMU encoder (parameter holder)
# ---------------------------------------------------------------------------------------------------------
class ImuEncoder(PretrainVisionTransformerEncoder):
    """Encoder for IMU data [B, D, L] treated as a (L, 1, 1) "video" with a (tubelet, 1, 1) patch
    (conjoined_vmae.py:1013-1147)."""
    default_num_imu_channels = 6

    def __init__(self, img_size=None, patch_size=None, sequence_length=200, num_frames=None, tubelet_size=8,
                 in_chans=default_num_imu_channels, use_learnable_pos_emb=False, frame_gap=None,
                 concat_dummy_token=True, use_campose=False, campose_in_chans=16, *args, **kwargs):
        if use_campose:
            raise NotImplementedError("camera-pose inputs are not used by any CWM factory")
        if use_learnable_pos_emb:
            raise NotImplementedError("use_learnable_pos_emb=True is not used by any CWM factory")
        kwargs.pop('spacetime_separable_pos_embed', None)
        kwargs.pop('embed_per_frame', None)
        super().__init__(img_size=(1, 1), patch_size=(1, 1), tubelet_size=tubelet_size, use_learnable_pos_emb=False,
                         in_chans=in_chans, num_frames=sequence_length, *args, **kwargs)
        self.in_dim = in_chans
        self.sequence_length = sequence_length
        self.num_tokens = self.num_patches
        self.num_frames = 0
        self.frame_gap = frame_gap
        self.timestamps = None
        self._concat_dummy_token = concat_dummy_token
        self._learnable_pos_embed = False
        # conjoined_vmae.py:1080-1094: built on first use in the reference; it only depends on the token count
        self.pos_embed = pos_embedding(self.num_tokens + int(self._concat_dummy_token), self.embed_dim)
        if self._concat_dummy_token:
            self.dummy_token = nn.Parameter(torch.zeros((1, self.in_dim, self.patch_size[0], 1, 1)),
                                            requires_grad=True)
            nn.init.trunc_normal_(self.dummy_token, mean=0., std=0.02, a=-0.02, b=0.02)

    @property
    def shape(self):
        return (self.sequence_length, 1, 1)

    def _get_dummy_token(self, B, device):
        return self.dummy_token.expand(B, -1, -1, -1, -1).to(device)

    def concat_dummy(self, imu, mask):
        """`tokenize` (conjoined_vmae.py:1110-1125): the always-visible learnable dummy token rides at the end."""
        if imu is None:
            raise NotImplementedError("imu=None (implicit fully-masked input) -- pass zeros and an all-True mask, as "
                                      "ImuConditionedFlowGenerator.get_fake_head_motion does (segmentation.py:814-832)")
        if self._concat_dummy_token:
            imu = torch.cat([imu, self._get_dummy_token(imu.size(0), imu.device).to(imu.dtype)], 2)
            mask = torch.cat([mask, torch.zeros_like(mask[:, -1:])], -1)
        return imu, mask


# ---------------------------------------------------------------------------------------------------------
# conjoined transformers
# ---------------------------------------------------------------------------------------------------------
class ConjoinedPretrainVisionTransformer(nn.Module):
    """Two parallel token streams conjoined by cross-attention blocks (conjoined_vmae.py:212-887)."""
    debug_mode = False
    default_cross_block_kwargs = {'num_heads': 4, 'mlp_ratio': 2.0, 'shared_similarity': False,
                                  'with_self_attention': False}
    default_model_kwargs = {'encoder_func': PretrainVisionTransformerEncoder, 'tubelet_size': 1}
    default_input_kwargs = {'unnormalize': True}

    def __init__(self, img_size=224, patch_size=(8, 8), context_img_size=None, context_patch_size=(8, 8),
                 num_frames=None, main_input='rgb02', main_input_kwargs=default_input_kwargs,
                 context_input='flow01', context_input_kwargs=default_input_kwargs,
                 main_model_func=PretrainVisionTransformer, main_model_kwargs=default_model_kwargs,
                 context_model_func=PretrainVisionTransformer, context_model_kwargs=default_model_kwargs,
                 conjoin_encoder_layers=[(0, 0), (-1, -1)], conjoin_decoder_layers=[(0, 0)],
                 conjoin_func=CrossAttentionTransformerBlock, encoder_cross_block_kwargs=default_cross_block_kwargs,
                 decoder_cross_block_kwargs=default_cross_block_kwargs, output_main=True, output_context=False,
                 context_mask_func=None, context_mask_kwargs={}, decode_main=True, decode_context=True,
                 use_flash_attention=False, *args, **kwargs):
        super().__init__()
        if conjoin_func is not CrossAttentionTransformerBlock:
            raise NotImplementedError("only CrossAttentionTransformerBlock conjoining is implemented")
        if not (decode_main and decode_context):
            raise NotImplementedError("decode_main / decode_context = False are not used by any CWM factory")
        if context_mask_func is not None:
            raise NotImplementedError("context_mask_func (model-internal mask sampling) is host-side RNG, out of scope")
        self.get_main_input = self._build_stream_input(main_input, **main_input_kwargs)
        self.get_context_input = self._build_stream_input(context_input, **context_input_kwargs)
        num_frames_main = self.get_main_input.get_num_frames()
        num_frames_context = self.get_context_input.get_num_frames()
        self.num_frames = num_frames or (num_frames_main or 0)

        main_kwargs = copy.deepcopy(kwargs)
        main_kwargs.update(main_model_kwargs)
        main_kwargs['use_flash_attention'] = use_flash_attention
        main_kwargs['encoder_in_chans'] = self.get_main_input.num_channels or 3
        if main_kwargs.get('decoder_num_classes', None) is None:
            num_out_chans = main_kwargs.get('tubelet_size', 1) * np.prod(patch_size)
            main_kwargs['decoder_num_classes'] = int(main_kwargs['encoder_in_chans'] * num_out_chans)
        context_kwargs = copy.deepcopy(kwargs)
        context_kwargs.update(context_model_kwargs)
        context_kwargs['use_flash_attention'] = use_flash_attention
        context_kwargs['encoder_in_chans'] = self.get_context_input.num_channels or 3
        if context_kwargs.get('decoder_num_classes', None) is None:
            num_out_chans = context_kwargs.get('tubelet_size', 1) * np.prod((context_patch_size or patch_size))
            context_kwargs['decoder_num_classes'] = int(context_kwargs['encoder_in_chans'] * num_out_chans)

        self.main_stream = main_model_func(img_size=img_size, patch_size=patch_size, num_frames=num_frames_main,
                                           **main_kwargs)
        context_img_size = (context_img_size or img_size)
        self.context_stream = context_model_func(img_size=context_img_size, patch_size=context_patch_size,
                                                 num_frames=num_frames_context, **context_kwargs)
        self._conjoin_func = conjoin_func
        # NB: the reference writes flash_attention=False into the (shared default) kwargs dicts
        # (conjoined_vmae.py:305-306); we copy instead of mutating.
        self._encoder_cross_block_kwargs = dict(copy.deepcopy(encoder_cross_block_kwargs), flash_attention=False)
        self._decoder_cross_block_kwargs = dict(copy.deepcopy(decoder_cross_block_kwargs), flash_attention=False)
        if conjoin_encoder_layers is True:
            conjoin_encoder_layers = list(range(min(self.main_stream.encoder.get_num_layers(),
                                                    self.context_stream.encoder.get_num_layers())))
        elif conjoin_encoder_layers in [False, None]:
            conjoin_encoder_layers = []
        if conjoin_decoder_layers is True:
            conjoin_decoder_layers = list(range(min(self.main_stream.decoder.get_num_layers(),
                                                    self.context_stream.decoder.get_num_layers())))
        elif conjoin_decoder_layers in [False, None]:
            conjoin_decoder_layers = []
        self._build_conjoining_attention_blocks(conjoin_encoder_layers, conjoin_decoder_layers)
        self._set_decoder_outputs(output_main, output_context)
        self.context_mask_generator = None
        self._decode_main = decode_main
        self._decode_context = decode_context
        self._context_input = context_input
        self._engines = (_StreamEngine(), _StreamEngine())
        self._cross_sig = None
        self._cross = None
        self._ws = _Workspace()
        self.last_forward_launches = 0
        self.last_aux = None

    def __getattr__(self, key):
        # conjoined_vmae.py:347-354: unknown attributes are looked up on the main stream; a value of None counts as
        # missing, which is how the wrappers tell "no padding mask set" (prediction.py:412-413)
        try:
            return super().__getattr__(key)
        except AttributeError:
            if key == 'main_stream':
                raise
            attr = getattr(self.main_stream, key, None)
            if attr is None:
                raise AttributeError("no attr %s in the module or the main transformer stream" % key)
            return attr

    @property
    def mask_size(self):
        return (self.num_frames // self.main_stream.patch_size[0],
                self.main_stream.image_size[-2] // self.main_stream.patch_size[-2],
                self.main_stream.image_size[-1] // self.main_stream.patch_size[-1])

    def _build_stream_input(self, func, temporal_dim=2, **kwargs):
        if isinstance(func, str):
            return preproc.get_preprocessor(func, temporal_dim=temporal_dim, **kwargs)
        elif isinstance(func, (partial, nn.Module)) or callable(func):
            return func(temporal_dim=temporal_dim, **kwargs)
        raise ValueError("%s is not a valid stream input function or name" % func)

    def _build_conjoining_block(self, layer_pair, encoder=True):
        main_idx, context_idx = layer_pair
        main_block = getattr(self.main_stream, 'encoder' if encoder else 'decoder').blocks[main_idx]
        context_block = getattr(self.context_stream, 'encoder' if encoder else 'decoder').blocks[context_idx]
        kwargs = self._encoder_cross_block_kwargs if encoder else self._decoder_cross_block_kwargs

        def _in_out_dim(block):
            return (block.attn.qkv.in_features, block.mlp.fc2.out_features)
        in_dim, out_dim = _in_out_dim(main_block)
        in_dim_src, out_dim_src = _in_out_dim(context_block)
        return self._conjoin_func(in_dim=in_dim, out_dim=out_dim, in_dim_src=in_dim_src, out_dim_src=out_dim_src,
                                  **kwargs)

    def _build_conjoining_attention_blocks(self, enc_layers, dec_layers):
        n_me, n_md = self.main_stream.encoder.get_num_layers(), self.main_stream.decoder.get_num_layers()
        n_ce, n_cd = self.context_stream.encoder.get_num_layers(), self.context_stream.decoder.get_num_layers()

        def keys(layers, n_m, n_c):
            out = []
            for pair in layers:
                if not hasattr(pair, '__len__'):
                    pair = (pair, pair)
                out.append((pair[0] % n_m, pair[1] % n_c))
            return out
        self.encoder_conjoining_blocks = nn.ModuleDict([
            ("{}-{}".format(*key), self._build_conjoining_block(key, encoder=True)) for key in keys(enc_layers, n_me, n_ce)])
        self.decoder_conjoining_blocks = nn.ModuleDict([
            ("{}-{}".format(*key), self._build_conjoining_block(key, encoder=False)) for key in keys(dec_layers, n_md, n_cd)])

    def _set_decoder_outputs(self, output_main=None, output_context=None):
        if output_main is not None:
            self._output_main = output_main
        if output_context is not None:
            self._output_context = output_context

    # ---- inputs (host-side frame selection / reshapes, conjoined_vmae.py:430-485) ----
    def get_stream_inputs(self, x, mask, timestamps=None, x_context=None, mask_context=None):
        B, _, T = x.shape[:3]
        if timestamps is None:
            timestamps = torch.arange(self.num_frames)[None].expand(B, -1).float().to(x.device)
        else:
            assert list(timestamps.shape) == [B, self.num_frames], (timestamps.shape, [B, self.num_frames])
        x_m = self.get_main_input(x, timestamps=timestamps)
        x_c = self.get_context_input(x_context if x_context is not None else x, timestamps=timestamps)
        ts_m = self.get_main_input.get_output_frames(timestamps, temporal_dim=1) if self.get_main_input.num_frames \
            else timestamps
        ts_c = self.get_context_input.get_output_frames(timestamps, temporal_dim=1) \
            if self.get_context_input.num_frames else timestamps
        assert (mask.size(-1) % T) == 0, mask.shape
        mask = mask.reshape(B, T, mask.size(-1) // T)
        mask_m = self.get_main_input.get_output_frames(mask, temporal_dim=1).reshape(B, -1)
        if mask_context is None:
            mask_c = self.get_context_input.get_output_frames(mask, temporal_dim=1).reshape(B, -1)
        elif self.get_context_input.num_frames in [0, None]:
            mask_c = mask_context
        else:
            mask_c = mask_context.view(B, T, mask_context.size(-1) // T)
            mask_c = self.get_context_input.get_output_frames(mask_c, temporal_dim=1).reshape(B, -1)
        return ((x_m, mask_m, ts_m), (x_c, mask_c, ts_c))

    def get_current_inputs(self, x, mask, *args, **kwargs):
        inp_m, inp_c = self.get_stream_inputs(x, mask, *args, **kwargs)
        if self._output_main and self._output_context:
            return (inp_m, inp_c)
        elif self._output_main:
            return (inp_m,)
        elif self._output_context:
            return (inp_c,)
        return (inp_m, inp_c)

    def get_masked_imu(self, imu, mask, preproc_func=None, postproc_func=None):
        """conjoined_vmae.py:840-850 -> :757-789 for the context stream (host-side indexing used by ImuGenerator)."""
        x = imu[..., None, None]
        if preproc_func is not None:
            x = preproc_func(x)
        pt, ph, pw = self.context_stream.patch_size
        B, C, T, H, W = x.shape
        x = x.reshape(B, C, T // pt, pt, H // ph, ph, W // pw, pw).permute(0, 2, 4, 6, 3, 5, 7, 1)
        x = x.reshape(B, (T // pt) * (H // ph) * (W // pw), pt * ph * pw, C)
        if postproc_func is not None:
            x = postproc_func(x)
        full_mask = getattr(self.context_stream, 'full_input_mask', None)
        if full_mask is None:
            targets = x[mask]
        else:
            num_pad = full_mask.size(1) - mask.size(1)
            targets = torch.cat([x, torch.zeros(B, num_pad, x.shape[2], x.shape[3]).to(x)], 1)[full_mask]
        return targets.reshape(mask.size(0), -1 if targets.size(0) > 0 else 0, int(np.prod(list(x.shape[2:]))))

    # ---- engine plumbing ----
    def _ensure_cross(self, device):
        blocks = list(self.encoder_conjoining_blocks.values()) + list(self.decoder_conjoining_blocks.values())
        sig = (str(device),) + tuple((p.data_ptr(), p._version) for b in blocks for p in b.parameters())
        if sig != self._cross_sig:
            pk = _Packer(device)
            enc = {k: _pack_cross_block(pk, b) for k, b in self.encoder_conjoining_blocks.items()}
            dec = {k: _pack_cross_block(pk, b) for k, b in self.decoder_conjoining_blocks.items()}
            self._cross = (pk, enc, dec)
            self._cross_sig = sig
        return self._cross[1], self._cross[2]

    def _stream_masks(self, mask_m, x_c, mask_c):
        """Full (token + padding position) masks of both streams; the non-padded context encoder appends its dummy
        token (conjoined_vmae.py:1121-1123, :595-609)."""
        enc_c = self.context_stream.encoder
        if isinstance(enc_c, ImuEncoder):
            x_c, mask_c = enc_c.concat_dummy(x_c, mask_c)
        return mask_m, x_c, mask_c

    def _set_context_pos_embed(self):
        """`_set_decoder_inputs` + `_concat_excess_tokens` (conjoined_vmae.py:578-609): an IMU context stream switches
        its decoder table to the torch-fp32 `pos_embedding` (vmae.py:446-449, timestamps never set on the stream), one
        extra row for the dummy token."""
        cs = self.context_stream
        if isinstance(cs.encoder, ImuEncoder):
            n = cs.num_patches + int(cs.encoder._concat_dummy_token)
            if getattr(cs, '_imu_pos_rows', None) != n:
                cs.pos_embed = pos_embedding(n, cs.decoder.embed_dim)
                cs._imu_pos_rows = n

    def _padded_full_mask(self, stream, mask):
        return mask

    def _zero_pads(self, stream):
        return False

    @torch.no_grad()
    def forward(self, x, mask, timestamps=None, x_context=None, mask_context=None, output_main=None,
                output_context=None, *args, input_norm=None, **kwargs):
        """conjoined_vmae.py:852-887 (+ :918-1011 for the padded subclass).  ``input_norm=(mean, std)`` (extension, used
        by the B200 `PredictorBasedGenerator.predict`): ``x`` holds raw [0, 1] frames and imagenet_normalize is fused
        into the main stream's patch gather."""
        _require_cuda(x)
        lib = _lib.load()
        ms, cs = self.main_stream, self.context_stream
        ms.device = x.device
        cs.device = x_context.device if x_context is not None else x.device
        B = x.shape[0]
        mask = mask.reshape(B, -1).to(x.device)
        if mask_context is not None:
            mask_context = mask_context.to(x.device)
        (x_m, mask_m, _), (x_c, mask_c, _) = self.get_stream_inputs(
            x, mask, timestamps, x_context=(x_context if x_context is not None else x), mask_context=mask_context)
        self._set_decoder_outputs(output_main, output_context)
        with torch.cuda.device(x.device):
            lib.cwm_launch_count_reset()
            stream = _stream_ptr(x.device)
            self._set_context_pos_embed()
            mask_m, x_c, mask_c = self._stream_masks(mask_m, x_c, mask_c)
            full_m = self._padded_full_mask(ms, mask_m.bool())
            full_c = self._padded_full_mask(cs, mask_c.bool())
            se_m = self._engines[0].ensure(ms, x.device)
            se_c = self._engines[1].ensure(cs, x.device)
            cross_enc, cross_dec = self._ensure_cross(x.device)
            perm_m, inv_m, nvis_m = compact_mask(full_m)
            perm_c, inv_c, nvis_c = compact_mask(full_c)
            assert perm_m.shape[1] == se_m.n_tokens + se_m.P, (perm_m.shape, se_m.n_tokens, se_m.P)
            assert perm_c.shape[1] == se_c.n_tokens + se_c.P, (perm_c.shape, se_c.n_tokens, se_c.P)
            n_m = _visible_count(nvis_m, B, "main-stream")
            n_c = _visible_count(nvis_c, B, "context-stream")
            Nx_m, Nx_c = perm_m.shape[1], perm_c.shape[1]

            # ---- encoders: tokenise, then blocks with the cross block *before* the paired blocks (:543-576)
            xe = _tokenize_visible(se_m, x_m, ms.patch_size, perm_m, n_m, B, stream, input_norm)
            xc = _tokenize_visible(se_c, x_c, cs.patch_size, perm_c, n_c, B, stream)
            i = j = 0
            for pair in self.encoder_conjoining_blocks.keys():
                pi, pj = (int(p) for p in pair.split('-'))
                while i < pi:
                    _run_block(self._ws, se_m.enc_blocks, i, xe, B, n_m, se_m.Ce, se_m.enc_dims, se_m.ln_eps, stream)
                    i += 1
                while j < pj:
                    _run_block(self._ws, se_c.enc_blocks, j, xc, B, n_c, se_c.Ce, se_c.enc_dims, se_c.ln_eps, stream)
                    j += 1
                w, g = cross_enc[pair]
                _run_cross(self._ws, w, g, xe, xc, B, n_m, n_c, stream)
            for _i in range(i, len(ms.encoder.blocks)):
                _run_block(self._ws, se_m.enc_blocks, _i, xe, B, n_m, se_m.Ce, se_m.enc_dims, se_m.ln_eps, stream)
            for _j in range(j, len(cs.encoder.blocks)):
                _run_block(self._ws, se_c.enc_blocks, _j, xc, B, n_c, se_c.Ce, se_c.enc_dims, se_c.ln_eps, stream)

            # ---- encoder norm + encoder_to_decoder + decoder-input assembly (:876-877, :620-635 / :956-977)
            xd = _to_decoder(se_m, xe, perm_m, n_m, B, stream)
            xdc = _to_decoder(se_c, xc, perm_c, n_c, B, stream)
            self.B, self.N, self.C = B, n_m, se_m.Cd
            self.M, self.D = n_c, se_c.Cd

            # ---- decoders: the cross block comes *after* the paired blocks (:688-720)
            i = j = 0
            for pair in self.decoder_conjoining_blocks.keys():
                pi, pj = (int(p) for p in pair.split('-'))
                while i <= pi:
                    _run_block(self._ws, se_m.dec_blocks, i, xd, B, Nx_m, se_m.Cd, se_m.dec_dims, se_m.ln_eps, stream)
                    i += 1
                while j <= pj:
                    _run_block(self._ws, se_c.dec_blocks, j, xdc, B, Nx_c, se_c.Cd, se_c.dec_dims, se_c.ln_eps, stream)
                    j += 1
                w, g = cross_dec[pair]
                _run_cross(self._ws, w, g, xd, xdc, B, Nx_m, Nx_c, stream)
            for _i in range(i, len(ms.decoder.blocks)):
                _run_block(self._ws, se_m.dec_blocks, _i, xd, B, Nx_m, se_m.Cd, se_m.dec_dims, se_m.ln_eps, stream)
            for _j in range(j, len(cs.decoder.blocks)):
                _run_block(self._ws, se_c.dec_blocks, _j, xdc, B, Nx_c, se_c.Cd, se_c.dec_dims, se_c.ln_eps, stream)

            # ---- masked tokens of the requested streams (:673-686 / :987-1011)
            want_main = self._output_main or not self._output_context
            want_ctx = self._output_context or not self._output_main
            y = yc = None
            if want_main:
                if not (self._output_main or self._output_context):
                    raise NotImplementedError("output_main = output_context = False (all tokens of both streams)")
                y = _last_tokens(se_m, xd, perm_m, n_m, B, Nx_m - n_m, self._zero_pads(ms), stream)
            if want_ctx:
                yc = _last_tokens(se_c, xdc, perm_c, n_c, B, Nx_c - n_c, self._zero_pads(cs), stream)
            self.last_forward_launches = lib.cwm_last_forward_launches()
            self.last_aux = ((perm_m, inv_m, n_m), (perm_c, inv_c, n_c))
        if self._output_main and self._output_context:
            return (y, yc)
        elif self._output_main:
            return y
        return yc


class ConjoinedPaddedVisionTransformer(ConjoinedPretrainVisionTransformer):
    """conjoined_vmae.py:889-1011: both streams are PaddedVisionTransformers by default."""

    def __init__(self, main_model_func=PaddedVisionTransformer, context_model_func=PaddedVisionTransformer,
                 *args, **kwargs):
        super().__init__(main_model_func=main_model_func, context_model_func=context_model_func, *args, **kwargs)

    @property
    def _main_padded(self):
        return hasattr(self.main_stream, 'padding_mask')

    @property
    def _context_padded(self):
        return hasattr(self.context_stream, 'padding_mask')

    def _reset_padding_mask(self):
        self.main_stream._reset_padding_mask()
        self.context_stream._reset_padding_mask()

    def _set_padding_mask(self, mask, mask_context):
        self.main_stream.device = mask.device
        self.context_stream.device = mask_context.device
        if self._main_padded:
            self.main_stream._set_padding_mask(mask)
        if self._context_padded:
            self.context_stream._set_padding_mask(mask_context)

    def _stream_masks(self, mask_m, x_c, mask_c):
        enc_c = self.context_stream.encoder
        if isinstance(enc_c, ImuEncoder) and enc_c._concat_dummy_token:
            x_c, mask_c = enc_c.concat_dummy(x_c, mask_c)
        return mask_m, x_c, mask_c

    def _padded_full_mask(self, stream, mask):
        # conjoined_vmae.py:933-943: the padding mask is only computed when none is set -- the forward does NOT reset
        # it; the wrapper does after every call (prediction.py:451-452)
        if not hasattr(stream, 'padding_mask'):
            return mask
        if stream.padding_mask is None:
            stream._set_padding_mask(mask, device=mask.device)
        return stream.full_input_mask

    def _zero_pads(self, stream):
        return hasattr(stream, 'padding_mask')


# ---------------------------------------------------------------------------------------------------------
# scaffolds and factories (conjoined_vmae.py:1150-1243)
# ---------------------------------------------------------------------------------------------------------
def conjoined_full_videomae_base_224_scaffold(**kwargs):
    model = ConjoinedPretrainVisionTransformer(
        img_size=224, encoder_embed_dim=768, encoder_depth=12, encoder_num_heads=12, encoder_num_classes=0,
        decoder_embed_dim=384, decoder_num_heads=6, decoder_depth=4, mlp_ratio=4, qkv_bias=True,
        norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg()
    return model


def conjoined_padded_videomae_base_224_scaffold(**kwargs):
    model = ConjoinedPaddedVisionTransformer(
        img_size=224, encoder_embed_dim=768, encoder_depth=12, encoder_num_heads=12, encoder_num_classes=0,
        decoder_embed_dim=384, decoder_num_heads=6, decoder_depth=4, mlp_ratio=4, qkv_bias=True,
        norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg()
    return model


rgb_encoder_kwargs = copy.deepcopy(ConjoinedPretrainVisionTransformer.default_model_kwargs)
rgb_encoder_kwargs.update({'encoder_func': PretrainVisionTransformerEncoder, 'decoder_num_classes': None})
rgb_padded_encoder_kwargs = copy.deepcopy(rgb_encoder_kwargs)
rgb_padded_encoder_kwargs.update({'min_padding_tokens': 0, 'max_padding_tokens': 16})
rgb_4x4_padded_encoder_kwargs = copy.deepcopy(rgb_padded_encoder_kwargs)
rgb_4x4_padded_encoder_kwargs.update({'max_padding_tokens': 64})
imu_encoder_kwargs = copy.deepcopy(ConjoinedPretrainVisionTransformer.default_model_kwargs)
imu_encoder_kwargs.update({'encoder_func': ImuEncoder, 'spacetime_separable_pos_embed': True,
                           'encoder_embed_dim': 384, 'decoder_embed_dim': 192})
imu400_encoder_kwargs = copy.deepcopy(imu_encoder_kwargs)
imu400_encoder_kwargs.update({'sequence_length': 400, 'tubelet_size': 16, 'decoder_num_classes': 6 * 16})
imu400_padded_encoder_kwargs = copy.deepcopy(imu400_encoder_kwargs)
imu400_padded_encoder_kwargs.update({'min_padding_tokens': 0, 'max_padding_tokens': 25, 'concat_dummy_token': False})


def imu400_8x8patch_2frames_1tube_flowbackrgb01(**kwargs):
    """flow2imu (a17): 7-channel flow+rgb main stream, IMU context stream, predicts the IMU tokens.  The flow input needs
    `main_input_kwargs={'flow_model': ...}` (see preprocessor.FramePairFlow)."""
    return conjoined_full_videomae_base_224_scaffold(
        num_frames=2, main_input='flowback_rgb01', context_input='imu', main_model_kwargs=rgb_encoder_kwargs,
        context_model_kwargs=imu400_encoder_kwargs, conjoin_encoder_layers=[0, -1], conjoin_decoder_layers=True,
        **kwargs)


def imu400_base_4x4patch_2frames_1tube(**kwargs):
    """BASELINE config 5: IMU-conditioned ViT-base VMAE, 4x4 patches (conjoined_vmae.py:1230-1243)."""
    return conjoined_padded_videomae_base_224_scaffold(
        patch_size=(4, 4), main_model_func=PaddedVisionTransformer, main_model_kwargs=rgb_4x4_padded_encoder_kwargs,
        main_input='rgb01', main_input_kwargs={'unnormalize': False}, context_model_func=PaddedVisionTransformer,
        context_model_kwargs=imu400_padded_encoder_kwargs, context_input='imu',
        conjoin_encoder_layers=range(0, 12, 3), conjoin_decoder_layers=True, **kwargs)
