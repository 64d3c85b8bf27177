"""Drop-in replacement for the reference's VMAE predictor (``cwm/models/VideoMAE/vmae.py``).

Same constructor arguments, same ``state_dict`` keys, same ``forward(x, mask)`` signature and the same attributes
the reference wrappers read (``patch_size``, ``image_size``, ``num_frames``, ``mask_size``, ``num_patches``,
``encoder.patch_embed.proj.kernel_size`` ...).  The modules below only *hold parameters*;
all arithmetic of the forward pass runs in hand-written sm_100a CUDA behind the C ABI of ``libcwm_b200.so``
(``include/cwm_b200.h``).  There is no PyTorch or CPU fallback: calling ``forward`` without a B200 raises.

Reference citations are relative to /root/reference.
"""
import ctypes
import math
import os
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import _lib

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)  # cwm/models/utils.py:12
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)   # cwm/models/utils.py:13


def get_sinusoid_encoding_table(positions, d_hid):
    """Sinusoid table of cwm/models/VideoMAE/utils.py:251-268: float64 ``pos / 10000^(2*(j//2)/d)``, sin on even
    columns, cos on odd columns, then cast to float32.  Vectorised, same float64 operations."""
    pos = np.arange(positions, dtype=np.float64) if isinstance(positions, int) else np.asarray(positions, np.float64)
    j = np.arange(d_hid)
    denom = np.power(10000, 2 * (j // 2) / d_hid)
    table = pos[:, None] / denom[None, :]
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.FloatTensor(table).unsqueeze(0)


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def _no_eager(name):
    raise NotImplementedError(
        f"{name}.forward: the B200 build runs the whole VMAE forward inside libcwm_b200 "
        "(PretrainVisionTransformer.forward); sub-module eager forwards are not provided")


class Mlp(nn.Module):
    """Parameter holder for cwm/models/VideoMAE/utils.py:37-54."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        if act_layer is not nn.GELU:
            raise NotImplementedError("only GELU (erf) is fused into the fc1 epilogue")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x):
        _no_eager("Mlp")


class Attention(nn.Module):
    """Parameter holder for cwm/models/VideoMAE/utils.py:57-121 (qkv without bias + separate q/v biases)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.,
                 attn_head_dim=None, flash_attention=False):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if attn_head_dim is not None:
            head_dim = attn_head_dim
        all_head_dim = head_dim * self.num_heads
        self.head_dim = head_dim
        self.scale = qk_scale or head_dim ** -0.5
        self.flash_attention = flash_attention  # accepted for API parity; the CUDA kernel is always flash-style
        self.qkv = nn.Linear(dim, all_head_dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(all_head_dim))
            self.v_bias = nn.Parameter(torch.zeros(all_head_dim))
        else:
            self.q_bias = None
            self.v_bias = None
        self.proj = nn.Linear(all_head_dim, dim)

    def forward(self, x, attn_mask=None):
        _no_eager("Attention")


class Block(nn.Module):
    """Parameter holder for cwm/models/VideoMAE/utils.py:124-153."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., init_values=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm, attn_head_dim=None,
                 in_dim=None, flash_attention=False):
        super().__init__()
        if drop_path and drop_path > 0:
            raise NotImplementedError("drop_path_rate > 0 (training only)")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, attn_head_dim=attn_head_dim, flash_attention=flash_attention)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        if (init_values or 0) > 0:  # layer scale (utils.py:140-144): folded into the proj / fc2 weights when packed
            self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
            self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        else:
            self.gamma_1, self.gamma_2 = None, None

    def forward(self, x, attn_mask=None):
        _no_eager("Block")


class PatchEmbed(nn.Module):
    """Parameter holder for cwm/models/VideoMAE/utils.py:156-198 (Conv3d with kernel = stride = (tubelet, ph, pw))."""

    def __init__(self, img_size=224, patch_size=(16, 16), in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2):
        super().__init__()
        img_size = _to_2tuple(img_size)
        self.tubelet_size = int(tubelet_size)
        self.num_frames = int(num_frames)
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0]) * \
            (num_frames // self.tubelet_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.proj = nn.Conv3d(in_channels=in_chans, out_channels=embed_dim,
                              kernel_size=(self.tubelet_size, patch_size[0], patch_size[1]),
                              stride=(self.tubelet_size, patch_size[0], patch_size[1]))

    def forward(self, x, **kwargs):
        _no_eager("PatchEmbed")


def _init_weights(m):
    # vmae.py:100-107
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class PretrainVisionTransformerEncoder(nn.Module):
    """Parameter holder for cwm/models/VideoMAE/vmae.py:31-182."""

    def __init__(self, img_size=224, patch_size=(16, 16), in_chans=3, num_classes=0, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=None, tubelet_size=2,
                 use_learnable_pos_emb=False, num_frames=16, embed_per_frame=False,
                 spacetime_separable_pos_embed=False, block_func=Block, block_kwargs={}):
        super().__init__()
        if embed_per_frame:
            raise NotImplementedError("embed_per_frame=True is not exercised by any CWM factory")
        if num_classes:
            raise NotImplementedError("encoder_num_classes > 0 (classification head)")
        if block_func is not Block:
            raise NotImplementedError("custom encoder_block_func")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_size = (tubelet_size,) + tuple(patch_size)
        self.pt, self.ph, self.pw = self.patch_size
        self._embed_per_frame = False
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                      embed_dim=embed_dim, tubelet_size=tubelet_size, num_frames=num_frames)
        self.image_size = img_size
        self.num_patches = self.patch_embed.num_patches
        self.num_frames = num_frames
        if use_learnable_pos_emb:  # vmae.py:68-70, :87-88
            self._learnable_pos_embed = True
            self.pos_embed = nn.Parameter(torch.zeros(1, self.num_patches, embed_dim))
            nn.init.trunc_normal_(self.pos_embed, mean=0., std=.02, a=-.02, b=.02)
        else:
            self._learnable_pos_embed = False
            self.pos_embed = get_sinusoid_encoding_table(self.num_patches, embed_dim)  # plain tensor (vmae.py:75)
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, in_dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                  qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=0.,
                  norm_layer=norm_layer, init_values=init_values, **block_kwargs)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Identity()
        self.timestamps = None
        self.apply(_init_weights)

    def get_num_layers(self):
        return len(self.blocks)

    def forward(self, x, mask, *args, **kwargs):
        _no_eager("PretrainVisionTransformerEncoder")


class PretrainVisionTransformerDecoder(nn.Module):
    """Parameter holder for cwm/models/VideoMAE/vmae.py:184-255."""

    def __init__(self, patch_size=(16, 16), num_classes=768, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.,
                 qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, init_values=None, num_patches=196, tubelet_size=2, block_func=Block,
                 block_kwargs={}):
        super().__init__()
        if block_func is not Block:
            raise NotImplementedError("custom decoder_block_func")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_size = patch_size
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, in_dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                  qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=0.,
                  norm_layer=norm_layer, init_values=init_values, **block_kwargs)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(_init_weights)

    def get_num_layers(self):
        return len(self.blocks)

    def forward(self, x, return_token_num):
        _no_eager("PretrainVisionTransformerDecoder")


class _Packer:
    """Collects device copies of parameters (fp32 as is, GEMM weights as f16) and keeps them alive."""

    def __init__(self, device):
        self.device = device
        self.keep = []

    def f32(self, t):
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.keep.append(t)
        return t.data_ptr()

    def f16(self, t):
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous().to(_lib.act_dtype())   # f16 (bf16 build: bf16)
        self.keep.append(t)
        return t.data_ptr()

    def block_array(self, blocks, fold_ln=False):
        """``fold_ln``: also pack the LayerNorm-folded qkv / fc1 weights that let the library run the block without
        LayerNorm kernels (``cwm_block_weights.w_qkv_ln`` ...): W' = f16(W * gamma), s = row sums of W' (of the f16
        values, so that the mean term cancels exactly what the tensor cores accumulate), c = W @ beta + bias."""
        arr = (_lib.BlockWeights * max(1, len(blocks)))()
        for i, blk in enumerate(blocks):
            w = arr[i]
            w.ln1_g, w.ln1_b = self.f32(blk.norm1.weight), self.f32(blk.norm1.bias)
            w.w_qkv = self.f16(blk.attn.qkv.weight)
            if blk.attn.q_bias is not None:
                # qkv_bias = cat(q_bias, zeros_like(v_bias), v_bias)  (utils.py:89-91)
                w.b_qkv = self.f32(torch.cat([blk.attn.q_bias.detach(), torch.zeros_like(blk.attn.v_bias),
                                              blk.attn.v_bias.detach()]))
            else:
                w.b_qkv = None
            # layer scale x + gamma * f(x) (utils.py:151-152) == a row scaling of the last linear map of f
            g1 = getattr(blk, "gamma_1", None)
            g2 = getattr(blk, "gamma_2", None)
            wp, bp = blk.attn.proj.weight.detach().float(), blk.attn.proj.bias.detach().float()
            w2, b2 = blk.mlp.fc2.weight.detach().float(), blk.mlp.fc2.bias.detach().float()
            if g1 is not None:
                wp, bp = g1.detach().float()[:, None] * wp, g1.detach().float() * bp
            if g2 is not None:
                w2, b2 = g2.detach().float()[:, None] * w2, g2.detach().float() * b2
            w.w_proj, w.b_proj = self.f16(wp), self.f32(bp)
            w.ln2_g, w.ln2_b = self.f32(blk.norm2.weight), self.f32(blk.norm2.bias)
            w.w_fc1, w.b_fc1 = self.f16(blk.mlp.fc1.weight), self.f32(blk.mlp.fc1.bias)
            w.w_fc2, w.b_fc2 = self.f16(w2), self.f32(b2)
            if fold_ln:
                def fold(weight, bias, norm):
                    wt = weight.detach().to(device=self.device, dtype=torch.float32)
                    g = norm.weight.detach().to(device=self.device, dtype=torch.float32)
                    b = norm.bias.detach().to(device=self.device, dtype=torch.float32)
                    w16 = (wt * g[None, :]).to(_lib.act_dtype()).contiguous()
                    colsum = w16.float().sum(1).contiguous()
                    c = wt @ b
                    if bias is not None:
                        c = c + bias.detach().to(device=self.device, dtype=torch.float32)
                    c = c.contiguous()
                    self.keep += [w16, colsum, c]
                    return w16.data_ptr(), colsum.data_ptr(), c.data_ptr()
                qkv_bias = None
                if blk.attn.q_bias is not None:
                    qkv_bias = torch.cat([blk.attn.q_bias.detach(), torch.zeros_like(blk.attn.v_bias),
                                          blk.attn.v_bias.detach()])
                w.w_qkv_ln, w.s_qkv, w.c_qkv = fold(blk.attn.qkv.weight, qkv_bias, blk.norm1)
                w.w_fc1_ln, w.s_fc1, w.c_fc1 = fold(blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.norm2)
        self.keep.append(arr)
        return arr


class _Engine:
    """Device-side state of one model: f16 copies of the GEMM weights, fused qkv biases, positional tables, the
    ``cwm_vmae_model`` struct and a growable workspace.  Rebuilt whenever a parameter's storage or version
    changes (``load_state_dict``, ``.to()``, in-place edits)."""

    def __init__(self):
        self.signature = None
        self.keep = []          # tensors the struct points into
        self.model = None       # _lib.VmaeModel
        self.workspace = None
        self.aux = {}           # (B, Ntot) -> perm / inv_perm / n_visible buffers

    @staticmethod
    def _sig(module, device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in module.parameters())

    def ensure(self, module, device):
        sig = self._sig(module, device)
        if sig != self.signature:
            self._build(module, device)
            self.signature = sig
        return self.model

    def _build(self, m, device):
        pk = _Packer(device)
        keep, f32, f16, block_array = pk.keep, pk.f32, pk.f16, pk.block_array

        enc, dec = m.encoder, m.decoder
        s = _lib.VmaeModel()
        pe = enc.patch_embed
        s.in_chans = pe.proj.in_channels
        s.num_frames = m.num_frames
        s.img_h, s.img_w = int(m.image_size[-2]), int(m.image_size[-1])
        s.pt, s.ph, s.pw = [int(v) for v in m.patch_size]
        s.enc_dim, s.enc_depth = enc.embed_dim, len(enc.blocks)
        s.enc_heads = enc.blocks[0].attn.num_heads if len(enc.blocks) else 1
        s.enc_hidden = enc.blocks[0].mlp.fc1.out_features if len(enc.blocks) else enc.embed_dim
        s.dec_dim, s.dec_depth = dec.embed_dim, len(dec.blocks)
        s.dec_heads = dec.blocks[0].attn.num_heads if len(dec.blocks) else 1
        s.dec_hidden = dec.blocks[0].mlp.fc1.out_features if len(dec.blocks) else dec.embed_dim
        s.out_dim = dec.head.out_features
        s.ln_eps = float(enc.norm.eps)
        s.enc_qk_scale = float(enc.blocks[0].attn.scale) if len(enc.blocks) else 1.0
        s.dec_qk_scale = float(dec.blocks[0].attn.scale) if len(dec.blocks) else 1.0
        for blk in list(enc.blocks) + list(dec.blocks):
            if blk.attn.head_dim != 64:
                raise NotImplementedError(f"attention head_dim {blk.attn.head_dim}: only 64 is implemented on B200")
        s.w_patch = f16(pe.proj.weight.reshape(pe.proj.out_channels, -1))  # (c, kt, kh, kw) flattening
        s.b_patch = f32(pe.proj.bias)
        s.pos_enc = f32(enc.pos_embed[0])
        fold = os.environ.get("CWM_FUSE_LN", "1") != "0"  # LayerNorm folded into the consumer GEMM epilogues
        enc_arr = block_array(enc.blocks, fold_ln=fold)
        s.enc_blocks = ctypes.cast(enc_arr, ctypes.POINTER(_lib.BlockWeights))
        s.enc_norm_g, s.enc_norm_b = f32(enc.norm.weight), f32(enc.norm.bias)
        s.w_e2d = f16(m.encoder_to_decoder.weight)
        s.mask_token = f32(m.mask_token.reshape(-1))
        s.pos_dec = f32(m.pos_embed[0])
        dec_arr = block_array(dec.blocks, fold_ln=fold)
        s.dec_blocks = ctypes.cast(dec_arr, ctypes.POINTER(_lib.BlockWeights))
        s.dec_norm_g, s.dec_norm_b = f32(dec.norm.weight), f32(dec.norm.bias)
        s.w_head, s.b_head = f16(dec.head.weight), f32(dec.head.bias)
        self.keep = keep
        self.model = s
        self.workspace = None

    def get_workspace(self, nbytes, device):
        if self.workspace is None or self.workspace.numel() < nbytes or self.workspace.device != device:
            self.workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.workspace

    def get_aux(self, B, Ntot, device):
        key = (B, Ntot, str(device))
        if key not in self.aux:
            self.aux[key] = (torch.empty(B, Ntot, dtype=torch.int32, device=device),
                             torch.empty(B, Ntot, dtype=torch.int32, device=device),
                             torch.empty(B, dtype=torch.int32, device=device))
        return self.aux[key]


def compact_mask(mask, out=None):
    """a4: ``perm`` (visible ascending, then masked ascending), ``inv_perm`` and per-row visible counts for a bool
    mask [B, Ntot] (True = masked) -- bit-exact with ``torch.nonzero(~mask)`` (vmae.py:166-167, :555-556)."""
    lib = _lib.load()
    if mask.device.type != "cuda":
        raise RuntimeError("compact_mask: the mask must live on a CUDA (B200) device; there is no CPU fallback")
    B, Ntot = mask.shape
    m8 = mask.contiguous()
    if m8.dtype == torch.bool:
        m8 = m8.view(torch.uint8)
    elif m8.dtype != torch.uint8:
        m8 = (m8 != 0).view(torch.uint8)
    if out is None:
        perm = torch.empty(B, Ntot, dtype=torch.int32, device=mask.device)
        inv = torch.empty(B, Ntot, dtype=torch.int32, device=mask.device)
        nvis = torch.empty(B, dtype=torch.int32, device=mask.device)
    else:
        perm, inv, nvis = out
    stream = torch.cuda.current_stream(mask.device).cuda_stream
    _lib.check(lib.cwm_compact_mask(m8.data_ptr(), B, Ntot, perm.data_ptr(), inv.data_ptr(), nvis.data_ptr(), stream))
    return perm, inv, nvis


class PretrainVisionTransformer(nn.Module):
    """B200 implementation of cwm/models/VideoMAE/vmae.py:257-560 behind the same interface."""
    default_input_kwargs = {'unnormalize': True}

    def __init__(self,
                 img_size=224,
                 patch_size=(8, 8),
                 main_input=None,
                 main_input_kwargs=default_input_kwargs,
                 encoder_func=PretrainVisionTransformerEncoder,
                 encoder_in_chans=3,
                 encoder_num_classes=0,
                 encoder_embed_dim=768,
                 encoder_depth=12,
                 encoder_num_heads=12,
                 encoder_block_func=Block,
                 encoder_block_kwargs={},
                 decoder_num_classes=None,
                 decoder_embed_dim=512,
                 decoder_depth=8,
                 decoder_num_heads=8,
                 decoder_block_func=Block,
                 decoder_block_kwargs={},
                 mlp_ratio=4.,
                 qkv_bias=False,
                 qk_scale=None,
                 num_frames=2,
                 drop_rate=0.,
                 attn_drop_rate=0.,
                 drop_path_rate=0.,
                 norm_layer=nn.LayerNorm,
                 init_values=0.,
                 use_learnable_pos_emb=False,
                 spacetime_separable_pos_embed=False,
                 tubelet_size=1,
                 num_classes=0,
                 in_chans=0,
                 embed_per_frame=False,
                 use_flash_attention=False,
                 **kwargs):
        super().__init__()
        if main_input is not None:
            raise NotImplementedError("main_input preprocessors belong to the conjoined models (SURVEY 8a, a17)")
        if not (isinstance(encoder_func, type) and issubclass(encoder_func, PretrainVisionTransformerEncoder)):
            raise NotImplementedError("encoder_func must be PretrainVisionTransformerEncoder or a subclass (ImuEncoder)")
        if decoder_depth <= 0:
            raise NotImplementedError("decoder_depth=0 (encoder-only mode) is not exercised by any CWM factory")
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("dropout / drop-path are training-only; the B200 path is inference")
        patch_size = tuple(patch_size)
        # NB: the reference mutates the (shared default) block kwargs dicts (vmae.py:307-308); we copy instead.
        enc_kwargs = dict(encoder_block_kwargs, flash_attention=use_flash_attention)
        dec_kwargs = dict(decoder_block_kwargs, flash_attention=use_flash_attention)
        self.get_main_input = None
        self.encoder = encoder_func(
            img_size=img_size, patch_size=patch_size, in_chans=encoder_in_chans, num_classes=encoder_num_classes,
            embed_dim=encoder_embed_dim, depth=encoder_depth, num_heads=encoder_num_heads, mlp_ratio=mlp_ratio,
            qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
            drop_path_rate=drop_path_rate, norm_layer=norm_layer, init_values=init_values,
            tubelet_size=tubelet_size, use_learnable_pos_emb=use_learnable_pos_emb,
            spacetime_separable_pos_embed=spacetime_separable_pos_embed, num_frames=num_frames,
            embed_per_frame=embed_per_frame, block_func=encoder_block_func, block_kwargs=enc_kwargs, **kwargs)
        self.decoder = PretrainVisionTransformerDecoder(
            patch_size=patch_size, num_patches=self.encoder.num_patches,
            num_classes=3 * tubelet_size * (patch_size[0] * patch_size[1]) if decoder_num_classes is None
            else decoder_num_classes,
            embed_dim=decoder_embed_dim, depth=decoder_depth, num_heads=decoder_num_heads, mlp_ratio=mlp_ratio,
            qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
            drop_path_rate=drop_path_rate, norm_layer=norm_layer, init_values=init_values,
            tubelet_size=tubelet_size, block_func=decoder_block_func, block_kwargs=dec_kwargs)
        self.encoder_to_decoder = nn.Linear(encoder_embed_dim, decoder_embed_dim, bias=False)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self._learnable_pos_embed = False
        # The "spacetime separable" branch of the reference is dead code (vmae.py:422-441 would NameError on
        # `transformer`; it needs stream timestamps that are never set, vmae.py:446-449) but its Linear is a real
        # parameter of the shipped IMU checkpoints (conjoined_vmae.py:1198-1204), so it is kept for load_state_dict.
        self._spacetime_separable_pos_embed = spacetime_separable_pos_embed
        self.timestamps = None
        self.encoder.timestamps = None
        self.pos_embed = get_sinusoid_encoding_table(self.encoder.num_patches, decoder_embed_dim)  # vmae.py:366
        if self._spacetime_separable_pos_embed:
            self.pos_embed_encoder = nn.Linear(2 * decoder_embed_dim, decoder_embed_dim)  # vmae.py:368-369
        nn.init.trunc_normal_(self.mask_token, mean=0., std=.02, a=-.02, b=.02)  # vmae.py:25-26, :371
        self.num_frames = num_frames
        self.num_patches = self.encoder.num_patches
        if self.num_frames is not None:
            self.num_patches_per_frame = self.num_patches // self.num_frames
        else:
            self.num_patches_per_frame = self.num_patches
        self.patch_size = self.encoder.patch_size
        if isinstance(img_size, int):
            self.image_size = (img_size, img_size)
        else:
            assert hasattr(img_size, '__len__'), img_size
            self.image_size = img_size
        self._engine = _Engine()
        self.last_forward_launches = 0
        self.last_aux = None

    @property
    def mask_size(self):
        return (self.num_frames // self.patch_size[0],
                self.image_size[-2] // self.patch_size[-2],
                self.image_size[-1] // self.patch_size[-1])

    def get_num_layers(self):
        return len(self.encoder.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'mask_token'}

    def get_input(self, x, mask, timestamps=None, *args, **kwargs):
        return (x, mask)  # vmae.py:466-469 with main_input=None

    @torch.no_grad()
    def forward(self, x, mask, timestamps=None, *args, input_norm=None, num_visible=None, compaction=None, **kwargs):
        """``x`` float [B, C, T, H, W] (any strides), ``mask`` bool [B, Ntot] (True = masked) ->
        float32 [B, Nmask, D] predicted patches of the masked tokens in ascending token order (vmae.py:539-560).

        Extensions (keyword-only, not in the reference): ``input_norm=(mean, std)`` fuses ``imagenet_normalize``
        of the raw input into the patch gather; ``num_visible`` skips the device->host read of the per-row
        visible count (the caller vouches that every row has exactly that many visible tokens); ``compaction`` is the
        ``(perm, inv_perm, counts)`` triple ``compact_mask`` already returned for exactly this mask (the wrapper runs
        the compaction once to read the counts for its rectangulariser and hands the result on)."""
        lib = _lib.load()
        if x.device.type != "cuda":
            raise RuntimeError(
                "counterfactualworldmodels_b200: the VMAE forward only runs on a CUDA sm_100 (B200) device; "
                "there is no CPU or PyTorch fallback")
        assert x.dim() == 5, x.shape
        from .perturbation import CounterfactualVideo
        cf = x if isinstance(x, CounterfactualVideo) else None
        if cf is not None:   # virtual video [S, T, C, H, W], consumed as its [S, C, T, H, W] view
            B, T, C, H, W = x.shape
        else:
            B, C, T, H, W = x.shape
        self.device = x.device
        pt, ph, pw = self.patch_size
        assert (H % ph == 0) and (W % pw == 0), \
            f"Input image size({H},{W}) must be divisible by patch size ({ph},{pw})"
        assert T == self.num_frames, (T, self.num_frames)
        if (H, W) != (int(self.image_size[-2]), int(self.image_size[-1])):
            raise NotImplementedError(
                f"input size {(H, W)} differs from the model's image_size {tuple(self.image_size)}: the sinusoid "
                "tables are built for num_patches tokens (the reference fails the same way at vmae.py:165)")
        if cf is None and x.dtype != torch.float32:
            x = x.float()
        Ntot = self.num_patches
        mask = mask.reshape(B, -1)
        assert mask.shape[1] == Ntot, (mask.shape, Ntot)
        if mask.device != x.device:
            mask = mask.to(x.device)
        with torch.cuda.device(x.device):
            model = self._engine.ensure(self, x.device)
            aux = self._engine.get_aux(B, Ntot, x.device)
            if compaction is not None:
                perm, inv, nvis = compaction
                assert perm.shape == (B, Ntot) and perm.device == x.device, (perm.shape, perm.device)
            else:
                perm, inv, nvis = compact_mask(mask, out=aux)
            if num_visible is None:
                counts = nvis.cpu()
                n_vis = int(counts[0]) if B > 0 else 0
                if B > 0 and not bool((counts == n_vis).all()):
                    # same failure the reference hits at `x[~mask].reshape(B, -1, C)` (vmae.py:167)
                    raise RuntimeError(
                        f"shape '[{B}, -1, {self.encoder.embed_dim}]' is invalid: rows of the mask have different "
                        f"numbers of visible tokens {counts.tolist()} (rectangularize the masks first)")
            else:
                n_vis = int(num_visible)
            n_out = (Ntot - n_vis) if n_vis < Ntot else Ntot
            y = torch.empty(B, n_out, self.decoder.num_classes, dtype=torch.float32, device=x.device)
            ws_bytes = lib.cwm_vmae_workspace_bytes(ctypes.byref(model), B, n_vis)
            if ws_bytes == 0 and B > 0:
                _lib.check(-1)
            ws = self._engine.get_workspace(ws_bytes, x.device)
            mean = std = None
            if input_norm is not None:
                mean, std = _lib.float_array(input_norm[0]), _lib.float_array(input_norm[1])
            stream = torch.cuda.current_stream(x.device).cuda_stream
            if cf is not None:
                src, keep = cf.c_struct()
                _lib.check(lib.cwm_vmae_forward_cf(ctypes.byref(model), ctypes.byref(src), B, mean, std, perm.data_ptr(),
                                                   n_vis, y.data_ptr(), ws.data_ptr(), ws.numel(), stream))
                del keep
            else:
                _lib.check(lib.cwm_vmae_forward(ctypes.byref(model), x.data_ptr(), _lib.strides5(x), B, mean, std,
                                                perm.data_ptr(), n_vis, y.data_ptr(), ws.data_ptr(), ws.numel(),
                                                stream))
            # + the compaction kernel when it ran here
            self.last_forward_launches = lib.cwm_last_forward_launches() + (1 if compaction is None else 0)
            self.last_aux = (perm, inv, n_vis)
        return y


# ---- factories (vmae.py:563-619) ------------------------------------------------------------------------

def _cfg(url='', **kwargs):
    return {'url': url, 'num_classes': 400, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9,
            'interpolation': 'bicubic', 'mean': (0.5, 0.5, 0.5), 'std': (0.5, 0.5, 0.5), **kwargs}


def pretrain_videomae_large_224_scaffold(**kwargs):
    model = PretrainVisionTransformer(
        img_size=224, encoder_embed_dim=1024, encoder_depth=24, encoder_num_heads=16, encoder_num_classes=0,
        decoder_embed_dim=512, decoder_num_heads=8, decoder_depth=12, mlp_ratio=4, qkv_bias=True,
        norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg()
    return model


def pretrain_videomae_base_224_scaffold(**kwargs):
    model = PretrainVisionTransformer(
        img_size=224, encoder_embed_dim=768, encoder_depth=12, encoder_num_heads=12, encoder_num_classes=0,
        decoder_embed_dim=384, decoder_num_heads=6, decoder_depth=4, mlp_ratio=4, qkv_bias=True,
        norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg()
    return model


def base_16x16patch_2frames_1tube(**kwargs):
    return pretrain_videomae_base_224_scaffold(patch_size=(16, 16), num_frames=2, tubelet_size=1, **kwargs)


def base_8x8patch_2frames_1tube(**kwargs):
    return pretrain_videomae_base_224_scaffold(patch_size=(8, 8), num_frames=2, tubelet_size=1, **kwargs)


def base_4x4patch_2frames_1tube(**kwargs):
    """BASELINE config 3; the reference has no named factory for it."""
    return pretrain_videomae_base_224_scaffold(patch_size=(4, 4), num_frames=2, tubelet_size=1, **kwargs)


def large_4x4patch_2frames_1tube(**kwargs):
    return pretrain_videomae_large_224_scaffold(patch_size=(4, 4), num_frames=2, tubelet_size=1, **kwargs)
