"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference algorithm for the CWM VMAE forward path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / the CPU baseline -- never as the product path.

PARITY PIN: the reference has no tests or golden vectors for this path, so this oracle is
pinned against the reference module itself: ``oracle/make_golden.py`` (run in the build container, where
/root/reference is mounted) checks every function below against ``cwm.models.VideoMAE.vmae`` /
``cwm.models.prediction`` on identical weights, inputs and masks and writes the fixtures under ``tests/golden/``;
``tests/test_oracle.py`` re-checks the oracle against those fixtures everywhere (including the GPU box).
The known answers the reference's notebook records (parameter counts, token counts, mask counts) are pinned
in ``tests/test_known_answers.py``.

Integer steps are numpy (bit-exact); floating-point steps are plain torch CPU ops in float32 (or float64 when
``dtype=torch.float64``), one statement per reference line.  Citations are relative to /root/reference.
"""
import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)  # cwm/models/utils.py:12
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)   # cwm/models/utils.py:13


# ---------------------------------------------------------------------------------------------------------
# integer steps (bit-exact)
# ---------------------------------------------------------------------------------------------------------
def compact_mask(mask):
    """vmae.py:166-167 (`x[~mask]`), :555-557 (`[~mask]` then `[mask]`, concatenated visible-first).
    mask: bool array [B, Ntot], True = masked.  Returns perm [B, Ntot] int32 (visible token indices ascending,
    then masked ascending), inv_perm [B, Ntot] int32 and n_visible [B] int32."""
    mask = np.asarray(mask).astype(bool)
    B, N = mask.shape
    perm = np.zeros((B, N), np.int32)
    inv = np.zeros((B, N), np.int32)
    nvis = np.zeros((B,), np.int32)
    for b in range(B):
        vis = np.nonzero(~mask[b])[0]
        msk = np.nonzero(mask[b])[0]
        perm[b] = np.concatenate([vis, msk]).astype(np.int32)
        inv[b, perm[b]] = np.arange(N, dtype=np.int32)
        nvis[b] = len(vis)
    return perm, inv, nvis


def sinusoid_table(n_position, d_hid):
    """VideoMAE/utils.py:251-268, literally (python loops, float64), for pinning the vectorised product version."""
    def get_position_angle_vec(position):
        return [position / np.power(10000, 2 * (hid_j // 2) / d_hid) for hid_j in range(d_hid)]
    table = np.array([get_position_angle_vec(p) for p in range(n_position)])
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.FloatTensor(table).unsqueeze(0)


# ---------------------------------------------------------------------------------------------------------
# floating-point steps
# ---------------------------------------------------------------------------------------------------------
def preprocess(x, imagenet_normalize=True, t_dim=2, c_dim=1):
    """prediction.py:304-312 + models/utils.py:15-21.  x [B,T,C,H,W] -> [B,C,T,H,W] (a transposed view)."""
    x = x.transpose(t_dim, c_dim)
    if imagenet_normalize:
        mean = torch.as_tensor(IMAGENET_DEFAULT_MEAN)[None, None, :, None, None].to(x).transpose(1, 2)
        std = torch.as_tensor(IMAGENET_DEFAULT_STD)[None, None, :, None, None].to(x).transpose(1, 2)
        x = (x - mean) / std
    return x


def _round(t, operand_dtype):
    """Emulates feeding a GEMM/attention operand in a 16-bit type (fp32 accumulate).  None = exact."""
    return t if operand_dtype is None else t.to(operand_dtype).to(t.dtype)


def _linear(x, w, b, operand_dtype):
    return F.linear(_round(x, operand_dtype), _round(w, operand_dtype), b)


def block_forward(x, sd, prefix, num_heads, eps, operand_dtype=None, taps=None):
    """VideoMAE/utils.py:146-153 (gamma_* optional); Attention :87-121; Mlp :47-54."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], eps)
    qkv_bias = None
    if (prefix + "attn.q_bias") in sd:
        q_bias, v_bias = sd[prefix + "attn.q_bias"], sd[prefix + "attn.v_bias"]
        qkv_bias = torch.cat((q_bias, torch.zeros_like(v_bias), v_bias))          # utils.py:91
    qkv = _linear(h, sd[prefix + "attn.qkv.weight"], qkv_bias, operand_dtype)      # utils.py:93
    qkv = qkv.reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)               # utils.py:94
    q, k, v = qkv[0], qkv[1], qkv[2]
    head_dim = q.shape[-1]
    q = q * (head_dim ** -0.5)                                                      # utils.py:67,97
    attn = _round(q, operand_dtype) @ _round(k, operand_dtype).transpose(-2, -1)    # utils.py:108
    attn = attn.softmax(dim=-1)                                                     # utils.py:111
    a = _round(attn, operand_dtype) @ _round(v, operand_dtype)                      # utils.py:113
    a = a.transpose(1, 2).reshape(B, N, -1)                                         # utils.py:118
    if taps is not None:
        taps[prefix + "attn.core"] = a
    a = _linear(a, sd[prefix + "attn.proj.weight"], sd[prefix + "attn.proj.bias"], operand_dtype)  # :119
    if (prefix + "gamma_1") in sd:                                                  # layer scale, utils.py:151
        a = sd[prefix + "gamma_1"] * a
    x = x + a                                                                       # utils.py:148
    h = F.layer_norm(x, (C,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], eps)
    h = _linear(h, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"], operand_dtype)      # :48
    h = F.gelu(h)                                                                   # utils.py:49 (erf GELU)
    h = _linear(h, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"], operand_dtype)      # :52
    if (prefix + "gamma_2") in sd:                                                  # utils.py:152
        h = sd[prefix + "gamma_2"] * h
    x = x + h                                                                       # utils.py:149
    return x


def vmae_forward(sd, x, mask, cfg, operand_dtype=None, taps=None, dtype=torch.float32):
    """`PretrainVisionTransformer.forward` (vmae.py:539-560) on a state_dict.
    sd:   reference state_dict (fp32 tensors)
    x:    [B, C, T, H, W] (already preprocessed), mask: bool [B, Ntot]
    cfg:  dict(patch_size=(pt,ph,pw), enc_heads, dec_heads, eps)
    Returns [B, Nmask, D]."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    x = x.to(dtype)
    pt, ph, pw = cfg["patch_size"]
    eps = cfg.get("eps", 1e-6)
    B, C, T, H, W = x.shape
    w = sd["encoder.patch_embed.proj.weight"]                                       # [Ce, C, pt, ph, pw]
    Ce = w.shape[0]
    # PatchEmbed: Conv3d k = s = (pt,ph,pw) -> flatten(2).transpose(1,2)  (VideoMAE/utils.py:174-176,197)
    tok = F.conv3d(_round(x, operand_dtype), _round(w, operand_dtype), sd["encoder.patch_embed.proj.bias"],
                   stride=(pt, ph, pw))
    tok = tok.flatten(2).transpose(1, 2)                                            # [B, Ntot, Ce], order (t,h,w)
    Ntot = tok.shape[1]
    if "encoder.pos_embed" in sd:                                                   # use_learnable_pos_emb, vmae.py:68-70
        pos_e = sd["encoder.pos_embed"]
    else:
        pos_e = sinusoid_table_cached(Ntot, Ce).to(dtype)                           # vmae.py:75,162
    tok = tok + pos_e                                                               # vmae.py:165
    if taps is not None:
        taps["tokens"] = tok
    mask = torch.as_tensor(np.asarray(mask)).bool()
    x_vis = tok[~mask].reshape(B, -1, Ce)                                           # vmae.py:167
    n_enc = len({k.split(".")[2] for k in sd if k.startswith("encoder.blocks.")})
    for i in range(n_enc):                                                          # vmae.py:169-170
        x_vis = block_forward(x_vis, sd, f"encoder.blocks.{i}.", cfg["enc_heads"], eps, operand_dtype, taps)
        if taps is not None:
            taps[f"encoder.blocks.{i}"] = x_vis
    x_vis = F.layer_norm(x_vis, (Ce,), sd["encoder.norm.weight"], sd["encoder.norm.bias"], eps)   # vmae.py:172
    x_vis = _linear(x_vis, sd["encoder_to_decoder.weight"], None, operand_dtype)    # vmae.py:547
    Cd = x_vis.shape[-1]
    pos_d = sinusoid_table_cached(Ntot, Cd).to(dtype).expand(B, -1, -1)             # vmae.py:366,554
    pos_vis = pos_d[~mask].reshape(B, -1, Cd)                                       # vmae.py:555
    pos_mask = pos_d[mask].reshape(B, -1, Cd)                                       # vmae.py:556
    x_full = torch.cat([x_vis + pos_vis, sd["mask_token"] + pos_mask], dim=1)       # vmae.py:557
    if taps is not None:
        taps["decoder.input"] = x_full
    n_dec = len({k.split(".")[2] for k in sd if k.startswith("decoder.blocks.")})
    for i in range(n_dec):                                                          # vmae.py:247-248
        x_full = block_forward(x_full, sd, f"decoder.blocks.{i}.", cfg["dec_heads"], eps, operand_dtype, taps)
    n_ret = pos_mask.shape[1]
    if n_ret > 0:                                                                   # vmae.py:250-253
        x_full = x_full[:, -n_ret:]
    y = F.layer_norm(x_full, (Cd,), sd["decoder.norm.weight"], sd["decoder.norm.bias"], eps)
    y = _linear(y, sd["decoder.head.weight"], sd["decoder.head.bias"], operand_dtype)
    return y


_TABLES = {}


def sinusoid_table_cached(n, d):
    if (n, d) not in _TABLES:
        pos = np.arange(n, dtype=np.float64)
        j = np.arange(d)
        t = pos[:, None] / np.power(10000, 2 * (j // 2) / d)[None, :]
        t[:, 0::2] = np.sin(t[:, 0::2])
        t[:, 1::2] = np.cos(t[:, 1::2])
        _TABLES[(n, d)] = torch.FloatTensor(t).unsqueeze(0)
    return _TABLES[(n, d)]


def patchify(x, patch_size):
    """patches.py:67-74 with temporal_dim=1: 'b (t pt) c (h ph) (w pw) -> b (t h w) (pt ph pw c)'."""
    pt, ph, pw = patch_size
    B, T, C, H, W = x.shape
    x = x.reshape(B, T // pt, pt, C, H // ph, ph, W // pw, pw)
    x = x.permute(0, 1, 4, 6, 2, 5, 7, 3)        # b t h w pt ph pw c
    return x.reshape(B, (T // pt) * (H // ph) * (W // pw), pt * ph * pw * C)


def unpatchify(p, patch_size, shape):
    """patches.py:76-109: 'b (t h w) (pt ph pw) c -> b c (t pt) (h ph) (w pw)' then transpose(1,2)."""
    pt, ph, pw = patch_size
    B, T, C, H, W = shape
    p = p.reshape(B, T // pt, H // ph, W // pw, pt, ph, pw, C)
    p = p.permute(0, 1, 4, 7, 2, 5, 3, 6)        # b t pt c h ph w pw
    return p.reshape(B, T, C, H, W)


def pred_patches_to_video(y, x_raw, mask, patch_size):
    """prediction.py:245-259: raw input at visible positions, predictions at masked positions.
    y [B, Nmask, D]; x_raw [B, T, C, H, W]; mask bool [B, Ntot]."""
    mask = torch.as_tensor(np.asarray(mask)).bool()
    xp = patchify(x_raw, patch_size).to(y.dtype)
    out = torch.zeros_like(xp)
    D = xp.shape[-1]
    out[~mask] = xp[~mask].view(-1, D)
    out[mask] = y.reshape(-1, D)
    return unpatchify(out, patch_size, x_raw.shape)


def predict(sd, x_raw, mask, cfg, imagenet_normalize=True, frame=None, operand_dtype=None):
    """`PredictorBasedGenerator.predict` (prediction.py:406-454) for an already-rectangular mask."""
    y = vmae_forward(sd, preprocess(x_raw, imagenet_normalize), mask, cfg, operand_dtype=operand_dtype)
    v = pred_patches_to_video(y, x_raw, mask, cfg["patch_size"])
    if frame is not None:
        frame = frame % v.shape[1]
        v = v[:, frame:frame + 1]
    return v
