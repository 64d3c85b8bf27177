"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's padded / conjoined (IMU-conditioned) VMAE forward
(SURVEY.md section 8a rows a13-a17).  Same rules as ``vmae_oracle.py``: only ``tests/``, ``smoke()`` and the CPU
legs of ``bench.py`` may import it, never the product path.

PARITY PIN: ``oracle/make_golden_conjoined.py`` (build container, /root/reference mounted) checks every function here
against the real ``cwm.models.VideoMAE.conjoined_vmae`` modules on identical weights / inputs / masks (observed
max-abs difference < 2e-5, index arrays bit-identical) and writes ``tests/golden/conj_*.npz``; ``tests/test_oracle.py``
re-checks the oracle against those fixtures everywhere.  Known answers recorded by the reference's notebook
(148,265,040 and 135,730,048 parameters, 25 IMU tokens) are pinned in ``tests/test_known_answers.py``.

Plain torch CPU float32, one statement per reference line; citations relative to /root/reference/cwm/models.
"""
import numpy as np
import torch
import torch.nn.functional as F

import vmae_oracle as vo


def pos_embedding(n, dim):
    """transformer.py:37-52 (torch float32 -- differs from the numpy float64 table by up to 4.5e-4)."""
    positions = torch.arange(n).float()
    freqs = torch.arange(dim).float()
    freqs = torch.pow(10000, 2 * (torch.div(freqs, 2, rounding_mode='trunc')) / dim)
    out = positions[:, None] / freqs[None, :]
    out[:, 0::2] = torch.sin(out[:, 0::2])
    out[:, 1::2] = torch.cos(out[:, 1::2])
    return out.unsqueeze(0)


def padding_masks(mask, max_pad, min_pad=0):
    """VideoMAE/conjoined_vmae.py:49-116.  mask bool [B, N] (True = masked) ->
    padding_mask [B, P] (True = pad position NOT used), full_input_mask [B, N+P], null_mask [B, N+P-maxvis-min_pad]."""
    mask = torch.as_tensor(np.asarray(mask)).bool()
    B, N = mask.shape
    num_visible = torch.sum((~mask).int(), -1, keepdim=True)                      # :60
    max_vis = torch.max(num_visible)                                              # :62
    num_pad = max_vis - num_visible + min_pad                                     # :64
    padding_mask = torch.arange(max_pad)[None].expand(B, -1) < num_pad            # :65-67
    null_padding = torch.cat([torch.ones((B, 1), dtype=torch.bool),
                              torch.zeros((B, max_pad - 1), dtype=torch.bool)], -1)  # :69-72
    any_visible = (torch.sum(num_visible.float()) > 0).reshape(1, 1).expand(B, max_pad)  # :74-75
    padding_mask = torch.where(any_visible, padding_mask, null_padding)           # :81-82
    max_vis = torch.maximum(max_vis, torch.ones_like(max_vis))                    # :85
    padding_mask = ~padding_mask                                                  # :87
    min_masked = N - int(max_vis) - min_pad                                       # :88
    full_input_mask = torch.cat([mask, padding_mask], -1)                         # :99
    null_mask = torch.cat([torch.zeros_like(mask[:, :min_masked]), padding_mask], -1)  # :105
    return padding_mask, full_input_mask, null_mask


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _n_blocks(sd, prefix):
    return len({k[len(prefix):].split(".")[0] for k in sd if k.startswith(prefix)})


def tokenize(sd, x, scfg):
    """`encoder.tokenize` (VideoMAE/vmae.py:129-150; ImuEncoder.tokenize conjoined_vmae.py:1110-1125): Conv3d patch
    embedding + the stream's encoder positional table.  x [B,C,T,H,W] -> [B, Ntok, Ce]."""
    w = sd["encoder.patch_embed.proj.weight"]
    if scfg.get("dummy", False):
        x = torch.cat([x, sd["encoder.dummy_token"].expand(x.shape[0], -1, -1, -1, -1)], 2)   # :1122
    tok = F.conv3d(x, w, sd["encoder.patch_embed.proj.bias"], stride=tuple(w.shape[2:]))
    tok = tok.flatten(2).transpose(1, 2)
    n, Ce = tok.shape[1], tok.shape[2]
    pos = pos_embedding(n, Ce) if scfg["pos"] == "torch" else vo.sinusoid_table_cached(n, Ce)
    return tok + pos


def decoder_pos(sd, n, scfg):
    """Decoder table of a stream: numpy-f64 sinusoid (vmae.py:366) or, for an IMU context stream inside a conjoined
    forward, the torch-f32 `pos_embedding` (+1 row for the dummy token; vmae.py:446-449, conjoined_vmae.py:595-609)."""
    Cd = sd["mask_token"].shape[-1]
    return pos_embedding(n, Cd) if scfg["pos"] == "torch" else vo.sinusoid_table_cached(n, Cd)


def cross_block(x, src, sd, prefix, heads, eps):
    """transformer.py:559-583 (with_self_attention=False: gamma_1 = 0, norm1 = Identity) around
    BidirectionalCrossAttention.forward (:314-378)."""
    B, N, C = x.shape
    _, M, Cs = src.shape
    g = lambda k: sd[prefix + k]
    hx = F.layer_norm(x, (C,), g("norm1_cross.weight"), g("norm1_cross.bias"), eps)               # :541
    hs = F.layer_norm(src, (Cs,), g("norm1_src_cross.weight"), g("norm1_src_cross.bias"), eps)    # :542
    qk = F.linear(hx, g("cross_attention.qk.weight"))                                             # :333
    qk_s = F.linear(hs, g("cross_attention.qk_src.weight"))                                       # :334
    v = F.linear(hx, g("cross_attention.v.weight"))                                               # :335
    v_s = F.linear(hs, g("cross_attention.v_src.weight"))                                         # :336
    D = v.shape[-1]
    hd = D // heads
    scale = hd ** -0.5                                                                            # :273
    r = lambda t: t.reshape(t.shape[0], t.shape[1], heads, -1).permute(0, 2, 1, 3)                # :339 'b n (h d) -> b h n d'
    qk, qk_s, v, v_s = map(r, (qk, qk_s, v, v_s))
    attn = torch.einsum('bhnd,bhmd->bhnm', qk[..., 0:hd] * scale, qk_s[..., 0:hd]).softmax(-1)     # :358-361
    attn_s = torch.einsum('bhnd,bhmd->bhmn', qk[..., hd:] * scale, qk_s[..., hd:]).softmax(-1)     # :362-365
    ur = lambda t: t.permute(0, 2, 1, 3).reshape(t.shape[0], t.shape[2], -1)                       # :369
    y = ur(attn @ v_s)                                                                            # :370
    y_s = ur(attn_s @ v)                                                                          # :371
    y = F.linear(y, g("cross_attention.projection.weight"), g("cross_attention.projection.bias"))          # :374
    y_s = F.linear(y_s, g("cross_attention.projection_src.weight"), g("cross_attention.projection_src.bias"))  # :375
    x = x + 0.0 * x + 1.0 * y                                                                     # :569-571
    src = src + 0.0 * src + 1.0 * y_s                                                             # :573-575
    h = F.layer_norm(x, (C,), g("norm2.weight"), g("norm2.bias"), eps)
    h = F.linear(F.gelu(F.linear(h, g("mlp.trg.layers.0.weight"), g("mlp.trg.layers.0.bias"))),
                 g("mlp.trg.layers.2.weight"), g("mlp.trg.layers.2.bias"))
    x = x + h                                                                                     # :578
    h = F.layer_norm(src, (Cs,), g("norm2_src.weight"), g("norm2_src.bias"), eps)
    h = F.linear(F.gelu(F.linear(h, g("mlp.src.layers.0.weight"), g("mlp.src.layers.0.bias"))),
                 g("mlp.src.layers.2.weight"), g("mlp.src.layers.2.bias"))
    src = src + h                                                                                 # :580
    return x, src


def _stream_prepare(sd, x, mask, scfg):
    """Tokens + masks of one stream: dummy token / padding positions appended.
    Returns tokens [B, Next, Ce] (pad rows = null_token_enc), full mask [B, Next], null_mask or None, n real tokens."""
    mask = torch.as_tensor(np.asarray(mask)).bool()
    tok = tokenize(sd, x, scfg)
    if scfg.get("dummy", False):
        mask = torch.cat([mask, torch.zeros_like(mask[:, -1:])], -1)              # conjoined_vmae.py:1123, :597-600
    n_real = tok.shape[1]
    null_mask = None
    if scfg.get("max_pad", 0) > 0:
        _, full, null_mask = padding_masks(mask, scfg["max_pad"], scfg.get("min_pad", 0))
        tok = torch.cat([tok, sd["null_token_enc"].expand(tok.shape[0], scfg["max_pad"], -1)], 1)   # :130-131
        mask = full
    return tok, mask, null_mask, n_real


def _decoder_input(sd, x_vis, mask, n_real, scfg):
    """conjoined_vmae.py:154-165, :620-635 / :956-977."""
    B = x_vis.shape[0]
    Cd = sd["mask_token"].shape[-1]
    pos = decoder_pos(sd, n_real, scfg).expand(B, -1, -1)
    if scfg.get("max_pad", 0) > 0:
        pos = torch.cat([pos, sd["null_token_dec"].expand(B, scfg["max_pad"], -1)], 1)            # :157-160
    pos_vis = pos[~mask].reshape(B, -1, Cd)
    pos_mask = pos[mask].reshape(B, -1, Cd)
    return torch.cat([x_vis + pos_vis, sd["mask_token"] + pos_mask], 1), pos_mask.shape[1]


def _last_tokens(sd, x, n_ret, eps):
    """vmae.py:238-244."""
    Cd = x.shape[-1]
    if n_ret > 0:
        x = x[:, -n_ret:]
    elif n_ret == 0:
        x = x[:, x.shape[1]:]
    y = F.layer_norm(x, (Cd,), sd["decoder.norm.weight"], sd["decoder.norm.bias"], eps)
    return F.linear(y, sd["decoder.head.weight"], sd["decoder.head.bias"])


def padded_forward(sd, x, mask, scfg, taps=None):
    """`PaddedVisionTransformer.forward` (conjoined_vmae.py:189-210)."""
    eps = scfg.get("eps", 1e-6)
    tok, full, null_mask, n_real = _stream_prepare(sd, x, mask, scfg)
    B, _, Ce = tok.shape
    x_vis = tok[~full].reshape(B, -1, Ce)                                         # :145
    for i in range(_n_blocks(sd, "encoder.blocks.")):
        x_vis = vo.block_forward(x_vis, sd, f"encoder.blocks.{i}.", scfg["enc_heads"], eps)
    x_vis = F.layer_norm(x_vis, (Ce,), sd["encoder.norm.weight"], sd["encoder.norm.bias"], eps)
    x_vis = F.linear(x_vis, sd["encoder_to_decoder.weight"])                      # :197
    x_full, n_ret = _decoder_input(sd, x_vis, full, n_real, scfg)
    for i in range(_n_blocks(sd, "decoder.blocks.")):
        x_full = vo.block_forward(x_full, sd, f"decoder.blocks.{i}.", scfg["dec_heads"], eps)
    y = _last_tokens(sd, x_full, n_ret if n_ret > 0 else -1, eps)                 # decoder.forward, vmae.py:250-253
    return y * ((~null_mask)[..., None].to(y))                                   # :208


def conjoined_forward(sd, x_main, mask_main, x_ctx, mask_ctx, cfg, output_main=True, output_context=False, taps=None):
    """`Conjoined(Padded)VisionTransformer.forward` (conjoined_vmae.py:852-887, :918-1011) on already-selected stream
    inputs: x_main [B,C,T,H,W], x_ctx [B,6,L,1,1], masks bool [B, Ntok] per stream.
    cfg: dict(main=stream cfg, ctx=stream cfg, enc_pairs=[(i,j)...], dec_pairs=[...], cross_heads, eps)."""
    eps = cfg.get("eps", 1e-6)
    sm, sc = _sub(sd, "main_stream."), _sub(sd, "context_stream.")
    mcfg, ccfg = cfg["main"], cfg["ctx"]
    tok_m, full_m, null_m, nreal_m = _stream_prepare(sm, x_main, mask_main, mcfg)
    tok_c, full_c, null_c, nreal_c = _stream_prepare(sc, x_ctx, mask_ctx, ccfg)
    B = tok_m.shape[0]
    x = tok_m[~full_m].reshape(B, -1, tok_m.shape[-1])                            # :133 / vmae.py:149
    xc = tok_c[~full_c].reshape(B, -1, tok_c.shape[-1])
    if taps is not None:
        taps["enc_in_main"], taps["enc_in_ctx"] = x, xc
    # encoder blocks, cross block BEFORE the paired blocks (:543-576)
    i = j = 0
    for (pi, pj) in cfg["enc_pairs"]:
        while i < pi:
            x = vo.block_forward(x, sm, f"encoder.blocks.{i}.", mcfg["enc_heads"], eps)
            i += 1
        while j < pj:
            xc = vo.block_forward(xc, sc, f"encoder.blocks.{j}.", ccfg["enc_heads"], eps)
            j += 1
        x, xc = cross_block(x, xc, sd, f"encoder_conjoining_blocks.{pi}-{pj}.", cfg["cross_heads"], eps)
    for _i in range(i, _n_blocks(sm, "encoder.blocks.")):
        x = vo.block_forward(x, sm, f"encoder.blocks.{_i}.", mcfg["enc_heads"], eps)
    for _j in range(j, _n_blocks(sc, "encoder.blocks.")):
        xc = vo.block_forward(xc, sc, f"encoder.blocks.{_j}.", ccfg["enc_heads"], eps)
    x = F.layer_norm(x, (x.shape[-1],), sm["encoder.norm.weight"], sm["encoder.norm.bias"], eps)      # :574
    xc = F.layer_norm(xc, (xc.shape[-1],), sc["encoder.norm.weight"], sc["encoder.norm.bias"], eps)   # :575
    if taps is not None:
        taps["enc_out_main"], taps["enc_out_ctx"] = x, xc
    x = F.linear(x, sm["encoder_to_decoder.weight"])                              # :876
    xc = F.linear(xc, sc["encoder_to_decoder.weight"])                            # :877
    x, nret_m = _decoder_input(sm, x, full_m, nreal_m, mcfg)
    xc, nret_c = _decoder_input(sc, xc, full_c, nreal_c, ccfg)
    if taps is not None:
        taps["dec_in_main"], taps["dec_in_ctx"] = x, xc
    # decoder blocks, cross block AFTER the paired blocks (:688-720)
    i = j = 0
    for (pi, pj) in cfg["dec_pairs"]:
        while i <= pi:
            x = vo.block_forward(x, sm, f"decoder.blocks.{i}.", mcfg["dec_heads"], eps)
            i += 1
        while j <= pj:
            xc = vo.block_forward(xc, sc, f"decoder.blocks.{j}.", ccfg["dec_heads"], eps)
            j += 1
        x, xc = cross_block(x, xc, sd, f"decoder_conjoining_blocks.{pi}-{pj}.", cfg["cross_heads"], eps)
    for _i in range(i, _n_blocks(sm, "decoder.blocks.")):
        x = vo.block_forward(x, sm, f"decoder.blocks.{_i}.", mcfg["dec_heads"], eps)
    for _j in range(j, _n_blocks(sc, "decoder.blocks.")):
        xc = vo.block_forward(xc, sc, f"decoder.blocks.{_j}.", ccfg["dec_heads"], eps)
    y = _last_tokens(sm, x, nret_m, eps)                                          # :675 / :989
    yc = _last_tokens(sc, xc, nret_c, eps)                                        # :677 / :991
    if null_m is not None:
        y = y * ((~null_m)[..., None].to(y))                                      # :999
    if null_c is not None:
        yc = yc * ((~null_c)[..., None].to(yc))                                   # :1002
    if output_main and output_context:
        return y, yc
    return y if output_main else yc
