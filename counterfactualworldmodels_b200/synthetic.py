"""Deterministic synthetic weights, videos and temporally-factored masks.

Shared by the tests, ``bench.py``, ``__graft_entry__.smoke()`` and ``oracle/make_golden.py`` so that the fixtures
under ``tests/golden/`` (produced by the reference in the build container) can be re-derived on the GPU box from
seeds alone: everything here is generated on the CPU with an explicit ``torch.Generator``.
"""
import math

import torch

# name -> constructor kwargs of PretrainVisionTransformer (ours or the reference's: same signature)
CONFIGS = {
    # small shapes that keep head_dim = 64 and exercise ragged tiles; cheap on the CPU oracle
    "tiny_4x4": dict(img_size=32, patch_size=(4, 4), encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2,
                     decoder_embed_dim=128, decoder_depth=1, decoder_num_heads=2, mlp_ratio=4, qkv_bias=True,
                     num_frames=2, tubelet_size=1),
    "tiny_8x8": dict(img_size=64, patch_size=(8, 8), encoder_embed_dim=256, encoder_depth=2, encoder_num_heads=4,
                     decoder_embed_dim=128, decoder_depth=2, decoder_num_heads=2, mlp_ratio=4, qkv_bias=True,
                     num_frames=2, tubelet_size=1),
    # 448 tokens per frame: more than one 256-row attention CTA and ragged 128-row tiles
    "small_4x4": dict(img_size=(64, 112), patch_size=(4, 4), encoder_embed_dim=192 + 64, encoder_depth=2,
                      encoder_num_heads=4, decoder_embed_dim=128, decoder_depth=2, decoder_num_heads=2, mlp_ratio=4,
                      qkv_bias=True, num_frames=2, tubelet_size=1),
    # constructor options no BASELINE config uses but the reference supports (last paragraph)
    "tiny_4x4_tube2": dict(img_size=32, patch_size=(4, 4), encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2,
                           decoder_embed_dim=128, decoder_depth=1, decoder_num_heads=2, mlp_ratio=4, qkv_bias=True,
                           num_frames=4, tubelet_size=2),
    "tiny_8x8_layerscale_learnpos": dict(img_size=64, patch_size=(8, 8), encoder_embed_dim=256, encoder_depth=2,
                                         encoder_num_heads=4, decoder_embed_dim=128, decoder_depth=2,
                                         decoder_num_heads=2, mlp_ratio=4, qkv_bias=True, num_frames=2, tubelet_size=1,
                                         init_values=0.5, use_learnable_pos_emb=True),
    # the BASELINE.json configurations (vmae.py:580-619)
    "base_8x8": dict(img_size=224, patch_size=(8, 8), encoder_embed_dim=768, encoder_depth=12, encoder_num_heads=12,
                     decoder_embed_dim=384, decoder_depth=4, decoder_num_heads=6, mlp_ratio=4, qkv_bias=True,
                     num_frames=2, tubelet_size=1),
    "base_4x4": dict(img_size=224, patch_size=(4, 4), encoder_embed_dim=768, encoder_depth=12, encoder_num_heads=12,
                     decoder_embed_dim=384, decoder_depth=4, decoder_num_heads=6, mlp_ratio=4, qkv_bias=True,
                     num_frames=2, tubelet_size=1),
    "large_4x4": dict(img_size=224, patch_size=(4, 4), encoder_embed_dim=1024, encoder_depth=24,
                      encoder_num_heads=16, decoder_embed_dim=512, decoder_depth=12, decoder_num_heads=8,
                      mlp_ratio=4, qkv_bias=True, num_frames=2, tubelet_size=1),
}


def model_kwargs(name):
    from functools import partial
    import torch.nn as nn
    kw = dict(CONFIGS[name])
    kw["norm_layer"] = partial(nn.LayerNorm, eps=1e-6)  # vmae.py:575,592
    kw["encoder_num_classes"] = 0
    return kw


def oracle_cfg(name):
    kw = CONFIGS[name]
    return dict(patch_size=(kw["tubelet_size"],) + tuple(kw["patch_size"]), enc_heads=kw["encoder_num_heads"],
                dec_heads=kw["decoder_num_heads"], eps=1e-6)


def image_hw(name):
    s = CONFIGS[name]["img_size"]
    return (s, s) if isinstance(s, int) else tuple(s)


def mask_size(name):
    kw = CONFIGS[name]
    h, w = image_hw(name)
    return (kw["num_frames"] // kw["tubelet_size"], h // kw["patch_size"][0], w // kw["patch_size"][1])


def init_weights_(model, seed=0, style="reference", skip=("get_main_input.flow_model.",)):
    """Overwrites every parameter of ``model`` (ours or the reference's) deterministically.

    ``style="reference"`` follows the reference's initialisation *distributions* (vmae.py:100-107, :371): xavier
    uniform Linear weights, zero biases, LayerNorm 1/0, Conv3d default, mask_token trunc-normal(0.02).
    ``style="perturbed"`` additionally gives every bias, q/v bias and LayerNorm affine a non-trivial value so that a
    dropped bias or a swapped gamma/beta cannot hide."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    with torch.no_grad():
        for name in sorted(sd.keys()):
            if name.startswith(tuple(skip)):  # a RAFT inside a flow preprocessor keeps its own (seeded) init
                continue
            t = sd[name]
            if name.endswith(("mask_token", "null_token_enc", "null_token_dec", "dummy_token")):
                v = torch.empty(t.shape).normal_(0, 0.02, generator=g).clamp_(-0.02, 0.02)
            elif name.endswith(("gamma_1", "gamma_2")):  # layer scale (VideoMAE/utils.py:140-144)
                v = 0.5 + torch.empty(t.shape).uniform_(-0.25, 0.25, generator=g)
            elif name.endswith("pos_embed") and t.dim() == 3:  # learnable positional embedding (vmae.py:68-70)
                v = torch.empty(t.shape).normal_(0, 0.2, generator=g)
            elif t.dim() == 1 and name.endswith(".weight"):  # LayerNorm weights (norm1 / norm2 / norm / norm*_cross)
                v = torch.ones(t.shape)
                if style == "perturbed":
                    v = v + torch.empty(t.shape).uniform_(-0.2, 0.2, generator=g)
            elif t.dim() == 1:  # biases (Linear, Conv3d, q_bias/v_bias, LayerNorm.bias)
                v = torch.zeros(t.shape)
                if style == "perturbed":
                    v = torch.empty(t.shape).uniform_(-0.2, 0.2, generator=g)
                elif name.endswith("patch_embed.proj.bias"):
                    fan_in = sd[name.replace("bias", "weight")][0].numel()
                    v = torch.empty(t.shape).uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in), generator=g)
            elif t.dim() == 2:  # Linear: xavier uniform
                bound = math.sqrt(6.0 / (t.shape[0] + t.shape[1]))
                v = torch.empty(t.shape).uniform_(-bound, bound, generator=g)
            elif t.dim() == 5:  # Conv3d default (kaiming uniform, a = sqrt(5)) -> U(-1/sqrt(fan_in), +)
                fan_in = t[0].numel()
                v = torch.empty(t.shape).uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in), generator=g)
            else:
                raise ValueError(f"unexpected parameter {name} {tuple(t.shape)}")
            t.copy_(v.to(t.dtype))
        if style == "trained_like":      # massive-activation channels + a mild drift of the token mean
            _add_trained_like_statistics_(sd, g)
        elif style == "mean_drift":      # stress: no outliers, the token mean runs away to several sigma
            _add_trained_like_statistics_(sd, g, mean_step=None, outliers=0)
    return model


def _add_trained_like_statistics_(sd, g, mean_step=0.9, outliers=4, outlier_scale=60.0):
    """Makes the residual streams look like a TRAINED ViT's instead of a freshly initialised one: every
    producer of the residual stream (patch embedding, each block's proj / fc2, the encoder->decoder map's input norm)
    gets a common per-channel offset so that the token mean drifts away from zero layer after layer (|mean| of a row
    reaches several of its standard deviations), and a handful of channels carry "massive activations" two orders of
    magnitude above the rest.  LayerNorm affine parameters are spread out as well.  On top of style="reference"."""
    def randn(shape):
        return torch.empty(shape).normal_(0, 1, generator=g)

    for stream in ("encoder", "decoder"):
        bias_names = sorted(n for n in sd if n.startswith(stream + ".blocks.") and
                            (n.endswith("attn.proj.bias") or n.endswith("mlp.fc2.bias")))
        if not bias_names:
            continue
        C = sd[bias_names[0]].numel()
        step = mean_step if mean_step is not None else 22.0 / len(bias_names)   # ends near 9 sigma at any depth
        big = torch.randperm(C, generator=g)[:outliers]
        sign = torch.where(torch.rand(outliers, generator=g) < 0.5, -1.0, 1.0)
        for n in bias_names:
            b = sd[n]
            b.add_(step + 0.1 * randn(b.shape))            # common drift: the row mean grows with depth
            if outliers:
                b[big] += sign * outlier_scale * (0.5 + torch.rand(outliers, generator=g))
        for n in sorted(n for n in sd if n.startswith(stream + ".") and n.endswith(("norm1.weight", "norm2.weight",
                                                                                    "norm.weight"))):
            sd[n].mul_(torch.empty(sd[n].shape).uniform_(0.5, 1.5, generator=g))
            sd[n.replace("weight", "bias")].add_(0.1 * randn(sd[n].shape))
    pe = "encoder.patch_embed.proj.bias"
    if pe in sd:
        sd[pe].add_(2.0 if outliers else 3.0)


def weights_checksum(model):
    """Order-independent float64 fingerprint used to check that the GPU box regenerated identical weights."""
    s = 0.0
    for name, t in sorted(model.state_dict().items()):
        s += float(t.double().sum()) + 0.5 * float(t.double().abs().sum())
    return s


def make_video(B, hw, seed=0, counterfactual_like=True, T=2, C=3):
    """Raw video in [0,1], layout [B, T, C, H, W] (the wrapper's input convention).  ``counterfactual_like``:
    frame 1 is a copy of frame 0 with a small shifted square pasted, like a motion counterfactual prompt."""
    g = torch.Generator().manual_seed(1000 + seed)
    H, W = hw
    x = torch.rand(B, T, C, H, W, generator=g)
    if counterfactual_like and T == 2:
        x[:, 1] = x[:, 0]
        s = max(4, H // 8)
        for b in range(B):
            y0 = int(torch.randint(0, H - 2 * s, (1,), generator=g))
            x0 = int(torch.randint(0, W - 2 * s, (1,), generator=g))
            x[b, 1, :, y0 + s // 2:y0 + s // 2 + s, x0 + s // 2:x0 + s // 2 + s] = x[b, 0, :, y0:y0 + s, x0:x0 + s]
    return x


def make_mask(B, msize, num_clumps=2, clump=2, seed=0):
    """Temporally-factored mask (README.md:21,68; masking.py:478-545): frame 0 fully visible, frame 1 fully masked
    except ``num_clumps`` visible ``clump x clump`` blocks per sample (distinct positions, so every row has
    exactly ``n_h*n_w + num_clumps*clump^2`` visible tokens).  bool [B, T*h*w], True = masked."""
    g = torch.Generator().manual_seed(2000 + seed)
    T, h, w = msize
    gh, gw = h // clump, w // clump
    mask = torch.zeros(B, T, h, w, dtype=torch.bool)
    mask[:, 1:] = True
    for b in range(B):
        cells = torch.randperm(gh * gw, generator=g)[:num_clumps]
        for c in cells.tolist():
            cy, cx = (c // gw) * clump, (c % gw) * clump
            mask[b, -1, cy:cy + clump, cx:cx + clump] = False
    return mask.reshape(B, -1)


# ---------------------------------------------------------------------------------------------------------
# conjoined / padded (IMU-conditioned) models -- SURVEY.md section 8a rows a13-a17, BASELINE config 5
# ---------------------------------------------------------------------------------------------------------
# name -> description of a conjoined model small enough for the CPU oracle; "imu400_base_4x4" is the real factory
CONJOINED = {
    # padded main (4x4 patches, 32 px, up to 8 null tokens) + padded IMU context (80 samples -> 5 tokens, 5 null tokens)
    "conj_padded_small": dict(
        kind="padded", img_size=32, patch_size=(4, 4), enc=(256, 4, 4), dec=(128, 2, 2), ctx_enc_dim=128,
        ctx_dec_dim=64, seq_len=80, main_pad=8, ctx_pad=5, enc_layers=[0, 3], dec_layers=True, main_chans=3,
        main_frames=[0, 1]),
    # the pair an ImuConditionedFlowGenerator drives, at 128 px so that RAFT's 4-level pyramid exists: an IMU-conditioned
    # padded predictor and a flow2imu model whose main-stream input is the real 'flowback_rgb01' preprocessor
    "conj_padded_128": dict(
        kind="padded", img_size=128, patch_size=(4, 4), enc=(256, 4, 4), dec=(128, 2, 2), ctx_enc_dim=128,
        ctx_dec_dim=64, seq_len=80, main_pad=8, ctx_pad=5, enc_layers=[0, 3], dec_layers=True, main_chans=3,
        main_frames=[0, 1]),
    "conj_flow2imu_128": dict(
        kind="full", img_size=128, patch_size=(8, 8), enc=(256, 3, 4), dec=(128, 2, 2), ctx_enc_dim=128,
        ctx_dec_dim=64, seq_len=80, main_pad=0, ctx_pad=0, enc_layers=[0, -1], dec_layers=True, main_chans=7,
        main_frames=[1], main_input='flowback_rgb01'),
    # non-padded, dummy-token IMU context, 7-channel single-frame main stream (the flow2imu topology, a17)
    "conj_flow2imu_small": dict(
        kind="full", img_size=32, patch_size=(8, 8), enc=(256, 3, 4), dec=(128, 2, 2), ctx_enc_dim=128,
        ctx_dec_dim=64, seq_len=80, main_pad=0, ctx_pad=0, enc_layers=[0, -1], dec_layers=True, main_chans=7,
        main_frames=[1]),
}


def build_conjoined(ns, name, preproc_ns=None, main_input_kwargs=None):
    """Builds the conjoined model `name` from the classes of module `ns` (ours: counterfactualworldmodels_b200.
    conjoined_vmae; or the reference's cwm.models.VideoMAE.conjoined_vmae -- same constructor signatures)."""
    import copy
    from functools import partial
    import torch.nn as nn
    if name == "imu400_base_4x4":
        return ns.imu400_base_4x4patch_2frames_1tube()
    c = CONJOINED[name]
    (Ce, Le, He), (Cd, Ld, Hd) = c["enc"], c["dec"]
    common = dict(img_size=c["img_size"], patch_size=c["patch_size"], encoder_embed_dim=Ce, encoder_depth=Le,
                  encoder_num_heads=He, encoder_num_classes=0, decoder_embed_dim=Cd, decoder_num_heads=Hd,
                  decoder_depth=Ld, mlp_ratio=4, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                  conjoin_encoder_layers=c["enc_layers"], conjoin_decoder_layers=c["dec_layers"], context_input='imu')
    ctx_kw = copy.deepcopy(ns.imu400_encoder_kwargs)
    ctx_kw.update(dict(encoder_embed_dim=c["ctx_enc_dim"], decoder_embed_dim=c["ctx_dec_dim"],
                       sequence_length=c["seq_len"]))
    main_kw = copy.deepcopy(ns.rgb_encoder_kwargs)
    if preproc_ns is None:
        import importlib
        preproc_ns = importlib.import_module(ns.__name__.rsplit(".", 1)[0] + ".preprocessor") \
            if ns.__name__.startswith("counterfactualworldmodels_b200") else importlib.import_module("cwm.models.preprocessor")
    if c.get("main_input"):  # a named preprocessor (e.g. 'flowback_rgb01': needs flow_model / flow_model_ckpt kwargs)
        main_input, main_input_kwargs = c["main_input"], dict(main_input_kwargs or {})
    elif c["main_chans"] == 3:
        main_input, main_input_kwargs = 'rgb01', {'unnormalize': False}
    else:  # a pass-through of frame 1 of a 7-channel input stands in for the RAFT-based 'flowback_rgb01' preprocessor
        main_input = partial(preproc_ns.Preprocessor, num_channels=c["main_chans"], frames_list=c["main_frames"])
        main_input_kwargs = {}
    if c["kind"] == "padded":
        main_kw.update(dict(min_padding_tokens=0, max_padding_tokens=c["main_pad"]))
        ctx_kw.update(dict(min_padding_tokens=0, max_padding_tokens=c["ctx_pad"], concat_dummy_token=False))
        return ns.ConjoinedPaddedVisionTransformer(
            main_model_func=ns.PaddedVisionTransformer, main_model_kwargs=main_kw, main_input=main_input,
            main_input_kwargs=main_input_kwargs, context_model_func=ns.PaddedVisionTransformer,
            context_model_kwargs=ctx_kw, **common)
    return ns.ConjoinedPretrainVisionTransformer(
        num_frames=2, main_model_kwargs=main_kw, main_input=main_input, main_input_kwargs=main_input_kwargs,
        context_model_kwargs=ctx_kw, **common)


def conjoined_oracle_cfg(name, model=None):
    """Structure description `oracle/conjoined_oracle.py` needs (it works on a bare state_dict)."""
    if name == "imu400_base_4x4":
        return dict(main=dict(enc_heads=12, dec_heads=6, max_pad=64, min_pad=0, pos="sinusoid", eps=1e-6),
                    ctx=dict(enc_heads=12, dec_heads=6, max_pad=25, min_pad=0, pos="torch", dummy=False, eps=1e-6),
                    enc_pairs=[(0, 0), (3, 3), (6, 6), (9, 9)], dec_pairs=[(0, 0), (1, 1), (2, 2), (3, 3)],
                    cross_heads=4, eps=1e-6, patch_size=(1, 4, 4))
    c = CONJOINED[name]
    Le, Ld = c["enc"][1], c["dec"][1]
    enc_pairs = [(l % Le, l % Le) for l in c["enc_layers"]]
    dec_pairs = [(l, l) for l in range(Ld)] if c["dec_layers"] is True else [(l % Ld, l % Ld) for l in c["dec_layers"]]
    return dict(main=dict(enc_heads=c["enc"][2], dec_heads=c["dec"][2], max_pad=c["main_pad"], min_pad=0,
                          pos="sinusoid", eps=1e-6),
                ctx=dict(enc_heads=c["enc"][2], dec_heads=c["dec"][2], max_pad=c["ctx_pad"], min_pad=0, pos="torch",
                         dummy=(c["kind"] == "full"), eps=1e-6),
                enc_pairs=enc_pairs, dec_pairs=dec_pairs, cross_heads=4, eps=1e-6,
                patch_size=(1,) + tuple(c["patch_size"]))


def make_imu(B, seq_len, seed=0, channels=6):
    """Synthetic IMU sequence [B, 6, L] (unit-variance noise plus a slow drift)."""
    g = torch.Generator().manual_seed(3000 + seed)
    t = torch.linspace(0, 1, seq_len)
    return torch.randn(B, channels, seq_len, generator=g) * 0.5 + torch.sin(6.0 * t)[None, None] * \
        torch.randn(B, channels, 1, generator=g)
