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


def init_weights_(model, seed=0, style="reference"):
    """Overwrites every parameter of ``model`` (ours or the reference's) deterministically.

    ``style="reference"`` follows the reference's initialisation *distributions* (vmae.py:100-107, :371): xavier
    uniform Linear weights, zero biases, LayerNorm 1/0, Conv3d default, mask_token trunc-normal(0.02).
    ``style="perturbed"`` additionally gives every bias, q/v bias and LayerNorm affine a non-trivial value so that a
    dropped bias or a swapped gamma/beta cannot hide."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    with torch.no_grad():
        for name in sorted(sd.keys()):
            t = sd[name]
            if name == "mask_token":
                v = torch.empty(t.shape).normal_(0, 0.02, generator=g).clamp_(-0.02, 0.02)
            elif name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("norm.weight"):
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
    return model


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
