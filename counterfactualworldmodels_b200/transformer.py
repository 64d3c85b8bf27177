"""Parameter holders for the conjoining blocks of the IMU-conditioned models (``cwm/models/transformer.py``):
``CrossAttentionTransformerBlock`` (:442-583) around ``BidirectionalCrossAttention`` (:253-378) and the two-layer
``Mlp`` (:77-110), plus the torch-fp32 sinusoid ``pos_embedding`` (:37-52) the IMU stream uses.

Same constructor arguments and ``state_dict`` keys as the reference for the configuration every shipped factory
instantiates (``with_self_attention=False``, ``shared_similarity=False``, no qkv bias, GELU, LayerNorm eps 1e-6:
``conjoined_vmae.py:215-220``); the arithmetic runs in ``cwm_cross_block_forward`` (include/cwm_b200.h).
"""
import torch
import torch.nn as nn


def pos_embedding(positions, hidden_dim, device='cpu'):
    """cwm/models/transformer.py:37-52: float32 torch ops (NOT the float64 numpy table of VideoMAE/utils.py:251-268;
    the two differ by up to 4.5e-4).  Evaluated on the host so every GPU sees the same table."""
    if isinstance(positions, int):
        positions = torch.arange(positions).float()
    elif isinstance(positions, torch.Tensor):
        positions = positions.clone().detach().float().cpu()
    else:
        assert hasattr(positions, '__len__')
        positions = torch.tensor(positions, dtype=torch.float)
    freqs = torch.arange(hidden_dim).float()
    freqs = torch.pow(10000, 2 * (torch.div(freqs, 2, rounding_mode='trunc')) / hidden_dim)
    out = positions[:, None] / freqs[None, :]
    out[:, 0::2] = torch.sin(out[:, 0::2])
    out[:, 1::2] = torch.cos(out[:, 1::2])
    return out.unsqueeze(0).to(device)


def _no_eager(name):
    raise NotImplementedError(f"{name}.forward: runs inside libcwm_b200 (cwm_cross_block_forward); no eager forward")


class Mlp(nn.Module):
    """transformer.py:77-110 with one hidden layer: ``layers = Sequential(Linear, GELU, Linear)``."""

    def __init__(self, in_dim, out_dim=None, hidden_dim=None, activation='gelu', dropout_prob=0.0):
        super().__init__()
        if activation != 'gelu':
            raise NotImplementedError("only GELU (erf) is fused into the fc1 epilogue")
        if dropout_prob:
            raise NotImplementedError("dropout is training-only")
        self.in_dim = in_dim
        self.out_dim = out_dim or in_dim
        hidden = in_dim if hidden_dim is None else hidden_dim
        self.hidden_dim = list(hidden) if hasattr(hidden, '__len__') else [hidden]
        if len(self.hidden_dim) != 1:
            raise NotImplementedError("conjoining blocks use a single hidden layer (conjoined_vmae.py:215-220)")
        self.layers = nn.Sequential(nn.Linear(self.in_dim, self.hidden_dim[0]), nn.GELU(),
                                    nn.Linear(self.hidden_dim[0], self.out_dim))

    def forward(self, x):
        _no_eager("Mlp")


class BidirectionalCrossAttention(nn.Module):
    """transformer.py:253-378 (shared_similarity=False, qkv_bias=False)."""

    def __init__(self, in_dim, num_heads, shared_similarity=False, in_dim_src=None, head_dim=None, out_dim=None,
                 out_dim_src=None, qkv_bias=False, qk_scale=None, attention_dropout_prob=0,
                 projection_dropout_prob=0, flash_attention=False):
        super().__init__()
        if shared_similarity:
            raise NotImplementedError("shared_similarity=True is not used by any CWM factory")
        if qkv_bias:
            raise NotImplementedError("qkv_bias=True of the cross attention is not used by any CWM factory "
                                      "(its biases are plain tensors, not parameters: transformer.py:287-289)")
        if attention_dropout_prob or projection_dropout_prob:
            raise NotImplementedError("dropout is training-only")
        self.in_dim = in_dim
        self.in_dim_src = in_dim_src or in_dim
        self.num_heads = self.H = num_heads
        self.head_dim = head_dim or (in_dim // num_heads)
        self.out_dim = out_dim or in_dim
        self.out_dim_src = out_dim_src or self.in_dim_src
        self.scale = qk_scale or (self.head_dim ** -0.5)
        self.shared_similarity = False
        D = self.D
        self.qk = nn.Linear(self.in_dim, D * 2, bias=False)
        self.qk_src = nn.Linear(self.in_dim_src, D * 2, bias=False)
        self.v = nn.Linear(self.in_dim, D, bias=False)
        self.v_src = nn.Linear(self.in_dim_src, D, bias=False)
        self.qkv_bias = self.qkv_bias_src = None
        self.flash_attention = False
        self.projection = nn.Linear(D, self.out_dim)
        self.projection_src = nn.Linear(D, self.out_dim_src)

    @property
    def D(self):
        return self.num_heads * self.head_dim

    def forward(self, x, src=None):
        _no_eager("BidirectionalCrossAttention")


class CrossAttentionTransformerBlock(nn.Module):
    """transformer.py:442-583 with ``with_self_attention=False`` (the only mode the factories use): norm1/norm1_src,
    the self-attention slots and the shortcuts are Identity (no parameters), gamma_1 = 0."""
    default_attention_func = BidirectionalCrossAttention

    def __init__(self, in_dim, num_heads, in_dim_src=None, head_dim=None, out_dim=None, out_dim_src=None,
                 mlp_ratio=4.0, drop_path_prob=0.0, init_values=None, activation='gelu',
                 normalization={'func': 'layer', 'eps': 1e-6}, attention_func=default_attention_func,
                 with_self_attention=True, shared_similarity=False, **kwargs):
        super().__init__()
        if with_self_attention:
            raise NotImplementedError("with_self_attention=True is not used by any CWM factory "
                                      "(conjoined_vmae.py:215-220)")
        if attention_func is not BidirectionalCrossAttention:
            raise NotImplementedError("only BidirectionalCrossAttention is implemented")
        if (init_values or 0) > 0 or drop_path_prob:
            raise NotImplementedError("layer-scale / drop-path are not used by any CWM factory")
        if not (mlp_ratio > 0.0):
            raise NotImplementedError("mlp_ratio must be > 0")
        norm = dict(normalization)
        if norm.pop('func', 'layer') != 'layer':
            raise NotImplementedError("only LayerNorm")
        kwargs.pop('flash_attention', None)  # forced to False by the reference (conjoined_vmae.py:305-306)
        self.in_dim = in_dim
        self.in_dim_src = in_dim_src or in_dim
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.out_dim = out_dim or self.in_dim
        self.out_dim_src = out_dim_src or self.in_dim_src
        if self.out_dim != self.in_dim or self.out_dim_src != self.in_dim_src:
            raise NotImplementedError("in_dim != out_dim (Linear shortcut) is not used by any CWM factory")
        self.norm1, self.norm1_src = nn.Identity(), nn.Identity()
        self.norm1_cross = nn.LayerNorm(self.in_dim, **norm)
        self.norm1_src_cross = nn.LayerNorm(self.in_dim_src, **norm)
        self.self_attention = nn.ModuleDict([('trg', nn.Identity()), ('src', nn.Identity())])
        self.shortcut = nn.ModuleDict([('trg', nn.Identity()), ('src', nn.Identity())])
        self.drop_path = nn.Identity()
        self.norm2 = nn.LayerNorm(self.out_dim, **norm)
        self.norm2_src = nn.LayerNorm(self.out_dim_src, **norm)
        self.cross_attention = BidirectionalCrossAttention(
            in_dim=self.in_dim, num_heads=num_heads, head_dim=head_dim, out_dim=self.out_dim,
            in_dim_src=self.in_dim_src, out_dim_src=self.out_dim_src, shared_similarity=shared_similarity, **kwargs)
        self.activation = activation
        self.mlp_ratio = mlp_ratio
        self.mlp = nn.ModuleDict([
            ('trg', Mlp(self.out_dim, hidden_dim=[int(self.out_dim * mlp_ratio)], activation=activation)),
            ('src', Mlp(self.out_dim_src, hidden_dim=[int(self.out_dim_src * mlp_ratio)], activation=activation))])
        self.gamma_1 = self.gamma_1_src = 0.0
        self.gamma_1_cross = self.gamma_1_src_cross = 1.0
        self.gamma_2 = self.gamma_2_src = 1.0

    def forward(self, x, src=None):
        _no_eager("CrossAttentionTransformerBlock")
