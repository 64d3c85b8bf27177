"""RAFT's non-convolution stages on device (first slice).

RAFT (``cwm/models/raft/``) is the flow network every counterfactual runs right after the VMAE path
(``cwm/models/segmentation.py:431``).  Its convolutions stay whatever torch module the caller supplies as
``flow_model``; this file mirrors the three pieces that are *not* convolutions, with the reference's names and
argument meaning, on the hand-written kernels of ``csrc/raftcorr.cu``:

  * ``CorrBlock(fmap1, fmap2, num_levels=4, radius=4)`` / ``corr_fn(coords)``  cwm/models/raft/corr.py:12-60
  * ``upsample_flow(flow, mask)``                                                cwm/models/raft/raft_model.py:175-186
  * ``coords_grid(batch, ht, wd, device)``                                       cwm/models/raft/utils.py:83-86

``use_in_reference(raft_model_module)`` shows how a maintainer switches the reference's RAFT over (INTEGRATION.md
section 7).  There is no CPU fallback: CPU tensors raise.
"""
import ctypes
import os

import torch

from . import _lib, ops


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _req(t, name, dims):
    if t.device.type != "cuda":
        raise RuntimeError(f"{name}: libcwm_b200 needs CUDA (B200) tensors; there is no CPU fallback")
    assert t.dim() == dims, (name, t.shape)
    if t.dtype != torch.float32:
        t = t.float()  # the reference runs these stages in fp32 (raft_model.py:224-225, corr.py:51)
    return t.contiguous()


def _ptr_table(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class CorrBlock:
    """All-pairs correlation pyramid + window lookup, ``cwm/models/raft/corr.py:12-60``.

    ``corr_pyramid[l]`` is fp32 ``[B*H*W, 1, H>>l, W>>l]`` like the reference's list; calling the block with
    ``coords [B, 2, H, W]`` returns fp32 ``[B, num_levels*(2r+1)^2, H, W]``.
    """

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        fmap1, fmap2 = _req(fmap1, "CorrBlock fmap1", 4), _req(fmap2, "CorrBlock fmap2", 4)
        assert fmap1.shape == fmap2.shape, (fmap1.shape, fmap2.shape)
        batch, dim, ht, wd = fmap1.shape
        self._shape = (batch, ht, wd)
        self.corr_pyramid = []
        h, w = ht, wd
        for _ in range(num_levels):
            self.corr_pyramid.append(torch.empty(batch * ht * wd, 1, h, w, dtype=torch.float32, device=fmap1.device))
            h, w = h // 2, w // 2
        lib = _lib.load()
        # level 0 on the tensor cores (3xTF32 split, fp32 accuracy) when the channel count allows the 32-wide k-slabs;
        # CWM_RAFT_CORR=simt keeps round 1's fp32 SIMT kernel (the A/B reference)
        tensor_cores = dim % 32 == 0 and os.environ.get("CWM_RAFT_CORR", "tc") != "simt"
        with torch.cuda.device(fmap1.device):
            if tensor_cores:
                ws = torch.empty(lib.cwm_raft_corr_tc_workspace_bytes(batch, dim, ht, wd), dtype=torch.uint8,
                                 device=fmap1.device)
                _lib.check(lib.cwm_raft_corr_pyramid_tc(fmap1.data_ptr(), fmap2.data_ptr(), batch, dim, ht, wd, num_levels,
                                                        _ptr_table(self.corr_pyramid), ws.data_ptr(), ws.numel(),
                                                        _stream(fmap1)))
            else:
                _lib.check(lib.cwm_raft_corr_pyramid(fmap1.data_ptr(), fmap2.data_ptr(), batch, dim, ht, wd, num_levels,
                                                     _ptr_table(self.corr_pyramid), _stream(fmap1)))

    @classmethod
    def from_rows(cls, rows1, rows2, batch, ht, wd, num_levels=4, radius=4, f16_pyramid=False):
        """The same block from f16 pixel-major feature rows (``FusedFeatureEncoder``'s output): ``rows2 [batch*ht*wd, D]``,
        ``rows1`` likewise or ``[ht*wd, D]`` for one image shared by the whole batch.  f16 products are exact in the fp32
        accumulator, so level 0 is one ``tcgen05.mma.kind::f16`` GEMM per sample straight from these rows: no fp32 cast, no
        transpose, no hi / lo split (``cwm_raft_corr_pyramid_rows_f16``).  ``f16_pyramid=True`` stores every level in f16
        (half the bytes: a 64-sample level 0 then stays in L2 across the 24 lookups); such a block serves the f16 row lookup of
        the fused recurrent block only (``__call__`` raises)."""
        self = cls.__new__(cls)
        self.num_levels, self.radius = num_levels, radius
        D = rows2.shape[1]
        hw = ht * wd
        n1 = rows1.shape[0] // hw
        assert rows1.dtype == rows2.dtype == _lib.act_dtype() and rows1.is_contiguous() and rows2.is_contiguous()
        assert rows2.shape == (batch * hw, D) and rows1.shape == (n1 * hw, D) and n1 in (1, batch) and D % 64 == 0
        self._shape = (batch, ht, wd)
        self.corr_pyramid = []
        h, w = ht, wd
        pyr_dtype = torch.float16 if f16_pyramid else torch.float32
        for _ in range(num_levels):
            self.corr_pyramid.append(torch.empty(batch * hw, 1, h, w, dtype=pyr_dtype, device=rows2.device))
            h, w = h // 2, w // 2
        lib = _lib.load()
        fn = lib.cwm_raft_corr_pyramid_rows_f16_pyr16 if f16_pyramid else lib.cwm_raft_corr_pyramid_rows_f16
        with torch.cuda.device(rows2.device):
            _lib.check(fn(rows1.data_ptr(), n1, rows2.data_ptr(), batch, D, ht, wd, num_levels, _ptr_table(self.corr_pyramid),
                          _stream(rows2)))
        return self

    def __call__(self, coords):
        coords = _req(coords, "CorrBlock coords", 4)
        if self.corr_pyramid[0].dtype != torch.float32:
            raise RuntimeError("an f16 correlation pyramid serves the fused recurrent block's f16 lookup only")
        batch, ht, wd = self._shape
        assert tuple(coords.shape) == (batch, 2, ht, wd), (coords.shape, self._shape)
        n1 = 2 * self.radius + 1
        out = torch.empty(batch, self.num_levels * n1 * n1, ht, wd, dtype=torch.float32, device=coords.device)
        with torch.cuda.device(coords.device):
            _lib.check(_lib.load().cwm_raft_corr_lookup(_ptr_table(self.corr_pyramid), self.num_levels, self.radius,
                                                        coords.data_ptr(), batch, ht, wd, out.data_ptr(),
                                                        _stream(coords)))
        return out

    @staticmethod
    def corr(fmap1, fmap2):
        """``fmap1^T fmap2 / sqrt(dim)`` as ``[B, H, W, 1, H, W]`` (corr.py:53-60)."""
        block = CorrBlock(fmap1, fmap2, num_levels=1, radius=0)
        batch, ht, wd = block._shape
        return block.corr_pyramid[0].view(batch, ht, wd, 1, ht, wd)


def upsample_flow(flow, mask):
    """Convex 8x upsampling ``[N, C, H, W] -> [N, C, 8H, 8W]`` (``RAFT.upsample_flow``, raft_model.py:175-186);
    ``mask`` is the update block's ``[N, 64*9, H, W]`` logits (already scaled by 0.25, update.py:137)."""
    flow, mask = _req(flow, "upsample_flow flow", 4), _req(mask, "upsample_flow mask", 4)
    N, C, H, W = flow.shape
    assert tuple(mask.shape) == (N, 576, H, W), (mask.shape, flow.shape)
    out = torch.empty(N, C, 8 * H, 8 * W, dtype=torch.float32, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.check(_lib.load().cwm_raft_upsample_flow(flow.data_ptr(), mask.data_ptr(), N, C, H, W, out.data_ptr(),
                                                      _stream(flow)))
    return out


def coords_grid(batch, ht, wd, device, dtype=torch.float32):
    """``[batch, 2, ht, wd]`` pixel grid, channel 0 = x, channel 1 = y (cwm/models/raft/utils.py:83-86)."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).to(dtype)[None].repeat(batch, 1, 1, 1)


def use_in_reference(raft_model_module):
    """Switches the reference's RAFT (the imported ``cwm.models.raft.raft_model`` module) to these kernels: its
    ``_forward_two_images`` looks ``CorrBlock`` up in the module namespace (raft_model.py:227-229) and calls
    ``self.upsample_flow`` (raft_model.py:266).  Returns the two replaced objects so the caller can restore them."""
    old = (raft_model_module.CorrBlock, raft_model_module.RAFT.upsample_flow)
    raft_model_module.CorrBlock = CorrBlock
    raft_model_module.RAFT.upsample_flow = lambda self, flow, mask: upsample_flow(flow, mask)
    return old


# ======================================================================================================
# The network around those stages.  RAFT's convolutions are cuDNN calls from torch (library code, like cuBLAS for a
# plain GEMM); the module tree below only exists so that (a) the published ``raft-large.pth`` / ``raft-small.pth``
# checkpoints load unchanged (same parameter names and shapes as cwm/models/raft/{extractor,update,raft_model}.py),
# (b) seeded random init reproduces the reference's (same construction order; pinned by tests/golden/raft_e2e_*.npz),
# and (c) ``FlowGenerator(flow_model=raft.RAFT(...))`` works on a B200 with the correlation block, the lookups of all
# GRU iterations and the convex upsampling on the kernels above.
# ======================================================================================================
import argparse  # noqa: E402
import os  # noqa: E402

import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def _make_norm(kind, channels, groups):
    if kind == 'group':
        return nn.GroupNorm(num_groups=groups, num_channels=channels)
    if kind == 'batch':
        return nn.BatchNorm2d(channels)
    if kind == 'instance':
        return nn.InstanceNorm2d(channels)
    assert kind == 'none', kind
    return nn.Sequential()


def _init_encoder(module):
    """extractor.py:151-158 / :229-236: He-normal conv weights (biases keep torch's default draw), unit norms."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
            if m.weight is not None:
                nn.init.constant_(m.weight, 1)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class _EncoderBlock(nn.Module):
    """``ResidualBlock`` (widths (p, p), kernels (3, 3); extractor.py:6-56) and ``BottleneckBlock`` (widths
    (p/4, p/4, p), kernels (1, 3, 1); :60-115) are the same pattern: conv-norm-relu chain, strided conv in the
    position ``strided``, optional 1x1 strided shortcut whose norm is ALSO registered as ``norm<k+1>``."""

    def __init__(self, in_planes, widths, kernels, strided, norm_fn, stride):
        super().__init__()
        planes = widths[-1]
        chain = [in_planes] + list(widths)
        for i, k in enumerate(kernels):
            setattr(self, f"conv{i + 1}", nn.Conv2d(chain[i], chain[i + 1], kernel_size=k, padding=k // 2,
                                                    stride=stride if i == strided else 1))
        self.relu = nn.ReLU(inplace=True)
        self.depth = len(kernels)
        for i, w in enumerate(widths):
            setattr(self, f"norm{i + 1}", _make_norm(norm_fn, w, planes // 8))
        self.downsample = None
        if stride != 1:
            shortcut_norm = _make_norm(norm_fn, planes, planes // 8)
            setattr(self, f"norm{self.depth + 1}", shortcut_norm)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride), shortcut_norm)

    def forward(self, x):
        y = x
        for i in range(1, self.depth + 1):
            y = self.relu(getattr(self, f"norm{i}")(getattr(self, f"conv{i}")(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class ResidualBlock(_EncoderBlock):
    def __init__(self, in_planes, planes, norm_fn='group', stride=1):
        super().__init__(in_planes, (planes, planes), (3, 3), 0, norm_fn, stride)


class BottleneckBlock(_EncoderBlock):
    def __init__(self, in_planes, planes, norm_fn='group', stride=1):
        super().__init__(in_planes, (planes // 4, planes // 4, planes), (1, 3, 1), 1, norm_fn, stride)


class _Encoder(nn.Module):
    """1/8-resolution feature pyramid: 7x7/2 stem, three 2-block stages (strides 1, 2, 2), 1x1 projection.
    ``BasicEncoder`` (extractor.py:117-190) and ``SmallEncoder`` (:193-267) differ in widths and block type."""

    block = None
    widths = None

    def __init__(self, output_dim=128, norm_fn='batch', dropout=0.0):
        super().__init__()
        self.norm_fn = norm_fn
        stem = self.widths[0]
        self.norm1 = _make_norm(norm_fn, stem, 8)
        self.conv1 = nn.Conv2d(3, stem, kernel_size=7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        self.in_planes = stem
        for i, (w, s) in enumerate(zip(self.widths, (1, 2, 2))):
            setattr(self, f"layer{i + 1}", self._make_layer(w, stride=s))
        self.dropout = None
        if type(self).dropout_first:
            self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
            self.conv2 = nn.Conv2d(self.widths[-1], output_dim, kernel_size=1)
        else:
            self.conv2 = nn.Conv2d(self.widths[-1], output_dim, kernel_size=1)
            self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        _init_encoder(self)

    def _make_layer(self, dim, stride=1):
        blocks = (self.block(self.in_planes, dim, self.norm_fn, stride=stride), self.block(dim, dim, self.norm_fn, stride=1))
        self.in_planes = dim
        return nn.Sequential(*blocks)

    def forward(self, x):
        pair = isinstance(x, (tuple, list))  # both frames in one batch (raft_model.py:221-222)
        if pair:
            n = x[0].shape[0]
            x = torch.cat(x, dim=0)
        x = self.relu1(self.norm1(self.conv1(x)))
        x = self.conv2(self.layer3(self.layer2(self.layer1(x))))
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return torch.split(x, [n, n], dim=0) if pair else x


class BasicEncoder(_Encoder):
    block, widths, dropout_first = ResidualBlock, (64, 96, 128), False


class SmallEncoder(_Encoder):
    block, widths, dropout_first = BottleneckBlock, (32, 64, 96), True


class FlowHead(nn.Module):
    """update.py:6-14."""

    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(self.relu(self.conv1(x)))


def _gru_step(h, x, convz, convr, convq):
    hx = torch.cat([h, x], dim=1)
    z = torch.sigmoid(convz(hx))
    r = torch.sigmoid(convr(hx))
    q = torch.tanh(convq(torch.cat([r * h, x], dim=1)))
    return (1 - z) * h + z * q


class ConvGRU(nn.Module):
    """update.py:16-31."""

    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        for gate in "zrq":
            setattr(self, "conv" + gate, nn.Conv2d(hidden_dim + input_dim, hidden_dim, 3, padding=1))

    def forward(self, h, x):
        return _gru_step(h, x, self.convz, self.convr, self.convq)


class SepConvGRU(nn.Module):
    """update.py:33-60: a horizontal (1x5) then a vertical (5x1) GRU step."""

    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        for idx, (k, p) in enumerate((((1, 5), (0, 2)), ((5, 1), (2, 0))), start=1):
            for gate in "zrq":
                setattr(self, f"conv{gate}{idx}", nn.Conv2d(hidden_dim + input_dim, hidden_dim, k, padding=p))

    def forward(self, h, x):
        h = _gru_step(h, x, self.convz1, self.convr1, self.convq1)
        return _gru_step(h, x, self.convz2, self.convr2, self.convq2)


class SmallMotionEncoder(nn.Module):
    """update.py:62-77."""

    def __init__(self, args):
        super().__init__()
        cor_planes = args.corr_levels * (2 * args.corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(cor_planes, 96, 1, padding=0)
        self.convf1 = nn.Conv2d(2, 64, 7, padding=3)
        self.convf2 = nn.Conv2d(64, 32, 3, padding=1)
        self.conv = nn.Conv2d(128, 80, 3, padding=1)

    def forward(self, flow, corr):
        cor = F.relu(self.convc1(corr))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        return torch.cat([F.relu(self.conv(torch.cat([cor, flo], dim=1))), flow], dim=1)


class BasicMotionEncoder(nn.Module):
    """update.py:79-98."""

    def __init__(self, args):
        super().__init__()
        cor_planes = args.corr_levels * (2 * args.corr_radius + 1) ** 2
        self.convc1 = nn.Conv2d(cor_planes, 256, 1, padding=0)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def forward(self, flow, corr):
        cor = F.relu(self.convc2(F.relu(self.convc1(corr))))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        return torch.cat([F.relu(self.conv(torch.cat([cor, flo], dim=1))), flow], dim=1)


class SmallUpdateBlock(nn.Module):
    """update.py:100-113."""

    def __init__(self, args, hidden_dim=96):
        super().__init__()
        self.encoder = SmallMotionEncoder(args)
        self.gru = ConvGRU(hidden_dim=hidden_dim, input_dim=82 + 64)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=128)

    def forward(self, net, inp, corr, flow):
        net = self.gru(net, torch.cat([inp, self.encoder(flow, corr)], dim=1))
        return net, None, self.flow_head(net)


class BasicUpdateBlock(nn.Module):
    """update.py:115-139; the upsampling logits are scaled by 0.25 ("to balance gradients", :137)."""

    def __init__(self, args, hidden_dim=128, input_dim=128):
        super().__init__()
        self.args = args
        self.encoder = BasicMotionEncoder(args)
        self.gru = SepConvGRU(hidden_dim=hidden_dim, input_dim=128 + hidden_dim)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(256, 64 * 9, 1, padding=0))

    def forward(self, net, inp, corr, flow, upsample=True):
        net = self.gru(net, torch.cat([inp, self.encoder(flow, corr)], dim=1))
        return net, (.25 * self.mask(net) if upsample else None), self.flow_head(net)


class FusedFeatureEncoder:
    """``BasicEncoder`` (RAFT's feature network with ``norm_fn='instance'`` and its context network with ``'batch'``,
    extractor.py:118-190) for the mixed-precision inference path, on f16 pixel-major rows ``[S*H*W, C]``:

    * every convolution is an implicit GEMM on the repo's tcgen05 kernel (``cwm_conv2d_strided_f16``: 2-D tiles for the
      112 / 56 pixel maps, stride 2 through the tensor map's traversal stride); the 7x7 / 2 stem on the 3-channel frame is an
      im2col (``cwm_im2col_nchw_f16``, K = 147 -> 152) followed by the same kernel as a 1x1 convolution;
    * instance norm: a per-channel bias in front of an affine-free instance norm cancels exactly and is dropped; every
      ``norm -> relu [-> + shortcut -> relu]`` is one statistics + one transform launch of ``cwm_instnorm_f16``;
    * batch norm (inference: running statistics) is folded into the convolution weights and biases, relu runs in the
      convolution epilogue and the residual join is one ``cwm_add_act_f16``.

    ``conv_impl='cudnn'`` (``CWM_RAFT_ENCODER_CONV=cudnn``) keeps round 1's route for A/B runs: cuDNN convolutions on
    channels-last views of the same rows (instance norm only).  Same arithmetic as the module it wraps up to the f16
    rounding of the activations (which autocast applies as well)."""

    K_STEM = 152   # 7 * 7 * 3 = 147 im2col columns padded to a multiple of 8 (16-byte rows)

    def __init__(self, enc, device, conv_impl=None):
        assert isinstance(enc, BasicEncoder) and enc.norm_fn in ('instance', 'batch')
        _lib.require_f16("raft.FusedFeatureEncoder")
        self.norm = enc.norm_fn
        self.conv_impl = conv_impl or (os.environ.get("CWM_RAFT_ENCODER_CONV", "tcgen05") if self.norm == 'instance' else "tcgen05")
        assert self.conv_impl in ("tcgen05", "cudnn") and (self.conv_impl == "tcgen05" or self.norm == 'instance')
        self.eps = 1e-5

        def fold(conv, bn):
            """-> (weight fp32 [Cout, Cin, kh, kw], bias fp32 [Cout] or None) with the batch norm folded in"""
            w = conv.weight.detach().float()
            if self.norm == 'instance':
                return w, None
            g = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            return w * g.view(-1, 1, 1, 1), (conv.bias.detach().float() - bn.running_mean.detach().float()) * g + bn.bias.detach().float()

        def pack(conv, bn):
            w, b = fold(conv, bn)
            w16 = w.to(device=device, dtype=torch.float16)
            b = None if b is None else b.to(device).contiguous()
            if self.conv_impl == "cudnn":
                return w16.contiguous(memory_format=torch.channels_last), b, conv.kernel_size[0], conv.stride[0]
            return ops.pack_conv_weight(w16), b, conv.kernel_size[0], conv.stride[0]

        k = enc.conv1.kernel_size[0]
        assert enc.conv1.in_channels * k * k <= self.K_STEM
        w, b = fold(enc.conv1, enc.norm1)
        wk = torch.zeros(w.shape[0], self.K_STEM, 1, 1)
        wk[:, :w.shape[1] * k * k, 0, 0] = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).cpu()
        self.stem_geom = (k, enc.conv1.stride[0], enc.conv1.padding[0])
        self.stem_w = ops.pack_conv_weight(wk.to(device=device, dtype=torch.float16))
        self.stem_b = None if b is None else b.to(device).contiguous()
        self.stem_cudnn = w.to(device=device, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        self.blocks = []
        for layer in (enc.layer1, enc.layer2, enc.layer3):
            for blk in layer:
                down = None if blk.downsample is None else pack(blk.downsample[0], blk.downsample[1])
                self.blocks.append((pack(blk.conv1, blk.norm1), pack(blk.conv2, blk.norm2), down))
        self.out_conv = (ops.pack_conv_weight(enc.conv2.weight.detach().to(device=device, dtype=torch.float16)),
                         enc.conv2.bias.detach().float().to(device).contiguous(), enc.conv2.out_channels)
        self.out_cudnn = (enc.conv2.weight.detach().to(device=device, dtype=torch.float16),
                          enc.conv2.bias.detach().to(device=device, dtype=torch.float16))
        self._ws = None

    def _inorm(self, rows, S, relu_inner, add=None, relu_outer=False):
        """rows f16 [S*HW, C] -> instance-normalised rows (statistics per sample and channel)."""
        lib = _lib.load()
        C = rows.shape[1]
        need = lib.cwm_instnorm_workspace_bytes(S, C)
        if self._ws is None or self._ws.numel() < need or self._ws.device != rows.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=rows.device)
        out = torch.empty_like(rows)
        assert rows.is_contiguous() and rows.dtype == torch.float16 and (add is None or (add.shape == rows.shape and add.is_contiguous()))
        _lib.check(lib.cwm_instnorm_f16(rows.data_ptr(), S, rows.shape[0] // S, C, self.eps, int(relu_inner),
                                        None if add is None else add.data_ptr(), int(relu_outer), out.data_ptr(),
                                        self._ws.data_ptr(), self._ws.numel(), _stream(rows)))
        return out

    def _conv(self, rows, S, H, W, packed, relu):
        """-> (output rows, Ho, Wo); bias / relu in the epilogue when the norm is folded (batch norm)."""
        w, b, k, stride = packed
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        Cin = rows.shape[1]
        if self.conv_impl == "cudnn":
            y = F.conv2d(rows.view(S, H, W, Cin).permute(0, 3, 1, 2), w, None, stride, k // 2)
            if not y.is_contiguous(memory_format=torch.channels_last):
                y = y.contiguous(memory_format=torch.channels_last)
            return y.permute(0, 2, 3, 1).reshape(S * Ho * Wo, -1), Ho, Wo
        Cout = w.shape[0]
        out = torch.empty(S * Ho * Wo, Cout, dtype=torch.float16, device=rows.device)
        _lib.check(_lib.load().cwm_conv2d_strided_f16(rows.data_ptr(), rows.stride(0), S, H, W, Cin, w.data_ptr(), Cout, k, k,
                                                      k // 2, k // 2, stride, None if b is None else b.data_ptr(),
                                                      int(bool(relu and b is not None)), out.data_ptr(), Cout, _stream(rows)))
        return out, Ho, Wo

    def __call__(self, x, scale=1.0, shift=0.0):
        """x: fp32 / f16 ``[S, 3, H, W]`` or a list of such tensors (encoded as one batch, in order); the network sees
        ``scale * x + shift`` -- the input normalisation 2 (x / 255) - 1 folded into the stem's im2col, so frame slices of a
        movie are neither copied nor normalised by separate kernels.  -> f16 [S, output_dim, H/8, W/8] (channels-last)."""
        lib = _lib.load()
        inst = self.norm == 'instance'
        parts = list(x) if isinstance(x, (list, tuple)) else [x]
        with torch.cuda.device(parts[0].device):
            _, _, H, W = parts[0].shape
            S = sum(t.shape[0] for t in parts)
            k, stride, pad = self.stem_geom
            H1, W1 = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
            if self.conv_impl == "cudnn":
                xin = torch.cat([t.float() for t in parts], 0) * scale + shift
                y = F.conv2d(xin.to(torch.float16).contiguous(memory_format=torch.channels_last), self.stem_cudnn, None, stride, pad)
                y = y.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(S * H1 * W1, -1)
            else:
                cols = torch.empty(S * H1 * W1, self.K_STEM, dtype=torch.float16, device=parts[0].device)
                row = 0
                for t in parts:
                    t = t.float()
                    if t.stride()[1:] != (H * W, W, 1):
                        t = t.contiguous()
                    n = t.shape[0] * H1 * W1
                    ops.im2col_nchw_f16(t, k, stride, pad, self.K_STEM, scale, shift, out=cols[row:row + n])
                    row += n
                y, _, _ = self._conv(cols, S, H1, W1, (self.stem_w, self.stem_b, 1, 1), relu=True)
            if inst:
                y = self._inorm(y, S, relu_inner=True)
            H, W = H1, W1
            for c1, c2, down in self.blocks:
                z, Ho, Wo = self._conv(y, S, H, W, c1, relu=True)
                if inst:
                    z = self._inorm(z, S, relu_inner=True)
                z, _, _ = self._conv(z, S, Ho, Wo, c2, relu=True)
                if down is not None:
                    y, _, _ = self._conv(y, S, H, W, down, relu=False)
                    if inst:
                        y = self._inorm(y, S, relu_inner=False)
                if inst:
                    y = self._inorm(z, S, relu_inner=True, add=y, relu_outer=True)     # relu(x + relu(norm2(conv2(.))))
                else:
                    out = torch.empty_like(z)
                    _lib.check(lib.cwm_add_act_f16(y.data_ptr(), z.data_ptr(), z.numel(), 1, out.data_ptr(), _stream(z)))
                    y = out
                H, W = Ho, Wo
            if self.conv_impl == "cudnn":
                return F.conv2d(y.view(S, H, W, -1).permute(0, 3, 1, 2), *self.out_cudnn)
            w, b, Cout = self.out_conv
            out, _, _ = self._conv(y, S, H, W, (w, b, 1, 1), relu=False)
            return out.view(S, H, W, Cout).permute(0, 3, 1, 2)


class FusedBasicUpdate:
    """``BasicUpdateBlock`` (update.py:115-139) for the mixed-precision path, on f16 pixel-major rows ``[M = N*H*W, C]``:
    cuDNN convolutions WITHOUT bias on channels-last views of those rows, and everything between them -- bias, relu,
    the ``torch.cat``s, the GRU gates, the flow update -- on the ``cwm_raft_*_f16`` kernels, which write each result
    straight into its slot of the next convolution's input.  Per iteration: 1 lookup + 11 convolutions + 11 kernels
    (the eager block: ~80 launches).  z|r and flow-head|mask-head convolutions are stacked along the output channels.

    Row buffers (f16):  HX  [M, 384] = [h | inp | motion features (126) , flow (2)]   -> convz / convr input
                        RHX [M, 384] = [r*h | inp | motion features, flow]             -> convq input
                        CORFLO [M, 256] = [cor (192) | flo (64)]                       -> encoder.conv input
    """
    CORR_LD = 328   # 4 * 81 = 324 correlation channels padded to a multiple of 8 (16-byte rows)

    def __init__(self, ub, device, conv_impl=None):
        assert isinstance(ub, BasicUpdateBlock)
        _lib.require_f16("raft.FusedBasicUpdate")
        # 'tcgen05' (default): every convolution is an implicit GEMM on the repo's tensor-core kernel (cwm_conv2d_f16);
        # 'cudnn': the library convolutions of round 1, kept as the A/B reference (CWM_RAFT_CONV=cudnn)
        self.conv_impl = conv_impl or os.environ.get("CWM_RAFT_CONV", "tcgen05")
        assert self.conv_impl in ("tcgen05", "cudnn"), self.conv_impl
        self.fuse_gru = os.environ.get("CWM_RAFT_FUSE_GRU", "1") != "0" and ub.gru.convz1.out_channels % 64 == 0
        self._packed = {}

        def cl(w, out_pad=None, in_pad=None):
            w = w.detach().float()
            if out_pad is not None or in_pad is not None:
                full = torch.zeros(out_pad or w.shape[0], in_pad or w.shape[1], *w.shape[2:], device=w.device)
                full[:w.shape[0], :w.shape[1]] = w
                w = full
            out = w.to(device=device, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
            if self.conv_impl == "tcgen05":
                Cout, Cin, kh, kw = out.shape
                cin_pad = (Cin + 63) // 64 * 64
                pk = torch.zeros(Cout, kh, kw, cin_pad, dtype=torch.float16, device=device)
                pk[..., :Cin] = out.permute(0, 2, 3, 1)
                self._packed[id(out)] = (pk.reshape(Cout, -1).contiguous(), Cout, Cin, kh, kw)
            return out

        def f32(*bs, pad=None):
            b = torch.cat([x.detach().float().reshape(-1) for x in bs])
            if pad is not None:
                b = torch.cat([b, b.new_zeros(pad - b.numel())])
            return b.to(device).contiguous()

        enc, gru, fh, mk = ub.encoder, ub.gru, ub.flow_head, ub.mask
        self.C = gru.convz1.out_channels                       # hidden width (128)
        assert enc.convc1.in_channels <= self.CORR_LD and self.C % 8 == 0
        self.w_c1, self.b_c1 = cl(enc.convc1.weight, in_pad=self.CORR_LD), f32(enc.convc1.bias)
        self.w_c2, self.b_c2 = cl(enc.convc2.weight), f32(enc.convc2.bias)
        self.w_f1, self.b_f1 = cl(enc.convf1.weight, in_pad=8), f32(enc.convf1.bias)
        # the 7x7 convolution of the 2-channel flow as ONE K = 128 GEMM over im2col-ed rows (49 taps x 2 channels = 98)
        k7 = enc.convf1.kernel_size[0]
        wk = torch.zeros(enc.convf1.out_channels, 128, device=enc.convf1.weight.device)
        wk[:, :2 * k7 * k7] = enc.convf1.weight.detach().float().permute(0, 2, 3, 1).reshape(enc.convf1.out_channels, -1)
        self.w_f1k, self.k_f1 = cl(wk.view(-1, 128, 1, 1)), k7
        self.w_f2, self.b_f2 = cl(enc.convf2.weight), f32(enc.convf2.bias)
        self.w_cv, self.b_cv = cl(enc.conv.weight, out_pad=128), f32(enc.conv.bias, pad=128)
        self.w_zr1, self.b_zr1 = cl(torch.cat([gru.convz1.weight, gru.convr1.weight])), f32(gru.convz1.bias, gru.convr1.bias)
        self.w_zr2, self.b_zr2 = cl(torch.cat([gru.convz2.weight, gru.convr2.weight])), f32(gru.convz2.bias, gru.convr2.bias)
        self.w_q1, self.b_q1 = cl(gru.convq1.weight), f32(gru.convq1.bias)
        self.w_q2, self.b_q2 = cl(gru.convq2.weight), f32(gru.convq2.bias)
        self.w_fh1, self.b_fh1 = cl(fh.conv1.weight), f32(fh.conv1.bias)
        self.w_fh1m = cl(torch.cat([fh.conv1.weight, mk[0].weight]))
        self.b_m0 = f32(mk[0].bias)
        self.b_fh1m = f32(fh.conv1.bias, mk[0].bias)
        self.w_fh2, self.b_fh2 = cl(fh.conv2.weight, out_pad=8), f32(fh.conv2.bias)
        # the flow head's 3x3 / 256 -> 2 convolution as a 1x1 GEMM to its 18 per-tap products (row (ky*3 + kx)*2 + co), summed
        # over the 3x3 neighbourhood inside cwm_raft_flow_update_taps: a 3x3 implicit GEMM with N = 2 is 97 % padding
        wt = fh.conv2.weight.detach().float().permute(2, 3, 0, 1).reshape(-1, fh.conv2.in_channels)     # [(ky, kx, co), c]
        self.fuse_tail = (os.environ.get("CWM_RAFT_FUSE_TAIL", "1") != "0" and tuple(fh.conv2.kernel_size) == (3, 3)
                          and fh.conv2.out_channels == 2 and enc.conv.out_channels == self.C - 2)
        self.w_fh2t = cl(wt.view(wt.shape[0], -1, 1, 1), out_pad=24)
        self.w_m2 = cl(mk[2].weight)
        self.b_m2 = mk[2].bias.detach().to(device=device, dtype=torch.float16)
        self.w_m2q, self.b_m2q = cl(0.25 * mk[2].weight), f32(0.25 * mk[2].bias)   # .25 * mask(net), update.py:137

    class State:
        pass

    def begin(self, N, H, W, net, inp, flow_init=None):
        """net / inp: [N or 1, C, H, W] (tanh / relu already applied, raft_model.py:236-237)."""
        st, dev, C = self.State(), net.device, self.C
        st.N, st.H, st.W, st.M = N, H, W, N * H * W
        buf = lambda c, zero=False: (torch.zeros if zero else torch.empty)(st.M, c, dtype=torch.float16, device=dev)  # noqa: E731
        st.HX, st.RHX, st.CORFLO = buf(3 * C), buf(3 * C), buf(256)
        st.cor1, st.flo1, st.Z, st.Hd, st.fh1, st.mh = buf(256), buf(128), buf(C), buf(C), buf(256), buf(256)
        st.corr16, st.flow16 = buf(self.CORR_LD), buf(8, zero=True)
        # the implicit-GEMM kernel tiles whole image rows of <= 32 pixels (224 px frames: 28); wider maps use cuDNN
        st.tc = self.conv_impl == "tcgen05" and W <= 32
        st.flowcols = buf(128) if st.tc else None
        st.fhm = buf(512) if st.tc else None
        rows = lambda t: t.expand(N, -1, -1, -1).permute(0, 2, 3, 1).reshape(st.M, -1)  # noqa: E731
        st.HX[:, :C].copy_(rows(net))
        st.HX[:, C:2 * C].copy_(rows(inp))
        st.RHX[:, C:2 * C].copy_(st.HX[:, C:2 * C])
        if flow_init is not None:
            st.flow16[:, :2].copy_(rows(flow_init))
        st.taps = buf(24) if st.tc else None
        return st

    def _conv(self, st, rows, weight, padding, bias=None, relu=False, out=None):
        """f16 rows [M, Cin] (a column slice is fine) -> convolution output rows [M, Cout].  On the implicit-GEMM kernel
        ``bias`` / ``relu`` run in the epilogue and ``out`` may be a column slice of the next convolution's input; the
        cuDNN route returns the raw output (its callers apply bias / activation with the elementwise kernels)."""
        if st.tc:
            packed, Cout, Cin, kh, kw = self._packed[id(weight)]
            rows = rows[:, :Cin] if rows.shape[1] > Cin else rows
            assert rows.shape[1] == Cin and (kh // 2, kw // 2) == tuple(padding if isinstance(padding, tuple) else (padding, padding))
            if out is None:
                out = torch.empty(st.M, Cout, dtype=torch.float16, device=rows.device)
            assert out.shape == (st.M, Cout) and out.stride(1) == 1
            _lib.check(_lib.load().cwm_conv2d_f16(rows.data_ptr(), rows.stride(0), st.N, st.H, st.W, Cin, packed.data_ptr(),
                                                  Cout, kh, kw, kh // 2, kw // 2,
                                                  None if bias is None else bias.data_ptr(), int(bool(relu)),
                                                  out.data_ptr(), out.stride(0), _stream(rows)))
            return out
        x = rows.view(st.N, st.H, st.W, rows.shape[1]).permute(0, 3, 1, 2)
        y = F.conv2d(x, weight, None, 1, padding)
        if not y.is_contiguous(memory_format=torch.channels_last):
            y = y.contiguous(memory_format=torch.channels_last)
        return y.permute(0, 2, 3, 1).reshape(st.M, weight.shape[0])

    def step(self, st, corr_fn, coords1, emit):
        """One GRU iteration; ``coords1`` (fp32 [N, 2, H, W]) is updated in place.  Returns the upsampling logits
        ``0.25 * mask(net)`` ([N, 576, H, W], f16) when ``emit`` else None."""
        lib, M, C = _lib.load(), st.M, self.C
        s = _stream(coords1)
        p = lambda t: t.data_ptr()  # noqa: E731

        def bias_act(raw, ldx, bias, C_, dst, ld, dst2=None, ld2=0, tail=None, x_ptr=None):
            _lib.check(lib.cwm_raft_bias_act_f16(x_ptr or p(raw), ldx, p(bias), 1, C_, M, p(dst), ld,
                                                 p(dst2) if dst2 is not None else None, ld2,
                                                 p(tail) if tail is not None else None, 8, 2, s))

        with torch.cuda.device(coords1.device):
            lookup = (lib.cwm_raft_corr_lookup_f16_pyr16 if corr_fn.corr_pyramid[0].dtype == torch.float16
                      else lib.cwm_raft_corr_lookup_f16)
            _lib.check(lookup(_ptr_table(corr_fn.corr_pyramid), corr_fn.num_levels, corr_fn.radius, p(coords1), st.N, st.H, st.W,
                              p(st.corr16), self.CORR_LD, s))
            # BasicMotionEncoder (update.py:90-98)
            if st.tc:
                # bias + relu in the convolution epilogues, every result written straight into its consumer's input slot
                self._conv(st, st.corr16, self.w_c1, 0, self.b_c1, True, st.cor1)
                self._conv(st, st.cor1, self.w_c2, 1, self.b_c2, True, st.CORFLO[:, :192])
                _lib.check(lib.cwm_raft_im2col_flow(p(st.flow16), 8, st.N, st.H, st.W, self.k_f1, p(st.flowcols), 128, s))
                self._conv(st, st.flowcols, self.w_f1k, 0, self.b_f1, True, st.flo1)
                self._conv(st, st.flo1, self.w_f2, 1, self.b_f2, True, st.CORFLO[:, 192:])
            else:
                raw = self._conv(st, st.corr16, self.w_c1, 0)
                bias_act(raw, 256, self.b_c1, 256, st.cor1, 256)
                raw = self._conv(st, st.cor1, self.w_c2, 1)
                bias_act(raw, 192, self.b_c2, 192, st.CORFLO, 256)
                raw = self._conv(st, st.flow16, self.w_f1, 3)
                bias_act(raw, 128, self.b_f1, 128, st.flo1, 128)
                raw = self._conv(st, st.flo1, self.w_f2, 1)
                bias_act(raw, 64, self.b_f2, 64, st.CORFLO[:, 192:], 256)
            if st.tc and self.fuse_tail:
                # bias + relu in the epilogue, the 126 motion features + the flow stored into BOTH GRU input buffers
                pk = self._packed[id(self.w_cv)][0]
                _lib.check(lib.cwm_conv2d_dual_f16(p(st.CORFLO), 256, st.N, st.H, st.W, 256, p(pk), 128, 3, 3, 1, 1, p(self.b_cv), 1,
                                                   p(st.flow16), 8, p(st.HX[:, 2 * C:]), 3 * C, p(st.RHX[:, 2 * C:]), 3 * C, s))
            else:
                raw = self._conv(st, st.CORFLO, self.w_cv, 1)
                bias_act(raw, 128, self.b_cv, 128, st.HX[:, 2 * C:], 3 * C, st.RHX[:, 2 * C:], 3 * C, tail=st.flow16)
            # SepConvGRU (update.py:43-60): horizontal then vertical
            for w_zr, b_zr, w_q, b_q, pad, dense in ((self.w_zr1, self.b_zr1, self.w_q1, self.b_q1, (0, 2), None),
                                                     (self.w_zr2, self.b_zr2, self.w_q2, self.b_q2, (2, 0), st.Hd)):
                if st.tc and self.fuse_gru:
                    # gate / update arithmetic in the epilogues of the two convolutions: no raw tensors, no extra launches
                    pk_zr, pk_q = self._packed[id(w_zr)][0], self._packed[id(w_q)][0]
                    kh, kw = 2 * pad[0] + 1, 2 * pad[1] + 1
                    _lib.check(lib.cwm_conv2d_gru_gate_f16(p(st.HX), 3 * C, st.N, st.H, st.W, 3 * C, p(pk_zr), C, kh, kw,
                                                           pad[0], pad[1], p(b_zr), p(st.HX), 3 * C, p(st.Z), C,
                                                           p(st.RHX), 3 * C, s))
                    _lib.check(lib.cwm_conv2d_gru_update_f16(p(st.RHX), 3 * C, st.N, st.H, st.W, 3 * C, p(pk_q), C, kh, kw,
                                                             pad[0], pad[1], p(b_q), p(st.Z), C, p(st.HX), 3 * C,
                                                             p(dense) if dense is not None else None, s))
                    continue
                raw = self._conv(st, st.HX, w_zr, pad)
                _lib.check(lib.cwm_raft_gru_gate_f16(p(raw), p(b_zr), p(st.HX), 3 * C, C, M, p(st.Z), p(st.RHX), 3 * C, s))
                raw = self._conv(st, st.RHX, w_q, pad)
                _lib.check(lib.cwm_raft_gru_update_f16(p(raw), p(b_q), p(st.Z), p(st.HX), 3 * C, C, M,
                                                       p(dense) if dense is not None else None, s))
            # FlowHead (update.py:13-14) and, where a prediction is emitted, the mask head (:131-137)
            if st.tc:
                if emit:   # flow head conv1 | mask head conv0 stacked: one convolution, both biases + relu in the epilogue
                    self._conv(st, st.Hd, self.w_fh1m, 1, self.b_fh1m, True, st.fhm)
                    fh1, mh = st.fhm[:, :256], st.fhm[:, 256:]
                else:
                    fh1, mh = self._conv(st, st.Hd, self.w_fh1, 1, self.b_fh1, True, st.fh1), None
                if self.fuse_tail:
                    self._conv(st, fh1, self.w_fh2t, 0, None, False, st.taps)
                    _lib.check(lib.cwm_raft_flow_update_taps(p(st.taps), 24, p(self.b_fh2), p(coords1), st.N, st.H, st.W,
                                                             p(st.flow16), None, 0, None, 0, s))
                    raw = None
                else:
                    raw = self._conv(st, fh1, self.w_fh2, 1)
            else:
                raw = self._conv(st, st.Hd, self.w_fh1m if emit else self.w_fh1, 1)
                ldx = raw.shape[1]
                bias_act(raw, ldx, self.b_fh1, 256, st.fh1, 256)
                if emit:
                    bias_act(raw, ldx, self.b_m0, 256, st.mh, 256, x_ptr=p(raw[:, 256:]))
                mh = st.mh
                raw = self._conv(st, st.fh1, self.w_fh2, 1)
            if raw is not None:
                _lib.check(lib.cwm_raft_flow_update(p(raw), 8, p(self.b_fh2), p(coords1), st.N, st.H, st.W, p(st.flow16), s))
            if not emit:
                return None
            if st.tc:
                # mask head's 1x1 convolution with its bias in the epilogue; 0.25 is folded into weights and bias (exact)
                out = self._conv(st, mh, self.w_m2q, 0, self.b_m2q, False)
                return out.view(st.N, st.H, st.W, out.shape[1]).permute(0, 3, 1, 2)
            mh = st.mh.view(st.N, st.H, st.W, 256).permute(0, 3, 1, 2)
            return .25 * F.conv2d(mh, self.w_m2, self.b_m2)


def get_args(cmd=None):
    """raft_model.py:36-51 (same option names and defaults)."""
    parser = argparse.ArgumentParser()
    parser.add_argument('--corr_levels', type=int, default=4)
    parser.add_argument('--corr_radius', type=int, default=4)
    parser.add_argument('--output_dim', type=int, default=None)
    parser.add_argument('--iters', type=int, default=None)
    parser.add_argument('--dropout', type=float, default=0.0)
    parser.add_argument('--mixed_precision', action='store_true')
    parser.add_argument('--small', action='store_true')
    parser.add_argument('--gpus', type=int, nargs='+', default=[0])
    return parser.parse_args() if cmd is None else parser.parse_args(cmd)


def upflow8(flow, mode='bilinear'):
    """utils.py:88-90 (RAFT-small has no upsampling mask)."""
    return 8 * F.interpolate(flow, size=(8 * flow.shape[2], 8 * flow.shape[3]), mode=mode, align_corners=True)


class RAFT(nn.Module):
    """``cwm.models.raft.raft_model.RAFT`` (raft_model.py:114-301): same ``args``, attributes (``iters``,
    ``multiframe``, ``scale_inputs``, ``hidden_dim``, ``context_dim``), parameter names and forward semantics.
    The correlation pyramid, its per-iteration lookups and the convex upsampling run on ``csrc/raftcorr.cu``; in
    ``test_mode`` only the last iteration is upsampled (the reference upsamples all of them and returns the last).
    ``args.mixed_precision`` (the reference's autocast option, raft_model.py:216-222): the encoders run under autocast
    and the recurrent block on f16 channels-last tensors (``args.half_update=False`` restores plain autocast); for
    RAFT-large inference that block is ``FusedBasicUpdate`` (``args.fused_update=False``: the same in eager torch ops)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.multiframe = getattr(args, 'multiframe', True)
        self.scale_inputs = getattr(args, 'scale_inputs', True)
        self._iters = None
        if getattr(args, 'iters', None) is not None:
            self.iters = args.iters
        small = bool(getattr(args, 'small', False))
        self.hidden_dim = hdim = 96 if small else 128
        self.context_dim = cdim = 64 if small else 128
        args.corr_levels = 4
        args.corr_radius = 3 if small else 4
        for name, default in (('dropout', 0), ('alternate_corr', False), ('mixed_precision', False), ('output_dim', None)):
            if not hasattr(args, name):
                setattr(args, name, default)
        if args.alternate_corr:
            raise NotImplementedError("alternate_corr needs the reference's optional alt_cuda_corr extension "
                                      "(corr.py:5-9); CorrBlock is the supported path")
        if small:
            self.fnet = SmallEncoder(output_dim=128, norm_fn='instance', dropout=args.dropout)
            self.cnet = SmallEncoder(output_dim=hdim + cdim, norm_fn='none', dropout=args.dropout)
            self.update_block = SmallUpdateBlock(args, hidden_dim=hdim)
        else:
            self.fnet = BasicEncoder(output_dim=256, norm_fn='instance', dropout=args.dropout)
            self.cnet = BasicEncoder(output_dim=hdim + cdim, norm_fn='batch', dropout=args.dropout)
            self.update_block = BasicUpdateBlock(args, hidden_dim=hdim)
        self.output_block = None
        if args.output_dim is not None:
            self.output_block = nn.Sequential(nn.Conv2d(hdim, 192 if small else 256, 3, padding=1), nn.ReLU(inplace=True),
                                              nn.Conv2d(192 if small else 256, args.output_dim, 1, padding=0))

    @property
    def iters(self):
        return getattr(self, '_iters', None)

    @iters.setter
    def iters(self, value=None):
        self._iters = value

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def initialize_flow(self, img):
        N, _, H, W = img.shape
        grid = coords_grid(N, H // 8, W // 8, device=img.device, dtype=img.dtype)
        return grid, grid.clone()

    def upsample_flow(self, flow, mask):
        return upsample_flow(flow, mask)

    def _half_update_block(self):
        """f16 channels-last twin of ``update_block``, rebuilt when the fp32 parameters change."""
        import copy
        key = tuple((p.data_ptr(), p._version) for p in self.update_block.parameters())
        cached = getattr(self, '_half_ub', None)
        if cached is None or cached[0] != key:
            twin = copy.deepcopy(self.update_block).half().to(memory_format=torch.channels_last).eval().requires_grad_(False)
            object.__setattr__(self, '_half_ub', (key, twin))  # not a submodule: keeps state_dict() the reference's
            cached = self._half_ub
        return cached[1]

    def _fused_encoder(self, which):
        """Packed twin of ``fnet`` / ``cnet`` (rebuilt when its parameters or batch-norm statistics change)."""
        enc = getattr(self, which)
        key = tuple((p.data_ptr(), p._version) for p in list(enc.parameters()) + list(enc.buffers()))
        cached = getattr(self, '_fused_' + which, None)
        if cached is None or cached[0] != key:
            device = next(enc.parameters()).device
            object.__setattr__(self, '_fused_' + which, (key, FusedFeatureEncoder(enc, device)))
            cached = getattr(self, '_fused_' + which)
        return cached[1]

    def _fused_feature_encoder(self):
        return self._fused_encoder('fnet')

    def _fused_update_block(self):
        key = tuple((p.data_ptr(), p._version) for p in self.update_block.parameters())
        cached = getattr(self, '_fused_ub', None)
        if cached is None or cached[0] != key:
            device = next(self.update_block.parameters()).device
            object.__setattr__(self, '_fused_ub', (key, FusedBasicUpdate(self.update_block, device)))
            cached = self._fused_ub
        return cached[1]

    def _forward_two_images(self, image1, image2, iters=24, flow_init=None, upsample=True, test_mode=True, **kwargs):
        """raft_model.py:199-277.  One of the two images may have batch 1 while the other has N (a counterfactual
        sweep shares its first frame): that image goes through the encoders once and is broadcast -- same result,
        the per-sample instance norm makes the encoders independent of what else is in the batch."""
        if self.iters is not None:
            iters = self.iters
        # images arrive in [0, 255] / in_scale: the multiframe front end hands its [0, 1] frames over unscaled (in_scale = 255)
        # so that the fused encoders can fold the whole normalisation 2 (x / 255) - 1 into their first kernel
        in_scale = float(kwargs.pop('_in_scale', 1.0))
        raw1, raw2 = image1, image2
        normalise = lambda im: (2 * ((im * in_scale if in_scale != 1.0 else im) / 255.0) - 1.0).contiguous()  # noqa: E731
        n1, n2 = image1.shape[0], image2.shape[0]
        N = max(n1, n2)
        assert n1 in (1, N) and n2 in (1, N), (image1.shape, image2.shape)
        amp = bool(self.args.mixed_precision or image1.dtype in (torch.float16, torch.bfloat16))
        fused_fnet = (amp and test_mode and isinstance(self.fnet, BasicEncoder) and self.fnet.norm_fn == 'instance'
                      and bool(getattr(self.args, 'fused_encoder', True)) and os.environ.get("CWM_RAFT_ENCODER", "fused") != "eager")
        if fused_fnet:
            # f16 pixel-major rows, implicit-GEMM convolutions, instance norm + relu (+ shortcut) in cwm_instnorm_f16;
            # the input normalisation rides on the stem's im2col
            fmaps = self._fused_feature_encoder()([raw1, raw2], scale=2.0 * in_scale / 255.0, shift=-1.0)
        else:
            image1, image2 = normalise(raw1), normalise(raw2)
            with torch.autocast("cuda", enabled=amp):
                fmaps = self.fnet(torch.cat([image1, image2], dim=0))  # both frames in one batch (raft_model.py:221-222)
        rows_ok = (fused_fnet and fmaps.dtype == torch.float16 and fmaps.shape[1] % 64 == 0 and n2 == N
                   and fmaps.permute(0, 2, 3, 1).is_contiguous() and os.environ.get("CWM_RAFT_CORR", "tc") == "tc")
        # will the recurrent block run fused (f16 rows, f16 lookups)?  Decided here because it also picks the pyramid's type
        use_fused_update = (amp and bool(getattr(self.args, 'half_update', True)) and test_mode
                            and isinstance(self.update_block, BasicUpdateBlock) and self.output_block is None and iters > 0
                            and bool(getattr(self.args, 'fused_update', True)))
        if rows_ok:
            # the fused encoder's output IS the operand layout of the volume: f16 pixel-major rows, products exact in fp32
            hw = fmaps.shape[2] * fmaps.shape[3]
            rows = fmaps.permute(0, 2, 3, 1).reshape(-1, fmaps.shape[1])
            pyr16 = (use_fused_update and self.args.corr_radius == 4 and self.args.corr_levels <= 4 and hw % 8 == 0
                     and fmaps.shape[3] <= 32 and os.environ.get("CWM_RAFT_PYR16", "0") != "0"   # opt-in: measured 11.09 -> 11.07 ms only
                     and os.environ.get("CWM_RAFT_CONV", "tcgen05") == "tcgen05"
                     and os.environ.get("CWM_RAFT_LOOKUP", "fast")[0] not in "e0")
            corr_fn = CorrBlock.from_rows(rows[:n1 * hw], rows[n1 * hw:], N, fmaps.shape[2], fmaps.shape[3],
                                          num_levels=self.args.corr_levels, radius=self.args.corr_radius, f16_pyramid=pyr16)
        else:
            fmap1 = fmaps[:n1].float().expand(N, -1, -1, -1)
            fmap2 = fmaps[n1:].float().expand(N, -1, -1, -1)
            corr_fn = CorrBlock(fmap1, fmap2, num_levels=self.args.corr_levels, radius=self.args.corr_radius)
        # the context network: batch norm folded into its convolutions (inference statistics only)
        fused_cnet = (fused_fnet and isinstance(self.cnet, BasicEncoder) and self.cnet.norm_fn == 'batch'
                      and not self.cnet.training)
        if not fused_cnet and image1 is raw1:      # (not normalised yet: the fused feature encoder took the raw frames)
            image1 = normalise(raw1)
        with torch.autocast("cuda", enabled=amp):
            ctx = (self._fused_encoder('cnet')(raw1, scale=2.0 * in_scale / 255.0, shift=-1.0) if fused_cnet
                   else self.cnet(image1))
            net, inp = torch.split(ctx, [self.hidden_dim, self.context_dim], dim=1)
            net = torch.tanh(net).expand(N, -1, -1, -1)
            inp = torch.relu(inp).expand(N, -1, -1, -1)
        coords0, coords1 = self.initialize_flow(image1 if n1 == N else image2)
        if flow_init is not None:
            coords1 = coords1 + flow_init
        has_mask_head = isinstance(self.update_block, BasicUpdateBlock)
        update_block = self.update_block
        half_update = amp and bool(getattr(self.args, 'half_update', True))
        if use_fused_update:
            # RAFT-large, inference: the recurrent block on cuDNN convolutions + the cwm_raft_*_f16 kernels
            fused = self._fused_update_block()
            st = fused.begin(N, coords0.shape[2], coords0.shape[3], net, inp, flow_init)
            coords1 = coords1.float().contiguous().clone()
            up_mask = None
            for itr in range(iters):
                up_mask = fused.step(st, corr_fn, coords1, emit=(itr + 1 == iters))
            flow_low = coords1 - coords0
            return flow_low, self.upsample_flow(flow_low, up_mask)
        if half_update:
            # mixed precision without autocast's per-call casts and cuDNN's NCHW<->NHWC transforms: the recurrent
            # block runs on f16 channels-last activations and an f16 channels-last copy of its weights
            update_block = self._half_update_block()
            to16 = lambda t: t.to(dtype=torch.float16, memory_format=torch.channels_last)  # noqa: E731
            net, inp = to16(net), to16(inp)
        predictions = []
        flow_up = None
        for itr in range(iters):
            coords1 = coords1.detach()
            corr = corr_fn(coords1)
            flow = coords1 - coords0
            if half_update:
                corr, flow = to16(corr), to16(flow)
            emit = (not test_mode) or itr + 1 == iters  # test_mode only returns the last prediction
            with torch.autocast("cuda", enabled=amp and not half_update):
                if has_mask_head:  # the upsampling logits are only needed where a prediction is emitted
                    net, up_mask, delta_flow = update_block(net, inp, corr, flow, upsample=emit)
                else:
                    net, up_mask, delta_flow = update_block(net, inp, corr, flow)
            coords1 = coords1 + delta_flow.float()
            if not emit:
                continue
            out = self.output_block(net.float()) if self.output_block is not None else coords1 - coords0
            flow_up = upflow8(out) if up_mask is None else self.upsample_flow(out, up_mask)
            predictions.append(flow_up)
        if test_mode:
            return coords1 - coords0, flow_up
        return predictions

    def forward(self, *args, **kwargs):
        """raft_model.py:279-301: ``[B, T, 3, H, W]`` frames in [0, 1] -> flows ``[B, T-1, 2, H, W]``.
        ``shared_frame=t`` (extension): frame t is identical across the batch and is encoded once."""
        shared = kwargs.pop('shared_frame', None)
        if not self.multiframe:
            return self._forward_two_images(*args, **kwargs)
        x = args[0]
        kwargs['_in_scale'] = 255.0 if self.scale_inputs else 1.0     # applied inside _forward_two_images
        if x.dim() == 4:
            x = x.unsqueeze(1)
        assert x.dim() == 5, x.shape
        if x.size(1) == 1:  # a single frame is repeated
            x = x.repeat(1, 2, 1, 1, 1)
        backward = kwargs.get('backward', False)
        frame = lambda t: x[:1, t] if shared is not None and t == shared else x[:, t]  # noqa: E731
        flows = []
        for t in range(x.size(1) - 1):
            pair = (frame(t + 1), frame(t)) if backward else (frame(t), frame(t + 1))
            flow = self._forward_two_images(*pair, *args[1:], **kwargs)[-1]
            if backward:
                flows.insert(0, flow)
            else:
                flows.append(flow)
        return torch.stack(flows, 1)


def load_raft_model(load_path=None, ignore_prefix=None, multiframe=True, scale_inputs=True, output_dim=None, **kwargs):
    """raft_model.py:55-99: builds a RAFT and loads a published checkpoint (``module.`` prefixes stripped)."""
    if ((load_path is None) or (not os.path.exists(load_path))) and (output_dim is None):
        raise ValueError(f"{load_path} is not a valid raft checkpoint (the reference fetches them with "
                         "cwm/models/raft/download_raft_checkpoints.sh)")
    args = get_args("")
    for k, v in kwargs.items():
        setattr(args, k, v)
    args.multiframe, args.scale_inputs, args.output_dim = multiframe, scale_inputs, output_dim
    model = RAFT(args)
    if load_path is not None:
        weights = torch.load(load_path, map_location=torch.device("cpu"))
        weights = {k.replace('module.', ''): v for k, v in weights.items()}
        if ignore_prefix is not None:
            weights = {k.replace(ignore_prefix, ''): v for k, v in weights.items()}
        print(model.load_state_dict(weights, strict=False), type(model).__name__, load_path)
    return model
