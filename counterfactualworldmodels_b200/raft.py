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

import torch

from . import _lib


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
        with torch.cuda.device(fmap1.device):
            _lib.check(_lib.load().cwm_raft_corr_pyramid(fmap1.data_ptr(), fmap2.data_ptr(), batch, dim, ht, wd,
                                                         num_levels, _ptr_table(self.corr_pyramid), _stream(fmap1)))

    def __call__(self, coords):
        coords = _req(coords, "CorrBlock coords", 4)
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
