"""Host-side mirror of ``cwm/models/sampling.py:128-286`` (``FlowSampleFilter``), SURVEY.md section 8(f) rank 2.

Same constructor arguments, method names and return shapes; the reductions run in libcwm_b200 (``csrc/flowstats.cu``):
one pass over every flow sample produces all per-sample statistics, the filter mask is derived from them, and the
rejected samples are zeroed in place on the strided ``[B, 2, H, W, S]`` view the caller hands over.  No CPU fallback.
"""
import ctypes

import torch
from torch import nn

from . import _lib
from .masking import EnergySamplingMaskingGenerator, RotatedTableEnergyMaskingGenerator  # noqa: F401  (sampling.py:11-126)

_METHOD_BITS = {'patch_magnitude': 1, 'flow_area': 2, 'num_corners': 4}


_WS = {}


def _workspace(B, H, W, S, device):
    """Scratch of the two reductions (per device, grown on demand)."""
    need = int(_lib.load().cwm_flow_stats_workspace_bytes(B, H, W, S))
    ws = _WS.get(str(device))
    if ws is None or ws.numel() < need:
        ws = _WS[str(device)] = torch.empty(need, dtype=torch.uint8, device=device)
    return ws


def _strides(t, n):
    assert t.dim() == n, t.shape
    return (ctypes.c_int64 * n)(*t.stride())


def flow_sample_stats(flow_samples, active_patches=None, magnitude_threshold=0.0):
    """-> float32 [B, S, 5]: patch_flow_mag, flow_area, num_corners, min |flow|, max |flow| per sample."""
    lib = _lib.load()
    if flow_samples.device.type != "cuda":
        raise RuntimeError("flow statistics: tensors must live on a CUDA (B200) device; there is no CPU fallback")
    assert flow_samples.dim() == 5 and flow_samples.size(1) == 2, flow_samples.shape
    if flow_samples.dtype != torch.float32:
        flow_samples = flow_samples.float()
    B, _, H, W, S = flow_samples.shape
    stats = torch.empty(B, S, 5, dtype=torch.float32, device=flow_samples.device)
    a_ptr, a_str, h, w = None, None, 0, 0
    if active_patches is not None:
        _, num_patches, s_a = active_patches.shape
        assert s_a == S, (active_patches.shape, S)
        assert H == W, "the inference of patch size assumes H == W"
        h = w = int((num_patches / 2) ** 0.5)  # num_patches counts the patches of 2 frames (sampling.py:184-186)
        a = active_patches.to(flow_samples.device)
        a = a.view(torch.uint8) if a.dtype == torch.bool else (a != 0).view(torch.uint8)
        a_ptr, a_str = a.data_ptr(), _strides(a, 3)
    with torch.cuda.device(flow_samples.device):
        ws = _workspace(B, H, W, S, flow_samples.device)
        stream = torch.cuda.current_stream(flow_samples.device).cuda_stream
        _lib.check(lib.cwm_flow_sample_stats(flow_samples.data_ptr(), _strides(flow_samples, 5), B, H, W, S, a_ptr, a_str,
                                             h, w, float(magnitude_threshold), stats.data_ptr(), ws.data_ptr(),
                                             ws.numel(), stream))
    return stats


class FlowSampleFilter(nn.Module):
    """Filter out flow samples based on predefined heuristics (sampling.py:128-286)."""

    def __init__(self, filter_methods=['patch_magnitude', 'flow_area', 'num_corners'], flow_magnitude_threshold=5.0,
                 flow_area_threshold=0.75, num_corners_threshold=2):
        super().__init__()
        self.filter_methods = filter_methods
        self.flow_magnitude_threshold = flow_magnitude_threshold
        self.flow_area_threshold = flow_area_threshold
        self.num_corners_threshold = num_corners_threshold

    def __repr__(self):
        return ("filtering by %s\nusing flow_magnitude_threshold %0.1f\n" +
                "using flow_area_threshold %0.2f\n" +
                "using num_corners_threshold %d") % \
            (self.filter_methods, self.flow_magnitude_threshold, self.flow_area_threshold, self.num_corners_threshold)

    def compute_flow_statistics(self, flow_samples, active_patches=None):
        """All per-sample quantities the filters need in one pass: dict of [B, S] tensors."""
        st = flow_sample_stats(flow_samples, active_patches, self.flow_magnitude_threshold)
        return dict(patch_flow_mag=st[..., 0], flow_area=st[..., 1], num_corners=st[..., 2], min=st[..., 3],
                    max=st[..., 4], _raw=st)

    def filter_by_patch_magnitude(self, patch_flow_mag):
        assert self.flow_magnitude_threshold is not None
        return patch_flow_mag < self.flow_magnitude_threshold

    def filter_by_flow_area(self, flow_area):
        """NB: takes the per-sample area fraction (``compute_flow_statistics``), not the full magnitude map."""
        assert self.flow_magnitude_threshold is not None
        assert self.flow_area_threshold is not None
        return flow_area > self.flow_area_threshold

    def filter_by_num_corners(self, num_corners):
        assert self.flow_magnitude_threshold is not None
        return num_corners >= self.num_corners_threshold

    def filter_mask(self, flow_samples, active_patches):
        """-> (bool [B, S], statistics): 1 means the sample is filtered out (sampling.py:263-279)."""
        lib = _lib.load()
        methods = 0
        for method in self.filter_methods:
            if method not in _METHOD_BITS:
                raise ValueError(f'Filter method must be one of {self.filter_methods}, but got {method}')
            methods |= _METHOD_BITS[method]
        st = self.compute_flow_statistics(flow_samples, active_patches)
        B, S = st["_raw"].shape[:2]
        mask = torch.empty(B, S, dtype=torch.bool, device=flow_samples.device)
        with torch.cuda.device(flow_samples.device):
            stream = torch.cuda.current_stream(flow_samples.device).cuda_stream
            _lib.check(lib.cwm_flow_filter_mask(st["_raw"].data_ptr(), B, S, methods,
                                                float(self.flow_magnitude_threshold), float(self.flow_area_threshold),
                                                float(self.num_corners_threshold), mask.data_ptr(), stream))
        return mask, st

    def forward(self, flow_samples, active_patches):
        """flow_samples [B, 2, H, W, S], active_patches [B, num_patches, S] -> (flow_samples with the rejected samples
        set to zero IN PLACE, filter mask expanded to the shape of flow_samples) -- sampling.py:252-286.

        Stride contract (differs from the reference): the reference returns ``flow_samples.contiguous()``,
        a 411 MB copy for a 1024-sample sweep because its input is the permuted ``'(b s) c h w -> b c h w s'`` view;
        here the SAME strided view comes back (every kernel downstream takes strides).  Call ``.contiguous()`` on the
        result before ``.view()``-ing it."""
        lib = _lib.load()
        B, _, H, W, S = flow_samples.shape
        if flow_samples.dtype != torch.float32:
            # e.g. fp16 flows from an autocast flow network: the statistics run on an fp32 copy, the rejected samples
            # are zeroed in the caller's tensor (the reference accepts any float dtype, sampling.py:252-286)
            mask, _ = self.filter_mask(flow_samples.float(), active_patches)
            flow_samples.mul_((~mask).view(B, 1, 1, 1, S).to(flow_samples.dtype))
            return flow_samples, mask.view(B, 1, 1, 1, S).expand_as(flow_samples)
        mask, _ = self.filter_mask(flow_samples, active_patches)
        with torch.cuda.device(flow_samples.device):
            stream = torch.cuda.current_stream(flow_samples.device).cuda_stream
            _lib.check(lib.cwm_flow_zero_filtered(flow_samples.data_ptr(), _strides(flow_samples, 5), B, H, W, S,
                                                  mask.data_ptr(), stream))
        return flow_samples, mask.view(B, 1, 1, 1, S).expand_as(flow_samples)


def flow_magnitude_sum(flow_samples, filter_mask=None, stats=None, normalize_per_sample=False, eps=1e-2, out=None):
    """Partial numerator of ``flow_mags.mean(-1)`` (segmentation.py:257-267): float32 [B, H, W] sum over this call's
    samples; ``out`` accumulates (chunks of a sweep, or the local shard before the all-reduce)."""
    lib = _lib.load()
    if flow_samples.device.type != "cuda":
        raise RuntimeError("flow statistics: tensors must live on a CUDA (B200) device; there is no CPU fallback")
    if flow_samples.dtype != torch.float32:
        flow_samples = flow_samples.float()
    B, _, H, W, S = flow_samples.shape
    if normalize_per_sample and stats is None:
        stats = flow_sample_stats(flow_samples)
    accumulate = out is not None
    if out is None:
        out = torch.empty(B, H, W, dtype=torch.float32, device=flow_samples.device)
    with torch.cuda.device(flow_samples.device):
        ws = _workspace(B, H, W, S, flow_samples.device)
        stream = torch.cuda.current_stream(flow_samples.device).cuda_stream
        _lib.check(lib.cwm_flow_magnitude_sum(
            flow_samples.data_ptr(), _strides(flow_samples, 5), B, H, W, S,
            None if filter_mask is None else filter_mask.contiguous().data_ptr(),
            None if stats is None else stats.contiguous().data_ptr(), int(bool(normalize_per_sample)), float(eps),
            int(accumulate), out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out


def motion_map_finalize(sums, count, normalize=True, eps=1e-2):
    """sums float32 [B, H, W] -> motion map [B, 1, H, W] (segmentation.py:268-276)."""
    lib = _lib.load()
    B, H, W = sums.shape
    sums = sums.contiguous()
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=sums.device)
    with torch.cuda.device(sums.device):
        stream = torch.cuda.current_stream(sums.device).cuda_stream
        _lib.check(lib.cwm_motion_map_finalize(sums.data_ptr(), B, H, W, float(count), int(bool(normalize)), float(eps),
                                               out.data_ptr(), stream))
    return out


def flow_corrs(flow_samples, downsample=1, use_covariance=False):
    """``FlowGenerator.compute_flow_corrs`` with default options (segmentation.py:478-547) ->
    float32 [B, 1, H/ds, W/ds, H/ds, W/ds]."""
    lib = _lib.load()
    if flow_samples.device.type != "cuda":
        raise RuntimeError("flow statistics: tensors must live on a CUDA (B200) device; there is no CPU fallback")
    if flow_samples.dtype != torch.float32:
        flow_samples = flow_samples.float()
    B, C, H, W, S = flow_samples.shape
    assert C == 2, flow_samples.shape
    ds = int(downsample)
    n_h, n_w = H // ds, W // ds
    out = torch.empty(B, n_h * n_w, n_h * n_w, dtype=torch.float32, device=flow_samples.device)
    with torch.cuda.device(flow_samples.device):
        ws = torch.empty(int(lib.cwm_flow_corrs_workspace_bytes(B, H, W, S, ds)), dtype=torch.uint8,
                         device=flow_samples.device)
        stream = torch.cuda.current_stream(flow_samples.device).cuda_stream
        _lib.check(lib.cwm_flow_corrs(flow_samples.data_ptr(), _strides(flow_samples, 5), B, H, W, S, ds,
                                      int(bool(use_covariance)), out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out.view(B, 1, n_h, n_w, n_h, n_w)
