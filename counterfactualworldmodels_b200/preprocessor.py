"""Stream preprocessors of the conjoined models (``cwm/models/preprocessor.py``): which frames of the input video a
stream sees and how the IMU sequence is shaped.  Pure index-select / reshape -- no arithmetic -- so they stay on the
host side.  The RAFT-based ``FramePairFlow`` family (``flowback_rgb01`` & co.) runs on a
``raft.RAFT`` (or any RAFT-like module) handed in by the caller or loaded from a checkpoint path, see ``FramePairFlow``.
"""
import copy
from functools import partial

import torch
import torch.nn as nn

from .vmae import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD


class Preprocessor(nn.Module):
    """preprocessor.py:18-136: select ``frames_list`` of the input along the temporal dim."""
    num_channels = None

    def __init__(self, frames_list=None, temporal_dim=2, channel_dim=None, preproc_func=None, preproc_kwargs={},
                 num_frames=None, num_channels=None, stack=False, unnormalize=False, *args, **kwargs):
        super().__init__()
        if stack:
            raise NotImplementedError("stacked frame inputs (rgb01stack) are not used by any CWM factory")
        if preproc_func is not None:
            raise NotImplementedError("custom preproc_func")
        self.set_frames_list(frames_list)
        self.temporal_dim = temporal_dim
        self.channel_dim = (1 if temporal_dim == 2 else 2) if channel_dim is None else channel_dim
        assert self.channel_dim != self.temporal_dim
        self.num_frames = num_frames
        if num_channels is not None:
            self.num_channels = num_channels
        self.stack = False
        # get_preprocessor(name, unnormalize=True) composes imagenet_unnormalize (preprocessor.py:364-367)
        self.unnormalize = unnormalize

    @property
    def t_dim(self):
        return self.temporal_dim

    @property
    def c_dim(self):
        return self.channel_dim

    def set_frames_list(self, frames_list):
        """An int t means the frame pair (t, t + 1); any other iterable is taken as it is (preprocessor.py:56-64)."""
        if isinstance(frames_list, int):
            frames_list = [frames_list, frames_list + 1]
        elif frames_list is not None:
            frames_list = list(frames_list)
        self.frames_list = frames_list
        if frames_list is not None:
            self.num_input_frames = len(frames_list)

    def get_num_frames(self):
        if self.num_frames is not None:
            return self.num_frames
        return None if self.frames_list is None else len(self.frames_list)

    def get_num_channels(self, x):
        return x.shape[self.c_dim] if 0 <= self.c_dim < x.dim() else 0

    def set_input_dims(self, x):
        """Records the input's channel / frame counts; the frame list defaults to all frames and is taken modulo T."""
        self.num_input_channels = self.get_num_channels(x)
        self.T = T = x.shape[self.t_dim]
        if self.frames_list is None:
            self.frames_list, self.num_input_frames = list(range(T)), T
        self.frames_list = [t % T for t in self.frames_list]

    def set_output_dims(self, x):
        """First call fixes the output's channel / frame counts, later calls check them."""
        for attr, got in (("num_channels", self.get_num_channels(x)), ("num_frames", x.shape[self.t_dim])):
            want = getattr(self, attr)
            if want is None:
                setattr(self, attr, got)
            else:
                assert want == got, (attr, want, got)

    def _select_frames(self, x, frames, dim):
        return torch.index_select(x, dim=dim, index=torch.as_tensor(frames, dtype=torch.long, device=x.device))

    def get_input_frames(self, x):
        return self._select_frames(x, self.frames_list, self.temporal_dim)

    def get_output_frames(self, y, temporal_dim=None):
        return self._select_frames(y, self.frames_list[-self.num_frames:], self.t_dim if temporal_dim is None else temporal_dim)

    def forward(self, x, *args, **kwargs):
        self.set_input_dims(x)
        x = self.get_input_frames(x)
        if self.unnormalize:
            # imagenet_unnormalize (models/utils.py:23-31).  No shipped factory takes this branch for an RGB stream
            # (imu400_base_4x4 passes unnormalize=False, conjoined_vmae.py:1236), so it is plain host-side input
            # construction rather than a fused kernel.
            shape = [1] * x.dim()
            shape[self.c_dim] = 3
            mean = torch.as_tensor(IMAGENET_DEFAULT_MEAN, device=x.device, dtype=x.dtype).view(shape)
            std = torch.as_tensor(IMAGENET_DEFAULT_STD, device=x.device, dtype=x.dtype).view(shape)
            x = x * std + mean
        self.set_output_dims(x)
        return x


class IMU(Preprocessor):
    """preprocessor.py:169-206: [B, 6, L] -> [B, 6, L, 1, 1]; no frames."""
    num_frames = None
    num_channels = 6

    def __init__(self, sequence_length=None, frames_list=None, temporal_dim=2, channel_dim=None, *args, **kwargs):
        super().__init__(frames_list=frames_list, temporal_dim=temporal_dim, channel_dim=channel_dim)
        self.num_frames = None
        self.sequence_length = sequence_length

    def set_output_dims(self, x):
        super().set_output_dims(x)
        self.num_frames = None
        if self.sequence_length is not None:
            assert self.sequence_length == x.shape[self.t_dim], (self.sequence_length, x.shape[self.t_dim])

    def get_sequence_length(self):
        return self.sequence_length

    def forward(self, imu=None, timestamps=None, *args, **kwargs):
        if imu is None:
            return None
        imu = imu.unsqueeze(-1).unsqueeze(-1)
        self.set_input_dims(imu)
        self.set_output_dims(imu)
        return imu


class FramePairFlow(Preprocessor):
    """preprocessor.py:208-285: the stream input is RAFT optical flow between the two selected frames, optionally with
    the backward flow and the second frame's RGB appended -- ``flowback_rgb01`` (7 channels, one output frame) is the
    main stream of the flow2imu model.

    ``flow_model``: a RAFT-like ``nn.Module`` (``counterfactualworldmodels_b200.raft.RAFT`` or the reference's:
    ``flow_model([B, T, 3, H, W] in [0, 1], iters=, backward=) -> [B, T-1, 2, H, W]``), or ``flow_model_ckpt`` = path of
    a published RAFT checkpoint (the reference's only way in, :263-264).  The pipeline is the reference's:
    unnormalise -> cat([flow, backward flow, normalised rgb of frame 1]) -> flow / (size / 2).
    A plain callable (not an ``nn.Module``) keeps the older contract of this mirror: it receives the selected frames and
    returns the finished ``[B, num_channels, 1, H, W]`` stream input."""
    num_channels = 2

    def __init__(self, iters=24, backward=False, unnormalize_rgb=True, normalize_flow=True, concat_backward=False,
                 concat_rgb=False, flow_model_ckpt=None, flow_model=None, frames_list=None, temporal_dim=2, **kwargs):
        super().__init__(frames_list=frames_list, temporal_dim=temporal_dim)
        if flow_model is None and flow_model_ckpt is not None:
            from .raft import load_raft_model
            flow_model = load_raft_model(flow_model_ckpt).eval().requires_grad_(False)
        if isinstance(flow_model, nn.Module):
            self.flow_model = flow_model.eval().requires_grad_(False)
        else:
            object.__setattr__(self, 'flow_model', flow_model)
        self.iters, self.backward = iters, backward
        self.unnormalize_rgb, self.normalize_flow = unnormalize_rgb, normalize_flow
        self._concat_backward, self._concat_rgb = concat_backward, concat_rgb
        self.num_channels = 2 + (2 if concat_backward else 0) + (3 if concat_rgb else 0)
        self.unnormalize = False
        if self.frames_list is not None:
            self.num_frames = self.num_input_frames - 1

    def get_num_frames(self):
        if self.num_frames is None:
            return len(self.frames_list) - 1 if self.frames_list is not None else None
        return self.num_frames

    def _imagenet(self, x, inverse):
        shape = [1] * x.dim()
        shape[self.c_dim] = 3
        mean = torch.as_tensor(IMAGENET_DEFAULT_MEAN, device=x.device, dtype=x.dtype).view(shape)
        std = torch.as_tensor(IMAGENET_DEFAULT_STD, device=x.device, dtype=x.dtype).view(shape)
        return x * std + mean if inverse else (x - mean) / std      # models/utils.py:15-31

    def get_flow(self, x, **kwargs):
        """preprocessor.py:279-285: the flow network takes [B, T, C, H, W]."""
        if self.t_dim == 2 and self.c_dim == 1:
            return self.flow_model(x.transpose(self.t_dim, self.c_dim), **kwargs).transpose(self.t_dim, self.c_dim)
        return self.flow_model(x, **kwargs)

    def _normalize_flow(self, flow):
        """preprocessor.py:266-277: flow in units of half the image size; the rgb channels are left alone."""
        h, w = flow.shape[-2:]
        size = torch.as_tensor([w, h]).to(flow).view(1, 2, 1, 1, 1)
        if self._concat_backward:
            size = torch.cat([size, size], 1)
        if self._concat_rgb:
            size = torch.cat([size, 2 * torch.ones((1, 3, 1, 1, 1)).to(flow)], 1)
        if self.c_dim == 2:
            size = size.transpose(self.t_dim, self.c_dim)
        return flow / (size / 2.0)

    def forward(self, x, *args, **kwargs):
        if self.flow_model is None:
            raise NotImplementedError(
                "flow-based stream inputs need an optical-flow network: pass main_input_kwargs={'flow_model': "
                "raft.RAFT(...)} or {'flow_model_ckpt': '<raft-large.pth>'} (a plain callable producing the "
                "[B, %d, 1, H, W] stream input is accepted too)" % self.num_channels)
        self.set_input_dims(x)
        x = self.get_input_frames(x)
        if not isinstance(self.flow_model, nn.Module):
            y = self.flow_model(x)
        else:
            if self.unnormalize_rgb:
                x = self._imagenet(x, inverse=True)
            parts = [self.get_flow(x, iters=self.iters, backward=self.backward)]
            if self._concat_backward:
                parts.append(self.get_flow(x, iters=self.iters, backward=(not self.backward)))
            if self._concat_rgb:
                rgb = self._imagenet(x, inverse=False) if self.unnormalize_rgb else x
                idx = torch.tensor(self.frames_list[1:]).long().to(x.device)
                parts.append(torch.index_select(rgb, dim=self.t_dim, index=idx))
            y = torch.cat(parts, self.c_dim)
            if self.normalize_flow:
                y = self._normalize_flow(y)
        self.set_output_dims(y)
        return y


_REGISTRY = {
    'rgb01': partial(Preprocessor, num_channels=3, frames_list=[0, 1]),
    'rgb02': partial(Preprocessor, num_channels=3, frames_list=[0, -1]),
    'rgb0': partial(Preprocessor, num_channels=3, frames_list=[0]),
    'rgb1': partial(Preprocessor, num_channels=3, frames_list=[1]),
    'rgb12': partial(Preprocessor, num_channels=3, frames_list=[1, -1]),
    'rgb012': partial(Preprocessor, num_channels=3, frames_list=[0, 1, -1]),
    'flow01': partial(FramePairFlow, frames_list=[0, 1]),
    'flow_rgb01': partial(FramePairFlow, frames_list=[0, 1], concat_rgb=True),
    'flowback01': partial(FramePairFlow, frames_list=[0, 1], concat_backward=True),
    'flowback_rgb01': partial(FramePairFlow, frames_list=[0, 1], concat_backward=True, concat_rgb=True),
    'imu': IMU,
}


def get_preprocessor(name, temporal_dim=2, unnormalize=True, **kwargs):
    """preprocessor.py:364-387."""
    kwargs = copy.copy(kwargs)
    if 'imu' not in name:
        kwargs['unnormalize'] = unnormalize
    return _REGISTRY[name](temporal_dim=temporal_dim, **kwargs)
