"""Host-side mirror of the counterfactual entry points of ``cwm/models/segmentation.py`` (``FlowGenerator`` :23-547),
SURVEY.md section 8(f): the callers right above the VMAE hot path.

Built so far (rank 1): ``create_motion_counterfactuals`` (:279-343) and ``predict_counterfactual_videos_and_flows``
(:345-432) -- the per-sample Python loop over ``ShiftPatchesAndMask`` becomes one mask kernel plus a *virtual* video
the VMAE forward reads directly, so the S prompts ``x_mocos`` (1.2 MB each) are never written to HBM.  The flow
network is ``raft.RAFT`` (rank 3) or any ``nn.Module`` the caller hands in (the reference's own constructor accepts one,
:70-78), or RAFT loaded from ``flow_model_load_path``; without one the flow entry points raise.

Same method names, argument meaning and error behaviour as the reference.
"""
import copy

import numpy as np
import torch
from torch import nn

from . import sampling
from .masking import RotatedTableEnergyMaskingGenerator, boltzmann
from .perturbation import CounterfactualVideo, shift_patches_and_masks
from .prediction import PredictorBasedGenerator
from .sampling import FlowSampleFilter


def _with_sample_axis(t, num_samples=None):
    """[B, N] -> [B, N, 1] (or expanded to ``num_samples``); a trailing axis of length 1 is expanded likewise."""
    if t.dim() == 2:
        t = t.unsqueeze(-1)
    if num_samples is not None and t.size(-1) == 1 and num_samples != 1:
        t = t.expand(-1, -1, num_samples)
    return t


def _two_frame_movie(x):
    """[C, H, W] / [B, C, H, W] / [B, 1, C, H, W] / [B, T >= 2, C, H, W] -> ([B, 2, C, H, W], input was a still image)."""
    still = x.dim() in (3, 4)
    if x.dim() == 3:
        x = x.unsqueeze(0)
    if x.dim() == 4:
        x = x.unsqueeze(1)
    assert x.dim() == 5, x.shape
    if x.size(1) == 1:
        x = x.expand(-1, 2, -1, -1, -1)
    return x[:, 0:2], still


_DEFAULT_FILTER = object()   # sentinel: "build the default FlowSampleFilter"


class FlowGenerator(PredictorBasedGenerator):
    """A wrapper for masked predictors that builds motion counterfactuals and (with a caller-supplied flow network)
    runs the counterfactual movies through it (segmentation.py:23-41)."""

    default_flow_filter_params = {  # segmentation.py:29-34
        'filter_methods': ['patch_magnitude', 'flow_area', 'num_corners'],
        'flow_magnitude_threshold': 5.0,
        'flow_area_threshold': 0.75,
        'num_corners_threshold': 2
    }

    default_patch_sampling_kwargs = {  # segmentation.py:36-41
        'energy_power': 1,
        'eps': 1e-16,
        'pool_mode': 'mean',
        'resize': False
    }

    def __init__(self, *args, flow_model=None, flow_model_load_path=None, flow_model_kwargs={}, raft_iters=24,
                 flow_sample_filter=_DEFAULT_FILTER,
                 patch_sampling_func=RotatedTableEnergyMaskingGenerator,
                 patch_sampling_kwargs=default_patch_sampling_kwargs, device_masks=False, **kwargs):
        super().__init__(*args, **kwargs)
        # device_masks=True (extension): patches are sampled and masks rectangularised on the device
        # from a counter-based RNG (device_masks.py) -- reproducible for any number of GPUs, but not the reference's
        # host RNG streams, hence opt-in
        self.device_masks = bool(device_masks)
        if self.device_masks:
            from .device_masks import DeterministicRectangularizeMasks, DeviceEnergyMaskingGenerator
            if patch_sampling_func is RotatedTableEnergyMaskingGenerator:
                patch_sampling_func = DeviceEnergyMaskingGenerator
            self.mask_rectangularizer = DeterministicRectangularizeMasks(seed=self.seed)
        self._sweep_counter = 0
        # submodule for sampling patches (segmentation.py:65-69): consumes one draw of self.rng like the reference
        self._patch_sampling_func = patch_sampling_func
        self._patch_sampling_kwargs = copy.deepcopy(self.default_patch_sampling_kwargs)
        self._patch_sampling_kwargs.update(patch_sampling_kwargs)
        self.patch_sampler = None
        self.set_patch_sampler()
        # the reference's default argument is a filter instance and an explicit None means "no filtering"
        # (segmentation.py:48,63)
        self.flow_sample_filter = FlowSampleFilter(**self.default_flow_filter_params) \
            if flow_sample_filter is _DEFAULT_FILTER else flow_sample_filter
        if flow_model is not None or flow_model_load_path is not None:
            self.set_flow_model(flow_model=flow_model, flow_model_load_path=flow_model_load_path, **flow_model_kwargs)
        else:
            self.flow_model = None  # the reference insists on a checkpoint here (:70-75); this mirror defers the error
        self.raft_iters = raft_iters
        self.set_raft_iters(raft_iters)
        self.shifts = None

    # ---- small helpers (segmentation.py:130-140, :247-248) ----
    @staticmethod
    def batch_to_samples(flows, t=0, B=1):
        assert len(flows.shape) == 5, flows.shape
        x = flows[:, t]
        return x.reshape(B, -1, *x.shape[1:]).permute(0, 2, 3, 4, 1)  # '(b s) c h w -> b c h w s'

    def _batch_to_samples(self, flows, t=0):
        assert self.x is not None
        if len(flows.shape) != 5:
            flows = flows.unsqueeze(1)
            t = 0
        return self.batch_to_samples(flows, t=t, B=self.x.size(0))

    def reset_shifts(self):
        self.shifts = []

    def predict_flow(self, vid, backward=False, iters=None, **kwargs):
        """segmentation.py:142-153."""
        if self.flow_model is None:
            raise RuntimeError("FlowGenerator.predict_flow needs a flow_model: pass flow_model=raft.RAFT(...) / any "
                               "nn.Module, or flow_model_load_path=<raft-large.pth>, to the constructor")
        if iters is not None:
            self.set_raft_iters(iters)  # every RAFT inside this generator, like the reference (:147-148)
            if hasattr(self.flow_model, 'iters'):
                self.flow_model.iters = iters
        from .raft import RAFT
        if isinstance(self.flow_model, RAFT) and 'shared_frame' not in kwargs and vid.dim() == 5 and vid.size(0) > 1 \
                and vid.size(1) == 2 and bool((vid[:, 0] == vid[:1, 0]).all()):
            # a counterfactual sweep: every sample keeps the input's first frame (all of frame 0 is visible, so the
            # predictor returns it bit for bit) -> RAFT encodes that frame once instead of once per sample
            kwargs['shared_frame'] = 0
        return self.flow_model(vid, backward=backward, **kwargs).to(vid)

    def set_flow_model(self, flow_model=None, flow_model_load_path=None, **kwargs):
        """segmentation.py:71-84: a given module, or RAFT loaded from a published checkpoint."""
        if flow_model is None:
            from .raft import load_raft_model
            flow_model = load_raft_model(load_path=flow_model_load_path, multiframe=True, scale_inputs=True, **kwargs)
        else:
            assert isinstance(flow_model, nn.Module)
        self.flow_model = flow_model.eval().requires_grad_(False)

    def set_raft_iters(self, iters=None):
        """segmentation.py:86-90."""
        from .raft import RAFT
        for m in self.modules():
            if isinstance(m, RAFT) or type(m).__name__ == 'RAFT':
                m.iters = iters

    def predict_video_and_flow(self, x=None, mask=None, backward=False, propagate_error=False, **kwargs):
        """segmentation.py:170-197: roll the predictor over the movie, then the flow of (frame t, predicted t+1)."""
        x = self.x if x is None else x
        mask = self.mask if mask is None else mask
        num_frames, dt = x.size(1), self.sequence_length
        x_pred = [x[:, 0:1]]
        for t in range(num_frames - dt + 1):
            x_pred.append(self.predict(x[:, t:t + dt], mask, frame=1, **kwargs))
        x_pred = torch.cat(x_pred, 1)
        if propagate_error:
            return x_pred, self.predict_flow(x_pred, backward, **kwargs)
        f_pred = []
        for t in range(num_frames - dt + 1):
            _x = torch.cat([x[:, t:t + 1], x_pred[:, t + 1:t + 2], x[:, t + 2:t + dt]], 1)
            f_pred.append(self.predict_flow(_x, backward, **kwargs))
        return x_pred, torch.cat(f_pred, 1)

    def predict_flow_per_sample(self, x, masks, x_context=None, mask_context=None, timestamps=None, backward=False,
                                **kwargs):
        """segmentation.py:199-208: flows of the S sample predictions as [B, T-1, 2, H, W, S]."""
        S = masks.size(-1)
        x_preds = self.predict_per_sample(x, masks, x_context=x_context, mask_context=mask_context,
                                          timestamps=timestamps, frame=None, split_samples=False)
        flow = self.predict_flow(x_preds, backward, **kwargs)
        p_dims = tuple(range(2, len(flow.shape) + 1))
        return flow.view(-1, S, *flow.shape[1:]).permute(0, *p_dims, 1)

    def predict_video_and_flow_per_sample(self, x, masks, x_context=None, mask_context=None, timestamps=None,
                                          backward=False, **kwargs):
        """segmentation.py:210-245."""
        assert len(masks.shape) == 3
        B, _, S = masks.shape
        tile = lambda z: self.sample_tile(z, S) if (z is not None and z.size(0) != B * S) else z  # noqa: E731
        ys = self.predict_per_sample(x, masks, x_context=tile(x_context), mask_context=tile(mask_context),
                                     timestamps=tile(timestamps), frame=None, split_samples=False, **kwargs)
        flows = self.predict_flow(ys, backward)
        p_dims = tuple(range(2, len(flows.shape) + 1))
        ys = ys.view(-1, S, *ys.shape[1:]).permute(0, *p_dims, 1)
        flows = flows.view(-1, S, *flows.shape[1:]).permute(0, *p_dims, 1)
        return ys, flows

    def compute_flow_samples_magnitude(self, flows, normalize=True, dim=-4, eps=1e-2):
        """segmentation.py:250-255 (the per-sample normalised magnitudes themselves; the mean motion map fuses this
        into ``cwm_flow_magnitude_sum`` instead of materialising it)."""
        flow_mags = flows.square().sum(dim, True).sqrt().to(flows.dtype)
        if normalize:
            flow_mags = flow_mags - flow_mags.amin((-3, -2), True)
            flow_mags = flow_mags / flow_mags.amax((-3, -2), True).clamp(min=eps)
        return flow_mags

    # ---- which patches to move (host-side mask bookkeeping, reference RNG streams) ----
    def set_patch_sampler(self, num_visible=1, mask_ratio=None, **kwargs):
        """segmentation.py:98-116."""
        if (getattr(self, 'patch_sampler', None) is None) or len(kwargs.keys()):
            _kwargs = copy.deepcopy(self._patch_sampling_kwargs)
            _kwargs.update(kwargs)
            try:
                mask_shape = self.mask_shape
            except Exception:
                mask_shape = self.predictor.mask_size
            self.patch_sampler = self._patch_sampling_func(
                input_size=mask_shape, mask_ratio=(mask_ratio or 0), seed=self.rng.randint(9999), always_batch=True,
                **_kwargs)
        if mask_ratio is not None:
            self.patch_sampler.mask_ratio = mask_ratio
        elif num_visible is not None:
            self.patch_sampler.num_visible = num_visible * self.patch_sampler.clumping_factor ** 2

    def sample_patches_from_energy(self, energy=None, num_samples=10, num_visible=1, beta=None, batched=False,
                                   **kwargs):
        """segmentation.py:118-128: bool [B, N, num_samples], False = the sampled (visible / active) patches.

        ``batched=True`` (extension for large sweeps): all ``num_samples`` draws come from ONE multinomial call on the
        energy's device instead of ``num_samples`` sequential sampler calls (0.3 ms each: as long as the whole sweep's
        forward at S = 1024).  Same distribution, same seeding discipline, but a different RNG stream -- the masks are
        not the ones the reference would draw, so it is opt-in."""
        self.set_patch_sampler(num_visible, **kwargs)
        if num_visible == 0:
            return torch.stack([self.get_zeros_mask() for _ in range(num_samples)], -1)
        if energy is None:
            assert self.x is not None
            energy = torch.ones_like(self.x[:, 0, 0:1])
        energy = boltzmann(energy, beta)
        torch.manual_seed(self.rng.randint(99999))
        if self.device_masks and hasattr(self.patch_sampler, 'sample'):
            # one table + one sampling launch for the whole sweep; sample s of sweep k is global sample (k << 20) + s
            masks = self.patch_sampler.sample(energy, num_samples, sample_offset=self._sweep_counter << 20)
            self._sweep_counter += 1
            return masks
        if batched:
            return self._sample_patches_batched(energy, num_samples)
        return torch.stack([self.patch_sampler(energy) for _ in range(num_samples)], -1)

    def _sample_patches_batched(self, energy, num_samples):
        ps = self.patch_sampler
        if ps.randomize_num_visible or ps.temperature is not None:
            raise NotImplementedError("batched patch sampling covers the default sampler settings")
        B, S = energy.shape[0], num_samples
        e = energy.reshape(B, 1, *energy.shape[-2:]).float()
        H, W = e.shape[-2:]
        cf = ps.cf
        gh, gw = ps.height // cf, ps.width // cf
        if (H, W) != (gh, gw):  # pool to the clump grid (sampling.py:63-70)
            e = ps._pool(e, (H // gh, W // gw))
        p = torch.pow(e, ps.energy_power).reshape(B, gh * gw)
        p = p - p.amin(-1, True)                                   # utils.py:160-163 (normalize=True)
        p = torch.relu(p + ps.eps)
        p = p / p.sum(-1, keepdim=True).clamp(min=ps.eps)
        n_points = max((ps.num_patches_per_frame - ps.num_masks_per_frame) // (cf ** 2), 1)
        idx = torch.multinomial(p, S * n_points, replacement=True).view(B, S, n_points)
        vis = torch.zeros(B, S, gh * gw, dtype=torch.bool, device=p.device)
        vis.scatter_(2, idx, True)
        vis = vis.view(B, S, gh, gw)
        if cf > 1:
            vis = vis.repeat_interleave(cf, 2).repeat_interleave(cf, 3)
        mask = (~vis).reshape(B, S, ps.height * ps.width)
        if ps.visible_frames > 0:
            front = torch.zeros(B, S, ps.visible_frames * ps.height * ps.width, dtype=torch.bool, device=p.device)
            mask = torch.cat([front, mask], -1)
        return mask.permute(0, 2, 1).contiguous()

    def sample_counterfactual_motion_map(self, x, active_sampling_distribution=None,
                                         passive_sampling_distribution=None, active_patches=None,
                                         passive_patches=None, num_active_patches=1, num_passive_patches=0,
                                         num_samples=8, sample_batch_size=8, patch_sampling_kwargs={}, do_filter=True,
                                         **kwargs):
        """segmentation.py:434-476: sample the patches, predict the counterfactual movies and their flows, filter."""
        self.set_input(x)

        def _sample_patches(dist, num_visible):
            return self.sample_patches_from_energy(energy=dist, num_samples=num_samples, num_visible=num_visible,
                                                   **patch_sampling_kwargs)

        if active_patches is None:
            active_patches = _sample_patches(active_sampling_distribution, num_active_patches)
        if passive_patches is None:
            passive_patches = _sample_patches(passive_sampling_distribution, num_passive_patches)
        ys, flows = self.predict_counterfactual_videos_and_flows(
            x, active_patches=active_patches, passive_patches=passive_patches, num_samples=num_samples,
            sample_batch_size=sample_batch_size, fix_passive=True, **kwargs)
        flows = self._batch_to_samples(flows)
        if (self.flow_sample_filter is not None) and do_filter:
            flows, filter_mask = self.flow_sample_filter(flows, active_patches)
        return (flows, active_patches, passive_patches)

    def set_flow_sample_filter(self, params=None):
        """segmentation.py:92-96."""
        self.flow_sample_filter = None if params is None else FlowSampleFilter(**params)

    # ---- flow-derived statistics ----
    def compute_mean_motion_map(self, flows, normalize_per_sample=False, normalize=True, dim=-4, eps=1e-2,
                                group=None, num_samples_total=None):
        """segmentation.py:257-276: mean over the samples of the flow magnitude, normalised to [0, 1] per image.
        ``flows`` [B, 2, H, W, S] (any strides), or an already computed distribution [B, 1, H, W] that is only
        normalised.  Extension for sharded sweeps: with ``group`` (a torch.distributed process group) every rank passes
        its local samples and the partial sums are combined with ONE all-reduce; ``num_samples_total`` is the S of
        the whole sweep."""
        if len(flows.shape) == 5:
            if dim not in (-4, 1):
                raise NotImplementedError("the flow channel axis must be dim 1 of [B, 2, H, W, S]")
            sums = sampling.flow_magnitude_sum(flows, normalize_per_sample=normalize_per_sample, eps=eps)
            count = flows.shape[-1]
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
                count = num_samples_total if num_samples_total is not None else count * dist.get_world_size(group)
        else:  # just normalize the input distribution
            assert len(flows.shape) == 4 and flows.shape[1] == 1, flows.shape
            sums, count, normalize = flows[:, 0].float(), 1, True
        return sampling.motion_map_finalize(sums, count, normalize=normalize, eps=eps)

    @staticmethod
    def compute_flow_corrs(flow_samples, flow_samples_swap=None, downsample=1, take_top_k=None, do_spearman=False,
                           distance_func=None, thresh=None, use_covariance=False, eps=1e-12, binarize=False,
                           normalize=False, zscore=False, range_thresh=None):
        """segmentation.py:478-547: covariance / correlation of the flow magnitude between image locations over the
        samples.  The options the reference's only caller uses (interface.py:27-29, :466-493: ``downsample``,
        ``use_covariance``, ``take_top_k``) run on the hand-written kernels; the other options take the torch-op route of
        ``_flow_corrs_general``."""
        custom_distance = distance_func is not None and not FlowGenerator._is_channel_rms(distance_func)
        if flow_samples_swap is not None or do_spearman or custom_distance or thresh is not None or \
                binarize or normalize or zscore or range_thresh is not None:
            # the options the reference's only caller never sets: the same feature transforms as torch ops on the
            # device tensors, then torch.cov / torch.corrcoef (library calls; no hand-written kernel for these)
            if flow_samples.device.type != "cuda":
                raise RuntimeError("compute_flow_corrs: tensors must live on a CUDA (B200) device; there is no CPU fallback")
            return FlowGenerator._flow_corrs_general(
                flow_samples, flow_samples_swap=flow_samples_swap, downsample=downsample, take_top_k=take_top_k,
                do_spearman=do_spearman, distance_func=distance_func, thresh=thresh, use_covariance=use_covariance,
                eps=eps, binarize=binarize, normalize=normalize, zscore=zscore, range_thresh=range_thresh)
        B, C, H, W, S = flow_samples.shape
        if S == 0:  # segmentation.py:494-497
            flow_samples = torch.zeros(list(flow_samples.shape)[:-1] + [1], device=flow_samples.device).float()
        if take_top_k is not None:
            flow_samples = flow_samples[..., :take_top_k]
        return sampling.flow_corrs(flow_samples, downsample=downsample, use_covariance=use_covariance)

    @staticmethod
    def _is_channel_rms(distance_func):
        """True only for the reference's default ``ChannelMSE(dim=1)`` (utils.py:510-521): the channel-RMS the kernels
        hard-code.  Any other class, or the same class reducing another dim, takes the general route."""
        return type(distance_func).__name__ == "ChannelMSE" and getattr(distance_func, 'dim', 1) in (1, -4)

    @staticmethod
    def _flow_corrs_general(flow_samples, flow_samples_swap=None, downsample=1, take_top_k=None, do_spearman=False,
                            distance_func=None, thresh=None, use_covariance=False, eps=1e-12, binarize=False,
                            normalize=False, zscore=False, range_thresh=None):
        """segmentation.py:478-547 with every option, as device-agnostic torch ops (pinned against the live reference in
        tests/test_flowstats.py)."""
        import torch.nn.functional as F
        B, C, H, W, S = flow_samples.shape
        if S == 0:
            flow_samples = torch.zeros(list(flow_samples.shape)[:-1] + [1]).to(flow_samples.device).float()
            S = 1
        if flow_samples_swap is not None:
            assert list(flow_samples_swap.shape) == [B, C, H, W, S]
        K = S if take_top_k is None else take_top_k
        ds = downsample

        def _ds(fs):
            return F.avg_pool3d(fs[..., :K].permute(0, 1, 4, 2, 3), (1, ds, ds), stride=(1, ds, ds)).permute(0, 1, 3, 4, 2)

        flow_inp = _ds(flow_samples)
        if flow_samples_swap is not None:
            flow_inp = torch.cat([flow_inp, _ds(flow_samples_swap)], -1)
        if distance_func is None or FlowGenerator._is_channel_rms(distance_func):
            flow_inp = torch.sqrt(flow_inp.square().mean(1, True).float())      # ChannelMSE(dim=1), utils.py:510-521
        else:
            flow_inp = distance_func(flow_inp, torch.zeros_like(flow_inp))
        flow_inp = flow_inp.reshape(B, -1, flow_inp.size(-1))
        flow_corrs = []
        for b in range(B):
            f = torch.argsort(flow_inp[b], -1).float() if do_spearman else flow_inp[b]
            if (thresh is not None) and (binarize is False):
                f = f * (f > thresh).float()
            elif thresh is not None:
                f = (f > thresh).float()
            elif range_thresh is not None:
                f = f - f.amin(0, True)
                f = (f > (range_thresh * f.amax(0, True))).float()
            if normalize:
                f = f / f.amax(0, True).clamp(min=eps)
            if zscore:
                mn, std = f.mean(0), f.std(0).clamp(min=eps)
                f = (f - mn[None]) / std[None]
            c = torch.cov(f) if use_covariance else torch.corrcoef(f)
            c[torch.isnan(c)] = 0
            flow_corrs.append(c)
        return torch.stack(flow_corrs, 0).view(B, 1, H // ds, W // ds, H // ds, W // ds)

    def filter_flow_samples(self, flows, active_patches, do_filter=True):
        """The tail of ``sample_counterfactual_motion_map`` (segmentation.py:470-476): flows [(b s), T, 2, H, W] or
        [(b s), 2, H, W] from the flow network -> [B, 2, H, W, S] view with the rejected samples zeroed."""
        flows = self._batch_to_samples(flows)
        if (self.flow_sample_filter is not None) and do_filter:
            flows, _ = self.flow_sample_filter(flows, active_patches)
        return flows

    # ---- SURVEY 8(f) rank 1 ----
    def create_motion_counterfactuals(self, x, masks, active_patches=None, shifts=None, frame=1, num_samples=None,
                                      fix_passive=True, reset_shifts=False, virtual=False):
        """Create motion counterfactuals by applying shifts to active_patches and no shifts to masks
        (segmentation.py:279-343).  Returns ``(x_shift [B*S, T, C, H, W], mask_shift [B*S, N])`` after the mask
        rectangulariser.  ``virtual=True`` (extension) returns the prompts as a ``CounterfactualVideo``."""
        if reset_shifts or getattr(self, 'shifts', None) is None:
            self.reset_shifts()
        # every patch tensor ends up [B, N, S]: 2-D inputs are one pattern shared by all samples (:289-304)
        if masks.dim() == 2:
            assert num_samples is not None, "Choose how many samples to shift with arg num_samples"
            masks = _with_sample_axis(masks, num_samples)
        B, N, S = masks.shape
        num_samples = S
        if active_patches is None:
            active_patches = torch.ones_like(masks)
        else:
            active_patches = _with_sample_axis(active_patches, S)
            assert active_patches.size(-1) in (1, S)
            active_patches = _with_sample_axis(active_patches, S)

        # `make_static_movie(x[:,0:1], T=2)` + `sample_tile(x, S)` (:309-312) are index arithmetic here: every
        # frame of every sample reads frame 0 of image i // S
        if fix_passive and x.size(1) != 2:
            x = x[:, 0:1].expand(-1, 2, -1, -1, -1)
        static_frame = 0 if fix_passive else -1
        masks_bs = masks.transpose(1, 2).reshape(B * S, N)          # 'b n s -> (b s) n' (:313-314)
        active_bs = active_patches.transpose(1, 2).reshape(B * S, N)

        # set the shifts (:317-319)
        self.shifter.set_shapes(x, mask=masks_bs)
        self.shifter.set_num_shifts(S)
        shifts = self.shifter._preprocess_shifts_sequence(shifts, is_mask_shift=True)
        BS = B * S
        if len(shifts) < BS:
            # the reference indexes `shifts[i]` for i < B*S (:329): IndexError for B > 1 unless B*S shifts are given
            raise IndexError("list index out of range")
        shifts = [[int(s[0]), int(s[1])] for s in shifts[:BS]]
        sample_image = torch.arange(BS, dtype=torch.int32, device=x.device) // S
        video, mask_shift = shift_patches_and_masks(x, masks_bs, active_bs, shifts, self.patch_size, frame=frame,
                                                    static_frame=static_frame, sample_image=sample_image)
        for s in shifts:  # `self.shifts.append(np.array(self.shift))` per sample (:335-338)
            px = (s[0] * self.patch_size[-2], s[1] * self.patch_size[-1])
            self.shifter.shift = px
            self.shift = [px[0] // self.patch_size[-2], px[1] // self.patch_size[-1]]
            self.shifts.append(np.array(self.shift))
        mask_shift = self.mask_rectangularizer(mask_shift)
        return ((video if virtual else video.materialize()), mask_shift)

    def predict_counterfactual_videos(self, x, active_patches, passive_patches=None, shifts=None, num_samples=8,
                                      sample_batch_size=8, fix_passive=True, max_shift_fraction=None, frame=1,
                                      **kwargs):
        """The video half of ``predict_counterfactual_videos_and_flows`` (segmentation.py:345-430): the predicted
        counterfactual movies ``y_mocos`` [B*S, T, C, H, W]."""
        # whatever comes in becomes a 2-frame movie (:364-373); a single image means the passive patches stay put
        x, was_image = _two_frame_movie(x)
        fix_passive = fix_passive or was_image
        self.set_input(x)
        self.reset_shifts()

        # patches and shifts are brought to one common number of samples (:379-410)
        active_patches = _with_sample_axis(active_patches)
        passive_patches = (self.get_zeros_mask().unsqueeze(-1) if passive_patches is None
                           else _with_sample_axis(passive_patches))
        S = max(active_patches.size(-1), passive_patches.size(-1))
        if S == 1 and num_samples > 1:
            S = num_samples
        self.shifter.set_shapes(x, mask=active_patches[..., 0])
        if shifts is not None:
            self.shifter.set_num_shifts(shifts.shape[-1] if hasattr(shifts, 'shape') else len(shifts))
        else:
            self.shifter.set_num_shifts(S)
            if max_shift_fraction is not None:
                self.shifter.max_shift_fraction = max_shift_fraction
        shifts = self.shifter._preprocess_shifts_sequence(shifts, is_mask_shift=True)
        num_samples = len(shifts)
        if num_samples > 1:
            active_patches, passive_patches = (t.expand(-1, -1, num_samples) if t.size(-1) == 1 else t
                                               for t in (active_patches, passive_patches))
        assert active_patches.size(-1) == passive_patches.size(-1) == num_samples, \
            (active_patches.shape, passive_patches.shape, num_samples)

        x_mocos, masks_mocos = self.create_motion_counterfactuals(
            x, masks=passive_patches, active_patches=active_patches, shifts=shifts, num_samples=num_samples,
            fix_passive=fix_passive, reset_shifts=False, frame=frame, virtual=True)
        # batch predict (:421-430)
        return self.batch_predict_per_sample(x_mocos, masks=masks_mocos, frame=None,
                                             batch_size=(sample_batch_size or x_mocos.size(0)), sample_dim=0,
                                             **kwargs)

    def predict_counterfactual_videos_and_flows(self, x, active_patches, passive_patches=None, shifts=None,
                                                num_samples=8, sample_batch_size=8, fix_passive=True,
                                                max_shift_fraction=None, frame=1, raft_iters=None, backward=False,
                                                **kwargs):
        """segmentation.py:345-432."""
        y_mocos = self.predict_counterfactual_videos(
            x, active_patches, passive_patches=passive_patches, shifts=shifts, num_samples=num_samples,
            sample_batch_size=sample_batch_size, fix_passive=fix_passive, max_shift_fraction=max_shift_fraction,
            frame=frame, **kwargs)
        flow_mocos = self.predict_flow(y_mocos, backward=backward, iters=raft_iters)
        return (y_mocos, flow_mocos)


__all__ = ["FlowGenerator", "CounterfactualVideo"]
