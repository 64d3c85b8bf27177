"""Host-side mirror of the reference's predictor wrapper for the hot path
(``cwm/models/prediction.py``: ``PredictorBasedGenerator._preprocess / predict / pred_patches_to_video /
predict_per_sample / batch_predict_per_sample``; ``cwm/models/masking.py:90-132`` ``RectangularizeMasks``).

Same names, argument meaning and error behaviour as the reference for this path; the arithmetic
(normalise, gather, VMAE forward, scatter + unpatchify) runs in libcwm_b200.  Patch perturbations (the counterfactual prompts) live in
``perturbation.py`` / ``segmentation.py``; mask *generation*, RAFT flow and the statistics built on top are out of
scope ("next").
"""
import numpy as np
import torch
from torch import nn

import ctypes

from . import _lib, perturbation
from .perturbation import CounterfactualVideo
from .conjoined_vmae import ConjoinedPretrainVisionTransformer, PaddedVisionTransformer
from .vmae import (IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, PretrainVisionTransformer, compact_mask)


class RectangularizeMasks(nn.Module):
    """Make sure all masks in a batch have the same number of 1s and 0s (cwm/models/masking.py:90-132).
    Host-side integer bookkeeping; mutates ``masks`` in place like the reference and draws from the global torch
    RNG (``torch.randperm``) only for rows that actually need changes."""

    def __init__(self, truncation_mode='min'):
        super().__init__()
        self._mode = truncation_mode
        assert self._mode in ['min', 'max', 'mean', 'full', 'none', None], (self._mode)

    def set_mode(self, mode):
        self._mode = mode

    def __call__(self, masks):
        if self._mode in ['none', None]:
            return masks
        assert isinstance(masks, torch.Tensor), type(masks)
        if self._mode == 'full':
            return torch.ones_like(masks)
        shape = masks.shape
        masks = masks.flatten(1)
        B, N = masks.shape
        num_masked = masks.float().sum(-1)
        M = {'min': torch.amin, 'max': torch.amax, 'mean': torch.mean}[self._mode](num_masked).long()
        num_changes = (num_masked.long() - M).cpu()
        if not bool(num_changes.any()):
            return masks.view(*shape) if list(masks.shape) != list(shape) else masks
        for b in range(B):
            nc = int(num_changes[b])
            if nc > 0:
                inds = torch.where(masks[b])[0]
                inds = inds[torch.randperm(inds.size(0))[:nc].to(inds.device)]
                masks[b, inds] = 0
            elif nc < 0:
                inds = torch.where(~masks[b])[0]
                inds = inds[torch.randperm(inds.size(0))[:-nc].to(inds.device)]
                masks[b, inds] = 1
        if list(masks.shape) != list(shape):
            masks = masks.view(*shape)
        return masks


def unpatchify_scatter(y, x_raw, inv_perm, n_vis, patch_size):
    """a12: ``pred_patches_to_video`` (prediction.py:245-259): predictions at masked patches, the raw input at
    visible patches (bit-exact copy), laid out back as a video [B, T, C, H, W]."""
    lib = _lib.load()
    B, T, C, H, W = x_raw.shape
    pt, ph, pw = patch_size
    if isinstance(x_raw, CounterfactualVideo):
        # fused path: visible patches are read from the virtual counterfactual video
        out = torch.empty(B, T, C, H, W, dtype=torch.float32, device=x_raw.device)
        if y is not None:
            y = y.contiguous().float()
        src, keep = x_raw.c_struct()
        stream = torch.cuda.current_stream(x_raw.device).cuda_stream
        _lib.check(lib.cwm_unpatchify_scatter_cf(y.data_ptr() if y is not None else None, ctypes.byref(src),
                                                 inv_perm.data_ptr(), B, T, C, H, W, pt, ph, pw, int(n_vis),
                                                 out.data_ptr(), stream))
        del keep
        return out
    if x_raw.dtype != torch.float32:
        x_raw = x_raw.float()
    out = torch.empty(B, T, C, H, W, dtype=torch.float32, device=x_raw.device)
    if y is not None:
        y = y.contiguous()
        if y.dtype != torch.float32:
            y = y.float()
    stream = torch.cuda.current_stream(x_raw.device).cuda_stream
    _lib.check(lib.cwm_unpatchify_scatter(y.data_ptr() if y is not None else None, x_raw.data_ptr(),
                                          _lib.strides5(x_raw), inv_perm.data_ptr(), B, T, C, H, W, pt, ph, pw,
                                          int(n_vis), out.data_ptr(), stream))
    return out


class PredictorBasedGenerator(nn.Module):
    """The slice of cwm/models/prediction.py:16-540 that brackets the predictor call."""

    def __init__(self, predictor=None, imagenet_normalize_inputs=False, temporal_dim=2, seed=0,
                 mask_generator=None, max_shift_fraction=0.15, error_func=nn.MSELoss(reduction='none'),
                 keypoint_predictor=None, **kwargs):
        super().__init__()
        self.error_func = error_func
        self.keypoint_predictor = keypoint_predictor  # prediction.py:66-71 (any module: video -> [B, 1, 1, H, W] logits)
        if predictor is None:
            raise ValueError("There is no predictor set for this generator and no model to load to")
        self.predictor = predictor
        self.imagenet_normalize_inputs = imagenet_normalize_inputs
        self.set_temporal_dim(temporal_dim)
        self.rng = np.random.RandomState(seed=seed)
        self.torch_rng = torch.manual_seed(seed)  # prediction.py:44-45: the reference seeds the global generator
        self.seed = seed
        self.mask_generator = mask_generator
        self.mask_rectangularizer = RectangularizeMasks('min')
        # submodules of prediction.py:51-58 (the multi-patch shifter is not on the counterfactual path)
        self.make_static = perturbation.MakeStatic(patch_size=self.predictor.patch_size)
        self.shifter = perturbation.ShiftPatchesAndMask(
            patch_size=self.predictor.patch_size, padding_mode='constant', max_shift_fraction=max_shift_fraction,
            allow_fractional_shifts=False)
        self.x = self.mask = self.inp_shape = None

    # ---- attributes the callers read (prediction.py:131-214) ----
    @property
    def patch_size(self):
        if hasattr(self.predictor, 'patch_size'):
            return self.predictor.patch_size
        return self.predictor.encoder.patch_embed.proj.kernel_size

    @property
    def image_size(self):
        return self.predictor.image_size

    @property
    def sequence_length(self):
        return getattr(self.predictor, 'num_frames', 2)

    @property
    def mask_shape(self):
        pt, ph, pw = self.patch_size
        h, w = self.image_size[-2:] if self.inp_shape is None else self.inp_shape[-2:]
        return (self.sequence_length // pt, h // ph, w // pw)

    def set_temporal_dim(self, t_dim=1):
        if t_dim not in (1, 2):
            raise ValueError("temporal_dim must be 1 or 2")
        self.predictor.t_dim, self.predictor.c_dim = t_dim, 3 - t_dim       # time and channel axes are 1 and 2, either way

    @property
    def t_dim(self):
        return self.predictor.t_dim

    @property
    def c_dim(self):
        return self.predictor.c_dim

    def set_image_size(self, *args, **kwargs):
        if hasattr(self.predictor, 'set_image_size'):
            self.predictor.set_image_size(*args, **kwargs)
        else:
            self.predictor.image_size = args[0]

    def get_zeros_mask(self, x=None, frame=-1):
        """All-visible mask [B, N] with (optionally) one fully masked token frame."""
        x = self.x if x is None else x
        self.inp_shape = x.shape
        mask = torch.zeros(self.mask_shape, device=x.device, dtype=torch.bool)
        if frame is not None:
            mask[frame] = True
        return mask.reshape(1, -1).expand(x.shape[0], -1)

    def generate_mask(self, x=None):
        assert self.mask_generator is not None
        if x is None:
            x = self.x
        mask = self.mask_generator(x).view(x.size(0), -1).to(x.device)
        return self.mask_rectangularizer(mask)

    def reset_padding_masks(self):
        """prediction.py:121-129."""
        p = self.predictor
        streams = [p.main_stream, p.context_stream] if hasattr(p, 'main_stream') else [p]    # conjoined: per stream
        for stream in streams:
            if hasattr(stream, 'padding_mask'):
                stream._reset_padding_mask()

    # ---- a1 ----
    def _preprocess(self, x):
        """prediction.py:304-312 (kept for callers that want the normalised tensor; ``predict`` itself fuses the
        transpose + normalisation into the patch gather)."""
        if self.t_dim != 1:
            x = x.transpose(self.t_dim, self.c_dim)
        if not self.imagenet_normalize_inputs:
            return x
        shape = [1] * 5
        shape[1 if self.t_dim == 2 else 2] = 3          # the channel axis after the transpose above
        mean = torch.as_tensor(IMAGENET_DEFAULT_MEAN).to(x).view(shape)
        std = torch.as_tensor(IMAGENET_DEFAULT_STD).to(x).view(shape)
        return (x - mean) / std

    # ---- a12 ----
    def pred_patches_to_video(self, y, x, mask):
        """input at visible positions, preds at masked positions (prediction.py:245-259)."""
        mask = mask.reshape(x.shape[0], -1)
        _, inv, nvis = compact_mask(mask.to(x.device))
        counts = nvis.cpu()
        if not bool((counts == counts[0]).all()):
            raise RuntimeError("shape mismatch: rows of the mask have different numbers of masked tokens")
        return unpatchify_scatter(y, x, inv, int(counts[0]), self.patch_size)

    # ---- the call into the predictor (prediction.py:406-454) ----
    @torch.no_grad()
    def predict(self, x=None, mask=None, frame=-1, reset_masks=True, *args, **kwargs):
        if x is None:
            x = self.x
        if mask is None:
            mask = self.generate_mask(x)
        # extension used by HostPipeline: the caller already rectangularised the mask on the host and knows the
        # per-row visible count, so neither step needs a device->host read here
        num_visible = kwargs.pop('_num_visible', None)
        self.inp_shape = x.shape
        self.set_image_size(x.shape[-2:])
        plain_vmae = isinstance(self.predictor, PretrainVisionTransformer) and not isinstance(
            self.predictor, (ConjoinedPretrainVisionTransformer, PaddedVisionTransformer))
        if num_visible is None and plain_vmae and x.size(0) > 1 and mask.device.type == "cuda" and \
                self.mask_rectangularizer._mode == 'min':
            # One device->host read serves both the rectangulariser and the forward: the compaction kernel counts
            # the visible tokens of every row; when they agree (every sweep whose prompts are well formed) the
            # rectangulariser is the identity and draws nothing from the RNG (masking.py:117-128).
            compaction = compact_mask(mask.reshape(x.size(0), -1))
            counts = compaction[2].cpu()
            if bool((counts == counts[0]).all()):
                num_visible = int(counts[0])
                kwargs['compaction'] = compaction   # the forward reuses it: one compaction launch per call
        if num_visible is None:
            mask = mask if (x.size(0) == 1) else self.mask_rectangularizer(mask)
        elif isinstance(self.predictor, PretrainVisionTransformer) and \
                not isinstance(self.predictor, PaddedVisionTransformer):
            kwargs['num_visible'] = num_visible
        if isinstance(x, CounterfactualVideo) and not (plain_vmae and self.t_dim == 2):
            x = x.materialize()  # other predictors take the materialised prompts (still one kernel, not a loop)
        if isinstance(self.predictor, (ConjoinedPretrainVisionTransformer, PaddedVisionTransformer)):
            y = self._predict_padded_or_conjoined(x, mask, *args, **kwargs)
        elif isinstance(x, CounterfactualVideo):
            # fused: the gather and the unpatchify read the virtual video; x_shift never touches HBM
            norm = (IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD) if self.imagenet_normalize_inputs else None
            y = self.predictor(x, mask, *args, input_norm=norm, **kwargs)
            _, inv, n_vis = self.predictor.last_aux
            y = unpatchify_scatter(y, x, inv, n_vis, self.patch_size)
        elif isinstance(self.predictor, PretrainVisionTransformer):
            xin = x.transpose(self.t_dim, self.c_dim) if self.t_dim != 1 else x  # a view, never materialised
            norm = (IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD) if self.imagenet_normalize_inputs else None
            y = self.predictor(xin, mask, *args, input_norm=norm, **kwargs)
            _, inv, n_vis = self.predictor.last_aux
            y = unpatchify_scatter(y, x, inv, n_vis, self.patch_size)
        else:
            y = self.predictor(self._preprocess(x), mask, *args, **kwargs)
            if len(y.shape) != 5:
                y = self.pred_patches_to_video(y, x, mask=mask)
        if frame is not None:
            frame = frame % y.size(1)
            y = y[:, frame:frame + 1]
        if reset_masks:
            self.reset_padding_masks()
        return y

    def _predict_padded_or_conjoined(self, x, mask, *args, **kwargs):
        """prediction.py:412-446 for padded / conjoined predictors: the rows of the (max - min) padding positions are
        stripped (:424-432) and the main-stream frames + mask go to `pred_patches_to_video` (:435-446).  The imagenet
        normalisation is fused into the main stream's patch gather and the visible patches are copied from the raw
        input (the reference round-trips them through normalise -> unnormalise, a <= 6e-8 difference)."""
        P = self.predictor
        is_padded = hasattr(P, 'padding_mask') and bool((mask.sum(-1).amax() != mask.sum(-1).amin()).item())
        if is_padded:
            print("Warning: passed a batch of images with different numbers of visible tokens.")
        xin = x.transpose(self.t_dim, self.c_dim) if self.t_dim != 1 else x
        norm = (IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD) if self.imagenet_normalize_inputs else None
        y = P(xin, mask, *args, input_norm=norm, **kwargs)
        conjoined = hasattr(P, 'main_stream')
        if hasattr(P, 'padding_mask'):
            ms = P.main_stream if conjoined else P
            num_pad = ms.max_padding_tokens - ms.min_padding_tokens
            y = y[:, :-num_pad]
        if len(y.shape) == 5:
            return y
        if conjoined:
            (_x, _mask, _), _ = P.get_stream_inputs(xin, mask.reshape(x.shape[0], -1), *args,
                                                    **{k: v for k, v in kwargs.items()
                                                       if k in ('timestamps', 'x_context', 'mask_context')})
            if self.t_dim == 2:
                _x = _x.transpose(1, 2)
            patch_size = P.main_stream.patch_size
        else:
            _x, _mask, patch_size = x, mask, P.patch_size
        _, inv, nvis = compact_mask(_mask.reshape(_x.shape[0], -1).to(_x.device))
        counts = nvis.cpu()
        if not bool((counts == counts[0]).all()):
            raise RuntimeError("shape mismatch: rows of the mask have different numbers of masked tokens")
        return unpatchify_scatter(y, _x, inv, int(counts[0]), patch_size)

    # ---- checkpoint loading and the `interface.py` entry points around `predict` ----
    def load_predictor(self, load_path=None, model=None, map_location='cpu'):
        """prediction.py:81-107: load a checkpoint (``{'model': state_dict}`` or a bare state_dict) into the predictor."""
        if (getattr(self, 'predictor', None) is None) and (model is None):
            raise ValueError("There is no predictor set for this generator and no model to load to")
        if load_path is None:
            return
        weights = torch.load(load_path, map_location=torch.device(map_location))
        if 'model' in weights.keys():
            weights = weights['model']
        target = self.predictor if model is None else model
        print(target.load_state_dict(weights), load_path)
        if model is None:
            self._predictor_load_path = load_path

    def forward(self, x, mask=None, frame=None, *args, **kwargs):
        return self.predict(x, mask, frame, *args, **kwargs)

    def predict_error(self, x=None, mask=None, target=None, frame=None, dim=-3):
        """Factual prediction error map (prediction.py:331-343, `interface.py`'s entry point): `error_func` between
        `predict(x, mask)` and the target (default: the input itself), summed over ``dim`` (channels)."""
        x = self.x if x is None else x
        mask = self.generate_mask(x) if mask is None else mask
        prediction = self.predict(x, mask, frame=frame)
        target = x if target is None else target
        if frame is not None:
            target = target[:, frame].unsqueeze(1)
        return self.error_func(prediction, target).sum(dim, True)

    # ---- counterfactual prompts ----
    def set_input(self, x, mask=None, make_mask=False, timestamps=None):
        """Remembers the movie the following calls work on (prediction.py:703-724): a single frame ``[B, C, H, W]``
        becomes a one-frame movie; ``mask`` / ``timestamps`` are stored when given, ``make_mask`` draws a fresh mask."""
        if x.dim() == 4:
            x = x.unsqueeze(1)
        elif x.dim() != 5:
            raise AssertionError("Input must be a movie of shape [B,T,C,H,W] or a single frame of shape [B,C,H,W]")
        self.x, self.inp_shape = x, x.shape
        self.B, self.T, self.C = (int(v) for v in x.shape[:3])
        if mask is not None:
            self.mask = mask
        elif make_mask:
            if self.mask_generator is None:
                raise AssertionError("You need to have a mask generator to set a new mask")
            self.mask = self.generate_mask(x)
        if timestamps is not None:
            self.timestamps = timestamps

    def make_static_movie(self, x=None, T=None, frame=0):
        """``T`` copies of one frame of ``x`` as a ``[B, T, C, H, W]`` movie (prediction.py:731-740)."""
        x = self.x if x is None else x
        T = getattr(self.predictor, 'num_frames', 2) if T is None else T
        if x.dim() == 4:
            x = x.unsqueeze(1)
        if x.dim() != 5:
            raise AssertionError("x must be of shape [B,C,H,W] or [B,T,C,H,W], but is %s" % (tuple(x.shape),))
        return x[:, frame % x.size(1)].unsqueeze(1).repeat(1, T, 1, 1, 1)

    def _shift(self, x, mask, active_patches=None, shift=None, frame=1, virtual=False):
        """prediction.py:756-779: ``shift`` is a mask shift (patch units)."""
        if getattr(self, 'shifts', None) is None:
            self.shifts = []
        if active_patches is None:
            active_patches = torch.ones_like(mask)
        points = ~active_patches
        x_shift, mask_shift = self.shifter(x, mask=torch.minimum(mask, active_patches), mask_shift=shift,
                                           perturbation_points=points, frame=frame, virtual=virtual)
        mask_shift = self.mask_rectangularizer(mask_shift)
        self.shift = self.shifter.shift
        self.shift = [self.shift[0] // self.patch_size[-2], self.shift[1] // self.patch_size[-1]]
        self.shifts.append(np.array(self.shift))
        return (x_shift, mask_shift)

    def get_counterfactual_prediction(self, x, mask=None, active_patches=None, shift=None, fix_passive=False,
                                      **kwargs):
        """prediction.py:781-813."""
        if len(x.shape) == 4:  # make into a 2-frame movie
            x = x[:, None]
        elif len(x.shape) == 3:
            x = x[None, None]
        if x.size(1) == 1:
            x = self.make_static_movie(x, T=2)
        if mask is None:
            mask = self.get_zeros_mask(x)
        if active_patches is None:
            active_patches = self.get_zeros_mask(x)
        if fix_passive:
            x, _ = self.make_static(x, mask)
        x_p, mask_p = self._shift(x, mask=mask, active_patches=active_patches, shift=shift, frame=1, virtual=True)
        return self.predict(x_p, mask_p, frame=None, **kwargs)

    def predict_per_sample(self, x, masks, frame=-1, batch_size=None, split_samples=True, *args, **kwargs):
        """Run predictions in parallel for S sample masks (prediction.py:456-482)."""
        assert len(masks.shape) == 3, masks.shape
        S = masks.size(-1)
        if x is None:
            x = self.x
        B = x.size(0)
        BS = B * S
        x = x[:, None].expand(-1, S, -1, -1, -1, -1).reshape(BS, *x.shape[1:])
        masks = masks.transpose(1, 2).reshape(BS, -1)
        y = self.predict(x=x, mask=masks, frame=frame, *args, **kwargs)
        if not split_samples:
            return y
        p_dims = tuple(range(2, len(y.shape) + 1))
        return y.view(B, S, *y.shape[1:]).permute(0, *p_dims, 1)

    def batch_predict_per_sample(self, x, masks, frame=-1, batch_size=None, sample_dim=None, **kwargs):
        """prediction.py:497-540 for ``sample_dim=0`` (the layout every caller on the path uses,
        segmentation.py:423-430): ``x`` [S, T, C, H, W] and ``masks`` [S, N], processed in chunks."""
        if sample_dim != 0:
            raise NotImplementedError("only sample_dim=0 is used by the counterfactual path")
        S = masks.size(0)
        if batch_size is None:
            batch_size = S
        batch_size = max(1, batch_size)
        ys = []
        for b0 in range(0, S, batch_size):
            xb = x[b0:b0 + batch_size]
            ys.append(self.predict(xb, mask=masks[b0:b0 + batch_size], frame=frame, reset_masks=True,
                                   **self.sample_tile_all_tensors(xb.size(0), **kwargs)))
            self.reset_padding_masks()
        return torch.cat(ys, 0)

    def sample_tile(self, z, num_samples):
        """prediction.py:484-487."""
        S = num_samples
        rank = len(z.shape)
        return z[:, None].expand(-1, S, *([-1] * (rank - 1))).reshape(-1, *z.shape[1:])

    def sample_tile_all_tensors(self, num_samples, **kwargs):
        """prediction.py:489-495: per-image tensors (e.g. the IMU context) are repeated for every sample of a chunk."""
        return {kw: self.sample_tile(val, num_samples) if isinstance(val, torch.Tensor) else val
                for kw, val in kwargs.items()}


class HostPipeline:
    """Host-buffer front end of ``PredictorBasedGenerator.predict`` for sweeps that do not fit (or do not start) in
    device memory: batches live in pinned host memory, the predicted videos go back to pinned host memory, and the
    host->device copy of batch i+1 and the device->host copy of batch i-1 overlap the kernels of batch i (three CUDA
    streams, two device slots, events for ordering).  Numerically identical to calling ``predict`` per batch.

        pipe = HostPipeline(G, batch_shape=(64, 2, 3, 224, 224), n_tokens=1568)
        for x_host, mask_host, out_host in batches:
            pipe.submit(x_host, mask_host, out_host, frame=None)
        pipe.finish()            # all outputs have landed in their out_host buffers
    """

    def __init__(self, generator, batch_shape, n_tokens, device=None, post=None, **predict_kwargs):
        self.G = generator
        self.device = torch.device(device) if device is not None else next(generator.predictor.parameters()).device
        self.kwargs = predict_kwargs
        self.post = post  # optional callable(video) run on the compute stream right after predict (e.g. a gather)
        with torch.cuda.device(self.device):
            self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()
            self.x = [torch.empty(batch_shape, dtype=torch.float32, device=self.device) for _ in range(2)]
            self.m = [torch.empty(batch_shape[0], n_tokens, dtype=torch.bool, device=self.device) for _ in range(2)]
            self.loaded = [torch.cuda.Event() for _ in range(2)]
            self.computed = [torch.cuda.Event() for _ in range(2)]
            self.stored = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self.h2d_bytes = self.d2h_bytes = 0

    def submit(self, x_host, mask_host, out_host, frame=None):
        # host-side integer bookkeeping (what `predict` would do on the device with a sync): rectangularise in place
        # like the reference (masking.py:100-132) and read the per-row visible count
        if x_host.size(0) > 1:
            mask_host = self.G.mask_rectangularizer(mask_host)
        counts = (~mask_host.reshape(mask_host.size(0), -1)).sum(-1)
        if not bool((counts == counts[0]).all()):
            raise RuntimeError("rows of the mask have different numbers of visible tokens")
        slot = self.i & 1
        compute = torch.cuda.current_stream(self.device)
        if self.i >= 2:
            self.h2d.wait_event(self.computed[slot])   # the slot's previous batch has been consumed
        with torch.cuda.stream(self.h2d):
            self.x[slot].copy_(x_host, non_blocking=True)
            self.m[slot].copy_(mask_host, non_blocking=True)
            self.loaded[slot].record(self.h2d)
        compute.wait_event(self.loaded[slot])
        video = self.G.predict(self.x[slot], self.m[slot], frame=frame, _num_visible=int(counts[0]), **self.kwargs)
        if self.post is not None:
            self.post(video)
        self.computed[slot].record(compute)
        self.d2h.wait_event(self.computed[slot])
        with torch.cuda.stream(self.d2h):
            out_host.copy_(video, non_blocking=True)
            self.stored[slot].record(self.d2h)
        video.record_stream(self.d2h)
        self.h2d_bytes = x_host.numel() * x_host.element_size() + mask_host.numel() * mask_host.element_size()
        self.d2h_bytes = out_host.numel() * out_host.element_size()
        self.i += 1

    def finish(self):
        """Makes the current stream wait for every outstanding copy (no host synchronisation)."""
        compute = torch.cuda.current_stream(self.device)
        for ev in self.stored[:min(self.i, 2)]:
            compute.wait_event(ev)
