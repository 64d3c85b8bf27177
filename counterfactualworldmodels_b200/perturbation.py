"""Host-side mirror of the reference's patch perturbations on the counterfactual path
(``cwm/models/perturbation.py``: ``PatchPerturbation`` :13-112, ``MakeStatic`` :120-150, ``ShiftPatchesAndMask``
:152-289), SURVEY.md section 8(f) rank 1.

Same class names, constructor arguments, attributes (``shift``, ``num_shifts``, ``rng`` ...) and error behaviour; the
arithmetic runs in libcwm_b200 (``csrc/counterfactual.cu``): masks of all samples in one launch, pixels either
materialised by one HBM-write-bound kernel or -- ``CounterfactualVideo`` -- never materialised at all: the VMAE patch
gather and the final unpatchify read the *virtual* video directly (``cwm_vmae_forward_cf``).

Masks are bit-exact with the reference; so are the pixels (the blends are evaluated literally in fp32).
There is no CPU fallback.
"""
import ctypes

import numpy as np
import torch
from torch import nn

from . import _lib


def _as_u8(mask):
    m = mask.contiguous()
    if m.dtype == torch.bool:
        return m.view(torch.uint8)
    if m.dtype != torch.uint8:
        return (m != 0).view(torch.uint8)
    return m


def _require_cuda(t, who):
    if t.device.type != "cuda":
        raise RuntimeError(f"{who}: tensors must live on a CUDA (B200) device; there is no CPU fallback")


class CounterfactualVideo:
    """The S motion-counterfactual prompts of a sweep as a *virtual* float32 video [S, T, C, H, W]:

        v[i, t]      = x[sample_image[i], t']                               t != frame
        v[i, frame]  = shift(x[.., frame'], shift_px[i]) * (1 - m) + x[.., frame'] * m,  m = shifted_active[i] per patch

    with t' = ``static_frame`` when the input is made static (``make_static_movie``, prediction.py:731-740).  Behaves
    like a tensor where the predictor wrappers need it (``shape``, ``size``, ``device``, slicing of the sample axis);
    ``materialize()`` writes the videos out (= ``x_shift`` of segmentation.py:339)."""

    def __init__(self, x, sample_image, shift_px, shifted_active, frame, static_frame, patch_size):
        assert x.dim() == 5, x.shape
        self.x = x if x.dtype == torch.float32 else x.float()
        self.sample_image = sample_image      # int32 [S]
        self.shift_px = shift_px              # int32 [S, 2]
        self.shifted_active = shifted_active  # uint8 [S, h*w]
        self.frame = int(frame)
        self.static_frame = int(static_frame)
        self.patch_size = tuple(int(p) for p in patch_size)

    # ---- the slice of the tensor interface the wrappers use ----
    @property
    def shape(self):
        return torch.Size((self.sample_image.shape[0],) + tuple(self.x.shape[1:]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 5

    @property
    def device(self):
        return self.x.device

    @property
    def dtype(self):
        return torch.float32

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        if not isinstance(idx, slice):
            raise TypeError("CounterfactualVideo supports slicing of the sample axis only; call materialize() first")
        return CounterfactualVideo(self.x, self.sample_image[idx], self.shift_px[idx], self.shifted_active[idx],
                                   self.frame, self.static_frame, self.patch_size)

    def c_struct(self):
        """-> (cwm_cf_source, keep-alive list)."""
        s = _lib.CfSource()
        sample_image = self.sample_image.contiguous()
        shift_px = self.shift_px.contiguous()
        shifted_active = self.shifted_active.contiguous()
        s.x = self.x.data_ptr()
        for k, v in enumerate(self.x.stride()):
            s.xs[k] = v
        s.sample_image = sample_image.data_ptr()
        s.shift_px = shift_px.data_ptr()
        s.shifted_active = shifted_active.data_ptr()
        s.frame, s.static_frame = self.frame, self.static_frame
        return s, (sample_image, shift_px, shifted_active, self.x)

    def materialize(self):
        lib = _lib.load()
        S, T, C, H, W = self.shape
        out = torch.empty(S, T, C, H, W, dtype=torch.float32, device=self.device)
        src, keep = self.c_struct()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(lib.cwm_cf_build_videos(ctypes.byref(src), S, T, C, H, W, self.patch_size[-2],
                                               self.patch_size[-1], out.data_ptr(), stream))
        del keep
        return out


def shift_patches_and_masks(x, passive, active, mask_shifts, patch_size, frame=1, static_frame=-1, sample_image=None):
    """All S samples of ``FlowGenerator.create_motion_counterfactuals``' loop (segmentation.py:321-338) at once.

    x [B_img, T, C, H, W] raw frames; passive / active bool [S, N] (True = masked / not active); mask_shifts S pairs
    (dy, dx) in patch units; ``sample_image`` [S] (default: all samples read image 0).
    Returns (CounterfactualVideo, mask [S, N] bool) -- the mask is *not* rectangularised."""
    lib = _lib.load()
    _require_cuda(x, "shift_patches_and_masks")
    pt, ph, pw = patch_size
    if pt != 1:
        raise NotImplementedError("motion counterfactuals need a temporal patch size of 1")
    B, T, C, H, W = x.shape
    h, w = H // ph, W // pw
    S, N = passive.shape
    assert active.shape == passive.shape, (active.shape, passive.shape)
    assert N == T * h * w, f"mask has {N} tokens but the video has {T}x{h}x{w} patches"
    frame = frame % T
    dev = x.device
    ms = np.asarray(mask_shifts, dtype=np.int64).reshape(-1, 2)
    assert ms.shape[0] == S, (ms.shape, S)
    # `shift = mask_shift * patch` (perturbation.py:254-256); the mask padding divides it back with the axes
    # crossed (:234-238) -- identical for square patches
    shift_px = np.stack([ms[:, 0] * ph, ms[:, 1] * pw], 1)
    mshift = np.stack([shift_px[:, 0] // pw, shift_px[:, 1] // ph], 1)
    packed = torch.from_numpy(np.concatenate([shift_px, mshift], 1).astype(np.int32)).to(dev, non_blocking=True)
    shift_px_d, mshift_d = packed[:, :2].contiguous(), packed[:, 2:].contiguous()
    if sample_image is None:
        sample_image = torch.zeros(S, dtype=torch.int32, device=dev)
    p8, a8 = _as_u8(passive.to(dev)), _as_u8(active.to(dev))
    shifted_active = torch.empty(S, h * w, dtype=torch.uint8, device=dev)
    mask_out = torch.empty(S, N, dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.cwm_cf_shift_masks(p8.data_ptr(), a8.data_ptr(), mshift_d.data_ptr(), S, T, h, w, frame,
                                          shifted_active.data_ptr(), mask_out.data_ptr(), stream))
    video = CounterfactualVideo(x, sample_image.to(torch.int32), shift_px_d, shifted_active, frame, static_frame,
                                patch_size)
    return video, mask_out


class PatchPerturbation(nn.Module):
    """cwm/models/perturbation.py:13-112."""

    def __init__(self, patch_size, seed=0, frame=None, use_image_coordinates=True, **kwargs):
        super().__init__()
        self.patch_size = tuple(patch_size)
        self.frame = frame
        self.seed = seed
        self.rng = np.random.RandomState(seed=seed)
        self.use_image_coordinates = use_image_coordinates

    @property
    def T(self):
        return self.sequence_length

    @property
    def C(self):
        return self.num_channels

    @property
    def H(self):
        return self.image_size[0]

    @property
    def W(self):
        return self.image_size[1]

    @property
    def mask_shape(self):
        return (self.sequence_length // self.patch_size[0], self.image_size[0] // self.patch_size[1],
                self.image_size[1] // self.patch_size[2])

    @property
    def mask_image_size(self):
        return self.mask_shape[-2:]

    def _check_shapes(self, x, mask):
        if mask is not None:
            self.inp_mask_shape = mask.shape

    def set_shapes(self, x, mask):
        assert len(x.shape) == 5, x.shape
        self.inp_shape = x.shape
        self.B = self.inp_shape[0]
        self.image_size = self.inp_shape[-2:]
        self.sequence_length = self.inp_shape[1]
        self.num_channels = self.inp_shape[2]
        self.num_patches = np.prod(self.mask_shape)
        self._check_shapes(x, mask)

    def reshape_mask_to_video(self, mask):
        mask = mask.view(self.B, -1, *self.mask_image_size)
        self.T_mask = mask.size(1)
        return mask

    def sample_random_patch(self, batch_size, frames=[0]):
        patch_idx_list = []
        if frames is None:
            frames = list(range(self.T))
        elif not isinstance(frames, (list, tuple)):
            frames = [frames]
        for b_idx in range(batch_size):
            t_idx = self.rng.choice(frames)
            h_idx = self.rng.randint(self.mask_image_size[0])
            w_idx = self.rng.randint(self.mask_image_size[1])
            patch_idx_list.append([b_idx, t_idx, h_idx, w_idx])
        return patch_idx_list

    def image_to_patch_inds(self, inds):
        return [inds[-i] // self.patch_size[-i] for i in range(1, len(inds) + 1)]

    def perturb(self, x, mask, **kwargs):
        raise NotImplementedError("Do the perturbation")

    def forward(self, x, mask=None, perturbation_points=None, **kwargs):
        self.set_shapes(x, mask)
        mask = mask.clone()
        if perturbation_points is None:
            perturbation_mask = mask
        else:  # remove the visible patches in common between mask and perturbation mask
            mask[perturbation_points] = 1
            perturbation_mask = torch.logical_not(perturbation_points)
        x_perturbed, mask_perturbed = self.perturb(x, perturbation_mask, **kwargs)
        if perturbation_points is not None:
            mask_perturbed = torch.minimum(mask, mask_perturbed)
        return x_perturbed, mask_perturbed


class NullPerturbation(PatchPerturbation):

    def perturb(self, x, mask, **kwargs):
        return (x, mask)


class MakeStatic(PatchPerturbation):
    """Make the visible patches in frames t > 0 identical to the spatially equivalent patches in frame t = 0
    (perturbation.py:120-150)."""

    def _check_shapes(self, x, mask):
        assert self.T > 1

    def perturb(self, x, mask):
        lib = _lib.load()
        _require_cuda(x, "MakeStatic")
        m = self.reshape_mask_to_video(mask.to(x.device))
        if self.T_mask != self.T:  # assume all other frames are masked, so they won't be altered (:137-141)
            T_vis = self.T - self.T_mask
            m_vis = torch.ones((self.B, T_vis, *self.mask_image_size), dtype=m.dtype, device=m.device)
            m = torch.cat([m_vis, m[:, -1:]], 1)
        if x.dtype != torch.float32:
            x = x.float()
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.cwm_cf_make_static(x.data_ptr(), _lib.strides5(x), _as_u8(m).data_ptr(), self.B, self.T,
                                              self.C, self.H, self.W, self.patch_size[-2], self.patch_size[-1],
                                              out.data_ptr(), stream))
        return (out, mask)  # original mask is unaltered


class ShiftPatchesAndMask(PatchPerturbation):
    """Shift the visible patches and mask in a target frame by some 2D vector (perturbation.py:152-289)."""

    def __init__(self, patch_size, max_shift_fraction=0.15, padding_mode='constant', allow_fractional_shifts=False,
                 **kwargs):
        super().__init__(patch_size, **kwargs)
        if padding_mode != 'constant':
            raise NotImplementedError("only padding_mode='constant' (the one prediction.py:53-58 uses)")
        if allow_fractional_shifts:
            raise NotImplementedError("allow_fractional_shifts=True belongs to MultiShiftPatchesAndMask")
        self.max_shift_fraction = max_shift_fraction
        self.padding_mode = padding_mode
        self.allow_fractional_shifts = allow_fractional_shifts
        self.set_num_shifts()

    def set_num_shifts(self, num_shifts=None):
        self._num_shifts = 1 if num_shifts is None else num_shifts

    @property
    def num_shifts(self):
        if getattr(self, '_num_shifts', None) is None:
            self.set_num_shifts()
        return self._num_shifts

    def _check_shapes(self, x, mask):
        self.inp_mask_shape = mask.shape

    def _preprocess_shifts_sequence(self, shifts_sequence, is_mask_shift=False):
        """perturbation.py:184-216."""
        if shifts_sequence is None:
            return [self.get_random_shift(is_mask_shift) for _ in range(self.num_shifts)]
        if hasattr(shifts_sequence, 'shape'):
            assert len(shifts_sequence.shape) == 2, shifts_sequence.shape
            D, S = shifts_sequence.shape
            assert D == 2, D
            assert S in (self.num_shifts, 1), (S, self.num_shifts)
            if isinstance(shifts_sequence, torch.Tensor):
                shifts_sequence = [shifts_sequence[..., s].detach().cpu().numpy() for s in range(S)]
            else:
                shifts_sequence = [shifts_sequence[..., s] for s in range(S)]
        if isinstance(shifts_sequence, (list, tuple)):
            if not isinstance(shifts_sequence[0], (list, tuple)):
                shifts_sequence = [shifts_sequence]
            assert all((len(s) == 2 for s in shifts_sequence))
            if len(shifts_sequence) == 1:  # all have same shift
                return shifts_sequence * self.num_shifts
            else:
                assert len(shifts_sequence) == self.num_shifts, (len(shifts_sequence), self.num_shifts)
        return shifts_sequence

    def get_random_shift(self, is_mask_shift=False):
        """perturbation.py:218-234: same draws from the same ``np.random.RandomState(seed)``."""
        def rect(s, p):
            q = 1 if is_mask_shift else p
            return int(s // p) * q

        max_shift = [int(self.max_shift_fraction * s) for s in self.image_size]
        random_shift = (0, 0)
        while sum(random_shift) == 0:
            random_shift = (
                rect(self.rng.randint(-max_shift[0], max_shift[0] + 1), self.patch_size[-2]),
                rect(self.rng.randint(-max_shift[1], max_shift[1] + 1), self.patch_size[-1]))
        return random_shift

    def perturb(self, x, mask, shift=None, mask_shift=None, frame=-1, passive=None, virtual=False):
        """perturbation.py:245-289.  ``mask`` is the perturbation mask (False = patch to move).  Extensions:
        ``passive`` (the mask `minimum`-ed in by ``forward``) and ``virtual`` (return a CounterfactualVideo)."""
        frame = (frame % self.T)
        if shift is not None:
            assert len(shift) == 2, shift
            assert (shift[0] % self.patch_size[-2]) == 0, shift
            assert (shift[1] % self.patch_size[-1]) == 0, shift
        elif mask_shift is not None:
            assert len(mask_shift) == 2, mask_shift
            shift = (mask_shift[0] * self.patch_size[-2], mask_shift[1] * self.patch_size[-1])
        else:
            shift = self.get_random_shift()
        self.shift = shift
        mask = self.reshape_mask_to_video(mask).reshape(self.B, -1)
        if self.T_mask != self.T:
            raise RuntimeError(f"shape '{list(self.inp_mask_shape)}' is invalid: the mask must cover all {self.T} "
                               "frames (the reference fails at perturbation.py:287)")
        if passive is None:
            passive = torch.ones_like(mask)
        ms = [[int(shift[0]) // self.patch_size[-2], int(shift[1]) // self.patch_size[-1]]] * self.B
        video, mask_shift_out = shift_patches_and_masks(
            x, passive, mask, ms, self.patch_size, frame=frame, static_frame=-1,
            sample_image=torch.arange(self.B, dtype=torch.int32, device=x.device))
        mask_shift_out = mask_shift_out.view(*self.inp_mask_shape)
        return (video if virtual else video.materialize(), mask_shift_out)

    def forward(self, x, mask=None, perturbation_points=None, **kwargs):
        """``PatchPerturbation.forward`` (perturbation.py:99-112) with the final ``minimum`` fused into the mask
        kernel: (mask | points) & shifted(~points)."""
        self.set_shapes(x, mask)
        if perturbation_points is None:
            return self.perturb(x, mask, **kwargs)
        return self.perturb(x, torch.logical_not(perturbation_points), passive=mask, **kwargs)


class ShiftPatches(ShiftPatchesAndMask):
    """Only shift the patches (perturbation.py:291-...): not on the counterfactual path."""

    def perturb(self, *args, **kwargs):
        raise NotImplementedError("ShiftPatches is not used by the counterfactual path")
