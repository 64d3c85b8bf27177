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


class _Geometry:
    """Shape record of one ``[B, T, C, H, W]`` video batch and its patch grid (what the reference keeps as a dozen
    attributes and properties on ``PatchPerturbation``, perturbation.py:13-72)."""
    __slots__ = ("batch", "frames", "channels", "height", "width", "grid")

    def __init__(self, video_shape, patch_size):
        if len(video_shape) != 5:
            raise AssertionError(tuple(video_shape))
        self.batch, self.frames, self.channels, self.height, self.width = (int(v) for v in video_shape)
        pt, ph, pw = patch_size
        self.grid = (self.frames // pt, self.height // ph, self.width // pw)

    def frame_masks(self, mask):
        """[B, N] -> [B, frames_in_mask, h, w] (the mask may cover fewer frames than the video)."""
        return mask.reshape(self.batch, -1, self.grid[1], self.grid[2])


class _Perturbation(nn.Module):
    """State shared by the two perturbations the counterfactual path uses: patch size, the seeded host RNG the
    reference draws random shifts from, and the geometry of the last input."""

    def __init__(self, patch_size, seed=0, **unused):
        super().__init__()
        self.patch_size = tuple(patch_size)
        self.seed = seed
        self.rng = np.random.RandomState(seed=seed)
        self.geometry = None
        self.inp_mask_shape = None

    def set_shapes(self, x, mask):
        self.geometry = _Geometry(x.shape, self.patch_size)
        if mask is not None:
            self.inp_mask_shape = mask.shape

    # the names other modules read
    @property
    def image_size(self):
        return (self.geometry.height, self.geometry.width)

    @property
    def mask_shape(self):
        return self.geometry.grid


class MakeStatic(_Perturbation):
    """Copies frame 0 into every visible patch of the later frames (cwm/models/perturbation.py:120-150): one launch of
    ``cwm_cf_make_static`` on the strided input, no Patchify round trip.  Returns ``(video, mask)``, the mask as given."""

    def forward(self, x, mask=None, **unused):
        lib = _lib.load()
        _require_cuda(x, "MakeStatic")
        self.set_shapes(x, mask)
        g = self.geometry
        if g.frames < 2:
            raise AssertionError("MakeStatic needs at least two frames")
        per_frame = g.frame_masks(mask.to(x.device))
        if per_frame.shape[1] != g.frames:
            # a mask that only covers the last frame: the earlier frames count as masked, i.e. are left alone
            lead = torch.ones((g.batch, g.frames - per_frame.shape[1]) + g.grid[1:], dtype=per_frame.dtype,
                              device=per_frame.device)
            per_frame = torch.cat([lead, per_frame[:, -1:]], 1)
        src = x if x.dtype == torch.float32 else x.float()
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
        with torch.cuda.device(src.device):
            _lib.check(lib.cwm_cf_make_static(src.data_ptr(), _lib.strides5(src), _as_u8(per_frame).data_ptr(),
                                              g.batch, g.frames, g.channels, g.height, g.width, self.patch_size[-2],
                                              self.patch_size[-1], out.data_ptr(),
                                              torch.cuda.current_stream(src.device).cuda_stream))
        return out, mask


class ShiftPatchesAndMask(_Perturbation):
    """Moves the patches selected by a mask, and the mask with them, by a 2-D vector inside one frame
    (cwm/models/perturbation.py:152-289), for a whole batch in one mask kernel + one (virtual) video.

    Shifts are ``(dy, dx)``; ``mask_shift`` counts patches, ``shift`` pixels (multiples of the patch size).
    Only what ``prediction.py:53-58`` configures is supported: constant padding, whole-patch shifts."""

    def __init__(self, patch_size, max_shift_fraction=0.15, padding_mode='constant', allow_fractional_shifts=False,
                 **kwargs):
        super().__init__(patch_size, **kwargs)
        if padding_mode != 'constant':
            raise NotImplementedError("only padding_mode='constant' (the one prediction.py:53-58 uses)")
        if allow_fractional_shifts:
            raise NotImplementedError("allow_fractional_shifts=True belongs to MultiShiftPatchesAndMask")
        self.max_shift_fraction = max_shift_fraction
        self.num_shifts = 1
        self.shift = None

    def set_num_shifts(self, num_shifts=None):
        self.num_shifts = num_shifts or 1

    def get_random_shift(self, is_mask_shift=False):
        """One non-zero-sum shift drawn like perturbation.py:218-234 (two ``randint`` per attempt from the seeded
        stream, floored to whole patches): pixel units, or patch units with ``is_mask_shift``."""
        limits = [int(self.max_shift_fraction * extent) for extent in self.image_size]
        patch = self.patch_size[-2:]
        while True:
            draw = [self.rng.randint(-lim, lim + 1) for lim in limits]
            whole = [int(d // p) for d, p in zip(draw, patch)]
            out = tuple(whole) if is_mask_shift else tuple(n * p for n, p in zip(whole, patch))
            if sum(out) != 0:
                return out

    def _preprocess_shifts_sequence(self, shifts_sequence, is_mask_shift=False):
        """One shift per sample from whatever the caller passed (perturbation.py:184-216): ``None`` -> random draws;
        one pair -> repeated; a list of ``num_shifts`` pairs -> itself; a ``[2, S]`` array -> its columns, which then
        go through the list rules (so, exactly like the reference, only S == 2 survives the pair check)."""
        count = self.num_shifts
        if shifts_sequence is None:
            return [self.get_random_shift(is_mask_shift) for _ in range(count)]
        seq = shifts_sequence
        if hasattr(seq, 'shape'):
            if len(seq.shape) != 2 or seq.shape[0] != 2 or seq.shape[1] not in (count, 1):
                raise AssertionError((tuple(seq.shape), count))
            cols = seq.detach().cpu().numpy() if isinstance(seq, torch.Tensor) else np.asarray(seq)
            seq = [cols[:, j] for j in range(cols.shape[1])]
        if not isinstance(seq, (list, tuple)):
            return seq
        pairs = seq if isinstance(seq[0], (list, tuple)) else [seq]
        if any(len(pair) != 2 for pair in pairs):
            raise AssertionError("every shift must be a (dy, dx) pair")
        if len(pairs) == 1:
            return list(pairs) * count
        if len(pairs) != count:
            raise AssertionError((len(pairs), count))
        return pairs

    def _resolve_shift(self, shift, mask_shift):
        ph, pw = self.patch_size[-2:]
        if shift is not None:
            if len(shift) != 2 or shift[0] % ph or shift[1] % pw:
                raise AssertionError(f"pixel shift {tuple(shift)} is not a pair of multiples of the patch size")
            return (shift[0], shift[1])
        if mask_shift is not None:
            if len(mask_shift) != 2:
                raise AssertionError(mask_shift)
            return (mask_shift[0] * ph, mask_shift[1] * pw)
        return self.get_random_shift()

    def forward(self, x, mask=None, perturbation_points=None, shift=None, mask_shift=None, frame=-1, virtual=False):
        """``(x_shifted, mask_shifted)`` for the batch.  With ``perturbation_points`` (True = patch to move) the
        result mask is ``(mask | points) & shifted(~points)`` -- the reference's clone / index-assign / ``minimum``
        sequence (perturbation.py:99-112) folded into the mask kernel; without them ``mask`` itself (False = move)
        selects the patches.  ``virtual=True`` returns the prompts as a ``CounterfactualVideo``."""
        self.set_shapes(x, mask)
        g = self.geometry
        moving = mask if perturbation_points is None else torch.logical_not(perturbation_points)
        if g.frame_masks(moving).shape[1] != g.frames:
            raise RuntimeError(f"shape '{list(self.inp_mask_shape)}' is invalid: the mask must cover all {g.frames} "
                               "frames (the reference fails at perturbation.py:287)")
        moving = moving.reshape(g.batch, -1)
        held = torch.ones_like(moving) if perturbation_points is None else mask.reshape(g.batch, -1)
        self.shift = self._resolve_shift(shift, mask_shift)
        ph, pw = self.patch_size[-2:]
        per_sample = [[int(self.shift[0]) // ph, int(self.shift[1]) // pw]] * g.batch
        video, shifted = shift_patches_and_masks(
            x, held, moving, per_sample, self.patch_size, frame=frame % g.frames, static_frame=-1,
            sample_image=torch.arange(g.batch, dtype=torch.int32, device=x.device))
        return (video if virtual else video.materialize()), shifted.view(*self.inp_mask_shape)
