"""Host-side mirror of the mask generators the counterfactual path draws its inputs from
(``cwm/models/masking.py``: ``upsample_masks`` :10-30, ``MaskingGenerator`` :267-401,
``RotatedTableUniformMaskingGenerator`` :478-545; ``cwm/models/sampling.py``: ``EnergySamplingMaskingGenerator`` :11-112,
``RotatedTableEnergyMaskingGenerator`` :114-126; ``cwm/models/utils.py``: ``boltzmann`` :91-95,
``sample_image_inds_from_probs`` :152-170, ``sample_from_energy`` :172-213) -- SURVEY.md section 8(f) rank 4.

Masks are *inputs* of the hot path: small integer bookkeeping driven by the reference's host RNG streams
(``np.random.RandomState(seed)``, the global torch generator).  A mask is only "the same mask" if it consumes those
streams in the same order, so this module is plain host code that makes exactly the reference's draws -- with the same
seeds it returns bit-identical masks (pinned by ``tests/golden/masks_*.npz``).  Nothing here touches pixels or tokens.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.distributions.categorical import Categorical

from .prediction import RectangularizeMasks  # noqa: F401  (re-exported: masking.RectangularizeMasks in the reference)


def upsample_masks(masks, size, thresh=0.5):
    """Nearest-neighbour up/down-sampling of patch masks by integer factors (masking.py:10-30)."""
    h, w = masks.shape[-2:]
    H, W = size
    if (H, W) == (h, w):
        return masks
    if H < h and W < w:
        return masks[..., ::h // H, ::w // W]
    if H % h or W % w:
        raise NotImplementedError("non-integer mask upsampling (bilinear resize + threshold) is not on the path")
    return masks.repeat_interleave(H // h, dim=-2).repeat_interleave(W // w, dim=-1)


class MaskingGenerator(nn.Module):
    """Uniformly random visible patches (clumps), per frame (masking.py:267-401)."""

    def __init__(self, input_size, mask_ratio, seed=0, visible_frames=0, clumping_factor=1,
                 randomize_num_visible=False, create_on_cpu=True, always_batch=False):
        super().__init__()
        self.frames = None
        if isinstance(input_size, int):
            self.height = self.width = input_size
        elif len(input_size) == 3:
            self.frames, self.height, self.width = input_size
        elif len(input_size) == 2:
            self.height, self.width = input_size
        else:
            self.height = self.width = input_size[0]
        self.clumping_factor = clumping_factor
        ch, cw = self.c
        self.pad_h, self.pad_w = self.height % ch, self.width % cw
        self.num_patches_per_frame = (self.height // ch) * (self.width // cw)
        self.mask_ratio = mask_ratio
        self.visible_frames = visible_frames
        self.always_batch = always_batch
        self.create_on_cpu = create_on_cpu
        self.rng = np.random.RandomState(seed=seed)
        self._set_torch_seed(seed)
        self.randomize_num_visible = randomize_num_visible

    # -- the coupled (ratio, count, visible) properties of the reference (:300-331) --
    @property
    def c(self):
        cf = self.clumping_factor
        return (cf, cf) if isinstance(cf, int) else tuple(cf[:2])

    @property
    def mask_ratio(self):
        return self._mask_ratio

    @mask_ratio.setter
    def mask_ratio(self, val):
        self._mask_ratio = val
        self._num_masks_per_frame = int(val * self.num_patches_per_frame)

    @property
    def num_masks_per_frame(self):
        if not hasattr(self, '_num_masks_per_frame'):
            self._num_masks_per_frame = int(self.mask_ratio * self.num_patches_per_frame)
        return self._num_masks_per_frame

    @num_masks_per_frame.setter
    def num_masks_per_frame(self, val):
        self._num_masks_per_frame = val
        self._mask_ratio = val / self.num_patches_per_frame

    @property
    def num_visible(self):
        return self.num_patches_per_frame - self.num_masks_per_frame

    @num_visible.setter
    def num_visible(self, val):
        self.num_masks_per_frame = self.num_patches_per_frame - val

    def _set_torch_seed(self, seed):
        self.seed = seed
        torch.manual_seed(self.seed)  # the reference seeds the GLOBAL generator (:333-335)

    def sample_mask_per_frame(self, *args, **kwargs):
        """One frame's mask [height*width] or [height, width] (:347-376).  RNG order: (randint), randperm,
        (choice, choice)."""
        n, k = self.num_patches_per_frame, self.num_masks_per_frame
        if self.randomize_num_visible:
            k = self.rng.randint(low=k, high=n + 1)
        ordered = torch.arange(n) >= (n - k)          # n-k visible slots first, then k masked
        mask = ordered[torch.randperm(n).long()]
        ch, cw = self.c
        if max(ch, cw) > 1:
            grid = mask.view(self.height // ch, self.width // cw)
            grid = grid.repeat_interleave(ch, 0).repeat_interleave(cw, 1)
            off_h = self.rng.choice(range(self.pad_h + 1))
            off_w = self.rng.choice(range(self.pad_w + 1))
            mask = F.pad(grid, (self.pad_w - off_w, off_w, self.pad_h - off_h, off_h), mode='constant', value=1)
            mask = mask.reshape(self.height, self.width)
        return mask

    def _stack_frames(self, num_frames):
        return torch.cat([self.sample_mask_per_frame() for _ in range(num_frames)], 0).flatten()

    def forward(self, x=None, num_frames=None):
        num_frames = (num_frames or self.frames) or 1
        if isinstance(x, torch.Tensor):
            batch_size = x.size(0)
            masks = torch.stack([self._stack_frames(num_frames) for _ in range(batch_size)], 0)
            if not self.create_on_cpu:
                masks = masks.to(x.device)
            if batch_size == 1 and not self.always_batch:
                masks = masks.squeeze(0)
        else:
            batch_size = 1
            masks = self._stack_frames(num_frames)
            if self.always_batch:
                masks = masks[None]
        if self.visible_frames > 0:
            vis = torch.zeros((batch_size, self.height * self.width), dtype=torch.bool).view(
                *masks.shape[:-1], -1).to(masks.device)
            masks = torch.cat(([vis] * self.visible_frames) + [masks], -1)
        return masks


class RotatedTableUniformMaskingGenerator(MaskingGenerator):
    """Leading frames fully visible, the last frame(s) masked at ``mask_ratio`` (masking.py:478-545) -- the
    temporally-factored mask of every factual prediction (README.md:21, ipynb cell 12)."""

    def __init__(self, input_size, mask_ratio, visible_frames=None, context_mask_ratio=None, seed=0,
                 clumping_factor=1, always_batch=True, randomize_num_visible=False, full_mask_prob=0):
        assert len(input_size) == 3, input_size
        if visible_frames is None:
            visible_frames = input_size[0] - 1
        super().__init__(input_size=(input_size[0] - visible_frames, *input_size[1:]), mask_ratio=mask_ratio,
                         visible_frames=visible_frames, seed=seed, clumping_factor=clumping_factor,
                         always_batch=always_batch, randomize_num_visible=randomize_num_visible)
        self.visible_frames = visible_frames
        self.full_mask_prob = full_mask_prob
        if context_mask_ratio is not None:
            self.context_mask_ratio = context_mask_ratio
            self.vis_frame_sampler = MaskingGenerator(
                input_size=(1, self.height, self.width), mask_ratio=context_mask_ratio, visible_frames=0,
                clumping_factor=1, create_on_cpu=self.create_on_cpu, always_batch=self.always_batch)
        else:
            self.context_mask_ratio = 0
            self.vis_frame_sampler = None

    def forward(self, x=None, *args, **kwargs):
        masks = super().forward(x=x, *args, **kwargs)
        n_frame = self.height * self.width
        if self.full_mask_prob > 0:
            n_vis = n_frame * self.visible_frames
            drop = (torch.rand((masks.size(0), 1)).to(masks.device) < self.full_mask_prob)
            full = torch.cat([torch.zeros(masks.size(0), n_vis, dtype=torch.bool, device=masks.device),
                              drop.expand(-1, masks.size(-1) - n_vis)], -1)
            masks = torch.maximum(masks, full)
        if self.vis_frame_sampler is not None:
            context = torch.cat([self.vis_frame_sampler(x) for _ in range(self.visible_frames)], -1)
            tail = masks.view(masks.size(0), self.frames, -1)[:, self.visible_frames:, :]
            masks = torch.cat([context, tail.reshape(masks.size(0), -1)], -1)
        return masks


# ---- energy-based sampling (cwm/models/utils.py:91-213, cwm/models/sampling.py:11-126) --------------------------

def boltzmann(x, beta=1, eps=1e-9):
    if beta is None:
        return x
    x = torch.exp(x * beta)
    return x / x.amax((-1, -2), keepdim=True).clamp(min=eps)


def sample_image_inds_from_probs(probs, num_points, eps=1e-9, normalize=False, seed=0):
    """[B, H, W] non-negative energies -> [B, P, 2] (row, col) indices drawn with replacement (utils.py:152-170)."""
    B, H, W = probs.shape
    p = probs.reshape(B, H * W)
    if normalize:
        p = p - p.amin(-1, True)
    p = F.relu(p + eps)
    p = p / p.to(p.dtype).sum(dim=-1, keepdim=True).clamp(min=eps)
    idx = Categorical(probs=p).sample([num_points]).permute(1, 0).to(torch.long)
    rows = torch.div(idx, W, rounding_mode='floor').clamp(0, H - 1)
    cols = torch.fmod(idx, W).clamp(0, W - 1)
    return torch.stack([rows, cols], dim=-1)


def sample_from_energy(probs, num_points=1, num_samples=1, binarize=False, normalize=False, eps=1e-9):
    """utils.py:172-213: an image per sample that is non-zero exactly at the drawn points."""
    shape = probs.shape
    if len(shape) == 5:
        B, T, _, H, W = shape
    elif len(shape) == 4:
        B, _, H, W = shape
        T = 1
        probs = probs[:, None]
    else:
        raise ValueError(probs.shape)
    assert probs.size(-3) == 1, probs.shape
    S, P = num_samples, num_points
    flat = probs.unsqueeze(1).expand(-1, S, -1, -1, -1, -1).reshape(B * S * T, H, W)
    inds = sample_image_inds_from_probs(flat, P, eps=eps, normalize=normalize)
    rows, cols = inds[..., 0], inds[..., 1]
    bidx = torch.arange(B * S * T, dtype=torch.long, device=inds.device)[:, None].expand(-1, P)
    values = torch.ones(B * S * T, P, dtype=flat.dtype, device=flat.device) if binarize else flat[bidx, rows, cols]
    activated = torch.zeros_like(flat)
    activated[bidx.flatten(), rows.flatten(), cols.flatten()] = values.flatten()
    activated = activated.view(B * S, T, 1, H, W)
    return activated[:, 0] if len(shape) == 4 else activated


class EnergySamplingMaskingGenerator(MaskingGenerator):
    """Sample unmasked patches where an energy map is high (sampling.py:11-112)."""

    def __init__(self, input_size, mask_ratio, seed=0, resize=True, temperature=None, clumping_factor=1,
                 pool_mode='mean', eps=1e-9, energy_power=1, **kwargs):
        super().__init__(input_size=input_size, mask_ratio=mask_ratio, clumping_factor=clumping_factor, seed=seed,
                         **kwargs)
        if resize:
            raise NotImplementedError("resize=True (bilinear resize of the energy) is not used by the CWM samplers "
                                      "(segmentation.py:36-41 sets resize=False)")
        self.pool_mode = pool_mode
        self.cf = clumping_factor
        self.temperature = temperature
        self.eps = eps
        self.energy_power = energy_power

    def _pool(self, energy, k):
        if self.pool_mode == 'mean':
            return F.avg_pool2d(energy, k, stride=k)
        if self.pool_mode == 'max':
            return F.max_pool2d(energy, k, stride=k)
        if self.pool_mode == 'min':
            return -F.max_pool2d(-energy, k, stride=k)
        raise ValueError(self.pool_mode)

    def sample_mask_per_frame(self, video):
        energy = video.reshape(-1, 1, *video.shape[-2:])
        H, W = energy.shape[-2:]
        assert (H % self.height == 0) and (W % self.width == 0)
        if (H != self.height) or (W != self.width):
            energy = self._pool(energy, ((H * self.cf) // self.height, (W * self.cf) // self.width))
        if self.temperature is not None:
            energy = torch.exp((energy - energy.amax((-2, -1), keepdim=True)) * self.temperature)
        num_points = (self.num_patches_per_frame - self.num_masks_per_frame) // (self.cf ** 2)
        if self.randomize_num_visible:
            num_points = self.rng.randint(low=0, high=(num_points + 1))
        visible = sample_from_energy(torch.pow(energy, self.energy_power), binarize=True,
                                     num_points=max(num_points, 1), eps=self.eps, normalize=True) > 0.5
        if num_points == 0:
            visible = torch.zeros_like(visible)
        if self.cf > 1:
            visible = upsample_masks(visible, size=(self.height, self.width))
        return torch.logical_not(visible).flatten(1)

    def forward(self, video, num_frames=None):
        if len(video.shape) == 4:
            video = video.unsqueeze(1)
        else:
            assert len(video.shape) == 5, video.shape
        B = video.size(0)
        masks = self.sample_mask_per_frame(video)
        masks = masks.view(B, -1, masks.shape[-1]).flatten(1)
        if B == 1 and not self.always_batch:
            masks = masks.squeeze(0)
        if self.visible_frames > 0:
            vis = torch.zeros((B, self.height * self.width), dtype=torch.bool).view(*masks.shape[:-1], -1).to(masks.device)
            masks = torch.cat(([vis] * self.visible_frames) + [masks], -1)
        return masks


class RotatedTableEnergyMaskingGenerator(EnergySamplingMaskingGenerator):
    """sampling.py:114-126."""

    def __init__(self, input_size, mask_ratio, visible_frames=1, seed=0, *args, **kwargs):
        super().__init__(input_size=(input_size[0] - visible_frames, *input_size[1:]), mask_ratio=mask_ratio,
                         visible_frames=visible_frames, seed=seed, *args, **kwargs)
        self.visible_frames = visible_frames


# ---- IMU token masks (masking.py:402-476): the head-motion stream of the conjoined models ------------------------

def patch_distance_transform(masks, self_mask=True):
    """For each patch the L-inf distance to the nearest visible patch, in units of half the grid (masking.py:32-56).
    masks bool [B, T, H, W] (True = masked) -> float [B, T, H, W]."""
    B, T, H, W = masks.shape
    flat = masks.view(B * T, H, W)
    hh, ww = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    grid = torch.stack([hh, ww], -1).to(masks.device)                                      # [H, W, 2]
    scale = torch.tensor([(H - 1) // 2, (W - 1) // 2], dtype=torch.float32, device=masks.device)
    dists = []
    for b in range(B * T):
        vis = torch.nonzero(~flat[b]).float()                                             # [N, 2]
        if vis.shape[0] == 0:
            dists.append(torch.zeros([H, W], dtype=torch.float32, device=masks.device))
            continue
        d = ((grid[None] - vis.view(-1, 1, 1, 2)) / scale).abs().amax(-1).amin(0)         # [H, W]
        if self_mask:
            d[~flat[b]] = d.amax()
        dists.append(d)
    return torch.stack(dists, 0).view(B, T, H, W)


def patches_adjacent_to_visible(masks, radius=1, size=None):
    """masking.py:58-71: patches within ``radius`` (L-inf, in patches) of a visible one; radius 0 -> a soft proximity."""
    if size is not None:
        masks = masks.view(-1, 1, *size)
    if radius is None:
        return masks
    H, W = masks.shape[-2:]
    dists = patch_distance_transform(masks)
    if radius != 0:
        return dists <= ((1 / ((min(H, W) - 1) // 2)) * radius)
    rmax = dists.amax((-1, -2), keepdim=True)
    return (rmax - dists) / rmax.clip(min=1.0)
