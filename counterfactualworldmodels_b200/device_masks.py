"""Masks on the device with a counter-based RNG (csrc/masks.cu) -- OPT-IN.

The reference generates masks on the host from sequential RNG streams (cwm/models/masking.py, sampling.py,
utils.py:152-213); ``masking.py`` of this package reproduces those streams bit for bit and stays the default.  For
sweeps of a thousand samples that host loop is as long as the forward itself, and its result depends on how many
draws were made before -- so a sweep sharded over GPUs cannot regenerate "its" masks locally.  The generators here
draw from Philox4x32-10 keyed by ``seed`` with the GLOBAL sample index in the counter: a sample's mask is a pure
function of (seed, sample index), whatever the batch split or the number of ranks.  They keep the reference
classes' constructor arguments and output conventions (bool, True = masked, ``[B, N]`` with the fully visible frames
first), not their random streams.  No CPU fallback.
"""
import torch
from torch import nn

from . import _lib


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(device, who):
    if torch.device(device).type != "cuda":
        raise RuntimeError(f"{who}: masks are generated on a CUDA (B200) device; there is no CPU fallback "
                           "(use counterfactualworldmodels_b200.masking for the host generators)")


class DeviceUniformMaskingGenerator(nn.Module):
    """``RotatedTableUniformMaskingGenerator`` (masking.py:478-545) on the device: the first ``visible_frames`` frames
    fully visible, every later frame masked except ``num_visible`` patches in ``clumping_factor``-sized square clumps
    chosen uniformly without replacement.  ``forward(x)`` -> bool ``[B, N]``; row b is global sample
    ``sample_offset + b`` of the stream ``seed``."""

    def __init__(self, input_size, mask_ratio, visible_frames=None, seed=0, clumping_factor=1, always_batch=True,
                 device=None):
        super().__init__()
        assert len(input_size) == 3, input_size
        self.frames, self.height, self.width = (int(v) for v in input_size)
        self.visible_frames = self.frames - 1 if visible_frames is None else int(visible_frames)
        self.clumping_factor = int(clumping_factor)
        if self.height % self.clumping_factor or self.width % self.clumping_factor:
            raise NotImplementedError("the patch grid must be a multiple of the clumping factor on the device path")
        self.num_patches_per_frame = (self.height // self.clumping_factor) * (self.width // self.clumping_factor)
        self.mask_ratio = mask_ratio
        self.seed = int(seed)
        self.always_batch = always_batch
        self.device = device

    # the reference's three coupled views of "how much is masked" (masking.py:300-335), in clump cells
    @property
    def mask_ratio(self):
        return self._mask_ratio

    @mask_ratio.setter
    def mask_ratio(self, value):
        self._mask_ratio = value
        self.num_masks_per_frame = int(value * self.num_patches_per_frame)

    @property
    def num_visible(self):
        return self.num_patches_per_frame - self.num_masks_per_frame

    @num_visible.setter
    def num_visible(self, value):
        self.num_masks_per_frame = self.num_patches_per_frame - int(value)
        self._mask_ratio = self.num_masks_per_frame / self.num_patches_per_frame

    def forward(self, x=None, batch_size=None, sample_offset=0):
        B = int(batch_size) if batch_size is not None else (x.size(0) if isinstance(x, torch.Tensor) else 1)
        device = self.device if self.device is not None else (x.device if isinstance(x, torch.Tensor) else "cuda")
        _require_cuda(device, type(self).__name__)
        mask_frames = self.frames - self.visible_frames
        out = torch.empty(B, self.frames * self.height * self.width, dtype=torch.uint8, device=device)
        with torch.cuda.device(out.device):
            _lib.check(_lib.load().cwm_mask_uniform(self.seed, int(sample_offset), B, self.visible_frames, mask_frames,
                                                    self.height, self.width, self.clumping_factor, self.num_visible,
                                                    out.data_ptr(), _stream(out.device)))
        masks = out.view(torch.bool)
        return masks if (B > 1 or self.always_batch) else masks[0]


class DeviceEnergyMaskingGenerator(DeviceUniformMaskingGenerator):
    """``RotatedTableEnergyMaskingGenerator`` / ``EnergySamplingMaskingGenerator`` (sampling.py:12-130) on the device: the
    visible clumps of the last frame are drawn WITH replacement from the categorical distribution the reference builds
    from an energy map -- pool to the clump grid (``pool_mode``), ``** energy_power``, shift by the minimum, ``+ eps``,
    normalise (sampling.py:63-90, utils.py:152-172).  ``sample(energy, num_samples)`` draws a whole sweep in two
    launches and returns the reference's ``[B, N, S]`` layout (as a view)."""

    def __init__(self, input_size, mask_ratio=0, visible_frames=None, seed=0, clumping_factor=1, always_batch=True,
                 energy_power=1, eps=1e-16, pool_mode='mean', resize=False, temperature=None, device=None, **unused):
        super().__init__(input_size, mask_ratio, visible_frames=visible_frames, seed=seed,
                         clumping_factor=clumping_factor, always_batch=always_batch, device=device)
        if resize:
            raise NotImplementedError("resize=True (bilinear resizing of the energy map) is not on the sweep path")
        if pool_mode not in ('mean', 'max', 'min'):
            raise ValueError(pool_mode)
        self.energy_power, self.eps, self.pool_mode, self.temperature = energy_power, eps, pool_mode, temperature

    def _cell_weights(self, energy):
        import torch.nn.functional as F
        e = energy.reshape(-1, 1, *energy.shape[-2:]).float()
        gh, gw = self.height // self.clumping_factor, self.width // self.clumping_factor
        H, W = e.shape[-2:]
        assert H % gh == 0 and W % gw == 0, (e.shape, gh, gw)
        if (H, W) != (gh, gw):
            k = (H // gh, W // gw)
            e = {'mean': F.avg_pool2d(e, k, stride=k), 'max': F.max_pool2d(e, k, stride=k),
                 'min': -F.max_pool2d(-e, k, stride=k)}[self.pool_mode]
        if self.temperature is not None:
            e = torch.exp(e / self.temperature)
        return torch.pow(e, self.energy_power).reshape(e.shape[0], gh * gw).contiguous()

    def sample(self, energy, num_samples, sample_offset=0):
        """energy ``[B, 1, H, W]`` (or ``[B, H, W]``) on the device -> bool ``[B, N, num_samples]`` (view of ``[B, S, N]``)."""
        _require_cuda(energy.device, type(self).__name__)
        lib = _lib.load()
        w = self._cell_weights(energy)
        B, n = w.shape
        S = int(num_samples)
        points = max(self.num_visible, 1)
        table = torch.empty(B, n, dtype=torch.int64, device=w.device)
        out = torch.empty(B * S, self.frames * self.height * self.width, dtype=torch.uint8, device=w.device)
        if self.frames - self.visible_frames != 1:
            raise NotImplementedError("energy sampling fills one masked frame (the counterfactual frame)")
        with torch.cuda.device(w.device):
            st = _stream(w.device)
            _lib.check(lib.cwm_mask_energy_table(w.data_ptr(), B, n, float(self.eps), table.data_ptr(), st))
            _lib.check(lib.cwm_mask_energy_sample(table.data_ptr(), B, self.height, self.width, self.clumping_factor,
                                                  self.seed, int(sample_offset), S, points if self.num_visible else 0,
                                                  self.visible_frames, out.data_ptr(), st))
        return out.view(torch.bool).view(B, S, -1).permute(0, 2, 1)

    def forward(self, energy, sample_offset=0):
        masks = self.sample(energy, 1, sample_offset=sample_offset)[..., 0]
        return masks if (masks.size(0) > 1 or self.always_batch) else masks[0]


class DeterministicRectangularizeMasks(nn.Module):
    """``RectangularizeMasks('min')`` (masking.py:100-132) without the global ``torch.randperm``: every row keeps as many
    masked tokens as the row with the fewest, and WHICH tokens of row r are revealed is a function of
    (seed, row_offset + r) alone.  In place on a CUDA bool/uint8 ``[B, N]`` tensor, like the reference.  Shards of a
    larger batch pass the global minimum as ``target_masked``."""

    def __init__(self, seed=0):
        super().__init__()
        self.seed = int(seed)
        self._mode = 'min'

    def forward(self, masks, row_offset=0, target_masked=None):
        _require_cuda(masks.device, type(self).__name__)
        assert masks.dtype in (torch.bool, torch.uint8) and masks.is_contiguous(), (masks.dtype, masks.is_contiguous())
        flat = masks.view(torch.uint8).view(masks.shape[0], -1)
        rows, N = flat.shape
        lib = _lib.load()
        ws = torch.empty(lib.cwm_mask_rectangularize_workspace_bytes(rows), dtype=torch.uint8, device=masks.device)
        with torch.cuda.device(masks.device):
            _lib.check(lib.cwm_mask_rectangularize(flat.data_ptr(), rows, N, int(row_offset), self.seed,
                                                   -1 if target_masked is None else int(target_masked), ws.data_ptr(),
                                                   ws.numel(), _stream(masks.device)))
        return masks
