"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's motion-counterfactual construction,
SURVEY.md section 8(f) rank 1.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

What it restates (citations relative to /root/reference):
  * ``PatchPerturbation.forward``                 cwm/models/perturbation.py:99-112
  * ``MakeStatic.perturb``                        cwm/models/perturbation.py:129-150
  * ``ShiftPatchesAndMask._get_padding/perturb``  cwm/models/perturbation.py:226-289
  * ``PredictorBasedGenerator._shift``            cwm/models/prediction.py:756-779
  * ``FlowGenerator.create_motion_counterfactuals``  cwm/models/segmentation.py:279-343

Parity pin: ``oracle/make_golden_counterfactual.py`` runs the REAL reference classes on seeded inputs in the build
container, asserts this file reproduces them bit for bit (videos and masks) and writes ``tests/golden/cf_*.npz``.

Conventions of the reference kept here: masks are bool with True = masked; ``active_patches`` is a mask whose
False entries are the patches to move; shifts are (dy, dx); mask shifts are in patch units; videos are
[B, T, C, H, W]; tokens are ordered (t, h, w); ``patch_size`` = (pt, ph, pw) with pt == 1.
"""
import numpy as np


def fingerprint(a):
    """Order-sensitive float64 fingerprint of an array (used by the golden fixtures for videos too big to store)."""
    a = np.asarray(a, np.float64)
    w = (np.arange(a.size, dtype=np.float64).reshape(a.shape) % 7.0) + 1.0
    return np.array([a.sum(), (a * a).sum(), (a * w).sum()])


def shift_zero_fill(a, dy, dx, fill):
    """``CenterCrop(size)(F.pad(a, padding, value=fill))`` with padding = (2p, 0) for p > 0 else (0, -2p) per axis
    (perturbation.py:226-231, :257-258, :263-264): out[..., y, x] = a[..., y - dy, x - dx] inside, ``fill`` outside."""
    H, W = a.shape[-2:]
    out = np.full_like(a, fill)
    ys0, ys1 = max(0, dy), min(H, H + dy)
    xs0, xs1 = max(0, dx), min(W, W + dx)
    if ys1 > ys0 and xs1 > xs0:
        out[..., ys0:ys1, xs0:xs1] = a[..., ys0 - dy:ys1 - dy, xs0 - dx:xs1 - dx]
    return out


def get_padding_shifts(shift, patch_size):
    """Pixel shift (dy, dx) and mask shift (my, mx) the reference derives from ``shift`` (perturbation.py:226-243,
    allow_fractional_shifts=False): the mask padding divides shift[1] by patch_size[-2] and shift[0] by
    patch_size[-1] -- the axes are crossed in the reference; identical for square patches."""
    sy, sx = int(shift[0]), int(shift[1])
    mx = sx // patch_size[-2]
    my = sy // patch_size[-1]
    return (sy, sx), (my, mx)


def _pixel_mask(m, ph, pw):
    """[.., h, w] patch mask -> [.., h*ph, w*pw] pixel mask."""
    return np.repeat(np.repeat(m, ph, axis=-2), pw, axis=-1)


def shift_patches_and_mask(x, mask, patch_size, shift=None, mask_shift=None, frame=-1):
    """``ShiftPatchesAndMask.perturb`` (perturbation.py:245-289).  x float32 [B,T,C,H,W]; mask bool [B, T*h*w]."""
    B, T, C, H, W = x.shape
    pt, ph, pw = patch_size
    assert pt == 1
    h, w = H // ph, W // pw
    frame = frame % T
    if shift is not None:
        assert len(shift) == 2 and shift[0] % ph == 0 and shift[1] % pw == 0, shift  # :249-253
    else:
        assert mask_shift is not None and len(mask_shift) == 2
        shift = (int(mask_shift[0]) * ph, int(mask_shift[1]) * pw)  # :254-256
    (sy, sx), (my, mx) = get_padding_shifts(shift, patch_size)
    x_f = shift_zero_fill(x[:, frame], sy, sx, 0.0)  # :262-263
    mv = mask.reshape(B, -1, h, w)
    T_mask = mv.shape[1]
    m = mv[:, frame] if T_mask > 1 else mv[:, 0]
    m_shift = shift_zero_fill(m.astype(np.float32), my, mx, 1.0).astype(bool)  # :268-269
    if T_mask > 1:
        mask_out = mv.copy()
        mask_out[:, frame] = m_shift  # :270-271
    else:
        mask_out = m_shift[:, None]
    # only shift visible patches in target frame (:273-285): x_shift * (1 - m) + x * m, literally in fp32
    mpx = _pixel_mask(m_shift, ph, pw)[:, None].astype(np.float32)  # [B,1,H,W]
    x_out = x.copy()
    x_out[:, frame] = x_f * (np.float32(1) - mpx) + x[:, frame] * mpx
    return x_out, mask_out.reshape(mask.shape), shift


def make_static(x, mask, patch_size):
    """``MakeStatic.perturb`` (perturbation.py:129-150): visible patches of frames t > 0 are replaced by the patch
    at the same position of frame 0; the mask is returned unchanged."""
    B, T, C, H, W = x.shape
    pt, ph, pw = patch_size
    assert pt == 1 and T > 1
    h, w = H // ph, W // pw
    m = mask.reshape(B, -1, h, w)
    if m.shape[1] != T:  # :137-141
        m = np.concatenate([np.ones((B, T - m.shape[1], h, w), bool), m[:, -1:]], 1)
    mpx = _pixel_mask(m, ph, pw)[:, :, None].astype(np.float32)  # [B,T,1,H,W]
    return (np.float32(1) - mpx) * x[:, 0:1] + mpx * x, mask


def perturbation_forward(x, mask, perturbation_points, patch_size, **kw):
    """``PatchPerturbation.forward`` around ``ShiftPatchesAndMask.perturb`` (perturbation.py:99-112)."""
    mask = mask.copy()
    if perturbation_points is None:
        return shift_patches_and_mask(x, mask, patch_size, **kw)[:2]
    mask[perturbation_points] = True
    pmask = np.logical_not(perturbation_points)
    x_p, mask_p, _ = shift_patches_and_mask(x, pmask, patch_size, **kw)
    return x_p, np.minimum(mask, mask_p)


def shift_one(x, mask, active_patches, patch_size, shift=None, frame=1):
    """``PredictorBasedGenerator._shift`` before the rectangulariser (prediction.py:756-771): ``shift`` is a MASK
    shift (patch units)."""
    if active_patches is None:
        active_patches = np.ones_like(mask)
    return perturbation_forward(x, np.minimum(mask, active_patches), ~active_patches, patch_size,
                                mask_shift=shift, frame=frame)


def create_motion_counterfactuals(x, masks, active_patches, shifts, patch_size, frame=1, fix_passive=True):
    """``FlowGenerator.create_motion_counterfactuals`` up to (not including) the mask rectangulariser
    (segmentation.py:279-341).  x [B,T,C,H,W]; masks / active_patches bool [B,N,S]; shifts: S*B mask shifts
    (the reference indexes ``shifts[i]`` for i < B*S, so B must be 1 unless len(shifts) == B*S)."""
    B, N, S = masks.shape
    if active_patches is None:
        active_patches = np.ones_like(masks)
    if fix_passive:
        x = np.repeat(x[:, 0:1], 2, axis=1)  # make_static_movie, prediction.py:731-740
    x = np.repeat(x[:, None], S, axis=1).reshape(B * S, *x.shape[1:])  # sample_tile, prediction.py:484-487
    masks = masks.transpose(0, 2, 1).reshape(B * S, N)
    active = active_patches.transpose(0, 2, 1).reshape(B * S, N)
    xs, ms = [], []
    for i in range(B * S):
        xi, mi = shift_one(x[i:i + 1], masks[i:i + 1], active[i:i + 1], patch_size, shift=shifts[i], frame=frame)
        xs.append(xi)
        ms.append(mi)
    return np.concatenate(xs, 0), np.concatenate(ms, 0)
