"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy fp32) of RAFT's non-convolution stages, SURVEY.md section
8(f) rank 3 (first slice).  Only tests/ and bench legs may import it; the product path is csrc/raftcorr.cu.

What it restates (citations relative to /root/reference):
  * ``CorrBlock.corr`` / ``CorrBlock.__init__``   cwm/models/raft/corr.py:12-28, 53-60   -> corr_pyramid()
  * ``CorrBlock.__call__``                         cwm/models/raft/corr.py:30-51         -> corr_lookup()
  * ``bilinear_sampler``                           cwm/models/raft/utils.py:60-80         (inside corr_lookup:
      pixel coordinates -> [-1, 1] -> ``F.grid_sample(align_corners=True)``, bilinear, zero padding; the
      un-normalisation and corner weights follow ATen's ``grid_sampler_2d``)
  * ``RAFT.upsample_flow``                         cwm/models/raft/raft_model.py:175-186  -> upsample_flow()

Parity pin: ``oracle/make_golden_raft.py`` runs the REAL reference classes on the seeded inputs below, asserts this
file reproduces them to 2e-6 of the output scale (numpy's BLAS order / exp differ from ATen's in the last bit, so the
pin is a tolerance, not ``array_equal``) and writes ``tests/golden/raft_*.npz`` with the reference's outputs.
"""
import numpy as np

F32 = np.float32


def make_fmaps(B, D, H, W, seed):
    """Two seeded unit-variance feature maps [B, D, H, W]; fmap2 is a shifted, noised copy of fmap1 so the volume has
    the peaked structure of a real matching problem (correlation ~ sqrt(D) on the true match, ~1 elsewhere)."""
    rng = np.random.RandomState(7000 + seed)
    f1 = rng.standard_normal((B, D, H, W)).astype(F32)
    f2 = np.roll(f1, shift=(1, -2), axis=(2, 3)) * F32(0.8) + rng.standard_normal((B, D, H, W)).astype(F32) * F32(0.6)
    return f1, f2.astype(F32)


def make_coords(B, H, W, seed, kind="random"):
    """[B, 2, H, W] fp32 lookup centres: 'grid' = the exact integer grid of RAFT's first iteration
    (raft_model.py:166-173), 'random' = grid + N(0, 2.5) flow with a few far-out-of-bounds and half-integer points."""
    ys, xs = np.meshgrid(np.arange(H, dtype=F32), np.arange(W, dtype=F32), indexing="ij")
    grid = np.broadcast_to(np.stack([xs, ys], 0)[None], (B, 2, H, W)).astype(F32).copy()
    if kind == "grid":
        return grid
    rng = np.random.RandomState(7100 + seed)
    c = grid + rng.standard_normal((B, 2, H, W)).astype(F32) * F32(2.5)
    flat = c.reshape(-1)
    idx = rng.choice(flat.size, size=max(4, flat.size // 16), replace=False)
    flat[idx[0::4]] += F32(40.0)           # far outside on the high side
    flat[idx[1::4]] -= F32(37.5)           # far outside on the low side
    flat[idx[2::4]] = np.round(flat[idx[2::4]]) + F32(0.5)  # exact half-integers
    flat[idx[3::4]] = np.round(flat[idx[3::4]])             # exact integers
    return c


def make_upsample_inputs(N, C, H, W, seed):
    rng = np.random.RandomState(7200 + seed)
    flow = (rng.standard_normal((N, C, H, W)) * 3.0).astype(F32)
    mask = (rng.standard_normal((N, 576, H, W)) * 2.0).astype(F32)
    return flow, mask


def corr_pyramid(fmap1, fmap2, num_levels=4):
    """-> list of fp32 [B*H*W, H>>l, W>>l] (corr.py:18-28, 53-60)."""
    B, D, H, W = fmap1.shape
    a = fmap1.reshape(B, D, H * W).astype(F32)
    b = fmap2.reshape(B, D, H * W).astype(F32)
    corr = np.matmul(a.transpose(0, 2, 1), b) / np.sqrt(F32(D))
    pyr = [corr.astype(F32).reshape(B * H * W, H, W)]
    for _ in range(num_levels - 1):
        c = pyr[-1]
        h, w = c.shape[1] // 2, c.shape[2] // 2
        c = c[:, :2 * h, :2 * w]
        s = ((c[:, 0::2, 0::2] + c[:, 0::2, 1::2]) + c[:, 1::2, 0::2]) + c[:, 1::2, 1::2]
        pyr.append((s / F32(4)).astype(F32))
    return pyr


def _axis(c, off, size):
    """Sample position of one window offset along one axis, exactly as the reference computes it in fp32."""
    pos = (c + F32(off)).astype(F32)                        # centroid_lvl + delta_lvl         corr.py:42-44
    g = (F32(2) * pos / F32(size - 1) - F32(1)).astype(F32)  # 2*x/(W-1) - 1                    utils.py:64-65
    ix = (((g + F32(1)) / F32(2)) * F32(size - 1)).astype(F32)  # grid_sampler_unnormalize, align_corners=True
    fl = np.floor(ix)
    return fl.astype(np.int64), ((fl + F32(1)) - ix).astype(F32), (ix - fl).astype(F32)


def corr_lookup(pyramid, coords, radius=4):
    """coords [B, 2, H, W] -> fp32 [B, L*(2r+1)^2, H, W] (corr.py:30-51).  Channel l*(2r+1)^2 + a*(2r+1) + b samples
    level l at (x/2^l + a - r, y/2^l + b - r): ``meshgrid(dy, dx)`` stacked as (.., 2) puts the FIRST window axis on x."""
    B, _, H, W = coords.shape
    P, r, n1 = B * H * W, radius, 2 * radius + 1
    cx = coords[:, 0].reshape(P).astype(F32)
    cy = coords[:, 1].reshape(P).astype(F32)
    rows = np.arange(P)
    outs = []
    for lvl, c in enumerate(pyramid):
        Hl, Wl = c.shape[1:]
        sx, sy = cx / F32(2 ** lvl), cy / F32(2 ** lvl)
        xs = [_axis(sx, a - r, Wl) for a in range(n1)]
        ys = [_axis(sy, b - r, Hl) for b in range(n1)]
        out_l = np.zeros((P, n1, n1), F32)
        for a, (x0, xlo, xhi) in enumerate(xs):
            for b, (y0, ylo, yhi) in enumerate(ys):
                acc = np.zeros(P, F32)
                for yy, xx, wt in ((y0, x0, xlo * ylo), (y0, x0 + 1, xhi * ylo), (y0 + 1, x0, xlo * yhi),
                                   (y0 + 1, x0 + 1, xhi * yhi)):  # nw, ne, sw, se
                    ok = (yy >= 0) & (yy < Hl) & (xx >= 0) & (xx < Wl)
                    v = np.where(ok, c[rows, np.clip(yy, 0, Hl - 1), np.clip(xx, 0, Wl - 1)], F32(0))
                    acc = (acc + v * wt).astype(F32)
                out_l[:, a, b] = acc
        outs.append(out_l.reshape(B, H, W, n1 * n1))
    return np.ascontiguousarray(np.concatenate(outs, -1).transpose(0, 3, 1, 2))


def upsample_flow(flow, mask):
    """[N, C, H, W], [N, 576, H, W] -> [N, C, 8H, 8W] (raft_model.py:175-186)."""
    N, C, H, W = flow.shape
    m = mask.reshape(N, 1, 9, 8, 8, H, W).astype(F32)
    e = np.exp(m - m.max(2, keepdims=True))
    p = (e / e.sum(2, keepdims=True)).astype(F32)
    f8 = np.pad(F32(8) * flow.astype(F32), ((0, 0), (0, 0), (1, 1), (1, 1)))
    nb = np.stack([f8[:, :, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)], 2)  # F.unfold order
    up = (p * nb[:, :, :, None, None]).sum(2, dtype=F32)                                        # [N, C, 8, 8, H, W]
    return np.ascontiguousarray(up.transpose(0, 1, 4, 2, 5, 3).reshape(N, C, 8 * H, 8 * W))


# ---- the same stages on the reference's own ATen ops (torch, all host threads): the CPU baseline that
# tools/raft_bench.py times beside the kernels.  Checked against the numpy restatement in tests/test_raft.py.
def torch_corr_block(fmap1, fmap2, coords_list, num_levels=4, radius=4):
    """fmaps [B, D, H, W] torch fp32, coords_list: list of [B, 2, H, W] -> list of [B, L*(2r+1)^2, H, W]."""
    import torch
    import torch.nn.functional as F
    B, D, H, W = fmap1.shape
    vol = torch.matmul(fmap1.reshape(B, D, H * W).transpose(1, 2), fmap2.reshape(B, D, H * W))
    vol = (vol / torch.sqrt(torch.tensor(D).float())).reshape(B * H * W, 1, H, W)
    pyr = [vol]
    for _ in range(num_levels - 1):
        pyr.append(F.avg_pool2d(pyr[-1], 2, stride=2))
    n1 = 2 * radius + 1
    d = torch.linspace(-radius, radius, n1)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, n1, n1, 2)  # [..., 0] varies along axis 0
    outs = []
    for coords in coords_list:
        centres = coords.permute(0, 2, 3, 1).reshape(B * H * W, 1, 1, 2)
        per_level = []
        for lvl, c in enumerate(pyr):
            pos = centres / 2 ** lvl + delta
            hl, wl = c.shape[-2:]
            grid = torch.cat([2 * pos[..., :1] / (wl - 1) - 1, 2 * pos[..., 1:] / (hl - 1) - 1], dim=-1)
            per_level.append(F.grid_sample(c, grid, align_corners=True).view(B, H, W, -1))
        outs.append(torch.cat(per_level, dim=-1).permute(0, 3, 1, 2).contiguous().float())
    return outs


def torch_upsample_flow(flow, mask):
    import torch
    import torch.nn.functional as F
    N, C, H, W = flow.shape
    w = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    nb = F.unfold(8 * flow, [3, 3], padding=1).view(N, C, 9, 1, 1, H, W)
    return torch.sum(w * nb, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(N, C, 8 * H, 8 * W)


# ---- elementwise half of the recurrent block (mixed-precision path): numpy fp32 on f16-rounded inputs ---------
def _h(x):
    """Round to f16 and back (what the kernels store between convolutions)."""
    return np.asarray(x, F32).astype(np.float16).astype(F32)


def bias_act(x, bias, relu=True, tail=None):
    """conv output + bias (+ relu), last columns replaced by ``tail`` -- ``F.relu(conv(x))`` / ``cat([out, flow])``
    (cwm/models/raft/update.py:90-98)."""
    y = x.astype(F32) + (0 if bias is None else bias.astype(F32))
    if relu:
        y = np.maximum(y, 0)
    if tail is not None:
        y[:, y.shape[1] - tail.shape[1]:] = tail
    return _h(y)


def gru_gate(zr, bias, h):
    """z = sigmoid(convz(hx)), r = sigmoid(convr(hx)) -> (z, r * h)  (update.py:46-47, :53-54); zr = [z | r] stacked."""
    C = h.shape[1]
    s = 1.0 / (1.0 + np.exp(-(zr.astype(F32) + bias.astype(F32))))
    return _h(s[:, :C]), _h(s[:, C:] * h.astype(F32))


def gru_update(q, bias, z, h):
    """h = (1 - z) * h + z * tanh(convq(...))  (update.py:48-49, :55-56)."""
    return _h((1 - z.astype(F32)) * h.astype(F32) + z.astype(F32) * np.tanh(q.astype(F32) + bias.astype(F32)))


def flow_update(delta, bias, coords1):
    """coords1 += delta_flow; flow = coords1 - coords0  (raft_model.py:249-254).  delta [M, >=2] rows, coords1
    [B, 2, H, W] -> (new coords1, flow rows [M, 2])."""
    B, _, H, W = coords1.shape
    d = (delta[:, :2].astype(F32) + bias.astype(F32)).reshape(B, H, W, 2).transpose(0, 3, 1, 2)
    new = (coords1.astype(F32) + d).astype(F32)
    grid = make_coords(B, H, W, 0, "grid")
    return new, _h((new - grid).transpose(0, 2, 3, 1).reshape(-1, 2))
