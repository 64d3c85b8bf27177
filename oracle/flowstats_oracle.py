"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32, the reference's own ATen ops) of the flow-derived
statistics of a counterfactual sweep, SURVEY.md section 8(f) rank 2.  Only tests/ and bench legs may import it.

What it restates (citations relative to /root/reference):
  * ``FlowSampleFilter.compute_flow_magnitude``      cwm/models/sampling.py:163-203
  * ``FlowSampleFilter.filter_by_* / forward``       cwm/models/sampling.py:205-286
  * ``FlowGenerator.compute_flow_samples_magnitude`` cwm/models/segmentation.py:250-255
  * ``FlowGenerator.compute_mean_motion_map``        cwm/models/segmentation.py:257-276

Parity pin: ``oracle/make_golden_flowstats.py`` runs the REAL reference classes on seeded inputs, asserts this file
reproduces them exactly (``torch.equal``) and writes ``tests/golden/fs_*.npz``.
"""
import torch
import torch.nn.functional as F


def make_flows(B, S, H, W, seed):
    """Seeded synthetic flow samples, layout [(b s), 2, H, W] like a flow network's output: a moving blob of random
    size / speed per sample (some tiny, some covering the whole image, some touching corners) on low-level noise."""
    g = torch.Generator().manual_seed(4000 + seed)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    out = torch.randn(B * S, 2, H, W, generator=g) * 0.3
    centers = []
    for i in range(B * S):
        kind = i % 4
        cy, cx = [float(v) for v in (torch.rand(2, generator=g) * torch.tensor([H, W]))]
        r = float(torch.rand(1, generator=g)) * H * (0.15 if kind else 1.5) + 2.0
        if kind == 3:
            cy, cx, r = 0.0, 0.0, H * 0.9  # covers corners
        speed = (torch.rand(2, generator=g) * 2 - 1) * (14.0 if kind != 2 else 2.0)
        blob = (((ys - cy) ** 2 + (xs - cx) ** 2) < r * r).float()
        out[i] += blob[None] * speed[:, None, None]
        centers.append((cy, cx))
    return out, centers


def make_active(B, S, n_h, n_w, centers, patch):
    """active_patches bool [B, 2*n_h*n_w, S] (False = active): one 2x2 clump per sample near the blob centre."""
    act = torch.ones(B, 2, n_h, n_w, S, dtype=torch.bool)
    for i, (cy, cx) in enumerate(centers):
        b, s = divmod(i, S)
        py = min(max(int(cy // patch), 0), n_h - 2)
        px = min(max(int(cx // patch), 0), n_w - 2)
        act[b, 1, py:py + 2, px:px + 2, s] = False
    return act.reshape(B, 2 * n_h * n_w, S)


def batch_to_samples(flows, B):
    """'(b s) c h w -> b c h w s' (segmentation.py:130-133) as a permuted VIEW."""
    x = flows.reshape(B, -1, *flows.shape[1:])
    return x.permute(0, 2, 3, 4, 1)


def compute_flow_magnitude(flow_samples, active_patches):
    """sampling.py:163-203."""
    flow_mag = flow_samples.norm(dim=1, p=2)
    B, _, H, W, num_samples = flow_samples.shape
    _, num_patches, _ = active_patches.shape
    h = w = int((num_patches / 2) ** 0.5)
    active_second = 1 - active_patches[:, (h * w):, :].float()
    active_second = active_second.permute(0, 2, 1)
    flow_mag_down = F.interpolate(flow_mag.permute(0, 3, 1, 2), size=[h, w], mode='bilinear').flatten(2, 3)
    patch_flow_mag = (flow_mag_down * active_second).sum(dim=-1) / (active_second.sum(-1) + 1e-12)
    return flow_mag, patch_flow_mag


def sample_statistics(flow_samples, active_patches, thr):
    """-> dict of [B, S] tensors: patch_flow_mag, flow_area (sampling.py:222-223), num_corners (:232-247), min, max."""
    flow_mag, patch_flow_mag = compute_flow_magnitude(flow_samples, active_patches)
    _, H, W, _ = flow_mag.shape
    flow_area = (flow_mag > thr).flatten(1, 2).sum(1) / (H * W)
    fb = (flow_mag > thr).float()
    corners = fb[:, 0, 0] + fb[:, 0, -1] + fb[:, -1, 0] + fb[:, -1, -1]
    return dict(patch_flow_mag=patch_flow_mag, flow_area=flow_area, num_corners=corners,
                min=flow_mag.amin((1, 2)), max=flow_mag.amax((1, 2)))


def filter_samples(flow_samples, active_patches, methods, mag_thr, area_thr, corners_thr):
    """``FlowSampleFilter.forward`` (sampling.py:252-286) -> (zeroed copy of the flows, filter mask [B, S])."""
    st = sample_statistics(flow_samples, active_patches, mag_thr)
    mask = torch.zeros_like(st["flow_area"], dtype=torch.bool)
    for m in methods:
        if m == 'patch_magnitude':
            mask = mask | (st["patch_flow_mag"] < mag_thr)
        elif m == 'flow_area':
            mask = mask | (st["flow_area"] > area_thr)
        elif m == 'num_corners':
            mask = mask | (st["num_corners"] >= corners_thr)
        else:
            raise ValueError(m)
    out = flow_samples.clone()
    out[mask[:, None, None, None, :].expand_as(out)] = 0.
    return out, mask, st


def mean_motion_map(flows, normalize_per_sample=False, normalize=True, eps=1e-2):
    """segmentation.py:250-276."""
    if flows.dim() == 5:
        mags = flows.square().sum(-4, True).sqrt()
        if normalize_per_sample:
            mags = mags - mags.amin((-3, -2), True)
            mags = mags / mags.amax((-3, -2), True).clamp(min=eps)
        mm = mags.mean(-1)
    else:
        mm, normalize = flows, True
    if normalize:
        mm = mm - mm.amin((-2, -1), True)
        mm = mm / mm.amax((-2, -1), True).clamp(min=eps)
    return mm


def flow_corrs(flow_samples, downsample=1, take_top_k=None, use_covariance=False):
    """``FlowGenerator.compute_flow_corrs`` with its default options (segmentation.py:478-547)."""
    B, C, H, W, S = flow_samples.shape
    K = S if take_top_k is None else take_top_k
    ds = downsample
    inp = F.avg_pool3d(flow_samples[..., :K].permute(0, 1, 4, 2, 3), (1, ds, ds), stride=(1, ds, ds)).permute(0, 1, 3, 4, 2)
    inp = torch.sqrt((inp - torch.zeros_like(inp)).square().mean(1, True).float())   # ChannelMSE(dim=1), utils.py:510-513
    inp = inp.reshape(B, -1, inp.size(-1))
    out = []
    for b in range(B):
        c = torch.cov(inp[b]) if use_covariance else torch.corrcoef(inp[b])
        c[torch.isnan(c)] = 0
        out.append(c)
    return torch.stack(out, 0).view(B, 1, H // ds, W // ds, H // ds, W // ds)
