"""Timing of the flow-statistics kernels on one B200 at 224 px: per-sample statistics +
filter mask, in-place zeroing, mean motion map; the torch-CPU oracle (reference ops) timed beside them on a bounded
sample.   python tools/flowstats_bench.py [S]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import flowstats_oracle as fso  # noqa: E402
from counterfactualworldmodels_b200 import sampling  # noqa: E402


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = "cuda:0"
    H = W = 224
    g = torch.Generator(device=dev).manual_seed(0)
    bufs = [torch.randn(S, 2, H, W, device=dev, generator=g) * 4 for _ in range(max(2, int(400e6 // (S * 2 * H * W * 4)) + 1))]
    active = torch.ones(1, 2 * 28 * 28, S, dtype=torch.bool, device=dev)
    active[0, 784 + 100:784 + 102, :] = False
    filt = sampling.FlowSampleFilter()
    views = [fso.batch_to_samples(b, 1) for b in bufs]
    i = [0]

    def nxt():
        i[0] += 1
        return views[i[0] % len(views)]

    nbytes = S * 2 * H * W * 4
    out = {"S": S, "bytes_per_pass": nbytes}
    t = timed(lambda: filt.filter_mask(nxt(), active))
    out["stats_and_filter_ms"] = t
    out["stats_GBps"] = nbytes / t / 1e6
    mask, _ = filt.filter_mask(views[0], active)
    t = timed(lambda: sampling.motion_map_finalize(sampling.flow_magnitude_sum(nxt(), filter_mask=mask), S))
    out["motion_map_ms"] = t
    out["motion_map_GBps"] = nbytes / t / 1e6
    t = timed(lambda: filt(nxt(), active))
    out["filter_forward_inplace_ms"] = t
    # motion covariance between image locations (segmentation.py:478-547) at downsample 4: N = 3136, 39 MB output
    Sc = min(S, 64)
    cview = fso.batch_to_samples(bufs[0][:Sc], 1)
    t = timed(lambda: sampling.flow_corrs(cview, downsample=4, use_covariance=True))
    out["flow_cov_ds4_ms"] = t
    out["flow_cov_ds4_samples"] = Sc
    out["flow_cov_ds4_GBps_written"] = 3136 * 3136 * 4 / t / 1e6
    t0 = time.perf_counter()
    fso.flow_corrs(fso.batch_to_samples(bufs[0][:Sc].cpu(), 1), downsample=4, use_covariance=True)
    out["flow_cov_ds4_cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    # CPU oracle (the reference's own torch ops, all host threads) on a bounded sample
    n = min(S, 32)
    fl = fso.batch_to_samples(bufs[0][:n].cpu(), 1)
    act = active[..., :n].cpu()
    t0 = time.perf_counter()
    zeroed, _, _ = fso.filter_samples(fl, act, ['patch_magnitude', 'flow_area', 'num_corners'], 5.0, 0.75, 2)
    fso.mean_motion_map(zeroed)
    out["cpu_oracle_ms_per_sample"] = (time.perf_counter() - t0) / n * 1e3
    out["cpu_threads"] = torch.get_num_threads()
    out["gpu_ms_per_sample"] = (out["filter_forward_inplace_ms"] + out["motion_map_ms"]) / S
    print(json.dumps(out))


if __name__ == "__main__":
    main()
