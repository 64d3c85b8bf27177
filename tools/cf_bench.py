"""Timing of the batched motion-counterfactual construction on one B200:
mask kernel, materialising kernel (HBM GB/s), fused vs materialised counterfactual prediction, and the CPU oracle's
per-sample construction time beside them.   python tools/cf_bench.py [S] [config]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from counterfactualworldmodels_b200 import perturbation, segmentation, synthetic, vmae  # noqa: E402


def timed(fn, iters=20, warmup=3):
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


def make_sweep(S, h, rng):
    active = torch.ones(S, 2, h, h, dtype=torch.bool)
    passive = torch.zeros(S, 2, h, h, dtype=torch.bool)
    passive[:, 1] = True
    shifts = []
    for s in range(S):
        ay, ax = rng.randint(4, h - 6), rng.randint(4, h - 6)
        active[s, 1, ay:ay + 2, ax:ax + 2] = False
        py, px = 2 * rng.randint(0, h // 2), 2 * rng.randint(0, h // 2)
        passive[s, 1, py:py + 2, px:px + 2] = False
        sh = [0, 0]
        while sh == [0, 0]:
            sh = [int(rng.randint(-2, 3)), int(rng.randint(-2, 3))]
        shifts.append(sh)
    return passive.reshape(S, -1), active.reshape(S, -1), shifts


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cfg = sys.argv[2] if len(sys.argv) > 2 else "base_8x8"
    dev = "cuda:0"
    P = synthetic.CONFIGS[cfg]["patch_size"][0]
    h = 224 // P
    rng = np.random.RandomState(0)
    passive, active, shifts = make_sweep(S, h, rng)
    x = synthetic.make_video(1, (224, 224), seed=0).to(dev)
    pd, ad = passive.to(dev), active.to(dev)
    out = {"S": S, "config": cfg}

    def build():
        return perturbation.shift_patches_and_masks(x, pd, ad, shifts, (1, P, P), frame=1, static_frame=0)

    video, mask = build()
    out["masks_ms"] = timed(lambda: build())
    ms = timed(lambda: video.materialize())
    nbytes = S * 2 * 3 * 224 * 224 * 4
    out["build_videos_ms"] = ms
    out["build_videos_GBps_written"] = nbytes / ms / 1e6

    model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg))
    synthetic.init_weights_(model, seed=0)
    model = model.to(dev).eval()
    G = segmentation.FlowGenerator(predictor=model, imagenet_normalize_inputs=True, temporal_dim=2)
    p3, a3 = pd.t().unsqueeze(0), ad.t().unsqueeze(0)   # [1, N, S]

    def fused():
        torch.manual_seed(0)  # the mask rectangulariser draws from the global RNG when rows differ
        return G.predict_counterfactual_videos(x, a3, passive_patches=p3, shifts=shifts, sample_batch_size=min(S, 64))

    def materialised():
        torch.manual_seed(0)
        G.set_input(x)
        xs, m = G.create_motion_counterfactuals(x, masks=p3, active_patches=a3, shifts=shifts, reset_shifts=True)
        return G.batch_predict_per_sample(xs, masks=m, frame=None, batch_size=min(S, 64), sample_dim=0)

    y0, y1 = fused(), materialised()
    assert torch.equal(y0, y1)
    out["predict_fused_ms"] = timed(fused, iters=10)
    out["predict_materialised_ms"] = timed(materialised, iters=10)
    out["frames_per_s_fused"] = S / out["predict_fused_ms"] * 1e3

    # CPU: the reference's per-sample construction loop, restated (oracle), bounded sample
    import counterfactual_oracle as cfo
    n = min(S, 16)
    xc = x.cpu().numpy()
    p_np = passive.numpy().T[None][:, :, :n]
    a_np = active.numpy().T[None][:, :, :n]
    t0 = time.time()
    cfo.create_motion_counterfactuals(xc, p_np, a_np, shifts[:n], (1, P, P), frame=1, fix_passive=True)
    out["cpu_oracle_ms_per_sample"] = (time.time() - t0) / n * 1e3
    out["gpu_ms_per_sample_materialised"] = (out["masks_ms"] + out["build_videos_ms"]) / S
    print(json.dumps(out))


if __name__ == "__main__":
    main()
