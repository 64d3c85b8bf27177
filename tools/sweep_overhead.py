"""Where the fixed (not per-chunk) time of a sharded counterfactual sweep goes: host-side stages of
dist.sharded_counterfactual_videos timed with a device synchronisation after each.   python tools/sweep_overhead.py [S]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from counterfactualworldmodels_b200 import segmentation, synthetic, vmae  # noqa: E402
from counterfactualworldmodels_b200 import dist as cdist  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    cfg = sys.argv[2] if len(sys.argv) > 2 else "base_8x8"
    dev = "cuda:0"
    model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg))
    synthetic.init_weights_(model, seed=0, style="reference")
    model = model.to(dev).eval()
    G = segmentation.FlowGenerator(predictor=model, imagenet_normalize_inputs=True, temporal_dim=2)
    a, p, shifts = bench.sweep_descriptors(model.mask_size, S, seed=0)
    x = synthetic.make_video(1, synthetic.image_hw(cfg), seed=7)[:, 0].to(dev)
    a, p = a.to(dev), p.to(dev)
    chunk = bench.SWEEP_CHUNK[cfg]

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for rep in range(3):
        t0 = sync()
        xx = x.unsqueeze(1).expand(-1, 2, -1, -1, -1)
        G.set_input(xx)
        G.reset_shifts()
        G.shifter.set_shapes(xx, mask=a[..., 0])
        G.shifter.set_num_shifts(len(shifts))
        sh = G.shifter._preprocess_shifts_sequence(shifts, is_mask_shift=True)
        t1 = sync()
        video, masks = G.create_motion_counterfactuals(xx, masks=p, active_patches=a, shifts=sh, num_samples=S, fix_passive=True,
                                                       reset_shifts=False, frame=1, virtual=True)
        t2 = sync()
        ys = []
        tc = []
        for b0 in range(0, S, chunk):
            c0 = sync()
            ys.append(G.predict(video[b0:b0 + chunk], mask=masks[b0:b0 + chunk], frame=-1, reset_masks=True))
            tc.append(sync() - c0)
        t3 = sync()
        y = torch.cat(ys, 0)
        t4 = sync()
        print(f"rep {rep}: prepare {1e3 * (t1 - t0):.2f} ms | create_motion_counterfactuals {1e3 * (t2 - t1):.2f} ms | "
              f"{len(tc)} chunks {1e3 * (t3 - t2):.2f} ms (median chunk {1e3 * sorted(tc)[len(tc) // 2]:.2f}, max {1e3 * max(tc):.2f}) | "
              f"cat {1e3 * (t4 - t3):.2f} ms | total {1e3 * (t4 - t0):.2f} ms")
    t0 = sync()
    for _ in range(2):
        cdist.sharded_counterfactual_videos(G, x, a, passive_patches=p, shifts=shifts, sample_batch_size=chunk, dst=0,
                                            predict_frame=-1)
    print(f"sharded_counterfactual_videos: {1e3 * (sync() - t0) / 2:.2f} ms per sweep")


if __name__ == "__main__":
    main()
