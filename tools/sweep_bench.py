"""One motion-counterfactual sweep end to end on one B200, stage by stage (CUDA events on the current stream):
(image, S active / passive patches, shifts) -> fused counterfactual construction + VMAE prediction
(`FlowGenerator.predict_counterfactual_videos`) -> RAFT flow (`raft.RAFT`: cuDNN convolutions + this repo's correlation /
lookup / upsampling kernels) -> flow-sample filter -> mean motion map.  Shows where a full movability iteration spends
its time once the VMAE path is fast.      python tools/sweep_bench.py [S] [config] [raft mode: fp32|tf32|f16]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from cf_bench import make_sweep  # noqa: E402
from counterfactualworldmodels_b200 import raft, segmentation, synthetic, vmae  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfg = sys.argv[2] if len(sys.argv) > 2 else "base_8x8"
    mode = sys.argv[3] if len(sys.argv) > 3 else "f16"
    dev = "cuda:0"
    P = synthetic.CONFIGS[cfg]["patch_size"][0]
    h = 224 // P
    passive, active, shifts = make_sweep(S, h, np.random.RandomState(0))
    x = synthetic.make_video(1, (224, 224), seed=0).to(dev)
    p3, a3 = passive.to(dev).t().unsqueeze(0), active.to(dev).t().unsqueeze(0)   # [1, N, S]
    model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg))
    synthetic.init_weights_(model, seed=0)
    torch.manual_seed(0)
    args = raft.get_args("")
    args.multiframe, args.scale_inputs, args.output_dim = True, True, None
    args.mixed_precision = mode == "f16"
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    flow_model = raft.RAFT(args)
    G = segmentation.FlowGenerator(predictor=model.to(dev).eval(), imagenet_normalize_inputs=True, temporal_dim=2,
                                   flow_model=flow_model.to(dev), raft_iters=24)
    G.set_input(x)
    chunk = min(S, 64)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def sweep(record):
        torch.manual_seed(0)
        if record:
            ev[0].record()
        ys = G.predict_counterfactual_videos(x, a3, passive_patches=p3, shifts=shifts, sample_batch_size=chunk)
        if record:
            ev[1].record()
        flows = torch.cat([G.predict_flow(ys[i:i + chunk], iters=24) for i in range(0, S, chunk)], 0)
        if record:
            ev[2].record()
        samples = G.filter_flow_samples(flows, a3)
        if record:
            ev[3].record()
        mm = G.compute_mean_motion_map(samples)
        if record:
            ev[4].record()
        return mm

    for _ in range(2):
        sweep(False)
    torch.cuda.synchronize()
    reps = 3
    acc = np.zeros(4)
    for _ in range(reps):
        mm = sweep(True)
        torch.cuda.synchronize()
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    acc /= reps
    total = float(acc.sum())
    out = {"S": S, "config": cfg, "raft_convs": mode, "ms": dict(zip(["counterfactual_predict", "raft_flow", "flow_filter",
                                                                      "motion_map"], [round(float(v), 3) for v in acc])),
           "ms_total": round(total, 3), "counterfactuals_per_s": round(S / total * 1e3, 1),
           "share": {k: round(float(v) / total, 4) for k, v in zip(["counterfactual_predict", "raft_flow", "flow_filter",
                                                                    "motion_map"], acc)},
           "motion_map_finite": bool(torch.isfinite(mm).all())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
