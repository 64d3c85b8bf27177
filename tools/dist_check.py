"""Multi-GPU check of the sharded counterfactual sweep (run under torchrun on N GPUs of one box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank predicts its slice of one sweep from the virtual counterfactual video; the gathered movies must equal the
single-GPU result of rank 0 bit for bit; the sharded mean motion map (ONE all-reduce) must match the unsharded one."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from counterfactualworldmodels_b200 import dist as cdist  # noqa: E402
from counterfactualworldmodels_b200 import sampling, segmentation, synthetic, vmae  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, S = "base_8x8", 52  # not divisible by 8: shards differ in size
    model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg))
    synthetic.init_weights_(model, seed=0)
    G = segmentation.FlowGenerator(predictor=model.to(dev).eval(), imagenet_normalize_inputs=True, temporal_dim=2)
    T, h, w = model.mask_size
    rng = np.random.RandomState(0)
    active = torch.ones(1, T, h, w, S, dtype=torch.bool)
    passive = torch.zeros(1, T, h, w, S, dtype=torch.bool)
    passive[:, -1] = True
    for s in range(S):
        ay, ax = rng.randint(0, h - 1), rng.randint(0, w - 1)   # border patches can be shifted out -> rectangulariser
        active[0, -1, ay:ay + 2, ax:ax + 2, s] = False
        py, px = rng.randint(0, h - 1), rng.randint(0, w - 1)
        passive[0, -1, py:py + 2, px:px + 2, s] = False
    shifts = [[int(rng.randint(-3, 4)), int(rng.randint(-3, 4))] for _ in range(S)]
    x = synthetic.make_video(1, (224, 224), seed=3)[:, 0].to(dev)
    a, p = active.reshape(1, -1, S).to(dev), passive.reshape(1, -1, S).to(dev)
    torch.manual_seed(7)
    y = cdist.sharded_counterfactual_videos(G, x, a, p, shifts=shifts, sample_batch_size=16)
    ok = True
    if rank == 0:
        torch.manual_seed(7)
        ref = G.predict_counterfactual_videos(x, a, passive_patches=p, shifts=shifts, sample_batch_size=16)
        ok = ok and torch.equal(ref, y)
    # sharded mean motion map from synthetic per-rank flows
    g = torch.Generator(device="cpu").manual_seed(11)
    flows_all = (torch.randn(S, 2, 224, 224, generator=g) * 3).to(dev)
    lo, hi = cdist.shard_bounds(S, rank, world)
    G.set_input(x[:, None])
    local_view = flows_all[lo:hi].reshape(1, hi - lo, 2, 224, 224).permute(0, 2, 3, 4, 1)
    mm = G.compute_mean_motion_map(local_view, group=dist.group.WORLD if world > 1 else None, num_samples_total=S)
    full = sampling.motion_map_finalize(sampling.flow_magnitude_sum(flows_all.reshape(1, S, 2, 224, 224).permute(0, 2, 3, 4, 1)), S)
    err = float((mm - full).abs().max())
    ok = ok and err < 1e-5
    # the whole iteration sharded: predict -> RAFT flow (fp32, same seeded weights on every rank) -> filter -> ONE
    # all-reduce of the magnitude sums; against the same sweep run on rank 0 alone
    from counterfactualworldmodels_b200 import raft
    torch.manual_seed(0)
    rargs = raft.get_args("")
    rargs.multiframe, rargs.scale_inputs, rargs.output_dim = True, True, None
    G.flow_model = raft.RAFT(rargs).to(dev).eval().requires_grad_(False)
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(7)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    mm_sharded = cdist.sharded_counterfactual_motion_map(G, x, a, p, shifts=shifts, sample_batch_size=16, raft_iters=6)
    ev1.record()
    torch.cuda.synchronize()
    sweep_ms = ev0.elapsed_time(ev1)
    err_sweep = 0.0
    if rank == 0:
        flows = torch.cat([G.predict_flow(ref[i:i + 16], iters=6) for i in range(0, S, 16)], 0)
        G.set_input(x[:, None])
        mm_single = G.compute_mean_motion_map(G.filter_flow_samples(flows, a))
        err_sweep = float((mm_sharded - mm_single).abs().max())
        ok = ok and err_sweep < 1e-4 and bool(torch.isfinite(mm_sharded).all())
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "samples": S, "sharded_equals_single_gpu": bool(flag.item()),
                          "motion_map_max_abs_diff": err, "sharded_sweep_with_flow_motion_map_max_abs_diff": err_sweep,
                          "sharded_sweep_with_flow_ms": round(sweep_ms, 2)}))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
