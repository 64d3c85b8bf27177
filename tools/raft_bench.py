"""Timing of the RAFT correlation / upsampling kernels (first slice) on one B200 at the sweep
shape: 224 px frames -> 28x28 maps, D = 256 (RAFT-large), S samples per chunk.  Per-kernel device times come from the
library's CUDA-event profiler (events on the launching stream); buffers rotate so every pass reads / writes more than
the 126 MB L2.  The reference's own ATen ops (oracle/raft_oracle.py torch port, all host threads) are timed beside them
on a bounded sample.     python tools/raft_bench.py [S]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import raft_oracle as ro  # noqa: E402
from counterfactualworldmodels_b200 import _lib, raft  # noqa: E402


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def profiled(fn, iters=6, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    _lib.profile_begin()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return {e["name"]: dict(ms=e["ms"] / e["launches"] * (e["launches"] / iters), launches=e["launches"] // iters,
                            flops=e["flops"] / iters, bytes=e["bytes"] / iters) for e in _lib.profile_end()}


def e2e(S):
    """The whole flow network (raft.RAFT, RAFT-large shapes, random init) on S frame pairs of 224 px, 24 iterations:
    samples/s and how the step splits between this repo's kernels and torch's cuDNN convolutions."""
    dev = "cuda:0"
    torch.manual_seed(0)
    args = raft.get_args("")
    args.multiframe, args.scale_inputs, args.output_dim = True, True, None
    model = raft.RAFT(args).eval().requires_grad_(False).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(S, 2, 3, 224, 224, device=dev, generator=g)
    out = {"mode": "e2e", "S": S, "iters": 24}
    if len(sys.argv) > 3 and sys.argv[3] == "ncu":   # one f16 sweep call for a launch list
        model.args.mixed_precision = True
        x[:, 0] = x[:1, 0]
        model(x, shared_frame=0)
        torch.cuda.synchronize()
        return
    xs = x.clone()
    xs[:, 0] = xs[:1, 0]                                   # a sweep: frame 0 shared by all samples
    variants = [("fp32", False, False, False, None, None), ("tf32_convs", True, False, False, None, None),
                ("f16_autocast_convs", True, True, False, None, None),
                ("f16_shared_frame0", True, True, False, 0, None), ("fp32_shared_frame0", False, False, False, 0, None),
                ("f16_half_update_shared_frame0", True, True, True, 0, None),
                ("f16_fused_update_shared_frame0_cudnn_convs", True, True, "fused", 0, "cudnn"),
                ("f16_fused_update_shared_frame0", True, True, "fused", 0, "tcgen05"),
                ("f16_fused_update_cudnn_convs", True, True, "fused", None, "cudnn"),
                ("f16_fused_update", True, True, "fused", None, "tcgen05")]
    if os.environ.get("RAFT_BENCH_ONLY"):     # e.g. RAFT_BENCH_ONLY=fused: only the variants whose name contains it
        variants = [v for v in variants if os.environ["RAFT_BENCH_ONLY"] in v[0]]
    for name, tf32, amp, half_update, shared, conv in variants:
        if conv is not None:   # which convolutions the fused recurrent block uses: the repo's implicit GEMMs or cuDNN
            os.environ["CWM_RAFT_CONV"] = conv
            object.__setattr__(model, '_fused_ub', None)
        torch.backends.cudnn.allow_tf32 = tf32
        model.args.mixed_precision = amp
        model.args.half_update = bool(half_update)
        model.args.fused_update = half_update == "fused"
        inp = xs if shared is not None else x
        kw = {} if shared is None else {"shared_frame": shared}
        for _ in range(2):
            model(inp, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            model(inp, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        _lib.profile_begin()
        y = model(inp, **kw)
        torch.cuda.synchronize()
        mine = sum(e["ms"] for e in _lib.profile_end())
        out[name] = {"ms_per_call": ms, "flow_samples_per_s": S / ms * 1e3, "ms_in_libcwm_kernels": mine,
                     "share_in_libcwm_kernels": mine / ms}
        if shared is not None:  # the broadcast path against the plain one on the same input
            model.args.half_update = False
            ref = model(inp)
            out[name]["max_abs_diff_vs_unshared"] = float((y - ref).abs().max())
            out[name]["flow_scale"] = float(ref.abs().max())
    print(json.dumps(out))


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    if len(sys.argv) > 2 and sys.argv[2] == "e2e":
        return e2e(S)
    dev = "cuda:0"
    D, H, W, L, r = 256, 28, 28, 4, 4
    hbm, src = peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    fmaps = [(torch.randn(S, D, H, W, device=dev, generator=g), torch.randn(S, D, H, W, device=dev, generator=g))
             for _ in range(2)]
    grid = raft.coords_grid(S, H, W, dev)
    coords = [grid + torch.randn(S, 2, H, W, device=dev, generator=g) * 2.5 for _ in range(2)]
    out = {"S": S, "shape": [D, H, W, L, r], "hbm_peak_GBps": hbm, "peak_source": src, "kernels": {}}
    i = [0]
    if len(sys.argv) > 2 and sys.argv[2] == "ncu":   # one launch of every kernel, for `ncu -k regex:raft_`
        block = raft.CorrBlock(*fmaps[0], num_levels=L, radius=r)
        block(coords[0])
        raft.upsample_flow(torch.randn(S, 2, H, W, device=dev), torch.randn(S, 576, H, W, device=dev))
        torch.cuda.synchronize()
        return

    def build():
        i[0] += 1
        return raft.CorrBlock(*fmaps[i[0] % 2], num_levels=L, radius=r)

    for name, e in profiled(build).items():
        out["kernels"][name] = e
    blocks = [build(), build()]   # 2 x 3.25 MB x S of pyramid: alternate so the lookups do not find their level in L2

    def lookup():
        i[0] += 1
        return blocks[i[0] % 2](coords[i[0] % 2])

    for name, e in profiled(lookup).items():
        out["kernels"][name] = e
    del blocks
    ups = [(torch.randn(S, 2, H, W, device=dev, generator=g) * 3, torch.randn(S, 576, H, W, device=dev, generator=g) * 2)
           for _ in range(2)]

    def upsample():
        i[0] += 1
        return raft.upsample_flow(*ups[i[0] % 2])

    for name, e in profiled(upsample).items():
        out["kernels"][name] = e
    for e in out["kernels"].values():
        e["GBps"] = e["bytes"] / e["ms"] / 1e6
        e["frac_of_hbm_peak"] = e["GBps"] / hbm
        if e["flops"]:
            e["TFLOPs"] = e["flops"] / e["ms"] / 1e9
    k = out["kernels"]
    pyr_ms = k["raft_corr_volume"]["ms"] + k["raft_corr_pool"]["ms"]
    out["per_sample_us"] = {"pyramid": pyr_ms / S * 1e3, "lookup_x24": 24 * k["raft_corr_lookup"]["ms"] / S * 1e3,
                            "upsample": k["raft_upsample"]["ms"] / S * 1e3}
    out["per_sample_us"]["total_24_iters"] = sum(out["per_sample_us"].values())
    # CPU: the reference's ATen ops, all host threads, bounded sample (n samples, 2 lookups scaled to 24)
    n = min(S, 8)
    f1, f2 = fmaps[0][0][:n].cpu(), fmaps[0][1][:n].cpu()
    cl = [c[:n].cpu() for c in coords]
    fl, mk = ups[0][0][:n].cpu(), ups[0][1][:n].cpu()
    ro.torch_corr_block(f1[:1], f2[:1], [cl[0][:1]], L, r)
    t0 = time.perf_counter()
    ro.torch_corr_block(f1, f2, [], L, r)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    ro.torch_corr_block(f1, f2, cl, L, r)
    t_look = (time.perf_counter() - t0 - t_build) / len(cl)
    t0 = time.perf_counter()
    ro.torch_upsample_flow(fl, mk)
    t_up = time.perf_counter() - t0
    out["cpu_reference_ops"] = {"threads": torch.get_num_threads(), "samples": n,
                                "per_sample_us": {"pyramid": t_build / n * 1e6, "lookup_x24": 24 * t_look / n * 1e6,
                                                  "upsample": t_up / n * 1e6}}
    out["cpu_reference_ops"]["per_sample_us"]["total_24_iters"] = sum(out["cpu_reference_ops"]["per_sample_us"].values())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
