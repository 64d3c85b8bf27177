#!/usr/bin/env python
"""bench.py -- counterfactual frames/s of the CWM VMAE hot path on B200 (contract: see the task prompt).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference algorithm on the host cores (CPU oracle)

A "step" is one pass of the hot path (a1..a12: normalise + gather, VMAE forward, scatter + unpatchify) over one
batch of synthetic counterfactual prompts.  Default workload = BASELINE.json configs[1]: ViT-base VMAE, 8x8 patches,
224 px, 2 frames, batch 64 motion counterfactuals per GPU (weak scaling: every rank runs its own batch of 64 and
only the predicted frames are gathered to rank 0 over NCCL at the end of each step).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# name -> (model config, per-GPU batch, visible 2x2 clumps in frame 1)
WORKLOADS = {
    "base_8x8_b64_counterfactual": ("base_8x8", 64, 1),      # BASELINE.json configs[1]
    "base_4x4_b32": ("base_4x4", 32, 2),                      # configs[2]
    "large_4x4_b32_movability": ("large_4x4", 32, 1),         # configs[3], per-GPU chunk of the 1024 sweep
    "imu400_base_4x4_b32": ("imu400_base_4x4", 32, 1),        # configs[4], per-GPU share of the 256 counterfactuals
}
DEFAULT_WORKLOAD = "base_8x8_b64_counterfactual"


def flops_per_frame_conjoined(n_vis, n_ctx_vis=25):
    """IMU-conditioned conjoined base 4x4 (BASELINE config 5): main stream as base 4x4 with a
    6336-token decoder, 12+4 context blocks on 25 / 50 tokens, 4 encoder + 4 decoder conjoining blocks
    (per block: qk|v 3C^2, projection C^2, MLP 2 x 2C^2 per token and stream, plus 8 N M C of cross attention)."""
    Ntot, P, D, Ce, Cd, Le, Ld = 6272, 64, 48, 768, 384, 12, 4
    Cse, Csd, M, Pc = 384, 192, 25, 25
    Nd, Md = Ntot + P, M + Pc
    blocks = lambda L, N, C: L * (24 * N * C ** 2 + 4 * N ** 2 * C)
    f = 2 * Ntot * D * Ce + blocks(Le, n_vis, Ce) + 2 * n_vis * Ce * Cd + blocks(Ld, Nd, Cd) + 2 * (Nd - n_vis) * Cd * D
    f += 2 * M * 96 * Cse + blocks(12, n_ctx_vis, Cse) + 2 * n_ctx_vis * Cse * Csd + blocks(4, Md, Csd) + \
        2 * (Md - n_ctx_vis) * Csd * 96

    def cross(N, Mc, C, Cs):
        trg = 2 * N * C * (3 * C + C + 4 * C)
        src = 2 * Mc * (Cs * 3 * C + C * Cs + 4 * Cs * Cs)
        return trg + src + 8 * N * Mc * C
    return f + 4 * cross(n_vis, n_ctx_vis, Ce, Cse) + 4 * cross(Nd, Md, Cd, Csd)


def flops_per_frame(cfg_name, n_vis):
    """SURVEY.md section 8d / BASELINE.md section 3 (multiply-add = 2; LN/GELU/softmax excluded)."""
    if cfg_name == "imu400_base_4x4":
        return flops_per_frame_conjoined(n_vis)
    from counterfactualworldmodels_b200 import synthetic
    kw = synthetic.CONFIGS[cfg_name]
    T, h, w = synthetic.mask_size(cfg_name)
    Ntot = T * h * w
    Nmask = Ntot - n_vis
    D = 3 * kw["tubelet_size"] * kw["patch_size"][0] * kw["patch_size"][1]
    Ce, Le, Cd, Ld = kw["encoder_embed_dim"], kw["encoder_depth"], kw["decoder_embed_dim"], kw["decoder_depth"]
    return (2 * Ntot * D * Ce + Le * (24 * n_vis * Ce ** 2 + 4 * n_vis ** 2 * Ce) + 2 * n_vis * Ce * Cd +
            Ld * (24 * Ntot * Cd ** 2 + 4 * Ntot ** 2 * Cd) + 2 * Nmask * Cd * D)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"],
                    tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


def bind_to_gpu_numa_node(device_index):
    """One process per GPU: keep the rank's host threads (and therefore its pinned staging buffers, first-touch) on the
    CPUs NVML reports as local to its GPU, so that the e2e host<->device copies of different ranks do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.proc = None
        self.lines = []
        self.idx = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_frames_per_s(cfg_name, n_clumps, sample_frames, repeats, threads):
    """The reference algorithm (CPU oracle port, fp32 eager) on the host cores: frames/s over `repeats` passes of a
    `sample_frames`-frame sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vmae_oracle as oracle
    from counterfactualworldmodels_b200 import synthetic, vmae
    torch.set_num_threads(threads)
    if cfg_name == "imu400_base_4x4":
        import conjoined_oracle as co
        from counterfactualworldmodels_b200 import conjoined_vmae
        m = conjoined_vmae.imu400_base_4x4patch_2frames_1tube()
        synthetic.init_weights_(m, seed=0, style="reference")
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        x = synthetic.make_video(sample_frames, (224, 224), seed=0)
        mask = synthetic.make_mask(sample_frames, m.mask_size, num_clumps=n_clumps, seed=0)
        imu = synthetic.make_imu(1, 400, seed=0).expand(sample_frames, -1, -1)
        mc = torch.zeros(sample_frames, 25, dtype=torch.bool)
        ocfg = synthetic.conjoined_oracle_cfg(cfg_name)
        times = []
        with torch.no_grad():
            for _ in range(repeats):
                t0 = time.perf_counter()
                y = co.conjoined_forward(sd, oracle.preprocess(x), mask, imu[..., None, None], mc, ocfg, True, False)
                oracle.pred_patches_to_video(y[:, :-64], x, mask, (1, 4, 4))
                times.append(time.perf_counter() - t0)
        return times
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
    synthetic.init_weights_(m, seed=0, style="reference")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = synthetic.make_video(sample_frames, synthetic.image_hw(cfg_name), seed=0)
    mask = synthetic.make_mask(sample_frames, synthetic.mask_size(cfg_name), num_clumps=n_clumps, seed=0)
    ocfg = synthetic.oracle_cfg(cfg_name)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            oracle.predict(sd, x, mask, ocfg, frame=None)
            times.append(time.perf_counter() - t0)
    return times


def measure_cf_sweep(dev, workload, peaks, with_cpu):
    """One counterfactual sweep per step: S = per-GPU batch samples of one image, each with one active 2x2 clump, one
    passive 2x2 clump and a preset shift (ipynb:1726), through segmentation.FlowGenerator (fused construction)."""
    import numpy as np
    from counterfactualworldmodels_b200 import segmentation, synthetic, vmae
    cfg_name, S, _ = WORKLOADS[workload]
    if cfg_name == "imu400_base_4x4":
        return {"workload": "cf_sweep", "skipped": "conjoined predictors take materialised prompts"}
    model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
    synthetic.init_weights_(model, seed=0, style="reference")
    model = model.to(dev).eval()
    G = segmentation.FlowGenerator(predictor=model, imagenet_normalize_inputs=True, temporal_dim=2)
    T, h, w = model.mask_size
    rng = np.random.RandomState(0)
    preset = [[2, 0], [0, 2], [-2, 0], [0, -2], [2, 2], [-2, -2], [2, -2], [-2, 2]]
    active = torch.ones(1, T, h, w, S, dtype=torch.bool)
    passive = torch.zeros(1, T, h, w, S, dtype=torch.bool)
    passive[:, -1] = True
    for s_ in range(S):
        ay, ax = 2 * rng.randint(2, h // 4), 2 * rng.randint(2, w // 4)                    # left half, away from borders
        py, px = 2 * rng.randint(2, h // 4), 2 * rng.randint(w // 4 + 2, w // 2 - 2)      # right half: never collides
        active[0, -1, ay:ay + 2, ax:ax + 2, s_] = False
        passive[0, -1, py:py + 2, px:px + 2, s_] = False
    shifts = [preset[s_ % 8] for s_ in range(S)]
    x_host = synthetic.make_video(1, synthetic.image_hw(cfg_name), seed=7)[:, 0].pin_memory()   # one image [1,3,H,W]
    a_host, p_host = active.reshape(1, -1, S).pin_memory(), passive.reshape(1, -1, S).pin_memory()

    def step():
        x = x_host.to(dev, non_blocking=True)
        a, p_ = a_host.to(dev, non_blocking=True), p_host.to(dev, non_blocking=True)
        return G.predict_counterfactual_videos(x, a, passive_patches=p_, shifts=shifts, sample_batch_size=S)

    for _ in range(3):
        y = step()
    torch.cuda.synchronize()
    steps = 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n_vis = h * w + 8
    out = {"workload": f"cf_sweep_{cfg_name}_s{S}_fused", "value": round(S / ms * 1e3, 2), "unit": "frames/s",
           "ms_per_step": round(ms, 3), "h2d_bytes_per_step": int(x_host.numel() * 4 + 2 * a_host.numel()),
           "tensor_frac_of_sustained_peak": round(S / ms * 1e3 * flops_per_frame(cfg_name, n_vis) / 1e12 /
                                                  peaks["tflops_sustained"], 4),
           "api": "segmentation.FlowGenerator.predict_counterfactual_videos(image, active, passive, shifts)"}
    # the same sweep continued through the flow network (raft.RAFT, RAFT-large shapes, random
    # init, mixed precision: cuDNN convolutions + the repo's correlation / lookup / fused recurrent-block kernels), the
    # flow-sample filter and the mean motion map.  Reported beside the headline, never part of it.
    try:
        from counterfactualworldmodels_b200 import raft
        torch.manual_seed(0)
        rargs = raft.get_args("")
        rargs.multiframe, rargs.scale_inputs, rargs.output_dim, rargs.mixed_precision = True, True, None, True
        G.flow_model = raft.RAFT(rargs).to(dev).eval().requires_grad_(False)

        def flow_step():
            x = x_host.to(dev, non_blocking=True)
            a, p_ = a_host.to(dev, non_blocking=True), p_host.to(dev, non_blocking=True)
            G.set_input(x)
            ys, flows = G.predict_counterfactual_videos_and_flows(x, a, passive_patches=p_, shifts=shifts,
                                                                  sample_batch_size=S, raft_iters=24)
            return G.compute_mean_motion_map(G.filter_flow_samples(flows, a))

        for _ in range(2):
            flow_step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            mm = flow_step()
        e1.record()
        torch.cuda.synchronize()
        ms_f = e0.elapsed_time(e1) / 3
        out["with_flow_and_statistics"] = {
            "value": round(S / ms_f * 1e3, 2), "unit": "counterfactuals/s", "ms_per_step": round(ms_f, 3),
            "motion_map_finite": bool(torch.isfinite(mm).all()),
            "api": "FlowGenerator.predict_counterfactual_videos_and_flows -> filter_flow_samples -> compute_mean_motion_map",
            "flow_model": "raft.RAFT (RAFT-large, 24 iterations, mixed precision, fused recurrent block, shared frame 0)"}
        G.flow_model = None
    except Exception as exc:  # the extra line must never take the headline down
        out["with_flow_and_statistics"] = {"error": repr(exc)[:300]}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import counterfactual_oracle as cfo
        n = min(S, 16)
        x2 = x_host.numpy()[:, None].repeat(2, axis=1)
        t0 = time.perf_counter()
        cfo.create_motion_counterfactuals(x2, p_host.numpy()[..., :n], a_host.numpy()[..., :n], shifts[:n],
                                          tuple(model.patch_size), frame=1, fix_passive=True)
        out["cpu_construction_ms_per_sample"] = round((time.perf_counter() - t0) / n * 1e3, 3)
        out["cpu_construction_kind"] = "port (oracle/counterfactual_oracle.py, numpy, 1 thread), 16-sample bound"
    del model, G
    torch.cuda.empty_cache()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name, B, n_clumps = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sample = args.ref_sample
    times = cpu_oracle_frames_per_s(cfg_name, n_clumps, sample, args.warmup + args.steps, threads)
    timed = times[args.warmup:]
    total = sum(timed)
    value = sample * len(timed) / total
    line = {
        "impl": "reference", "metric": "counterfactual frames/sec", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "model": cfg_name, "per_gpu_batch": B,
                   "sample": f"{sample} frames per step (bounded sample of the batch-{B} workload)"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} frames x {len(timed)} steps, CPU oracle (oracle/vmae_oracle.py, the "
                                   "reference algorithm in fp32 eager torch) on the host cores"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--ref-sample", type=int, default=2, help="frames per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--also", default="large_4x4_b32_movability,imu400_base_4x4_b32",
                    help="extra workloads measured briefly (comma separated, '' to skip)")
    ap.add_argument("--no-cf-sweep", dest="cf_sweep", action="store_false",
                    help="skip the fused counterfactual-sweep line in `also`")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    from counterfactualworldmodels_b200 import _lib, prediction, synthetic, vmae

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().cwm_device_check())
    if os.environ.get("CWM_ATTN_POLY"):  # tuning hook: eighths of the softmax exponentials evaluated on the FMA pipe
        _lib.load().cwm_debug_attention_poly.argtypes = [ctypes.c_int]
        _lib.load().cwm_debug_attention_poly(int(os.environ["CWM_ATTN_POLY"]))
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(workload, steps, warmup, with_e2e, with_profile, sample_clocks):
        cfg_name, B, n_clumps = WORKLOADS[workload]
        pred_kwargs = {}
        if cfg_name == "imu400_base_4x4":
            from counterfactualworldmodels_b200 import conjoined_vmae
            model = conjoined_vmae.imu400_base_4x4patch_2frames_1tube()
            hw = (224, 224)
            # one IMU context per image (segmentation.py:939-963), fully visible, tiled over the counterfactual samples
            pred_kwargs = dict(x_context=synthetic.make_imu(1, 400, seed=rank).expand(B, -1, -1).contiguous().to(dev),
                               mask_context=torch.zeros(B, 25, dtype=torch.bool, device=dev))
        else:
            model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
            hw = synthetic.image_hw(cfg_name)
        synthetic.init_weights_(model, seed=0, style="reference")
        model = model.to(dev).eval()
        G = prediction.PredictorBasedGenerator(predictor=model, imagenet_normalize_inputs=True, temporal_dim=2)
        n_rot = 3
        xs_host = [synthetic.make_video(B, hw, seed=100 * rank + i).pin_memory() for i in range(n_rot)]
        ms_host = [synthetic.make_mask(B, model.mask_size, num_clumps=n_clumps, seed=100 * rank + i).pin_memory()
                   for i in range(n_rot)]
        xs_dev = [x.to(dev) for x in xs_host]
        ms_dev = [m.to(dev) for m in ms_host]
        n_vis = int((~ms_host[0][0]).sum())
        gather_buf = [torch.empty(B, 1, *xs_host[0].shape[2:], device=dev) for _ in range(world)] \
            if (world > 1 and rank == 0) else None

        # The only exchange of the sharded sweep: predicted frames -> rank 0 (NCCL gather, 38.5 MB per rank and step).
        # It is issued asynchronously (NCCL's own stream, ordered after the frame copy) so that it overlaps the next
        # step's kernels; the previous gather is waited for -- on the device -- right before the next one is issued
        # (rank 0 reuses its receive buffers) and once more before the timed region closes.
        pending = []

        def gather_frames(video):
            frame = video[:, -1:].contiguous()
            drain_gather()
            work = dist.gather(frame, gather_buf, dst=0, async_op=True)
            pending.append((work, frame))

        def drain_gather():
            while pending:
                work, _ = pending.pop()
                work.wait()  # makes the current stream wait for the collective (no host synchronisation)

        def step(i, x, m):
            video = G.predict(x, m, frame=None, **pred_kwargs)
            if world > 1:
                gather_frames(video)
            return video

        for i in range(warmup):
            step(i, xs_dev[i % n_rot], ms_dev[i % n_rot])
        # + unpatchify_scatter + the compaction `predict` runs to read the per-row visible counts (plain VMAE only)
        launches_per_step = model.last_forward_launches + (2 if cfg_name != "imu400_base_4x4" else 1)
        barrier()
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()

        def timed_pass():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step(i, xs_dev[i % n_rot], ms_dev[i % n_rot])
            drain_gather()  # the last step's gather is inside the timed region
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        # three passes of exactly K steps each; the median pass is the reported one (a single pass is exposed to
        # one-off host / clock hiccups: one 2.5x outlier was seen in ~15 runs of the host-buffer leg)
        passes = [timed_pass() for _ in range(3 if with_profile else 1)]
        ms = sorted(passes)[len(passes) // 2]
        # Per-kernel device times: the same K steps once more with every launch bracketed by two CUDA events on the
        # launching stream (cwm_profile_begin/end).  The event records cost ~3 % of a step, so `value` comes from the
        # un-instrumented pass above and the instrumented pass is reported beside it (ms_per_step_profiled).
        prof, ms_prof = [], None
        if with_profile:
            _lib.profile_begin()
            ms_prof = timed_pass()
            prof = _lib.profile_end()
        clocks = sampler.stop() if sampler else None

        e2e = None
        if with_e2e:
            # the public host-buffer API (prediction.HostPipeline): pinned host inputs -> predict -> pinned host
            # outputs, every copy inside the timed region, copies of neighbouring steps overlapped with the kernels
            out_host = [torch.empty(xs_host[0].shape, dtype=torch.float32).pin_memory() for _ in range(2)]
            post = None
            if world > 1:
                post = gather_frames
            pipe = prediction.HostPipeline(G, tuple(xs_host[0].shape), ms_host[0].shape[1], device=dev, post=post,
                                           **pred_kwargs)

            def e2e_step(i):
                pipe.submit(xs_host[i % n_rot], ms_host[i % n_rot], out_host[i & 1], frame=None)

            for i in range(max(2, warmup // 2)):
                e2e_step(i)
            pipe.finish()
            barrier()
            # host-side copies are exposed to host jitter (one 2.5x outlier was seen in ~15 runs): three passes of
            # exactly K steps each, the median pass is reported and all three are listed
            runs = []
            for _ in range(3):
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for i in range(steps):
                    e2e_step(i)
                pipe.finish()
                drain_gather()
                f1.record()
                barrier()
                t2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
                if world > 1:
                    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
                runs.append(float(t2.item()))
            e2e_ms = sorted(runs)[1]
            e2e = {"value": world * B * steps / (e2e_ms * 1e-3), "unit": "frames/s",
                   "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                   "ms_per_step": e2e_ms / steps, "passes_ms_per_step": [round(r / steps, 3) for r in runs],
                   "api": "prediction.HostPipeline.submit(x_host, mask_host, out_host) per step (3 streams, 2 slots)"}
            # the pipelined result is bit-identical to a plain predict of the same batch
            chk = G.predict(xs_dev[(steps - 1) % n_rot], ms_dev[(steps - 1) % n_rot], frame=None, **pred_kwargs)
            torch.cuda.synchronize()
            assert torch.equal(out_host[(steps - 1) & 1], chk.cpu()), "HostPipeline output differs from predict"
        drain_gather()
        barrier()
        del model, G
        torch.cuda.empty_cache()
        return dict(cfg=cfg_name, B=B, n_vis=n_vis, ms=ms, passes=passes, ms_prof=ms_prof, steps=steps, prof=prof,
                    clocks=clocks, e2e=e2e,
                    launches_per_step=launches_per_step,
                    fps=world * B * steps / (ms * 1e-3), flops_frame=flops_per_frame(cfg_name, n_vis))

    r = measure(args.workload, args.steps, args.warmup, with_e2e=True, with_profile=True, sample_clocks=True)

    # ---- roofline of the dominant kernel class (device time measured live with CUDA events in the timed region)
    kernels = []
    for p in r["prof"]:
        per = p["ms"] / max(1, p["launches"])
        kernels.append({"name": p["name"], "launches": p["launches"], "ms_total": round(p["ms"], 3),
                        "share": round(p["ms"] / max(1e-9, sum(q["ms"] for q in r["prof"])), 4),
                        "tflops": round(p["flops"] / (p["ms"] * 1e-3) / 1e12, 1) if p["flops"] else None,
                        "gbs": round(p["bytes"] / (p["ms"] * 1e-3) / 1e9, 1), "avg_ms": round(per, 4)})
    roofline = None
    if r["prof"]:
        # "dominant kernel" = the kernel FUNCTION with the largest device time, as an ncu launch list groups it: the
        # epilogue-mode classes that run the same template instantiation are merged for this purpose
        # (gemm_f16_kernel<BN, f16-out>: qkv + fc1/GELU;  gemm_f16_kernel<BN, fp32-out>: proj / fc2 / head / embeds)
        merged = {}
        for p in r["prof"]:
            key = {"gemm_f16_out": "gemm_f16_kernel<f16-out>", "gemm_gelu_f16_out": "gemm_f16_kernel<f16-out>",
                   "gemm_residual_f32": "gemm_f16_kernel<fp32-out>", "gemm_f32_out": "gemm_f16_kernel<fp32-out>"}.get(
                       p["name"], p["name"])
            mm = merged.setdefault(key, {"name": key, "ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            for f in ("ms", "flops", "bytes", "launches"):
                mm[f] += p[f]
        top = max(merged.values(), key=lambda p: p["ms"])
        if top["flops"] > 0:
            achieved = top["flops"] / (top["ms"] * 1e-3) / 1e12
            roofline = {"kernel": top["name"], "bound": "tensor", "achieved": round(achieved, 1),
                        "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                        "frac": round(achieved / peaks["tflops_sustained"], 4), "traffic": None,
                        "launches": top["launches"], "avg_launch_ms": round(top["ms"] / max(1, top["launches"]), 4),
                        "share_of_step": round(top["ms"] / max(1e-9, sum(q["ms"] for q in r["prof"])), 4),
                        "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)"}
        else:
            achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
            roofline = {"kernel": top["name"], "bound": "hbm", "achieved": round(achieved, 1),
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4),
                        "traffic": None, "peak_source": f"{peaks['source']} HBM copy"}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                roofline["traffic"] = json.load(f).get(top["name"])

    also = []
    for w in [w for w in args.also.split(",") if w and w != args.workload]:
        ra = measure(w, max(2, min(args.steps, 4)), 3, with_e2e=False, with_profile=False, sample_clocks=False)
        also.append({"workload": w, "value": round(ra["fps"], 2), "unit": "frames/s",
                     "ms_per_step": round(ra["ms"] / ra["steps"], 3),
                     "tensor_frac_of_sustained_peak": round(ra["fps"] / world * ra["flops_frame"] / 1e12 /
                                                            peaks["tflops_sustained"], 4),
                     "gflop_per_frame": round(ra["flops_frame"] / 1e9, 1)})

    # the same workload driven from (image, patch descriptors, shifts) through the reference-facing
    # `FlowGenerator.predict_counterfactual_videos` -- masks built on device, the 64 prompts never materialised
    if world == 1 and args.cf_sweep:
        also.append(measure_cf_sweep(dev, args.workload, peaks, not args.no_cpu_baseline))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cfg_name, B, n_clumps = WORKLOADS[args.workload]
        times = cpu_oracle_frames_per_s(cfg_name, n_clumps, args.ref_sample, 4, threads)
        timed = times[1:]
        cpu_baseline = {"value": args.ref_sample * len(timed) / sum(timed), "unit": "frames/s", "cores": threads,
                        "kind": "port", "sample": f"{args.ref_sample} frames x {len(timed)} passes (+1 warm-up) of the "
                        f"same workload, CPU oracle (reference algorithm, fp32 eager torch, {threads} threads)"}

    if rank == 0:
        line = {
            "metric": "counterfactual frames/sec", "value": round(r["fps"], 2), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(r["ms"] / args.steps, 4),
            "passes_ms_per_step": [round(v / args.steps, 4) for v in r["passes"]],
            "ms_per_step_profiled": round(r["ms_prof"] / args.steps, 4) if r["ms_prof"] else None,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate/residual/softmax", "data": "synthetic",
            "config": {"workload": args.workload, "model": r["cfg"], "per_gpu_batch": r["B"],
                       "global_batch": r["B"] * world, "visible_tokens": r["n_vis"],
                       "parallelism": f"dp{world} replicas, samples sharded, NCCL gather of predicted frames",
                       "l2": "3 rotating 77 MB input batches; > 1 GB of workspace is rewritten every step (>> 126 MB L2)",
                       "gflop_per_frame": round(r["flops_frame"] / 1e9, 1),
                       "cpus_bound_per_rank": affinity},
            "tensor_frac_of_sustained_peak": round(r["fps"] / world * r["flops_frame"] / 1e12 /
                                                   peaks["tflops_sustained"], 4),
            "tensor_frac_of_burst_peak": round(r["fps"] / world * r["flops_frame"] / 1e12 / peaks["tflops_burst"], 4),
            "e2e": r["e2e"], "gpu_launches": int(r["launches_per_step"] * args.steps),
            "clocks": r["clocks"], "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels,
            "also": also,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
