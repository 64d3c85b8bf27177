#!/usr/bin/env python
"""bench.py -- counterfactual frames/s of the CWM VMAE hot path on B200 (contract: see the task prompt).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the REAL reference (baseline/_ref) on the host cores

Default workload = the north star (BASELINE.json configs[3]): ViT-large VMAE, 4x4 patches, 224 px, 2 frames, ONE
1024-counterfactual movability sweep per step.  A step starts from (image, active-patch descriptors, passive-patch
descriptors, shifts) and ends with the 1024 predicted counterfactual frames on rank 0: masks built on device
(rank 0) and broadcast, the prompts never materialised, every rank predicts its contiguous share in chunks of 32
through `dist.sharded_counterfactual_videos`, one NCCL gather of the predicted frames.  The sweep size is fixed, so
`--gpus N` is STRONG scaling.  The other workloads (`--workload`, and the brief `also` entries) are per-GPU batches of
materialised prompts through `PredictorBasedGenerator.predict` (weak scaling), e.g. BASELINE configs[1].
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

# name -> (model config, samples per step [per GPU for "batch" workloads, per job for "sweep" workloads],
#          visible 2x2 clumps in frame 1, kind)
WORKLOADS = {
    "large_4x4_movability_sweep1024": ("large_4x4", 1024, 2, "sweep"),   # BASELINE.json configs[3] = the north star
    "base_8x8_sweep1024": ("base_8x8", 1024, 2, "sweep"),
    "base_8x8_b64_counterfactual": ("base_8x8", 64, 1, "batch"),         # configs[1]
    "base_4x4_b32": ("base_4x4", 32, 2, "batch"),                         # configs[2]
    "large_4x4_b32_movability": ("large_4x4", 32, 1, "batch"),            # one per-GPU chunk of configs[3]
    "imu400_base_4x4_b32": ("imu400_base_4x4", 32, 1, "batch"),           # configs[4], per-GPU share of 256
}
DEFAULT_WORKLOAD = "large_4x4_movability_sweep1024"
SWEEP_CHUNK = {"large_4x4": 32, "base_4x4": 32, "base_8x8": 64}
METRIC = "counterfactual frames/sec"


def workload_config(workload, world):
    """The `config` object of the JSON line -- the same for both arms, so the driver can tell they ran the same job."""
    cfg_name, n, n_clumps, kind = WORKLOADS[workload]
    from counterfactualworldmodels_b200 import synthetic
    T, h, w = synthetic.mask_size(cfg_name) if cfg_name != "imu400_base_4x4" else (2, 56, 56)
    n_vis = h * w + 4 * n_clumps
    if kind == "sweep":
        return {"workload": workload, "model": cfg_name, "counterfactuals_per_step": n,
                "chunk": SWEEP_CHUNK[cfg_name], "visible_tokens": n_vis, "tokens": T * h * w,
                "gflop_per_frame": round(flops_per_frame(cfg_name, n_vis) / 1e9, 1),
                "inputs": "one image + per-sample (active 2x2 clump, passive 2x2 clump, preset shift) descriptors",
                "l2": "every 32-sample chunk rewrites > 3 GB of workspace (>> 126 MB L2)"}
    return {"workload": workload, "model": cfg_name, "per_gpu_batch": n, "visible_tokens": n_vis,
            "tokens": T * h * w, "gflop_per_frame": round(flops_per_frame(cfg_name, n_vis) / 1e9, 1),
            "l2": "3 rotating input batches; > 1 GB of workspace is rewritten every step (>> 126 MB L2)"}


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


def sweep_descriptors(mask_size, S, seed=0):
    """S motion counterfactuals of one image: per sample one active 2x2 clump (left half) that moves by a preset shift
    (ipynb:1726) and one passive 2x2 clump (right half) that stays -> n visible frame-1 patches = 8."""
    import numpy as np
    T, h, w = mask_size
    rng = np.random.RandomState(seed)
    preset = [[2, 0], [0, 2], [-2, 0], [0, -2], [2, 2], [-2, -2], [2, -2], [-2, 2]]
    active = torch.ones(1, T, h, w, S, dtype=torch.bool)
    passive = torch.zeros(1, T, h, w, S, dtype=torch.bool)
    passive[:, -1] = True
    for s_ in range(S):
        ay, ax = 2 * rng.randint(2, h // 4), 2 * rng.randint(2, w // 4)                    # left half, away from borders
        py, px = 2 * rng.randint(2, h // 4), 2 * rng.randint(w // 4 + 2, w // 2 - 2)      # right half: never collides
        active[0, -1, ay:ay + 2, ax:ax + 2, s_] = False
        passive[0, -1, py:py + 2, px:px + 2, s_] = False
    return active.reshape(1, -1, S), passive.reshape(1, -1, S), [preset[s_ % 8] for s_ in range(S)]


def cpu_reference_times(workload, sample_frames, repeats, threads):
    """The reference algorithm on the host cores, `repeats` passes over a `sample_frames`-frame sample of the workload.
    -> (kind, seconds per pass).  kind = "reference": the REAL, unmodified reference package staged under
    baseline/_ref (baseline/stage_reference.py; `cwm.models.prediction.PredictorBasedGenerator.predict` /
    `cwm.models.segmentation.FlowGenerator.predict_counterfactual_videos_and_flows`' video half, fp32 eager), or
    kind = "port": the CPU oracle restatement when the staged copy is missing."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    import vmae_oracle as oracle
    from counterfactualworldmodels_b200 import synthetic, vmae
    cfg_name, _, n_clumps, kind = WORKLOADS[workload]
    torch.set_num_threads(threads)
    times = []
    if cfg_name == "imu400_base_4x4":   # config 5: the port (its reference needs a RAFT checkpoint file on disk)
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
        with torch.no_grad():
            for _ in range(repeats):
                t0 = time.perf_counter()
                y = co.conjoined_forward(sd, oracle.preprocess(x), mask, imu[..., None, None], mc, ocfg, True, False)
                oracle.pred_patches_to_video(y[:, :-64], x, mask, (1, 4, 4))
                times.append(time.perf_counter() - t0)
        return "port", times
    hw = synthetic.image_hw(cfg_name)
    if ref_loader.available():
        import contextlib
        import io
        ref_vmae, ref_pred = ref_loader.import_reference()
        import cwm.models.segmentation as ref_seg
        with contextlib.redirect_stdout(io.StringIO()):     # the reference prints; stdout carries one JSON line
            torch.manual_seed(0)
            m = ref_vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name)).eval().requires_grad_(False)
            synthetic.init_weights_(m, seed=0, style="reference")
            if kind == "sweep":
                # the reference's own sweep: per-sample prompt construction on the host + chunked prediction
                # (segmentation.py:279-430); its flow half is not part of this metric
                G = ref_seg.FlowGenerator(predictor=m, flow_model=torch.nn.Identity(), imagenet_normalize_inputs=True,
                                          temporal_dim=2)
                x = synthetic.make_video(1, hw, seed=7)[:, 0]
                a, p_, shifts = sweep_descriptors(m.mask_size, sample_frames, seed=0)
                with torch.no_grad():
                    for _ in range(repeats):
                        t0 = time.perf_counter()
                        G.set_input(x[:, None].expand(-1, 2, -1, -1, -1))
                        G.reset_shifts()
                        G.shifter.set_shapes(G.x, mask=a[..., 0])
                        G.shifter.set_num_shifts(sample_frames)
                        xm, mm = G.create_motion_counterfactuals(G.x, masks=p_, active_patches=a, shifts=shifts,
                                                                 num_samples=sample_frames, fix_passive=True)
                        G.batch_predict_per_sample(xm, masks=mm, frame=None, batch_size=sample_frames, sample_dim=0)
                        times.append(time.perf_counter() - t0)
            else:
                G = ref_pred.PredictorBasedGenerator(predictor=m, imagenet_normalize_inputs=True, temporal_dim=2)
                x = synthetic.make_video(sample_frames, hw, seed=0)
                mask = synthetic.make_mask(sample_frames, m.mask_size, num_clumps=n_clumps, seed=0)
                with torch.no_grad():
                    for _ in range(repeats):
                        t0 = time.perf_counter()
                        G.predict(x, mask, frame=None)
                        times.append(time.perf_counter() - t0)
        return "reference", times
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
    synthetic.init_weights_(m, seed=0, style="reference")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = synthetic.make_video(sample_frames, hw, seed=0)
    mask = synthetic.make_mask(sample_frames, synthetic.mask_size(cfg_name), num_clumps=n_clumps, seed=0)
    ocfg = synthetic.oracle_cfg(cfg_name)
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            oracle.predict(sd, x, mask, ocfg, frame=None)
            times.append(time.perf_counter() - t0)
    return "port", times


def cpu_kind_text(kind, threads):
    if kind == "reference":
        return (f"the REAL reference (unmodified cwm package staged under baseline/_ref, fp32 eager, {threads} threads; "
                "timm / kornia / matplotlib import stubs)")
    return f"CPU oracle port (oracle/vmae_oracle.py, the reference algorithm in fp32 eager torch, {threads} threads)"


def measure_flow_sweep(dev, cfg_name, S, peaks, with_cpu):
    """SURVEY 8(f) ranks 1-3 beside the headline (never part of it): S counterfactuals of one image from
    (image, active, passive, shifts) through segmentation.FlowGenerator -- fused construction + VMAE prediction, then
    the flow network (raft.RAFT, RAFT-large shapes, random init, mixed precision), the flow-sample filter and the mean
    motion map."""
    from counterfactualworldmodels_b200 import segmentation, synthetic, vmae
    model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
    synthetic.init_weights_(model, seed=0, style="reference")
    model = model.to(dev).eval()
    G = segmentation.FlowGenerator(predictor=model, imagenet_normalize_inputs=True, temporal_dim=2)
    T, h, w = model.mask_size
    active, passive, shifts = sweep_descriptors(model.mask_size, S, seed=0)
    x_host = synthetic.make_video(1, synthetic.image_hw(cfg_name), seed=7)[:, 0].pin_memory()   # one image [1,3,H,W]
    a_host, p_host = active.pin_memory(), passive.pin_memory()

    def step():
        x = x_host.to(dev, non_blocking=True)
        a, p_ = a_host.to(dev, non_blocking=True), p_host.to(dev, non_blocking=True)
        return G.predict_counterfactual_videos(x, a, passive_patches=p_, shifts=shifts, sample_batch_size=S)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    steps = 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n_vis = h * w + 8
    out = {"workload": f"cf_sweep_{cfg_name}_s{S}_fused", "value": round(S / ms * 1e3, 2), "unit": "frames/s",
           "ms_per_step": round(ms, 3), "h2d_bytes_per_step": int(x_host.numel() * 4 + 2 * a_host.numel()),
           "tensor_frac_of_sustained_peak": round(S / ms * 1e3 * flops_per_frame(cfg_name, n_vis) / 1e12 /
                                                  peaks["tflops_sustained"], 4),
           "api": "segmentation.FlowGenerator.predict_counterfactual_videos(image, active, passive, shifts)"}
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


def run_reference_arm(args, out):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores, same workload,
    metric and unit; every step is a bounded sample (--ref-sample frames) of the workload.  Rank 0 alone runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_name, n, _, kind_w = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sample = args.ref_sample if args.ref_sample > 0 else (1 if "4x4" in cfg_name else 2)
    kind, times = cpu_reference_times(args.workload, sample, args.warmup + args.steps, threads)
    timed = times[args.warmup:]
    total = sum(timed)
    value = sample * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed),
        "higher_is_better": True, "scaling": "strong" if kind_w == "sweep" else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} counterfactual frame(s) per step x {len(timed)} timed steps "
                                   f"(+{args.warmup} warm-up) -- a bounded sample of the "
                                   f"{n}-{'counterfactual sweep' if kind_w == 'sweep' else 'sample batch'}; "
                                   + cpu_kind_text(kind, threads)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.emit(line)


class OneLineStdout:
    """stdout carries exactly ONE JSON line: everything else any library prints on fd 1 (NCCL's version banner, the
    reference's own progress prints) is diverted to stderr for the duration of the run."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, obj):
        sys.stdout.flush()
        os.write(self.real, (json.dumps(obj) + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


def main():
    with OneLineStdout() as out:
        _main(out)


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="frames per step of the CPU reference arm / baseline leg (0 = 1 for 4x4-patch models, else 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-operand-modes", action="store_true",
                    help="skip the f16 / bf16 operand-mode parity report (two short subprocesses after the timed runs)")
    ap.add_argument("--also", default="base_8x8_b64_counterfactual,imu400_base_4x4_b32",
                    help="extra workloads measured briefly (comma separated, '' to skip)")
    ap.add_argument("--no-cf-sweep", dest="cf_sweep", action="store_false",
                    help="skip the sweep-with-flow line in `also`")
    ap.add_argument("--profile-steps", type=int, default=4,
                    help="steps of the per-kernel CUDA-event pass of a sweep workload (a sweep is 32 chunks)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    if args.impl == "reference":
        run_reference_arm(args, out)
        return

    import torch.distributed as dist
    from counterfactualworldmodels_b200 import _lib, prediction, segmentation, synthetic, vmae
    from counterfactualworldmodels_b200 import dist as cwm_dist

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
    lib = _lib.load()
    _lib.check(lib.cwm_device_check())
    if os.environ.get("CWM_ATTN_POLY"):  # tuning hook: eighths of the softmax exponentials evaluated on the FMA pipe
        lib.cwm_debug_attention_poly(int(os.environ["CWM_ATTN_POLY"]))
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def measure(workload, steps, warmup, with_e2e, with_profile, sample_clocks):
        cfg_name, B, n_clumps, _ = WORKLOADS[workload]
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
        counted = {"launches": 0}
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
        barrier()
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()

        def timed_pass():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = lib.cwm_total_launches()
            e0.record()
            for i in range(steps):
                step(i, xs_dev[i % n_rot], ms_dev[i % n_rot])
            drain_gather()  # the last step's gather is inside the timed region
            e1.record()
            counted["launches"] = lib.cwm_total_launches() - n0
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
                    launches=counted["launches"] * world,
                    fps=world * B * steps / (ms * 1e-3), flops_frame=flops_per_frame(cfg_name, n_vis))


    def measure_sweep(workload, steps, warmup, profile_steps):
        """One counterfactual sweep per step, the S samples sharded over the ranks (strong scaling)."""
        cfg_name, S, _, _ = WORKLOADS[workload]
        chunk = SWEEP_CHUNK[cfg_name]
        model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg_name))
        synthetic.init_weights_(model, seed=0, style="reference")
        model = model.to(dev).eval()
        G = segmentation.FlowGenerator(predictor=model, imagenet_normalize_inputs=True, temporal_dim=2)
        T, h, w = model.mask_size
        hw = synthetic.image_hw(cfg_name)
        n_vis = h * w + 8
        n_rot = 3                                    # three different images / descriptor sets, rotated over the steps
        inputs_host = []
        for i in range(n_rot):
            a, p_, shifts = sweep_descriptors(model.mask_size, S, seed=i)
            inputs_host.append((synthetic.make_video(1, hw, seed=7 + i)[:, 0].pin_memory(), a.pin_memory(),
                                p_.pin_memory(), shifts))
        inputs_dev = [(x.to(dev), a.to(dev), p_.to(dev), sh) for x, a, p_, sh in inputs_host]

        def sweep(x, a, p_, shifts):
            # -> [S, 1, 3, H, W] predicted counterfactual frames on rank 0 (None elsewhere)
            return cwm_dist.sharded_counterfactual_videos(G, x, a, passive_patches=p_, shifts=shifts,
                                                          sample_batch_size=chunk, dst=0, predict_frame=-1)

        def timed_pass(step_fn, n_steps, finish=None):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            n0 = lib.cwm_total_launches()
            e0.record()
            for i in range(n_steps):
                step_fn(i)
            if finish is not None:
                finish()
            e1.record()
            n1 = lib.cwm_total_launches()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)), sum_over_ranks(n1 - n0)

        def resident_step(i):
            return sweep(*inputs_dev[i % n_rot])

        for i in range(warmup):
            resident_step(i)
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        ms, launches = timed_pass(resident_step, steps)
        clocks = sampler.stop() if sampler else None

        # per-kernel device times: `profile_steps` more sweeps with every launch bracketed by two CUDA events on the
        # launching stream (~3 % overhead, so `value` comes from the un-instrumented pass above)
        _lib.profile_begin()
        ms_prof, _ = timed_pass(resident_step, profile_steps)
        prof = _lib.profile_end()

        # ---- e2e: the same sweep through the same public call with HOST buffers: the image and the patch descriptors
        # start in pinned host memory on every rank, the predicted frames end in pinned host memory on rank 0; every
        # copy is inside the timed region.  The device->host copy of sweep i overlaps sweep i+1 (side stream, two host
        # buffers); the region closes only after the last copy has landed.
        frames_shape = (S, 1, 3) + tuple(hw)
        out_host = [torch.empty(frames_shape, dtype=torch.float32).pin_memory() for _ in range(2)] if rank == 0 else None
        d2h = torch.cuda.Stream(device=dev)
        stored = [torch.cuda.Event(), torch.cuda.Event()]
        keep = [None, None]

        def e2e_step(i):
            x, a, p_, shifts = inputs_host[i % n_rot]
            y = sweep(x.to(dev, non_blocking=True), a.to(dev, non_blocking=True), p_.to(dev, non_blocking=True), shifts)
            if rank == 0:
                slot = i & 1
                done = torch.cuda.Event()
                done.record()
                d2h.wait_event(done)
                with torch.cuda.stream(d2h):
                    out_host[slot].copy_(y, non_blocking=True)
                    stored[slot].record(d2h)
                y.record_stream(d2h)
                keep[slot] = y
            return y

        def e2e_finish():
            cur = torch.cuda.current_stream(dev)
            for ev in stored:
                cur.wait_event(ev)

        for i in range(2):
            e2e_step(i)
        e2e_finish()
        e2e_ms, _ = timed_pass(e2e_step, steps, finish=e2e_finish)
        x0, a0, p0, _ = inputs_host[0]
        h2d = world * (x0.numel() * 4 + a0.numel() + p0.numel())
        e2e = {"value": S * steps / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(S * 3 * hw[0] * hw[1] * 4), "ms_per_step": e2e_ms / steps,
               "api": "dist.sharded_counterfactual_videos(FlowGenerator, image, active, passive, shifts, "
                      f"sample_batch_size={chunk}, dst=0, predict_frame=-1) from pinned host buffers on every rank; the "
                      "predicted frames are copied to pinned host memory on rank 0 (overlapping the next sweep)"}
        if rank == 0:   # the host copy is what the resident sweep returns
            chk = resident_step(steps - 1)
            torch.cuda.synchronize()
            assert torch.equal(out_host[(steps - 1) & 1], chk.cpu()), "e2e output differs from the resident sweep"
        else:
            resident_step(steps - 1)
        barrier()
        del model, G
        torch.cuda.empty_cache()
        return dict(cfg=cfg_name, B=S, n_vis=n_vis, ms=ms, passes=[ms], ms_prof=ms_prof, steps=steps,
                    prof_steps=profile_steps, prof=prof, clocks=clocks, e2e=e2e, launches=launches,
                    fps=S * steps / (ms * 1e-3), flops_frame=flops_per_frame(cfg_name, n_vis))

    sweep_workload = WORKLOADS[args.workload][3] == "sweep"
    if sweep_workload:
        r = measure_sweep(args.workload, args.steps, args.warmup, max(1, min(args.profile_steps, args.steps)))
    else:
        r = measure(args.workload, args.steps, args.warmup, with_e2e=True, with_profile=True, sample_clocks=True)
        r["prof_steps"] = args.steps

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
        total_ms = max(1e-9, sum(q["ms"] for q in r["prof"]))
        if top["flops"] > 0:
            achieved = top["flops"] / (top["ms"] * 1e-3) / 1e12
            roofline = {"kernel": top["name"], "bound": "tensor", "achieved": round(achieved, 1),
                        "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                        "frac": round(achieved / peaks["tflops_sustained"], 4), "traffic": None,
                        "launches": top["launches"], "avg_launch_ms": round(top["ms"] / max(1, top["launches"]), 4),
                        "share_of_step": round(top["ms"] / total_ms, 4), "steps_profiled": r["prof_steps"],
                        "peak_source": f"{peaks['source']} bf16 sustained (kernel timed inside a long step)"}
        else:
            achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
            roofline = {"kernel": top["name"], "bound": "hbm", "achieved": round(achieved, 1),
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4),
                        "traffic": None, "peak_source": f"{peaks['source']} HBM copy"}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            roofline["traffic"] = tj.get(f"{top['name']}@{r['cfg']}", tj.get(top["name"]))
        # every kernel class against ITS roofline (tensor classes vs the sustained bf16 peak, the rest vs HBM copy)
        roofline["classes"] = [
            {"kernel": mm["name"], "share_of_step": round(mm["ms"] / total_ms, 4),
             **({"tflops": round(mm["flops"] / (mm["ms"] * 1e-3) / 1e12, 1),
                 "frac": round(mm["flops"] / (mm["ms"] * 1e-3) / 1e12 / peaks["tflops_sustained"], 4)}
                if mm["flops"] > 0 else
                {"gbs": round(mm["bytes"] / (mm["ms"] * 1e-3) / 1e9, 1),
                 "frac": round(mm["bytes"] / (mm["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)})}
            for mm in sorted(merged.values(), key=lambda q: -q["ms"])[:6]]

    also = []
    for w in [w for w in args.also.split(",") if w and w != args.workload]:
        if WORKLOADS[w][3] == "sweep":
            continue
        ra = measure(w, max(2, min(args.steps, 4)), 3, with_e2e=False, with_profile=False, sample_clocks=False)
        also.append({"workload": w, "value": round(ra["fps"], 2), "unit": "frames/s", "scaling": "weak",
                     "ms_per_step": round(ra["ms"] / ra["steps"], 3),
                     "tensor_frac_of_sustained_peak": round(ra["fps"] / world * ra["flops_frame"] / 1e12 /
                                                            peaks["tflops_sustained"], 4),
                     "gflop_per_frame": round(ra["flops_frame"] / 1e9, 1)})

    # a base-8x8 sweep continued through the flow network and the flow statistics
    if world == 1 and args.cf_sweep:
        also.append(measure_flow_sweep(dev, "base_8x8", 64, peaks, not args.no_cpu_baseline))

    # the bf16-operand build beside the f16 default: parity error of both on fixtures the REAL reference produced
    # (tools/dtype_error.py in a subprocess per build -- a process binds to one library at import time)
    operand_modes = None
    if rank == 0 and world == 1 and not args.no_operand_modes:
        operand_modes = {}
        for mode in ("f16", "bf16"):
            try:
                pr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dtype_error.py"), "base_8x8_b2_counterfactual",
                                     "large_4x4_b1_factual"], env=dict(os.environ, CWM_DTYPE=mode), capture_output=True,
                                    text=True, timeout=600)
                operand_modes[mode] = (json.loads(pr.stdout.strip().splitlines()[-1])["cases"] if pr.returncode == 0
                                       else {"error": pr.stderr[-300:]})
            except Exception as e:  # noqa: BLE001  (a report line, never fatal for the benchmark)
                operand_modes[mode] = {"error": repr(e)}
        operand_modes["note"] = ("max-abs / mean-abs of the predicted patches vs the real reference (fp32); bar 2e-2 / 2e-3. "
                                 "f16 is the shipped default; bf16 = the same kernels built with -DCWM_ACT_BF16 (CWM_DTYPE=bf16)")

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cfg_name = WORKLOADS[args.workload][0]
        sample = args.ref_sample if args.ref_sample > 0 else (1 if "4x4" in cfg_name else 2)
        big = "4x4" in cfg_name       # one large-4x4 reference forward is ~10-30 s of host time: a single timed pass
        kind, times = cpu_reference_times(args.workload, sample, 1 if big else 4, threads)
        timed = times if big else times[1:]
        cpu_baseline = {"value": sample * len(timed) / sum(timed), "unit": "frames/s", "cores": threads, "kind": kind,
                        "sample": f"{sample} counterfactual frame(s) x {len(timed)} pass(es)"
                                  f"{'' if big else ' (+1 warm-up)'} of the same workload; "
                                  + cpu_kind_text(kind, threads)}

    if rank == 0:
        per_gpu_fps = r["fps"] / world
        cfg = workload_config(args.workload, world)
        cfg.update({"global_batch": r["B"] if sweep_workload else r["B"] * world,
                    "parallelism": (f"dp{world} replicas, the sweep's samples sharded; masks broadcast from rank 0, "
                                    "one NCCL gather of the predicted frames" if sweep_workload else
                                    f"dp{world} replicas, one batch per rank, NCCL gather of predicted frames"),
                    "cpus_bound_per_rank": affinity})
        line = {
            "metric": METRIC, "value": round(r["fps"], 2), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(r["ms"] / args.steps, 4),
            "passes_ms_per_step": [round(v / args.steps, 4) for v in r["passes"]],
            "ms_per_step_profiled": round(r["ms_prof"] / r["prof_steps"], 4) if r["ms_prof"] else None,
            "higher_is_better": True, "scaling": "strong" if sweep_workload else "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate/residual/softmax", "data": "synthetic",
            "config": cfg,
            "tensor_frac_of_sustained_peak": round(per_gpu_fps * r["flops_frame"] / 1e12 / peaks["tflops_sustained"], 4),
            "tensor_frac_of_burst_peak": round(per_gpu_fps * r["flops_frame"] / 1e12 / peaks["tflops_burst"], 4),
            "e2e": r["e2e"], "gpu_launches": int(r["launches"]),
            "clocks": r["clocks"], "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels,
            "also": also, "operand_modes": operand_modes,
        }
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
