"""Stress test of the attention kernels for rare races: many repetitions per shape against an fp32 reference."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from counterfactualworldmodels_b200 import _lib, ops
lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
shapes = [(64, 130, 6), (64, 256, 6), (64, 200, 6), (64, 384, 6), (30, 788, 12), (40, 300, 4), (64, 788, 12), (9, 520, 8), (16, 1568, 6), (200, 100, 2), (8, 3168, 4), (6, 6272, 8), (32, 3140, 16)]
data = {}
for (B, N, H) in shapes:
    g = torch.Generator().manual_seed(B * 100 + N + H)
    qkv = torch.randn(B * N, 3 * H * 64, generator=g)
    qkv[:, :H * 64] *= 0.125 * 3.0
    qkv = qkv.to(torch.float16).cuda()
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    want = ((q @ k.transpose(-2, -1)).softmax(-1) @ v).transpose(1, 2).reshape(B * N, H * 64)
    data[(B, N, H)] = (qkv, want)
p = torch.cuda.get_device_properties(0)
print("GPU:", p.name, "SMs", p.multi_processor_count, "mem", p.total_memory >> 20, "MiB", "cc", p.major, p.minor, flush=True)
os.system("nvidia-smi --query-gpu=driver_version,clocks.sm,clocks.max.sm,clocks.mem,power.limit,compute_mode,mig.mode.current --format=csv,noheader")
os.system("nvidia-smi --query-compute-apps=pid,used_memory --format=csv,noheader")
for label, persistent, war in (("persistent (default)", 1, 1), ("one item per CTA", 3, 1)):
    lib.cwm_debug_attention_persistent(persistent)
    lib.cwm_debug_attention_war_safe(war)
    for shp in shapes:
        qkv, want = data[shp]
        bad, worst = 0, 0.0
        for r in range(reps):
            got = ops.attention_f16(qkv, *shp).float()
            e = float((got - want).abs().max())
            if not (e < 4e-3):
                bad += 1
            worst = max(worst, e) if e == e else float("nan")
        print(f"{label:22s} {str(shp):18s} failures {bad}/{reps} worst {worst:.3e}", flush=True)
