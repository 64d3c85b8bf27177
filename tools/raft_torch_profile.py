import sys, os, torch
sys.path.insert(0, os.getcwd())
from counterfactualworldmodels_b200 import raft
from torch.profiler import profile, ProfilerActivity
dev="cuda:0"
torch.manual_seed(0)
args = raft.get_args("")
args.multiframe, args.scale_inputs, args.output_dim, args.mixed_precision = True, True, None, True
model = raft.RAFT(args).eval().requires_grad_(False).to(dev)
S=64
x = torch.rand(S, 2, 3, 224, 224, device=dev); x[:,0]=x[:1,0]
for _ in range(3): model(x, shared_frame=0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    model(x, shared_frame=0)
    torch.cuda.synchronize()
rows = prof.key_averages()
rows = sorted(rows, key=lambda r: -r.device_time_total)
tot = sum(r.self_device_time_total for r in rows)
print("total device us", tot)
n=0
for r in sorted(rows, key=lambda r: -r.self_device_time_total):
    if r.self_device_time_total <= 0: continue
    print(f"{r.key[:90]:90s} n={r.count:4d} self_us={r.self_device_time_total:9.1f}")
    n+=1
    if n>=28: break
