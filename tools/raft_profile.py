import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from counterfactualworldmodels_b200 import _lib, raft
dev="cuda:0"
torch.manual_seed(0)
args = raft.get_args("")
args.multiframe, args.scale_inputs, args.output_dim, args.mixed_precision = True, True, None, True
model = raft.RAFT(args).eval().requires_grad_(False).to(dev)
S=int(sys.argv[1]) if len(sys.argv)>1 else 64
x = torch.rand(S, 2, 3, 224, 224, device=dev)
x[:,0]=x[:1,0]
for _ in range(3): model(x, shared_frame=0)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): model(x, shared_frame=0)
e1.record(); torch.cuda.synchronize()
print("ms per call", e0.elapsed_time(e1)/5)
_lib.profile_begin()
model(x, shared_frame=0)
torch.cuda.synchronize()
tot=0
for e in _lib.profile_end():
    print(f"{e['name']:28s} launches {e['launches']:4d} ms {e['ms']:.3f}  avg_us {1e3*e['ms']/e['launches']:.1f}")
    tot+=e['ms']
print("sum libcwm", tot)
