"""One mixed-precision flow call of raft.RAFT (RAFT-large shapes, random init) on S frame pairs sharing frame 0 -- the
command profiled under ncu for profiles/ (launch list, conv / correlation kernels).   python tools/raft_one_call.py [S]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from counterfactualworldmodels_b200 import raft  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda:0"
torch.manual_seed(0)
args = raft.get_args("")
args.multiframe, args.scale_inputs, args.output_dim, args.mixed_precision = True, True, None, True
model = raft.RAFT(args).eval().requires_grad_(False).to(dev)
x = torch.rand(S, 2, 3, 224, 224, device=dev)
x[:, 0] = x[:1, 0]
for _ in range(2):
    model(x, shared_frame=0)
torch.cuda.synchronize()
