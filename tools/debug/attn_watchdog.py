import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from counterfactualworldmodels_b200 import _lib, ops
lib = _lib.load()
lib.cwm_debug_attention_persistent(2)
B, N, H = (int(v) for v in sys.argv[1:4])
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 0.5).half()
o = ops.attention_f16(qkv, B, N, H)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
