"""Pipeline trace of CTA (0,0,0) of the attention kernel (clock64 at each hand-off), to see who waits for whom.
   python tools/attn_trace.py [B N H]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from counterfactualworldmodels_b200 import _lib, ops  # noqa: E402

B, N, H = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 6272, 8)
lib = _lib.load()
lib.cwm_debug_attention_trace.argtypes = [ctypes.c_void_p]
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 0.5).half()
ops.attention_f16(qkv, B, N, H)
torch.cuda.synchronize()
buf = torch.zeros(3 * 16 * 8, dtype=torch.int64, device="cuda")
lib.cwm_debug_attention_trace(buf.data_ptr())
ops.attention_f16(qkv, B, N, H)
torch.cuda.synchronize()
lib.cwm_debug_attention_trace(None)
t = buf.cpu().view(3, 16, 8)
t0 = int(t[t > 0].min())
names = {0: ["k_full ok", "QK0 issued", "P0+V ok", "PV0 issued", "QK1 issued", "P1 ok", "PV1 issued"],
         1: ["iter start", "S ready", "S loaded", "max done", "exp issued", "P stored", ""],
         2: ["iter start", "S ready", "S loaded", "max done", "exp issued", "P stored", ""]}
for role, label in ((0, "MMA issuer"), (1, "softmax WG0"), (2, "softmax WG1")):
    print(f"--- {label}: " + " | ".join(names[role]))
    for j in range(12):
        row = [int(v) - t0 if int(v) > 0 else -1 for v in t[role, j, :7]]
        print(f"  j={j:2d} " + " ".join(f"{v:7d}" for v in row))
