"""Times individual libcwm_b200 kernels at the shapes of the BASELINE workloads (CUDA events, L2 flushed between
launches by rotating over buffers > 126 MB).  Also the command profiled under `ncu --set full` for profiles/.

  python tools/kernel_bench.py gemm_qkv_base attn_enc_base ...      (no args = all)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from counterfactualworldmodels_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
F16, GELU, RES, F32 = _lib.EPI_F16, _lib.EPI_GELU_F16, _lib.EPI_RES_F32, _lib.EPI_F32

# name -> ("gemm", M, N, K, mode) | ("attn", B, N, H) | ("ln", M, C)
CASES = {
    # base 8x8, batch 64 (BASELINE configs[1]): encoder M = 64*788, decoder M = 64*1568
    "gemm_qkv_base_enc": ("gemm", 50432, 2304, 768, F16),
    "gemm_proj_base_enc": ("gemm", 50432, 768, 768, RES),
    "gemm_fc1_base_enc": ("gemm", 50432, 3072, 768, GELU),
    "gemm_fc2_base_enc": ("gemm", 50432, 768, 3072, RES),
    "gemm_qkv_base_dec": ("gemm", 100352, 1152, 384, F16),
    "gemm_proj_base_dec": ("gemm", 100352, 384, 384, RES),
    "gemm_fc1_base_dec": ("gemm", 100352, 1536, 384, GELU),
    "gemm_fc2_base_dec": ("gemm", 100352, 384, 1536, RES),
    # large 4x4, batch 32 chunk (configs[3]): encoder M = 32*3140, decoder M = 32*6272
    "gemm_qkv_large_enc": ("gemm", 100480, 3072, 1024, F16),
    "gemm_fc1_large_enc": ("gemm", 100480, 4096, 1024, GELU),
    "gemm_fc2_large_enc": ("gemm", 100480, 1024, 4096, RES),
    "gemm_fc2_large_dec": ("gemm", 200704, 512, 2048, RES),
    "attn_enc_base": ("attn", 64, 788, 12),
    "attn_dec_base": ("attn", 64, 1568, 6),
    "attn_enc_large": ("attn", 32, 3140, 16),
    "attn_dec_large": ("attn", 32, 6272, 8),
    "attn_dec_large_b6": ("attn", 6, 6272, 8),
    "ln_base_enc": ("ln", 50432, 768),
}


def run_case(name, iters=5):
    kind = CASES[name][0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    if kind == "gemm":
        _, M, N, K, mode = CASES[name]
        nbuf = max(2, int(200e6 // (M * K * 2)) + 1)
        a = [torch.randn(M, K, device=DEV).half() for _ in range(nbuf)]
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).half()
        bias = torch.randn(N, device=DEV)
        f16 = mode in (F16, GELU)
        outs = [torch.zeros(M, N, device=DEV, dtype=torch.float16 if f16 else torch.float32) for _ in range(nbuf)]
        def call(i):
            o = outs[i % nbuf]
            ops.gemm_f16(a[i % nbuf], w, mode, bias=bias, scale=0.125, scale_cols=N // 3 if mode == F16 else 0,
                         res=o if mode == RES else None, out=o)
        flops, byts = 2.0 * M * N * K, (M * K + N * K) * 2 + M * N * (2 if f16 else (8 if mode == RES else 4))
    elif kind == "attn":
        _, B, N, H = CASES[name]
        nbuf = max(2, int(200e6 // (B * N * 3 * H * 64 * 2)) + 1)
        qkv = [(torch.randn(B * N, 3 * H * 64, device=DEV) * 0.5).half() for _ in range(nbuf)]
        def call(i):
            ops.attention_f16(qkv[i % nbuf], B, N, H)
        flops, byts = 4.0 * B * H * N * N * 64, B * N * H * 64 * 2 * 4
    else:
        _, M, C = CASES[name]
        nbuf = max(2, int(200e6 // (M * C * 4)) + 1)
        x = [torch.randn(M, C, device=DEV) for _ in range(nbuf)]
        g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        def call(i):
            ops.layernorm_f16(x[i % nbuf], g, b, 1e-6)
        flops, byts = 0.0, M * C * 6
    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    for i in range(iters):
        ev[2 * i].record()
        call(i)
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(iters))
    med = ts[len(ts) // 2]
    print(f"{name:22s} {med * 1e3:9.1f} us  {flops / med / 1e9:8.1f} TFLOP/s  {byts / med / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    if os.environ.get("CWM_ATTN_POLY"):
        import ctypes
        lib = _lib.load()
        lib.cwm_debug_attention_poly.argtypes = [ctypes.c_int]
        lib.cwm_debug_attention_poly(int(os.environ["CWM_ATTN_POLY"]))
        print("attention poly eighths =", os.environ["CWM_ATTN_POLY"])
    if os.environ.get("CWM_ATTN_PERSIST"):
        lib = _lib.load()
        lib.cwm_debug_attention_persistent(int(os.environ["CWM_ATTN_PERSIST"]))
        print("attention mode (3 = one item per CTA, 1 = experimental multi-item) =", os.environ["CWM_ATTN_PERSIST"])
    if os.environ.get("CWM_ATTN_MAP"):
        _lib.load().cwm_debug_attention_persist_map(int(os.environ["CWM_ATTN_MAP"]))
        print("attention persist map =", os.environ["CWM_ATTN_MAP"])
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n)
