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
    # HBM-bound pixel <-> token kernels at base 8x8 batch 64 / large 4x4 batch 32 (a1+a2, a10, a12) and the
    # materialising counterfactual-construction kernel
    "gather_base": ("gather", 64, 8, 788),
    "gather_4x4": ("gather", 32, 4, 3140),
    "fillmask_base": ("fillmask", 64, 1568, 788, 384),
    "fillmask_large": ("fillmask", 32, 6272, 3140, 512),
    "unpatchify_base": ("unpatch", 64, 8, 788),
    "unpatchify_4x4": ("unpatch", 32, 4, 3140),
    "cf_build_256": ("cfbuild", 256, 8),
    # the small attentions of the IMU-conditioned model (config 5, batch 32): cross attention in both directions
    "xattn_enc_trg": ("xattn", 32, 3140, 25, 4, 192),
    "xattn_enc_src": ("xattn", 32, 25, 3140, 4, 192),
    "xattn_dec_trg": ("xattn", 32, 6336, 50, 4, 96),
    "xattn_dec_src": ("xattn", 32, 50, 6336, 4, 96),
}


def pixel_case(name):
    """The HBM-bound kernels, called through the C ABI on rotating buffers (> 126 MB in total)."""
    import ctypes
    from counterfactualworldmodels_b200 import perturbation, synthetic
    from counterfactualworldmodels_b200.vmae import compact_mask
    lib = _lib.load()
    kind = CASES[name][0]
    stream = lambda: torch.cuda.current_stream().cuda_stream
    if kind == "fillmask":
        _, B, Ntot, Nvis, C = CASES[name]
        mask = torch.ones(B, Ntot, dtype=torch.bool, device=DEV)
        mask[:, :Nvis] = False
        perm, _, _ = compact_mask(mask)
        tok, pos = torch.randn(C, device=DEV), torch.randn(Ntot, C, device=DEV)
        nbuf = max(2, int(300e6 // (B * Ntot * C * 4)) + 1)
        outs = [torch.empty(B, Ntot, C, device=DEV) for _ in range(nbuf)]
        def call(i):
            _lib.check(lib.cwm_fill_mask_tokens(tok.data_ptr(), pos.data_ptr(), perm.data_ptr(), B, Ntot, Nvis, C,
                                                outs[i % nbuf].data_ptr(), stream()))
        return call, 0.0, B * (Ntot - Nvis) * C * 4
    if kind == "cfbuild":
        _, S, P = CASES[name]
        h = 224 // P
        x = synthetic.make_video(1, (224, 224), seed=0).to(DEV)
        active = torch.ones(S, 2, h, h, dtype=torch.bool, device=DEV)
        active[:, 1, 5:7, 9:11] = False
        passive = torch.zeros(S, 2, h, h, dtype=torch.bool, device=DEV)
        passive[:, 1] = True
        video, _ = perturbation.shift_patches_and_masks(x, passive.reshape(S, -1), active.reshape(S, -1),
                                                        [[1, -2]] * S, (1, P, P), frame=1, static_frame=0)
        src, keep = video.c_struct()
        outs = [torch.empty(S, 2, 3, 224, 224, device=DEV) for _ in range(2)]
        def call(i, keep=keep):
            _lib.check(lib.cwm_cf_build_videos(ctypes.byref(src), S, 2, 3, 224, 224, P, P, outs[i % 2].data_ptr(),
                                               stream()))
        return call, 0.0, S * 2 * 3 * 224 * 224 * 4
    _, B, P, Nvis = CASES[name]
    n = (224 // P) ** 2
    nbuf = max(2, int(300e6 // (B * 2 * 3 * 224 * 224 * 4)) + 1)
    xs = [torch.rand(B, 2, 3, 224, 224, device=DEV) for _ in range(nbuf)]
    mask = torch.ones(B, 2 * n, dtype=torch.bool, device=DEV)
    mask[:, :n] = False
    mask[:, n + 7:n + 7 + (Nvis - n)] = False
    perm, inv, _ = compact_mask(mask)
    K = 3 * P * P
    mean, std = _lib.float_array((0.485, 0.456, 0.406)), _lib.float_array((0.229, 0.224, 0.225))
    if kind == "gather":
        outs = [torch.empty(B * Nvis, K, device=DEV, dtype=torch.float16) for _ in range(nbuf)]
        def call(i):
            xv = xs[i % nbuf].transpose(1, 2)  # the [B, C, T, H, W] view the reference hands over
            _lib.check(lib.cwm_patch_gather(xv.data_ptr(), _lib.strides5(xv), B, 3, 2, 224, 224, 1, P, P, perm.data_ptr(),
                                            2 * n, Nvis, mean, std, outs[i % nbuf].data_ptr(), stream()))
        return call, 0.0, B * Nvis * K * 6
    Nmask = 2 * n - Nvis
    ys = [torch.randn(B, Nmask, K, device=DEV) for _ in range(nbuf)]
    outs = [torch.empty(B, 2, 3, 224, 224, device=DEV) for _ in range(nbuf)]
    def call(i):
        x = xs[i % nbuf]
        _lib.check(lib.cwm_unpatchify_scatter(ys[i % nbuf].data_ptr(), x.data_ptr(), _lib.strides5(x), inv.data_ptr(), B,
                                              2, 3, 224, 224, 1, P, P, Nvis, outs[i % nbuf].data_ptr(), stream()))
    return call, 0.0, B * 2 * 3 * 224 * 224 * 8


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
    elif kind == "xattn":
        _, B, Nq, Nk, H, d = CASES[name]
        q = (torch.randn(B * Nq, H * d, device=DEV) * d ** -0.25).half()
        k = (torch.randn(B * Nk, H * d, device=DEV) * d ** -0.25).half()
        v = torch.randn(B * Nk, H * d, device=DEV).half()
        def call(i):
            ops.attention_generic_f16(q, k, v, B, Nq, Nk, H, d)
        flops, byts = 4.0 * B * H * Nq * Nk * d, (2.0 * B * Nq + 2.0 * B * Nk) * H * d * 2
    elif kind in ("gather", "unpatch", "fillmask", "cfbuild"):
        call, flops, byts = pixel_case(name)
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
    if os.environ.get("CWM_ATTN_SKIP"):
        _lib.load().cwm_debug_attention_skip_idle(int(os.environ["CWM_ATTN_SKIP"]))
        print("attention skip-idle-warps =", os.environ["CWM_ATTN_SKIP"])
    if os.environ.get("CWM_ATTN_STALE"):
        _lib.load().cwm_debug_attention_stale_max(int(os.environ["CWM_ATTN_STALE"]))
        print("attention stale-max =", os.environ["CWM_ATTN_STALE"])
    if os.environ.get("CWM_ATTN_MAP"):
        _lib.load().cwm_debug_attention_persist_map(int(os.environ["CWM_ATTN_MAP"]))
        print("attention persist map =", os.environ["CWM_ATTN_MAP"])
    if os.environ.get("CWM_ATTN_MMA_WIDE"):
        _lib.load().cwm_debug_attn_mma_wide(int(os.environ["CWM_ATTN_MMA_WIDE"]))
    if os.environ.get("CWM_ATTN_MMA_SPLIT"):
        _lib.load().cwm_debug_attn_mma_split(int(os.environ["CWM_ATTN_MMA_SPLIT"]))
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n)
