"""Per-shape timing of the recurrent block's convolutions at the sweep shape (S frame pairs of 224 px -> 28x28 maps):
the implicit-GEMM kernel (cwm_conv2d_f16) against cuDNN on the same f16 channels-last tensors.
    python tools/conv_bench.py [S]"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from counterfactualworldmodels_b200 import ops  # noqa: E402

SHAPES = [("convc1", 328, 256, 1, 1), ("convc2", 256, 192, 3, 3), ("convf1_gemm", 128, 128, 1, 1), ("convf2", 128, 64, 3, 3),
          ("conv", 256, 128, 3, 3), ("convz|r 1x5", 384, 256, 1, 5), ("convq 1x5", 384, 128, 1, 5),
          ("convz|r 5x1", 384, 256, 5, 1), ("convq 5x1", 384, 128, 5, 1), ("flow_head.conv1", 128, 256, 3, 3),
          ("flow_head.conv2", 256, 8, 3, 3)]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    H = W = 28
    dev = "cuda:0"
    out = {"S": S, "rows": S * H * W, "convs": []}
    tot_tc = tot_cudnn = tot_flop = 0.0
    for name, Cin, Cout, kh, kw in SHAPES:
        xs = [(torch.randn(S * H * W, Cin, device=dev) * 0.5).half() for _ in range(3)]
        w = (torch.randn(Cout, Cin, kh, kw, device=dev) / (Cin * kh * kw) ** 0.5).half()
        packed = ops.pack_conv_weight(w)
        dst = torch.empty(S * H * W, Cout, dtype=torch.float16, device=dev)
        i = [0]

        def tc():
            i[0] += 1
            ops.conv2d_f16(xs[i[0] % 3], S, H, W, w, packed=packed, out=dst)

        wcl = w.contiguous(memory_format=torch.channels_last)
        xcl = [x.view(S, H, W, Cin).permute(0, 3, 1, 2) for x in xs]

        def cudnn():
            i[0] += 1
            F.conv2d(xcl[i[0] % 3], wcl, None, 1, (kh // 2, kw // 2))

        t_tc, t_cd = timed(tc), timed(cudnn)
        flop = 2.0 * S * H * W * Cout * Cin * kh * kw
        out["convs"].append({"name": name, "Cin": Cin, "Cout": Cout, "k": [kh, kw], "ms_tcgen05": round(t_tc, 4),
                             "ms_cudnn": round(t_cd, 4), "tflops_tcgen05": round(flop / t_tc / 1e9, 1),
                             "tflops_cudnn": round(flop / t_cd / 1e9, 1)})
        tot_tc += t_tc
        tot_cudnn += t_cd
        tot_flop += flop
    out["sum_ms_tcgen05"], out["sum_ms_cudnn"] = round(tot_tc, 3), round(tot_cudnn, 3)
    out["tflops_tcgen05"], out["tflops_cudnn"] = round(tot_flop / tot_tc / 1e9, 1), round(tot_flop / tot_cudnn / 1e9, 1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
