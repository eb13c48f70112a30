"""N = 1408 residual GEMMs (ViT proj / fc2): 256-wide CTA-pair tiles (5.5 tiles, 8 % masked) vs 176-wide (8 exact tiles)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

M = 1100 * 257
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, N, K in [("vit_fc2", 1408, 6144), ("vit_proj", 1408, 1408), ("vit_qkv", 4224, 1408)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.zeros(N, device="cuda")
    res = torch.zeros(M, N, device="cuda", dtype=torch.float32) if N == 1408 else None
    out = None if N == 1408 else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for bn in (256, 176, 128):
        def fn():
            if res is not None:
                L.gemm(a, w, bias=bias, resid=res, out=res, force_bn=0x2000 | bn)
            else:
                L.gemm(a, w, bias=bias, out=out, force_bn=0x2000 | bn)
        for _ in range(3):
            fn()
        ts = []
        for _ in range(7):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ms = sorted(ts)[len(ts) // 2]
        print(f"{name} N={N} K={K} BN={bn}: {ms:.3f} ms {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
    del a, w, res, out
    torch.cuda.empty_cache()
