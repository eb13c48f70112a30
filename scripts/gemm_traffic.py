"""DRAM traffic and time of the residual / SwiGLU GEMM shapes of the bench step (ncu --metrics target and
CUDA-event timing): python scripts/gemm_traffic.py [time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B = 1100
shapes = [("llama_down", B * 72, 4096, 11008, "resid"), ("vit_fc2", B * 257, 1408, 6144, "resid"),
          ("llama_gateup", B * 72, 22016, 4096, "swiglu"), ("llama_qkv", B * 72, 12288, 4096, "none"),
          ("vit_fc1", B * 257, 6144, 1408, "gelu")]
timing = len(sys.argv) > 1 and sys.argv[1] == "time"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, M, N, K, kind in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    if kind == "resid":
        res = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        fn = lambda: L.gemm(a, w, resid=res, out=res)
    elif kind == "swiglu":
        out = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
        fn = lambda: L.gemm(a, w, act=L.ACT_SWIGLU, out=out)
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        bias = torch.zeros(N, device="cuda")
        fn = lambda: L.gemm(a, w, bias=bias, act=L.ACT_GELU if kind == "gelu" else L.ACT_NONE, out=out)
    fn()
    torch.cuda.synchronize()
    if timing:
        ts = []
        for _ in range(8):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ms = sorted(ts)[len(ts) // 2]
        print(f"{name:14s} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s  group_m={os.environ.get('CGPT_GEMM_GROUP_M', 'auto')}")
    else:
        fn()
        torch.cuda.synchronize()
    del a, w
    res = out = None
    torch.cuda.empty_cache()
