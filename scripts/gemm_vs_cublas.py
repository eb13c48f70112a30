"""Sustained (power-capped) throughput of libcgpt's tcgen05 GEMM against cuBLAS (torch.matmul) on the bench's top shapes:
plain bf16 C = A W^T without epilogue work on either side, each measured in its own ~1.5 s window, A B A B, CUDA events.
usage: python scripts/gemm_vs_cublas.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B = 1100
shapes = [("vit_fc1", B * 257, 6144, 1408), ("vit_fc2", B * 257, 1408, 6144), ("vit_qkv", B * 257, 4224, 1408),
          ("vit_proj", B * 257, 1408, 1408), ("llama_gateup", B * 72, 22016, 4096), ("llama_qkv", B * 72, 12288, 4096),
          ("llama_down", B * 72, 4096, 11008), ("llama_o", B * 72, 4096, 4096)]


def window(fn, seconds=1.5):
    fn(); torch.cuda.synchronize()
    t_end = time.time() + 0.4
    while time.time() < t_end:            # bring the clocks to their sustained level
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
    n = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    t_end = time.time() + seconds
    while time.time() < t_end:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for name, M, N, K in shapes:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ours = lambda: L.gemm(a, w, out=out)
    cublas = lambda: torch.matmul(a, w.t(), out=out)
    ms = {"ours": [], "cublas": []}
    for _ in range(2):
        ms["ours"].append(window(ours))
        ms["cublas"].append(window(cublas))
    fl = 2.0 * M * N * K
    o, c = min(ms["ours"]), min(ms["cublas"])
    print(f"{name:13s} M={M:6d} N={N:5d} K={K:5d}: libcgpt {fl / o / 1e9:7.1f} TFLOP/s ({o:.3f} ms)   cuBLAS {fl / c / 1e9:7.1f} TFLOP/s ({c:.3f} ms)   "
          f"ratio {c / o:.3f}", flush=True)
    del a, w, out
    torch.cuda.empty_cache()
