"""Sustained throughput of a residual GEMM with its in-place fp32 residual epilogue (llama_down: M = 79 200, N = 4096,
K = 11008; llama_o; vit_fc2) under the tile / epilogue switches: CGPT_GEMM_MT (1 = 256-row tiles, 2 = 512-row tiles) x direct / coalesced
red.global.add form.  usage: python scripts/gemm_down_probe.py [llama_down | llama_o | vit_fc2]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

SHAPES = {"llama_down": (1100 * 72, 4096, 11008), "llama_o": (1100 * 72, 4096, 4096), "vit_fc2": (1100 * 257, 1408, 6144)}
name = sys.argv[1] if len(sys.argv) > 1 else "llama_down"
M, N, K = SHAPES[name]
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
res = torch.zeros(M, N, device="cuda", dtype=torch.float32)
fn = lambda: L.gemm(a, w, resid=res, out=res)


def window(seconds=1.5):
    fn(); torch.cuda.synchronize()
    t_end = time.time() + 0.4
    while time.time() < t_end:
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


fl = 2.0 * M * N * K
for rep in range(2):
    for mt in ("1", "2"):
        for form in ("direct", "coalesced"):
            os.environ["CGPT_GEMM_MT"] = mt
            os.environ.pop("CGPT_GEMM_RED_COALESCED", None)
            if form == "coalesced":
                os.environ["CGPT_GEMM_RED_COALESCED"] = "1"
            ms = window()
            print(f"{name} + fp32 residual, {256 * int(mt)}-row tiles, {form:9s} reductions: {ms:.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
