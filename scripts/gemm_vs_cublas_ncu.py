"""One launch of libcgpt's GEMM and one of cuBLAS (torch.matmul) per shape, for an ncu --set full capture:
python scripts/gemm_vs_cublas_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B = 1100
for name, M, N, K in [("llama_down", B * 72, 4096, 11008), ("llama_gateup_plain", B * 72, 22016, 4096), ("vit_fc1_plain", B * 257, 6144, 1408)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, w, out=out)
    torch.cuda.synchronize()
    torch.matmul(a, w.t(), out=out)
    torch.cuda.synchronize()
    del a, w, out
    torch.cuda.empty_cache()
print("done")
