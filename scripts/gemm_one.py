"""One GEMM shape, a few launches (ncu --set full target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from certifiedgpt_b200 import _lib as L
M, N, K = [int(v) for v in sys.argv[1:4]]
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N // 2 if act == 2 else N, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    L.gemm(a, w, bias=bias, act=act, out=out)
torch.cuda.synchronize()
print("done")
