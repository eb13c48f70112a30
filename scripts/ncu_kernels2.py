"""One kernel at a time for `ncu --set full` (round 2): usage  python scripts/ncu_kernels2.py <what> [batch]
  gemm_qkv_rowmajor | gemm_qkv_headmajor   the ViT fused-QKV GEMM (M = batch * 257, N = 4224, K = 1408), both epilogues
  attn_vit                                   pipelined head-major attention, T = 257
  attn_long                                  multi-tile attention, T = 1025 (batch / 4 samples)
  noise                                      K1, Philox mode
Each kernel runs 3 times (2 warm-ups + the launch to look at: -s 2 -c 1 with -k regex:<kernel>)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

what = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
H, hd = 16, 88
D = H * hd
g = torch.Generator(device="cuda").manual_seed(0)
if what.startswith("gemm_qkv"):
    T = 257
    M = B * T
    a = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(3 * D, D, device="cuda", generator=g) / D ** 0.5).bfloat16()
    bias = torch.randn(3 * D, device="cuda", generator=g)
    out = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        L.gemm(a, w, bias=bias, out=out, headmajor=(T, H, hd) if what.endswith("headmajor") else None)
elif what in ("attn_vit", "attn_long"):
    T = 257 if what == "attn_vit" else 1025
    Bq = B if what == "attn_vit" else max(1, B // 4)
    q, k, v = (torch.randn(Bq * H * T * hd, device="cuda", generator=g).bfloat16() for _ in range(3))
    out = torch.empty(Bq * T, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        L.attention(q, k, v, out, B=Bq, H=H, Tq=T, Tk=T, head_dim=hd, scale=hd ** -0.5, head_major=True)
elif what == "noise":
    x = torch.rand(3, 224, 224, device="cuda")
    out = torch.empty(B * 256, 592, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        L.noise_patchify(x, B, 0.25, seed=1, out=out)
torch.cuda.synchronize()
print("done", what)
