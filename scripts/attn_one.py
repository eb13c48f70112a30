"""One attention shape, a few launches (ncu target). usage: attn_one.py vit|llm B"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from certifiedgpt_b200 import _lib as L
kind, B = sys.argv[1], int(sys.argv[2])
if kind == "vit":
    T, H, hd = 257, 16, 88
    D = H * hd
    qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).bfloat16()
    out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
    def run():
        L.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=hd ** -0.5)
    flops = 4.0 * T * T * hd * H * B
else:
    Tq, P, H, hd = 72, 7, 32, 128
    D = H * hd
    rows = 83
    qkv = (torch.randn(B * Tq, 3 * D, device="cuda") * 0.5).bfloat16()
    kc = (torch.randn(B, rows, D, device="cuda") * 0.5).bfloat16()
    vc = (torch.randn(B, rows, D, device="cuda") * 0.5).bfloat16()
    out = torch.empty(B * Tq, D, device="cuda", dtype=torch.bfloat16)
    flash = len(sys.argv) > 3
    def run():
        L.attention(qkv[:, :D], kc.view(-1, D), vc.view(-1, D), out, B=B, H=H, Tq=Tq, Tk=P + Tq, head_dim=hd,
                    scale=hd ** -0.5, kv_rows_per_batch=rows, causal=True, force_flash=flash)
    flops = 4.0 * Tq * (P + Tq) * hd * H * B / 2
for _ in range(3): run()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
for _ in range(5): run()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print(f"{kind} B={B}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TF/s (useful)")
