"""Times the ViT self-attention shape (B x 16 heads, 257 tokens, hd 88, fused QKV buffer): persistent tcgen05
kernel vs one CTA per item (CGPT_ATTN_ONE_SHOT=1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B, H, T, hd = int(os.environ.get("B", 1100)), 16, 257, 88
D = H * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g).bfloat16()
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    L.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=hd ** -0.5)


for _ in range(3):
    run()
ts = []
for _ in range(10):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = sorted(ts)[len(ts) // 2]
fl = 4.0 * B * H * T * T * hd
print(f"ViT attention B={B}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  {B * T * 4 * D * 2 / ms / 1e6:.0f} GB/s "
      f"kernel={'one CTA per item' if os.environ.get('CGPT_ATTN_ONE_SHOT') else 'persistent'}")
