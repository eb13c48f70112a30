"""Times the Llama-prefill attention shape (B x 32 heads, 72 queries x 79 keys, hd 128, KV-cache layout):
persistent pipelined kernel (attn_prefill.cu) vs the one-shot tcgen05 kernel (CGPT_ATTN_NO_PREFILL=1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B, H, Tq, Tk, rows, hd = int(os.environ.get("B", 1100)), 32, 72, 79, 83, 128
D = H * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * Tq, 3 * D, device="cuda", generator=g).bfloat16()
kc = torch.randn(B * rows, D, device="cuda", generator=g).bfloat16()
vc = torch.randn(B * rows, D, device="cuda", generator=g).bfloat16()
out = torch.empty(B * Tq, D, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    L.attention(qkv[:, :D], kc, vc, out, B=B, H=H, Tq=Tq, Tk=Tk, head_dim=hd, scale=hd ** -0.5, causal=True,
                kv_rows_per_batch=rows)


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
bytes_ = B * (2 * Tq * D * 2 + 2 * Tk * D * 2)
print(f"prefill attention B={B}: {ms:.3f} ms  ({bytes_ / ms / 1e6:.0f} GB/s of algorithmic Q+K+V+O bytes) "
      f"kernel={'one-shot' if os.environ.get('CGPT_ATTN_NO_PREFILL') else 'persistent'}")
