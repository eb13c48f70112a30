"""Single-token decode attention at the bench's shape (1100 samples x 32 heads x 128, 7 shared prefix rows + 73 own rows of an
83-row cache): time against the KV bytes it has to stream.  usage: python scripts/decode_attn_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B, H, hd, P, T_own, rows = 1100, 32, 128, 7, 73, 83
D = H * hd
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B, D, device="cuda", generator=g).bfloat16()
kc = torch.randn(B * rows, D, device="cuda", generator=g).bfloat16()
vc = torch.randn(B * rows, D, device="cuda", generator=g).bfloat16()
kp = torch.randn(P, D, device="cuda", generator=g).bfloat16()
vp = torch.randn(P, D, device="cuda", generator=g).bfloat16()
out = torch.empty(B, D, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    L.attention(q, kc, vc, out, B=B, H=H, Tq=1, Tk=P + T_own, head_dim=hd, scale=hd ** -0.5, q_rows_per_batch=1,
                kv_rows_per_batch=rows, kp=kp, vp=vp, P=P, decode=True)


for _ in range(3):
    run()
ts = []
for _ in range(20):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = sorted(ts)[len(ts) // 2]
byts = 2.0 * B * T_own * D * 2
print(f"decode attention B={B} Tk={P + T_own} lib={os.path.basename(L.LIB_PATH)}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s of KV "
      f"({byts / ms / 1e6 / 6546.2:.2f} of the measured HBM peak)")
