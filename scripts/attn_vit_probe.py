"""Times the ViT self-attention shape (B x 16 heads, 257 tokens, hd 88): the round-1 one-shot tcgen05 kernel on the
row-major fused QKV buffer against the pipelined head-major kernel (csrc/attn_vit.cu), L2 flushed between runs; then one
launch with CGPT_ATTN_DBG stamps: per-phase cycles of a steady-state item (the second) of every CTA (softmax thread 128 / MMA thread)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B, H, T, hd = int(os.environ.get("B", 1100)), 16, 257, 88
D = H * hd
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g).bfloat16()
hm = qkv.view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).contiguous().view(3, -1)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
out2 = torch.empty_like(out)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
scale = hd ** -0.5


def run_row():
    L.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=scale)


def run_hm():
    L.attention(hm[0], hm[1], hm[2], out2, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=scale, head_major=True)


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


fl = 4.0 * B * H * T * T * hd
bytes_alg = B * T * 4 * D * 2
for name, fn in (("row-major one-shot (round 1)", run_row), ("head-major pipelined", run_hm)):
    ms = timed(fn)
    print(f"ViT attention B={B} {name:30s}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  {bytes_alg / ms / 1e6:.0f} GB/s "
          f"({bytes_alg / ms / 1e6 / 6546.2:.2f} of the measured HBM peak)", flush=True)
print("max |row-major - head-major| =", (out.float() - out2.float()).abs().max().item())

# per-phase stamps of the pipelined kernel (last item of each CTA)
sms = torch.cuda.get_device_properties(0).multi_processor_count
dbg = torch.zeros(sms * 16, dtype=torch.int64, device="cuda")
os.environ["CGPT_ATTN_DBG"] = hex(dbg.data_ptr())
run_hm()
torch.cuda.synchronize()
del os.environ["CGPT_ATTN_DBG"]
d = dbg.view(sms, 16).cpu()
names = {5: "item start", 7: "S ready", 8: "row max done", 9: "P written", 12: "next cls scores", 10: "O ready", 11: "O stored"}
base = d[:, 5]
for k in (7, 8, 9, 12, 10, 11):
    v = (d[:, k] - base).float()
    print(f"  softmax thread 128: {names[k]:18s} +{v.median().item():8.0f} cycles (median over CTAs)")
print(f"  MMA thread: first S issued -> done = {(d[:, 3] - d[:, 2]).float().median().item():.0f} cycles for "
      f"{(B * H + sms - 1) // sms} items per CTA")
