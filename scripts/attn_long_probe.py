"""Times the 448 px ViT self-attention shape (B x 16 heads, 1025 tokens, hd 88) on the multi-tile tcgen05 kernel
(csrc/attn_long.cu), L2 flushed between runs, and checks it against an fp32 softmax(QK^T)V of the same bf16 inputs.
CGPT_LIB=<other libcgpt.so> times another build of the kernel on the same box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B, H, T, hd = int(os.environ.get("B", 256)), 16, int(os.environ.get("T", 1025)), 88
D = H * hd
g = torch.Generator(device="cuda").manual_seed(0)
hm = (torch.randn(3, B, H, T, hd, device="cuda", generator=g) * 1.5).bfloat16()
out = torch.zeros(B * T, D, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
scale = hd ** -0.5


def run():
    L.attention(hm[0].view(-1), hm[1].view(-1), hm[2].view(-1), out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=scale,
                head_major=True)


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
print(f"attn_long B={B} T={T} lib={os.path.basename(L.LIB_PATH)}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s (useful flops)", flush=True)
nb = min(B, 4)
q, k, v = (hm[i, :nb].float() for i in range(3))
ref = torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v                    # [nb, H, T, hd]
got = out.view(B, T, H, hd)[:nb].permute(0, 2, 1, 3).float()
print(f"max |kernel - fp32 reference| over {nb} samples = {(got - ref).abs().max().item():.4f} (bf16 P and output)")

# cycle stamps of every CTA's second unit (softmax thread 128 = group 0 row 0; the MMA-issuing thread)
sms = torch.cuda.get_device_properties(0).multi_processor_count
dbg = torch.zeros(sms * 16, dtype=torch.int64, device="cuda")
os.environ["CGPT_ATTN_DBG"] = hex(dbg.data_ptr())
run()
torch.cuda.synchronize()
del os.environ["CGPT_ATTN_DBG"]
d = dbg.view(sms, 16).cpu()
med = lambda v: v.float().median().item()
print(f"  softmax thread: key loop {med(d[:, 1] - d[:, 0]):7.0f} cycles (of which waiting for scores {med(d[:, 9]):6.0f}); "
      f"wait for O {med(d[:, 3] - d[:, 2]):6.0f}; epilogue {med(d[:, 4] - d[:, 3]):6.0f}; unit {med(d[:, 4] - d[:, 0]):7.0f}")
print(f"  MMA thread (group 0): unit {med(d[:, 6] - d[:, 5]):7.0f} cycles; waiting for Q {med(d[:, 13]):6.0f}")
print(f"  softmax thread, summed over the unit's 64-key steps: tcgen05.ld {med(d[:, 7]):6.0f}, maximum + P.V bookkeeping "
      f"{med(d[:, 8]):6.0f}, exp2 + pack + tcgen05.st + arrive {med(d[:, 10]):6.0f}")
print(f"  MMA thread, summed over the unit: P.V MMAs {med(d[:, 14]):6.0f} + their commits {med(d[:, 12]):6.0f}; S MMAs {med(d[:, 15]):6.0f} + their commits {med(d[:, 11]):6.0f}")
