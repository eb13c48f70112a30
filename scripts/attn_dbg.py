import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
B, T, H, hd = 64, 257, 16, 88
D = H * hd
dbg = torch.zeros(B * H * 16, dtype=torch.int64, device="cuda")
os.environ["CGPT_ATTN_DBG"] = str(dbg.data_ptr())
from certifiedgpt_b200 import _lib as L
qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).bfloat16()
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    L.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=hd ** -0.5)
torch.cuda.synchronize()
d = dbg.view(B * H, 16).cpu()
names = ["start", "setup_done", "qk_landed", "S_issued", "P0_ready", "P1_ready", "wg_pre_s", "wg_s_done", "pass1_done", "P_written", "O_done", "epi_done", "alloc_done", "end", "tma_issued", "xvec_done"]
for cta in (0, 1, 500, 900):
    t0 = d[cta, 0].item()
    print(f"CTA {cta}: " + "  ".join(f"{n}={d[cta, i].item() - t0}" for i, n in enumerate(names)))
rel = (d[:, :16] - d[:, :1]).float()
print("mean:  " + "  ".join(f"{n}={rel[:, i].mean().item():.0f}" for i, n in enumerate(names)))
