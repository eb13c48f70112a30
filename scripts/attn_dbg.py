"""Phase timestamps (clock64, relative to the item start) of the LAST item each persistent CTA of attn_umma_kernel
processed: steady state of the item loop.  ViT shape by default."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

B, T, H, hd = int(os.environ.get("B", 1100)), 257, 16, 88
D = H * hd
dbg = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
os.environ["CGPT_ATTN_DBG"] = str(dbg.data_ptr())
from certifiedgpt_b200 import _lib as L

qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).bfloat16()
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    dbg.zero_()
    L.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=hd ** -0.5)
torch.cuda.synchronize()
names = {0: "item_start(t0)", 15: "item_start(wg)", 2: "qk_landed+tmem_free", 3: "S_issued", 6: "wg_extra_done", 7: "wg_S_done",
         8: "pass1_done", 9: "P_written", 4: "P0_seen", 5: "P1_seen", 10: "wg_O_done", 11: "wg_epilogue_done",
         13: "t0_next_loads_issued"}
d = dbg.view(-1, 16)[:148].cpu()
rel = (d - d[:, :1]).float()
order = sorted(names, key=lambda i: rel[:, i].mean().item())
print("mean cycles since item start: " + "  ".join(f"{names[i]}={rel[:, i].mean().item():.0f}" for i in order))
