import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dbg = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
os.environ["CGPT_ATTN_DBG"] = str(dbg.data_ptr())
from certifiedgpt_b200 import _lib as L
names = ["start", "setup_done", "qk_landed", "S_issued", "P0_ready", "P1_ready", "wg_pre_s", "wg_s_done", "pass1_done", "P_written", "O_done", "epi_done", "alloc_done", "end", "tma_issued", "xvec_done"]
def run(B, T, H, hd, fused):
    D = H * hd
    if fused:
        qkv = (torch.randn(B * T, 3 * D, device="cuda") * 0.5).bfloat16()
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q = torch.randn(B * T, D, device="cuda").bfloat16(); k = torch.randn(B * T, D, device="cuda").bfloat16(); v = torch.randn(B * T, D, device="cuda").bfloat16()
    out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        dbg.zero_()
        L.attention(q, k, v, out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=hd ** -0.5)
    torch.cuda.synchronize()
    d = dbg.view(-1, 16)[:B * H].cpu()
    rel = (d - d[:, :1]).float()
    print(f"B={B} T={T} H={H} hd={hd} fused={fused}: " + "  ".join(f"{n}={rel[:, i].mean().item():.0f}" for i, n in enumerate(names) if d[:, i].abs().sum() > 0), flush=True)
run(64, 257, 16, 88, True)
run(64, 256, 16, 128, False)
run(64, 256, 16, 128, True)
run(64, 128, 16, 128, False)
run(8, 257, 16, 88, True)
run(64, 79, 32, 128, False)
