import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
os.environ["CGPT_GEMM_DBG"] = str(dbg.data_ptr())
from certifiedgpt_b200 import _lib as L
def run(M, N, K, act=0, resid=False):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if resid else None
    out = r if resid else torch.empty(M, N // 2 if act == 2 else N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        dbg.zero_(); L.gemm(a, w, bias=None if act == 2 else bias, act=act, resid=r, out=out)
    torch.cuda.synchronize()
    d = dbg.view(148, 8).float()
    lead = d[0::2]
    tiles = ((M + 255) // 256) * ((N + 255) // 256) / 74
    print(f"M={M} N={N} K={K} act={act} resid={resid}: tiles/pair={tiles:.1f}  MMA thread: total={lead[:,2].mean():.0f} wait_epilogue={lead[:,0].mean():.0f} wait_tma={lead[:,1].mean():.0f} | "
          f"epi warp4: wait_mma={d[:,4].mean():.0f} work={d[:,5].mean():.0f} bias+bar={d[:,6].mean():.0f}  per tile: work={d[:,5].mean()/tiles:.0f} mma_tile_ideal={K/16*128:.0f}", flush=True)
run(65792, 6144, 1408, act=1)
run(65792, 4224, 1408)
run(65792, 1408, 1408, resid=True)
run(65792, 1408, 6144, resid=True)
run(18432, 22016, 4096, act=2)
run(18432, 4096, 4096, resid=True)
