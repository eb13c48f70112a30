"""Skinny (decode-step / small per-rank batch) GEMM timing: weight-streaming bound shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from certifiedgpt_b200 import _lib as L
shapes = [(12288, 4096), (4096, 4096), (22016, 4096), (4096, 11008), (32000, 4096), (4224, 1408), (6144, 1408), (1408, 6144)]
for M in (13, 125, 250, 500, 1000):
    tot_new = tot_old = tot_b = 0.0
    for N, K in shapes:
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        res = []
        for flag in (0, 0x1000 | 256 if M <= 128 else 0x2000):
            for _ in range(3): L.gemm(a, w, out=out, force_bn=flag)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); s.record()
            for _ in range(10): L.gemm(a, w, out=out, force_bn=flag)
            e.record(); torch.cuda.synchronize()
            res.append(s.elapsed_time(e) / 10)
        gbs = N * K * 2 / res[0] / 1e6
        ref = (a.float() @ w.float().t())
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        tot_new += res[0]; tot_old += res[1]
        print(f"M={M:5d} N={N:6d} K={K:6d}: auto {res[0]*1e3:7.1f} us ({gbs:6.0f} GB/s weights)  previous {res[1]*1e3:7.1f} us  relerr {err:.1e}", flush=True)
    print(f"M={M}: sum auto {tot_new*1e3:.0f} us vs previous {tot_old*1e3:.0f} us", flush=True)
