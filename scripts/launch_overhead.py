import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from certifiedgpt_b200 import _lib as L
a = torch.randn(13, 64, device="cuda").bfloat16(); w = torch.randn(256, 64, device="cuda").bfloat16()
out = torch.empty(13, 256, device="cuda", dtype=torch.bfloat16)
for _ in range(10): L.gemm(a, w, out=out)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(1000): L.gemm(a, w, out=out)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"tiny gemm: host {1e6*(t1-t0)/1000:.1f} us/launch, total {1e6*(t2-t0)/1000:.1f} us/launch")
x = torch.randn(1000, 4096, device="cuda"); g = torch.ones(4096, device="cuda"); o = torch.empty(1000, 4096, device="cuda", dtype=torch.bfloat16)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(1000): L.norm_rows(x, g, None, 1e-5, o, rms=True)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"norm: host {1e6*(t1-t0)/1000:.1f} us/launch")
# GPU-side durations of skinny GEMMs via CUPTI
for M in (13, 125):
    for N, K in [(4096, 4096), (12288, 4096), (22016, 4096), (4096, 11008)]:
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        line = f"M={M} N={N} K={K}:"
        for name, flag in (("auto", 0), ("bn256", 0x1000 | 256), ("bn128", 0x1000 | 128), ("bn64", 0x1000 | 64), ("bn32", 0x1000 | 32)):
            for _ in range(2): L.gemm(a, w, out=out, force_bn=flag)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(5): L.gemm(a, w, out=out, force_bn=flag)
                torch.cuda.synchronize()
            ts = [ev.device_time for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and "gemm" in ev.name]
            line += f"  {name} {sorted(ts)[len(ts)//2]:.1f}us"
        print(line, flush=True)
