"""K1 throughput: algorithmic bytes (303104 B written per sample) / CUDA-event time."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from certifiedgpt_b200 import _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
x = torch.rand(3, 224, 224, device="cuda")
out = torch.empty(B * 256, 592, dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6546.2
for mode, eps in (("philox", None), ("injected", torch.randn(B, 3, 224, 224, device="cuda"))):
    for _ in range(3):
        L.noise_patchify(x, B, 0.25, eps=eps, seed=1, out=out)
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); L.noise_patchify(x, B, 0.25, eps=eps, seed=1, out=out); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    byts = B * 256 * 592 * 2 + (B * 3 * 224 * 224 * 4 if eps is not None else 0)
    print(f"K1 {mode}: B={B} {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s algorithmic = {byts / ms / 1e6 / peak:.2f} of measured HBM peak ({peak} GB/s)")
