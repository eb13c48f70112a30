"""Kernel-time breakdown of one full-size noisy batch (CUPTI via torch.profiler; no ncu replay)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from certifiedgpt_b200.config import ModelConfig
from certifiedgpt_b200.engine import MiniGPT4Engine
from certifiedgpt_b200.weights import random_state_dict
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cfg = ModelConfig.full(224)
sd = random_state_dict(cfg, seed=0, device="cuda")
prefix, suffix = bench.prompt_ids()
eng = MiniGPT4Engine(cfg, sd, prefix, suffix, bench.answer_table(32000, 3130), 3130, max_new_tokens=4)
del sd
x = bench.synthetic_image(0, 224).cuda()
for _ in range(2):
    eng.noisy_labels(x, B, 0.25, seed=1)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); eng.noisy_labels(x, B, 0.25, seed=1); e.record(); torch.cuda.synchronize()
print(f"batch of {B}: {s.elapsed_time(e):.1f} ms", flush=True)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.noisy_labels(x, B, 0.25, seed=1)
    torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = ev.name.split("(")[0].replace("void ", "").replace("cgpt::", "")[:60]
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot/1e3:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{v[1]/1e3:9.2f} ms {100*v[1]/tot:5.1f}% n={v[0]:5d}  {k}")
