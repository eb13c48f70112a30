"""CUPTI kernel-time breakdown of one full-size fine-tune step (forward + backward), batch B."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from certifiedgpt_b200.config import ModelConfig
from certifiedgpt_b200.engine import MiniGPT4Engine
from certifiedgpt_b200.train import LlamaProjTrainer
from certifiedgpt_b200.weights import random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda", 0)
cfg = ModelConfig.full(224)
sd = random_state_dict(cfg, seed=0, device=dev)
prefix, suffix = bench.prompt_ids(cfg.llm.vocab)
eng = MiniGPT4Engine(cfg, sd, prefix, suffix, bench.answer_table(cfg.llm.vocab, bench.NUM_CLASSES), bench.NUM_CLASSES,
                     max_new_tokens=8, device=dev, use_graphs=False)
del sd
tr = LlamaProjTrainer(eng, lr=1e-5, max_batch=B, max_answer=8)
g = torch.Generator().manual_seed(0)
images = torch.rand(B, 3, 224, 224, generator=g).to(dev)
answers = torch.randint(3, 32000, (B, 8), generator=g)
for s in range(2):
    tr.train_step(images, answers, 0.25, seed=1, step=s)
torch.cuda.synchronize()
for phase in ("forward", "backward"):
    if phase == "backward":
        tr.forward(images, answers, 0.25, seed=1, step=5)
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        (tr.forward(images, answers, 0.25, seed=1, step=5) if phase == "forward" else tr.backward())
        torch.cuda.synchronize()
    agg = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            n = ev.name.split("(")[0].replace("void ", "").replace("cgpt::", "")[:60]
            a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tot = sum(v[1] for v in agg.values())
    print(f"== {phase}: total kernel time {tot / 1e3:.2f} ms (batch {B})")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
        print(f"{v[1] / 1e3:9.2f} ms {100 * v[1] / tot:5.1f}% n={v[0]:5d}  {k}")
