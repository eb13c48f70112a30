"""Full-size noise-augmented fine-tune step on one B200: MiniGPT-4 (ViT-g 39L + Q-Former 12L +
Llama-2-7B shape 32L, random init), batch of B images, answers of 8 tokens, uniform noise 0.25; llama_proj trains.
python scripts/finetune_bench.py [B]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from certifiedgpt_b200.config import ModelConfig
from certifiedgpt_b200.engine import MiniGPT4Engine
from certifiedgpt_b200.train import LlamaProjTrainer
from certifiedgpt_b200.weights import random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = ModelConfig.full(224)
sd = random_state_dict(cfg, seed=0, device=dev)
prefix, suffix = bench.prompt_ids(cfg.llm.vocab)
eng = MiniGPT4Engine(cfg, sd, prefix, suffix, bench.answer_table(cfg.llm.vocab, bench.NUM_CLASSES), bench.NUM_CLASSES,
                     max_new_tokens=8, device=dev, use_graphs=False)
del sd
torch.cuda.empty_cache()
tr = LlamaProjTrainer(eng, lr=1e-5, max_batch=B, max_answer=8)
g = torch.Generator().manual_seed(0)
images = torch.rand(B, 3, 224, 224, generator=g).to(dev)
answers = torch.randint(3, 32000, (B, 8), generator=g)
answers[:, -1] = 2
for s in range(2):
    tr.train_step(images, answers, 0.25, seed=1, step=s)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
n = 5
tf = tb = ts = 0.0
losses = []
for s in range(n):
    ev[0].record()
    loss = tr.forward(images, answers, 0.25, seed=1, step=2 + s)
    ev[1].record()
    tr.backward()
    ev[2].record()
    tr.optimizer_step()
    ev[3].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2]); ts += ev[2].elapsed_time(ev[3])
    losses.append(float(loss.item()))
rows = B * (32 + len(suffix) + 8)
print(json.dumps({"workload": f"fine-tune step, full-size MiniGPT-4, batch {B} images, 8 answer tokens, uniform noise 0.25, "
                              f"{rows} decoder rows; llama_proj (3.1 M parameters) trains",
                  "ms_forward": tf / n, "ms_backward": tb / n, "ms_optimizer": ts / n, "ms_step": (tf + tb + ts) / n,
                  "images_per_s": B * n * 1e3 / (tf + tb + ts), "losses": losses,
                  "mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
