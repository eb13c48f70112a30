"""BASELINE.json configs[4]: black-box attack inner loop at full size on one B200 - 8 steps x (100 CLIP ViT-L/14
queries + cosine scoring + perturbation update + Smooth.predict(N=100) through the full MiniGPT-4 victim).
Random-init weights, synthetic images.  python scripts/attack_bench.py [--no-victim]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from certifiedgpt_b200.attack import BlackBoxAttack, ClipVisionConfig, ClipVisionEngine

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

ccfg = ClipVisionConfig()
torch.manual_seed(0)
hf = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=ccfg.hidden, intermediate_size=ccfg.mlp,
                                                    num_hidden_layers=ccfg.layers, num_attention_heads=ccfg.heads,
                                                    image_size=224, patch_size=14, projection_dim=ccfg.proj))
clip = ClipVisionEngine(ccfg, hf.state_dict(), device=dev)
del hf
smooth = None
if "--no-victim" not in sys.argv:
    from certifiedgpt_b200.config import ModelConfig
    from certifiedgpt_b200.native import NativeMiniGPT4Engine
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    from certifiedgpt_b200.weights import random_state_dict
    cfg = ModelConfig.full(224)
    sd = random_state_dict(cfg, seed=0, device=dev)
    prefix, suffix = bench.prompt_ids(cfg.llm.vocab)
    victim = NativeMiniGPT4Engine(cfg, sd, prefix, suffix, bench.answer_table(cfg.llm.vocab, bench.NUM_CLASSES),
                                  bench.NUM_CLASSES, max_new_tokens=4, device=dev)
    del sd
    smooth = Smooth(victim, bench.NUM_CLASSES, 0.25, noise_space="pixel", seed=42)
atk = BlackBoxAttack(clip, smooth, steps=8, queries=100, predict_n=100, predict_batch=100, seed=0)
x = torch.rand(3, 224, 224, generator=torch.Generator().manual_seed(1000))
target = torch.rand(3, 224, 224, generator=torch.Generator().manual_seed(2000))
atk.run(x, target)                       # warm-up (graph capture of the victim at batch 100)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
adv, info = atk.run(x, target)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e)
# CLIP encoder alone: 100 queries
for _ in range(2):
    clip.encode_perturbed(x.to(dev), 100, 8 / 255, seed=1)
s.record()
for _ in range(5):
    clip.encode_perturbed(x.to(dev), 100, 8 / 255, seed=1)
e.record()
torch.cuda.synchronize()
clip_ms = s.elapsed_time(e) / 5
D, T, Lyr = ccfg.hidden, ccfg.tokens, ccfg.layers
gf = (Lyr * (12 * D * D * T * 2 + 4 * T * T * D) + 2 * 588 * D * 256 + 2 * D * ccfg.proj) / 1e9
print(json.dumps({"workload": "configs[4] attack inner loop: 8 steps x (100 CLIP ViT-L/14 queries + cosine + RGF update"
                              + (" + Smooth.predict N=100 through MiniGPT-4 full size)" if smooth else ")"),
                  "ms_per_attack": ms, "ms_per_step": ms / 8, "steps_per_s": 8e3 / ms,
                  "clip_ms_per_100_queries": clip_ms, "clip_images_per_s": 100e3 / clip_ms,
                  "clip_tflops": gf * 100 / clip_ms, "clip_gflop_per_image": gf,
                  "final_score": info["final_score"], "score_start": info["steps"][0]["score_before"],
                  "predictions": [r.get("smoothed_prediction") for r in info["steps"]]}))
