"""Per-shape GEMM times of the encoder-only path (noise -> ViT-g -> Q-Former -> llama_proj) at one batch size, from the
event pairs libcgpt puts around every eager GEMM launch, next to the whole-pass time: what the GEMMs take, what is left
for attention / LayerNorm / the rest.  usage: python scripts/encoder_profile.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from certifiedgpt_b200 import _lib as L
from certifiedgpt_b200.config import LlmConfig, ModelConfig
from certifiedgpt_b200.engine import MiniGPT4Engine
from certifiedgpt_b200.weights import random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
cfg = ModelConfig.full(224)
cfg.llm = LlmConfig(layers=1)
sd = random_state_dict(cfg, seed=0, device=dev)
eng = MiniGPT4Engine(cfg, sd, [1], [3], [], 2, max_new_tokens=1, device=dev, use_graphs=False)
del sd
x = bench.synthetic_image(0, 224).to(dev)
for _ in range(2):
    eng.encode_noisy(x, B, 0.25, seed=1)
torch.cuda.synchronize()
iters = 3
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(iters):
    eng.encode_noisy(x, B, 0.25, seed=1, first_sample=i * B)
e.record()
torch.cuda.synchronize()
total = s.elapsed_time(e) / iters
L.gemm_profile_start()
for i in range(iters):
    eng.encode_noisy(x, B, 0.25, seed=1, first_sample=i * B)
prof = L.gemm_profile_stop()
by = {}
for ms, fl, shp in prof:
    a = by.setdefault((shp[1], shp[2]), [0.0, 0.0, 0])
    a[0] += ms / iters; a[1] += fl / iters; a[2] += 1
gemm = sum(v[0] for v in by.values())
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("CGPT_"))
print(f"encoder profile batch {B} [{tag or 'defaults'}]: pass {total:.1f} ms = {B / total * 1e3:.0f} samples/s; GEMMs {gemm:.1f} ms "
      f"({100 * gemm / total:.1f} %), everything else {total - gemm:.1f} ms")
for (n, k), v in sorted(by.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  N={n:5d} K={k:5d}: {v[0]:8.2f} ms per pass  {v[1] / v[0] / 1e9:7.1f} TFLOP/s  ({v[2] // iters} launches)")
