"""BASELINE.json config #4: encoder-only throughput sweep (noise -> EVA ViT-g/14 -> Q-Former -> llama_proj)
over noise batches 64..4096 against the bf16 tensor-core roofline (533.7 GFLOP per sample).
Under torchrun each rank runs the same per-rank batch (weak scaling); rank 0 prints the aggregate."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from certifiedgpt_b200.config import ModelConfig, LlmConfig
from certifiedgpt_b200.engine import MiniGPT4Engine
from certifiedgpt_b200.weights import random_state_dict

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = ModelConfig.full(224)
cfg.llm = LlmConfig(layers=1)     # the encoder sweep never runs the LLM; llama_proj keeps its 4096-wide output
sd = random_state_dict(cfg, seed=0, device=dev)
eng = MiniGPT4Engine(cfg, sd, [1], [3], [], 2, max_new_tokens=1, device=dev, use_graphs=False)
del sd
x = bench.synthetic_image(0, 224).to(dev)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops_sustained": 1379.5}
GF = 533.7
for B in [int(a) for a in sys.argv[1:]] or [64, 128, 256, 512, 1024, 2048, 4096]:
    for _ in range(2):
        eng.encode_noisy(x, B, 0.25, seed=1)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 3
    s.record()
    for i in range(iters):
        eng.encode_noisy(x, B, 0.25, seed=1, first_sample=i * B)
    e.record(); torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / iters], device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sps = world * B / (ms.item() / 1e3)
        tf = sps * GF / 1e3 / world
        print(f"encoder sweep: gpus={world} batch/gpu={B:5d}  {ms.item():9.2f} ms  {sps:9.1f} samples/s  "
              f"{tf:7.1f} TFLOP/s per GPU = {tf / peaks['bf16_tflops_sustained']:.2f} of measured sustained bf16 peak", flush=True)
if world > 1:
    dist.destroy_process_group()
