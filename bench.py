#!/usr/bin/env python
"""bench.py - Monte-Carlo smoothing hot path on B200 (contract: see the task's bench section).

Metric (BASELINE.json): noisy VLM samples/sec (and certified images/min at N=1000).
A "step" = one Smooth.certify of one synthetic 224x224 image: N0=100 + N=1000 noisy samples,
each = noise -> EVA ViT-g/14 -> Q-Former -> Llama-2-7B prefill + greedy short-answer decode ->
label -> histogram, then the Clopper-Pearson tail.  Random-init weights, synthetic image,
fixed question token ids (prefix 7 + 32 image + suffix 40 = 79 prompt tokens).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (libcgpt.so kernels)
  python bench.py --impl reference ...                          the reference's CPU path (oracle port)
Under torchrun (N>1) the N draws of each certify are sharded across ranks; only the int64 label
counts are all-reduced (NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N0, N, SIGMA, ALPHA = 100, 1000, 0.25, 0.001
NUM_CLASSES = 3130                   # VQAv2 answer vocabulary 3129 + "other"
PREFIX_LEN, SUFFIX_LEN = 7, 40
GF_PER_SAMPLE_ENCODER = 533.7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--max-new-tokens", type=int, default=4,
                    help="short-answer decode budget (BASELINE.md: ~3 decode steps per sample)")
    ap.add_argument("--batch-size", type=int, default=1100)
    ap.add_argument("--img-size", type=int, default=224)
    ap.add_argument("--n0", type=int, default=N0)
    ap.add_argument("--n", type=int, default=N)
    ap.add_argument("--cpu-samples", type=int, default=8,
                    help="bounded CPU-baseline sample (noisy samples; ~10 s of host work at ~1 sample/s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tiny", action="store_true", help="tiny model (debug only; not a bench number)")
    ap.add_argument("--engine", default="native", choices=["native", "python"],
                    help="native: the whole loop inside libcgpt (cgpt_certify); python: engine.py drives the kernels")
    return ap.parse_args()


def prompt_ids(vocab=32000):
    g = torch.Generator().manual_seed(7)
    prefix = [1] + torch.randint(3, vocab, (PREFIX_LEN - 1,), generator=g).tolist()
    suffix = torch.randint(3, vocab, (SUFFIX_LEN,), generator=g).tolist()
    return prefix, suffix


def synthetic_image(i, size, normalized=True):
    from certifiedgpt_b200 import _lib as L
    x = torch.rand(3, size, size, generator=torch.Generator().manual_seed(1000 + i))
    if normalized:
        m = torch.tensor(L.BLIP_MEAN).view(3, 1, 1)
        s = torch.tensor(L.BLIP_STD).view(3, 1, 1)
        x = (x - m) / s
    return x.contiguous()


def answer_table(vocab, num_classes):
    return [((t,), t % (num_classes - 1)) for t in range(3, vocab)]


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=5)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------- CPU reference
def aliased_full_state_dict(cfg):
    """Full-size fp32 random weights for the CPU oracle with every transformer layer ALIASING
    layer 0's tensors: identical shapes, FLOPs and bytes per layer, a fraction of the init time and
    RAM (timing sample only)."""
    from certifiedgpt_b200.weights import aliased_state_dict
    return aliased_state_dict(cfg, seed=0)


def cpu_reference_runner(cfg, max_new_tokens):
    """The reference path on the host cores: oracle port of Smooth._sample_noise over the oracle
    MiniGPT-4 classifier (fp32, all host threads)."""
    from oracle import model_oracle as mo
    from oracle import smoothing_oracle as so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = aliased_full_state_dict(cfg)
    prefix, suffix = prompt_ids(cfg.llm.vocab)
    clf = mo.MiniGPT4ClassifierOracle(sd, cfg, prefix, suffix, answer_table(cfg.llm.vocab, NUM_CLASSES),
                                      NUM_CLASSES, max_new_tokens=max_new_tokens)
    smooth = so.SmoothOracle(clf, NUM_CLASSES, SIGMA)
    x = synthetic_image(0, cfg.vit.img_size)

    def run(samples):
        t0 = time.perf_counter()
        counts = smooth._sample_noise(x, samples, samples)
        dt = time.perf_counter() - t0
        assert counts.sum() == samples
        return dt
    return run, cores


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 2 if args.steps + args.warmup > 6 else 4   # bounded: ~1 sample/s on 16 cores
    run, cores = cpu_reference_runner(cfg, args.max_new_tokens)
    for _ in range(args.warmup):
        run(per_step)
    t = sum(run(per_step) for _ in range(args.steps))
    sps = per_step * args.steps / t
    sample = (f"{per_step} noisy sample(s) per step through the oracle port (fp32 torch CPU, layer weights aliased), "
              f"same model shapes/prompt/max_new_tokens as the GPU arm")
    line = {
        "impl": "reference", "metric": "noisy_vlm_samples_per_sec", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "certified_images_per_min": sps * 60.0 / (args.n0 + args.n),
        "config": workload_config(args, cfg, 1),
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cfg, world):
    return {
        "workload": (f"Smooth.certify, 1 synthetic {cfg.vit.img_size}x{cfg.vit.img_size} image, N0={args.n0}, N={args.n}, "
                     f"sigma={SIGMA}, alpha={ALPHA}; MiniGPT-4 = EVA ViT-g/14 ({cfg.vit.depth}L) + Q-Former "
                     f"({cfg.qf.layers}L, {cfg.qf.n_query} queries) + Llama-2-7B shape ({cfg.llm.layers}L), random init; "
                     f"prompt {PREFIX_LEN}+{cfg.qf.n_query}+{SUFFIX_LEN} tokens; greedy decode max_new_tokens={args.max_new_tokens}; "
                     f"{NUM_CLASSES} classes"),
        "baseline_config": "BASELINE.json configs[1] (Smooth.certify on 1 B200, N0=100, N=1000, sigma=0.25)",
        "batch_size": args.batch_size, "samples_per_step": args.n0 + args.n,
        "parallelism": f"noise draws sharded over {world} GPU(s); int64 count all-reduce",
        "l2": "per-step working set (15.7 GB bf16 weights + GBs of activations) exceeds the 126 MB L2; no explicit flush",
    }


# ------------------------------------------------------------------------------- our arm
def run_ours(args, cfg):
    import torch.distributed as dist
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200.engine import MiniGPT4Engine
    from certifiedgpt_b200.native import NativeMiniGPT4Engine
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    from certifiedgpt_b200.weights import random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()

    sd = random_state_dict(cfg, seed=0, device=dev)
    prefix, suffix = prompt_ids(cfg.llm.vocab)
    Engine = NativeMiniGPT4Engine if args.engine == "native" else MiniGPT4Engine
    eng = Engine(cfg, sd, prefix, suffix, answer_table(cfg.llm.vocab, NUM_CLASSES), NUM_CLASSES,
                 max_new_tokens=args.max_new_tokens, device=dev, early_exit=True)
    del sd
    torch.cuda.empty_cache()
    smooth = Smooth(eng, NUM_CLASSES, SIGMA, seed=42, process_group=True if world > 1 else None)
    x_host = synthetic_image(0, cfg.vit.img_size).pin_memory()
    x_dev = x_host.to(dev)
    per_step = args.n0 + args.n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        return smooth.certify(x_dev, args.n0, args.n, ALPHA, args.batch_size)

    x_stage = torch.empty_like(x_dev)

    def step_e2e():
        if args.engine == "native":
            # HOST buffers straight through the C-ABI: cgpt_certify stages x (H2D from pinned memory) and reads
            # back (label, radius, cAHat, pABar, nA) - both copies inside the timed region
            return smooth.certify(x_host, args.n0, args.n, ALPHA, args.batch_size)
        x_stage.copy_(x_host, non_blocking=True)          # H2D of the step's input from pinned memory
        return smooth.certify(x_stage, args.n0, args.n, ALPHA, args.batch_size)   # D2H of (label, radius)

    for _ in range(args.warmup):
        step_resident()
    def n_launches():   # libcgpt counts eager launches and replayed graph nodes; the python engine counts its replays
        return L.launch_count() + getattr(eng, "replayed_launches", 0)

    launches0 = n_launches()
    with ClockSampler(local) as clocks:
        ms = timed(step_resident, args.steps)
    launches = n_launches() - launches0
    ms_e2e = timed(step_e2e, args.steps)
    result = step_resident()
    # roofline pass: the same steps launched eagerly (not as graph replays) so that every GEMM launch can be
    # bracketed by a CUDA-event pair on its stream; same kernels, same shapes, same data
    def set_graphs(on):
        if args.engine == "native":
            eng.set_option("use_graphs", on)
        else:
            eng.use_graphs = on

    set_graphs(False)
    step_resident()
    L.gemm_profile_start()
    ms_prof = timed(step_resident, args.steps)
    prof = L.gemm_profile_stop()
    set_graphs(True)

    value = per_step * args.steps / (ms / 1e3)
    e2e = per_step * args.steps / (ms_e2e / 1e3)

    # roofline of the dominant kernel (gemm_bf16_tcgen05_kernel, all launches of the timed region)
    gemm_ms = sum(p[0] for p in prof)
    gemm_fl = sum(p[1] for p in prof)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    by_shape = {}
    for t, fl, shp in prof:
        k = f"N{shp[1]}_K{shp[2]}"
        a = by_shape.setdefault(k, [0.0, 0.0, 0])
        a[0] += t; a[1] += fl; a[2] += 1
    top = sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:6]
    # DRAM traffic of the most expensive GEMM shape, from the committed ncu --set full capture
    traffic, traffic_note = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
        if top and top[0][0] in tr:
            t0 = tr[top[0][0]]
            traffic = t0["dram_read_bytes"] + t0["dram_write_bytes"]
            traffic_note = (f"{top[0][0]} (top shape, M={t0['M']}): {traffic / 1e9:.2f} GB DRAM per launch vs "
                            f"{t0['algorithmic_bytes'] / 1e9:.2f} GB algorithmic; {t0['capture']}")
    except Exception:
        pass

    if rank == 0:
        line = {
            "metric": "noisy_vlm_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args, cfg, world),
            "certified_images_per_min": value * 60.0 / per_step,
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": 32 if args.engine == "native" else 24,
                    "certified_images_per_min": e2e * 60.0 / per_step,
                    "api": ("Smooth.certify(x_host) -> cgpt_certify (C-ABI, host x, host label/radius)"
                            if args.engine == "native" else "Smooth.certify(x_dev) after an explicit pinned H2D copy")},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all GEMM launches of the timed steps)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "peak_source": peak_src, "traffic": traffic, "traffic_note": traffic_note,
                         "gemm_share_of_step": gemm_ms / ms_prof if ms_prof > 0 else None,
                         "measured_in": "eager replay of the timed steps (graph replays cannot be event-bracketed per kernel)",
                         "gemm_launches": len(prof),
                         "top_shapes": {k: {"ms": round(v[0], 3), "tflops": round(v[1] / (v[0] / 1e3) / 1e12, 1), "launches": v[2]}
                                        for k, v in top}},
            "result": {"label": result[0], "radius": result[1], "decode_steps": eng.last_steps},
            "engine": args.engine,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                run, cores = cpu_reference_runner(cfg, args.max_new_tokens)
                run(1)
                dt = run(args.cpu_samples)
                line["cpu_baseline"] = {
                    "value": args.cpu_samples / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                    "sample": (f"{args.cpu_samples} noisy samples of the same workload through the oracle port "
                               f"(fp32 torch CPU, {cores} threads, layer weights aliased), after 1 warm-up sample")}
            except Exception as ex:  # the GPU number must still be reported
                line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(ex).__name__}: {ex}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    from certifiedgpt_b200.config import ModelConfig
    cfg = ModelConfig.tiny() if args.tiny else ModelConfig.full(args.img_size)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
