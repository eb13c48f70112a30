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
    ap.add_argument("--cpu-samples", type=int, default=32,
                    help="CPU baseline = BASELINE.json configs[0]: Smooth.predict with this N (~30 s at ~1 sample/s)")
    ap.add_argument("--sigma", type=float, default=SIGMA)
    ap.add_argument("--images-per-pass", type=int, default=0,
                    help="images certified together per step (0 = one per GPU: the per-rank batch stays N0+N rows)")
    ap.add_argument("--no-decode-sweep", dest="decode_sweep", action="store_false",
                    help="skip the max_new_tokens=20 sub-record (the reference's shipped decode budget)")
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


def cpu_reference_runner(cfg, max_new_tokens, sigma=SIGMA, predict=False):
    """The reference path on the host cores: oracle port of Smooth._sample_noise / Smooth.predict over the oracle
    MiniGPT-4 classifier (fp32, all host threads)."""
    from oracle import model_oracle as mo
    from oracle import smoothing_oracle as so
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = aliased_full_state_dict(cfg)
    prefix, suffix = prompt_ids(cfg.llm.vocab)
    clf = mo.MiniGPT4ClassifierOracle(sd, cfg, prefix, suffix, answer_table(cfg.llm.vocab, NUM_CLASSES),
                                      NUM_CLASSES, max_new_tokens=max_new_tokens)
    smooth = so.SmoothOracle(clf, NUM_CLASSES, sigma)
    x = synthetic_image(0, cfg.vit.img_size)

    def run(samples):
        t0 = time.perf_counter()
        if predict:
            run.last = int(smooth.predict(x, samples, ALPHA, 8))
        else:
            counts = smooth._sample_noise(x, samples, samples)
            assert counts.sum() == samples
        return time.perf_counter() - t0
    run.last = None
    return run, cores


def run_reference(args, cfg):
    """--impl reference: the reference's CPU path on the box's host cores.  The reference is Python and /root/reference
    does not exist on the GPU box, so what runs is the oracle PORT (fp32, all host threads) - a reported baseline."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 2 if args.steps + args.warmup > 6 else 4   # bounded: ~1 sample/s on 16 cores
    run, cores = cpu_reference_runner(cfg, args.max_new_tokens, args.sigma)
    for _ in range(args.warmup):
        run(per_step)
    t = sum(run(per_step) for _ in range(args.steps))
    sps = per_step * args.steps / t
    sample = (f"{per_step} noisy sample(s) per step through the oracle port (fp32 torch CPU, layer weights aliased), "
              f"same model shapes/prompt/max_new_tokens as the GPU arm")
    line = {
        "impl": "reference", "metric": "noisy_vlm_samples_per_sec", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "certified_images_per_min": sps * 60.0 / (args.n0 + args.n),
        "config": workload_config(args, cfg, args.gpus),
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def images_per_pass(args, world):
    return args.images_per_pass if args.images_per_pass > 0 else world


def workload_config(args, cfg, world):
    """Identical for both arms of one (--gpus, flags) invocation."""
    K = images_per_pass(args, world)
    per = (args.n0 + args.n + world - 1) // world
    which = ("BASELINE.json configs[1] (Smooth.certify on 1 B200, N0=100, N=1000, sigma=0.25, 1 image)" if world == 1 and K == 1
             else "BASELINE.json configs[2] scheme (a VQAv2-shaped synthetic subset, N=1000, noise samples of every image "
                  "sharded across the GPUs; --gpus 8 --sigma 0.5 --steps 8 is the 64-image subset itself)")
    return {
        "workload": (f"Smooth.certify of {K} synthetic {cfg.vit.img_size}x{cfg.vit.img_size} image(s) per step, N0={args.n0}, "
                     f"N={args.n}, sigma={args.sigma}, alpha={ALPHA}; MiniGPT-4 = EVA ViT-g/14 ({cfg.vit.depth}L) + Q-Former "
                     f"({cfg.qf.layers}L, {cfg.qf.n_query} queries) + Llama-2-7B shape ({cfg.llm.layers}L), random init; "
                     f"prompt {PREFIX_LEN}+{cfg.qf.n_query}+{SUFFIX_LEN} tokens; greedy decode max_new_tokens={args.max_new_tokens}; "
                     f"{NUM_CLASSES} classes"),
        "baseline_config": which,
        "batch_size": args.batch_size, "samples_per_step": K * (args.n0 + args.n), "images_per_step": K,
        "parallelism": (f"{K} image(s) per pass; the {args.n0 + args.n} draws of every image sharded over {world} GPU(s) "
                        f"({per} per rank and image); one int64 all-reduce of the {K} count-vector pair(s)"),
        "l2": "per-step working set (15.7 GB bf16 weights + GBs of activations) exceeds the 126 MB L2; no explicit flush",
    }


def calibrated_answer_table(eng, x_dev, sigma, draws, max_new_tokens, vocab):
    """Answer vocabulary from a calibration pass (as VQAv2's 3 129 answers come from its training answers): the
    `max_new_tokens`-token answers the model gives to `draws` noisy copies of the calibration image, ranked by
    frequency -> class ids 0..; every other answer -> "other".  With random-init weights EOS never fires, so a table of
    1-token answers alone would send every draw to "other" and leave the histogram degenerate."""
    from certifiedgpt_b200 import _lib as L
    patches = L.noise_patchify(x_dev, draws, sigma, seed=4242, stream_id=0xC0FFEE)
    tokens = eng.vit_forward(patches)
    queries, _ = eng.qformer_forward(tokens, want_llm_embeds=False)
    ids, _, _ = eng.llm_prefill_decode(queries)
    seqs = {}
    for row in ids.cpu().tolist():
        seqs[tuple(row)] = seqs.get(tuple(row), 0) + 1
    ranked = sorted(seqs.items(), key=lambda kv: (-kv[1], kv[0]))
    table = answer_table(vocab, NUM_CLASSES)
    table += [(seq, rank % (NUM_CLASSES - 1)) for rank, (seq, _) in enumerate(ranked) if len(seq) > 1]
    return table, {"calibration_draws": draws, "distinct_answers": len(ranked),
                   "top_answer_share": ranked[0][1] / draws if ranked else None}


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
    native = args.engine == "native"
    K = images_per_pass(args, world)
    assert native or K == 1, "several images per pass need the native engine (cgpt_certify_batch)"
    sigma = args.sigma

    sd = random_state_dict(cfg, seed=0, device=dev)
    prefix, suffix = prompt_ids(cfg.llm.vocab)
    long_new = max(args.max_new_tokens, 20) if (native and args.decode_sweep) else args.max_new_tokens
    # the packed weights (and rotary tables long enough for the reference's max_new_tokens = 20) are built once
    src = MiniGPT4Engine(cfg, sd, prefix, suffix, answer_table(cfg.llm.vocab, NUM_CLASSES), NUM_CLASSES,
                         max_new_tokens=long_new, device=dev, early_exit=True, use_graphs=native is False)
    del sd
    torch.cuda.empty_cache()
    eng = NativeMiniGPT4Engine.from_engine(src, max_new_tokens=args.max_new_tokens, early_exit=True) if native else src
    x_hosts = [synthetic_image(i, cfg.vit.img_size).pin_memory() for i in range(K)]
    x_devs = [x.to(dev) for x in x_hosts]
    per_image = args.n0 + args.n
    per_step = K * per_image
    calib = None
    if native:
        table, calib = calibrated_answer_table(eng, x_devs[0], sigma, min(per_image, args.batch_size),
                                               args.max_new_tokens, cfg.llm.vocab)
        eng.set_answer_table(table)
    smooth = Smooth(eng, NUM_CLASSES, sigma, seed=42, process_group=True if world > 1 else None)

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

    def certify_step(sm, xs):
        if K == 1:
            return [sm.certify(xs[0], args.n0, args.n, ALPHA, args.batch_size)]
        return sm.certify_batch(xs, args.n0, args.n, ALPHA, args.batch_size)

    def step_resident():
        return certify_step(smooth, x_devs)

    x_stage = torch.empty_like(x_devs[0])

    def step_e2e():
        if native:
            # HOST buffers straight through the C-ABI: cgpt_certify(_batch) stages the image(s) (H2D from pinned
            # memory) and reads back (label, radius, cAHat, pABar, nA) per image - both copies inside the timed region
            return certify_step(smooth, x_hosts)
        x_stage.copy_(x_hosts[0], non_blocking=True)      # H2D of the step's input from pinned memory
        return certify_step(smooth, [x_stage])            # D2H of (label, radius)

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
    detail = smooth.last_batch_detail[0] if K > 1 else {"counts_estimation": smooth.last_counts_estimation}
    counts_est = detail["counts_estimation"]
    decode_steps = eng.last_steps

    # roofline pass: the same steps launched eagerly (not as graph replays) so that every GEMM launch can be
    # bracketed by a CUDA-event pair on its stream; same kernels, same shapes, same data
    def set_graphs(e_, on):
        if native:
            e_.set_option("use_graphs", on)
        else:
            e_.use_graphs = on

    set_graphs(eng, False)
    step_resident()
    L.gemm_profile_start()
    ms_prof = timed(step_resident, args.steps)
    prof = L.gemm_profile_stop()
    set_graphs(eng, True)

    value = per_step * args.steps / (ms / 1e3)
    e2e = per_step * args.steps / (ms_e2e / 1e3)

    # the reference's shipped decode budget (max_new_tokens = 20, configs/eval_configs/vqav2_eval_noise_0.yaml:44) next to
    # the short-answer budget of the headline: same weights, same image(s), workspace re-bound for the longer KV cache
    decode = {f"max_new_tokens_{args.max_new_tokens}": {"samples_per_s": value, "ms_per_step": ms / args.steps,
                                                        "decode_steps_run": decode_steps}}
    if native and args.decode_sweep and long_new > args.max_new_tokens:
        try:
            eng._ws = None
            del smooth, eng
            torch.cuda.empty_cache()
            eng = NativeMiniGPT4Engine.from_engine(src, max_new_tokens=long_new, early_exit=True)
            sm20 = Smooth(eng, NUM_CLASSES, sigma, seed=42, process_group=True if world > 1 else None)
            steps20 = max(1, min(args.steps, 3))
            for _ in range(2):
                certify_step(sm20, x_devs)
            ms20 = timed(lambda: certify_step(sm20, x_devs), steps20)
            decode[f"max_new_tokens_{long_new}"] = {"samples_per_s": per_step * steps20 / (ms20 / 1e3),
                                                    "ms_per_step": ms20 / steps20, "decode_steps_run": eng.last_steps,
                                                    "steps_timed": steps20,
                                                    "note": "random-init weights never emit EOS: all 20 steps always run"}
        except Exception as ex:   # the headline must still be reported
            decode[f"max_new_tokens_{long_new}"] = {"failed": f"{type(ex).__name__}: {ex}"}

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
        img_bytes = x_hosts[0].numel() * 4
        line = {
            "metric": "noisy_vlm_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args, cfg, world),
            "certified_images_per_min": value * 60.0 / per_image,
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": K * img_bytes,
                    "d2h_bytes_per_step": K * (36 if native and K > 1 else (32 if native else 24)),
                    "certified_images_per_min": e2e * 60.0 / per_image,
                    "api": ((("Smooth.certify(x_host) -> cgpt_certify" if K == 1 else
                              "Smooth.certify_batch(x_hosts) -> cgpt_certify_batch") +
                             " (C-ABI, host images, host labels / radii)")
                            if native else "Smooth.certify(x_dev) after an explicit pinned H2D copy")},
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
            "result": {"label": result[0][0], "radius": result[0][1], "decode_steps": decode_steps,
                       "classes_hit": int((counts_est > 0).sum()), "top_count": int(counts_est.max()),
                       "answer_table": calib},
            "decode": decode,
            "engine": args.engine,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                # BASELINE.json configs[0] as stated: Smooth.predict on the host cores, 1 image + question, N = 32,
                # sigma = 0.25, 1-token answer head, batch 8 (the reference's own CPU-runnable case)
                run, cores = cpu_reference_runner(cfg, 1, 0.25, predict=True)
                run(1)
                dt = run(args.cpu_samples)
                line["cpu_baseline"] = {
                    "value": args.cpu_samples / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                    "sample": (f"BASELINE.json configs[0]: Smooth.predict, 1 synthetic 224x224 image + question, N = "
                               f"{args.cpu_samples}, sigma = 0.25, batch 8, 1-token answer head, through the oracle port "
                               f"(fp32 torch CPU, {cores} threads, layer weights aliased), after 1 warm-up sample; "
                               f"result = {run.last}")}
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
