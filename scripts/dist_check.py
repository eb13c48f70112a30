"""Multi-GPU check (run under torchrun, one rank per GPU): the counts of a Smooth.certify whose
draws are sharded over W ranks equal the single-GPU counts bit for bit (noise keyed by global
sample index; only the int64 count vector is all-reduced over NCCL).  Default: the native engine
(cgpt_certify shards, histograms and all-reduces inside libcgpt through its own NCCL communicator);
--python: engine.py + torch.distributed.all_reduce."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from certifiedgpt_b200.config import ModelConfig
from certifiedgpt_b200.engine import MiniGPT4Engine
from certifiedgpt_b200.native import NativeMiniGPT4Engine
from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
from certifiedgpt_b200.weights import random_state_dict, round_to_bf16


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    full = "--full" in sys.argv
    cfg = ModelConfig.full(224) if full else ModelConfig.tiny()
    sd = random_state_dict(cfg, seed=0, device=dev) if full else round_to_bf16(random_state_dict(cfg, seed=3))
    V = cfg.llm.vocab
    g = torch.Generator().manual_seed(7)
    prefix = [1] + torch.randint(3, V, (6,), generator=g).tolist()
    suffix = torch.randint(3, V, (12,), generator=g).tolist()
    table = [((t,), t % 9) for t in range(3, V)]
    Engine = MiniGPT4Engine if "--python" in sys.argv else NativeMiniGPT4Engine
    eng = Engine(cfg, sd, prefix, suffix, table, 10, max_new_tokens=1, device=dev)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(1000)).to(dev)
    n0, n = (16, 96) if full else (100, 1000)
    sharded = Smooth(eng, 10, 0.25, seed=42, process_group=True)
    res = sharded.certify(x, n0, n, 0.001, 64)
    sel, est = sharded.last_counts_selection.clone(), sharded.last_counts_estimation.clone()
    single = Smooth(eng, 10, 0.25, seed=42)           # every rank recomputes the whole range alone
    res1 = single.certify(x, n0, n, 0.001, 64)
    ok = (res == res1 and torch.equal(sel, single.last_counts_selection)
          and torch.equal(est, single.last_counts_estimation) and int(est.sum()) == n)
    # several images per pass (cgpt_certify_batch): draws of EVERY image sharded over the ranks, one all-reduce of all the
    # count-vector pairs; per image the same counts / label / radius as single-GPU, one-image-at-a-time certification
    if Engine is NativeMiniGPT4Engine:
        xs = [torch.rand(3, S, S, generator=torch.Generator().manual_seed(2000 + i)).to(dev) for i in range(3)]
        many = Smooth(eng, 10, 0.25, seed=43, process_group=True)
        got = many.certify_batch(xs, n0, n, 0.001, 96)
        one = Smooth(eng, 10, 0.25, seed=43)
        for k, xk in enumerate(xs):
            want = one.certify(xk, n0, n, 0.001, 64)
            d = many.last_batch_detail[k]
            ok = ok and got[k] == want and torch.equal(d["counts_selection"], one.last_counts_selection) \
                and torch.equal(d["counts_estimation"], one.last_counts_estimation)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # data-parallel fine-tune step: every rank trains on its own images, the llama_proj gradient is averaged over the
    # ranks through libcgpt's NCCL communicator -> all ranks hold identical weights afterwards, equal to one process
    # that averages the per-rank gradients itself
    if not full:
        from certifiedgpt_b200.train import LlamaProjTrainer
        py = MiniGPT4Engine(cfg, sd, prefix, suffix, table, 10, max_new_tokens=4, device=dev, use_graphs=False)
        tr = LlamaProjTrainer(py, lr=1e-3, weight_decay=0.0, max_batch=2, max_answer=3, process_group=True)
        gi = torch.Generator().manual_seed(77)
        all_images = torch.rand(world * 2, 3, S, S, generator=gi)
        all_answers = torch.randint(3, V, (world * 2, 3), generator=gi)
        tr.forward(all_images[rank * 2:rank * 2 + 2].to(dev), all_answers[rank * 2:rank * 2 + 2], 0.0)
        gW, gb = tr.backward()
        local_gW = gW.clone()
        tr.optimizer_step()
        gathered = [torch.empty_like(local_gW) for _ in range(world)]
        dist.all_gather(gathered, local_gW)
        mean_gW = torch.stack(gathered).mean(0)
        same_w = [torch.empty_like(tr.Wp) for _ in range(world)]
        dist.all_gather(same_w, tr.Wp)
        train_ok = (all(torch.equal(same_w[0], t) for t in same_w)
                    and torch.allclose(tr.grads["W"] / world, mean_gW, rtol=1e-5, atol=1e-7))
        ok = ok and train_ok
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"data-parallel fine-tune step over {world} ranks: weights identical on all ranks and gradient = mean of the "
                  f"per-rank gradients: {'TRAIN_DIST_OK' if train_ok else 'TRAIN_DIST_MISMATCH'}", flush=True)
    if rank == 0:
        print(f"engine={Engine.__name__} world={world} sharded={res} single={res1} counts_est={est.tolist()} "
              f"{'DIST_OK' if flag.item() == 1 else 'DIST_MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
