"""Noise-augmented fine-tune agent: the call site of the training step
(reference: MiniGPT4FineTuneAgent, agents/minigpt4_finetune_agent.py:60-260).

Same loop as the reference, minus the TPU / wandb / checkpoint plumbing (out of scope): per epoch
`train` (maybe_add_noise -> forward -> backward -> reduce + AdamW -> lr_scheduler.step, :149-195) then `eval`
(validation loss under no_grad, :197-230), best-loss tracking with patience (:104-114), loss history (:121).
A dataset item is {"image": [3,S,S] fp32, "answer_ids": list[int], "suffix_ids": list[int] (optional)}: the answer
text already tokenised with the Llama tokenizer and terminated by the end symbol (minigpt_base.py:297-311), and the
item's OWN instruction as the token ids after <ImageHere> (data/vqav2.py::finetune_items) - every sample is trained on
its instruction_input, as MiniGPTBase.forward does.  Answers are padded with -100; a batch holds items whose
instructions have the same token length (bucketing inside the epoch's shuffled order), so no prompt padding exists.
Data parallelism (DistributedSampler(shuffle=True) in the reference): every epoch's order is a seeded permutation, rank r
of W takes batches r, r + W, ... and draws its noise from its own Philox sample range; the gradient all-reduce then
averages DIFFERENT batches.  The learning rate of a step is the one the scheduler set AFTER the previous optimizer step
(lr_scheduler.step(epoch, step) follows optimizer_step in the reference): the first step runs at init_lr.
"""
import math

import torch

from ..train import LlamaProjTrainer


def linear_warmup_cosine_lr(cur_epoch, cur_step, *, max_epoch, iters_per_epoch, min_lr, init_lr, warmup_steps=0,
                            warmup_start_lr=-1.0, warmup_max_lr=0.0):
    """LinearWarmupCosineLRScheduler.step (graphs/models/minigpt4/common/optims.py:11-56), as a pure function."""
    total = cur_epoch * iters_per_epoch + cur_step
    if total < warmup_steps:
        start = warmup_start_lr if warmup_start_lr >= 0 else init_lr
        return min(warmup_max_lr, start + (warmup_max_lr - start) * cur_step / max(warmup_steps, 1))      # optims.py:68-73
    return (init_lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * total / (max_epoch * iters_per_epoch))) + min_lr  # :58-65


def collate(items, pad=-100):
    """-> (images [B,3,S,S], answers [B,na] padded with -100, suffix ids [B,ns] or None)."""
    na = max(len(it["answer_ids"]) for it in items)
    ans = torch.full((len(items), na), pad, dtype=torch.long)
    for i, it in enumerate(items):
        ans[i, :len(it["answer_ids"])] = torch.as_tensor(it["answer_ids"], dtype=torch.long)
    sfx = None
    if all(it.get("suffix_ids") is not None for it in items):
        lens = {len(it["suffix_ids"]) for it in items}
        assert len(lens) == 1, "a batch must hold instructions of one token length (the agent buckets by length)"
        sfx = torch.tensor([list(it["suffix_ids"]) for it in items], dtype=torch.int32)
    else:
        assert all(it.get("suffix_ids") is None for it in items), "either every item carries suffix_ids or none does"
    return torch.stack([it["image"] for it in items]).float(), ans, sfx


class MiniGPT4FineTuneAgent:
    name = "image_text_finetune"

    def __init__(self, engine, train_set, val_set=None, *, noise_level=0.25, batch_size=4, max_epoch=1, init_lr=1e-5,
                 min_lr=1e-6, warmup_steps=0, warmup_start_lr=1e-6, warmup_max_lr=1e-5, weight_decay=0.05, beta1=0.9,
                 beta2=0.999, patience=3, seed=42, max_answer=8, process_group=None):
        self.engine, self.train_set, self.val_set = engine, train_set, val_set
        self.noise_level, self.batch_size, self.max_epoch = float(noise_level), batch_size, max_epoch
        self.sched = dict(max_epoch=max_epoch, iters_per_epoch=max(1, len(train_set) // batch_size), min_lr=min_lr,
                          init_lr=init_lr, warmup_steps=warmup_steps, warmup_start_lr=warmup_start_lr,
                          warmup_max_lr=warmup_max_lr)
        self.trainer = LlamaProjTrainer(engine, lr=init_lr, betas=(beta1, beta2), weight_decay=weight_decay,
                                        max_batch=batch_size, max_answer=max_answer, process_group=process_group)
        self.patience, self.seed = patience, seed
        self.process_group = process_group
        self.rank, self.world = 0, 1
        if process_group is not None:
            import torch.distributed as dist
            g = None if process_group is True else process_group
            self.rank, self.world = dist.get_rank(g), dist.get_world_size(g)
        self._prev_lr = init_lr       # AdamW is built with lr = init_lr (create_optimizer :338-347)
        self.loss_history = {"train_loss": [], "val_loss": [], "lr": []}
        self.best_val_loss, self.best_state = float("inf"), None

    @classmethod
    def setup_agent(cls, **kwargs):
        return cls(**kwargs)

    def _batch_indices(self, ds, epoch, shuffle):
        """This rank's batches of the epoch, as index lists: a seeded permutation (train) or the dataset order (eval),
        bucketed by instruction length so that a batch needs no prompt padding, full batches only (drop_last=True, :332),
        dealt round-robin to the ranks."""
        n = len(ds)
        order = list(range(n))
        if shuffle:
            g = torch.Generator().manual_seed(self.seed * 1000003 + epoch)
            order = torch.randperm(n, generator=g).tolist()
        buckets, batches = {}, []
        for j in order:
            sfx = ds[j].get("suffix_ids")          # items are prepared dicts (data/vqav2.py::finetune_items)
            key = len(sfx) if sfx is not None else -1
            b = buckets.setdefault(key, [])
            b.append(j)
            if len(b) == self.batch_size:
                batches.append(b)
                buckets[key] = []
        batches = batches[:len(batches) // self.world * self.world]       # every rank runs the same number of steps
        return batches[self.rank::self.world]

    def _batches(self, ds, epoch=0, shuffle=False):
        for idx in self._batch_indices(ds, epoch, shuffle):
            yield collate([ds[j] for j in idx])

    def train(self, epoch):
        total, nb, lr = 0.0, 0, self._prev_lr
        for step, (images, answers, sfx) in enumerate(self._batches(self.train_set, epoch, shuffle=True)):
            # the reference calls lr_scheduler.step(epoch, step) AFTER optimizer_step (:176-178): this step runs at the
            # rate set after the previous one, the very first at AdamW's init_lr
            lr = self._prev_lr
            gstep = epoch * self.sched["iters_per_epoch"] + step    # Philox stream of the uniform noise
            loss = self.trainer.train_step(images.to(self.engine.dev), answers, self.noise_level, seed=self.seed, step=gstep,
                                           lr=lr, suffix_ids=sfx, sample_offset=self.rank * self.batch_size)
            self._prev_lr = linear_warmup_cosine_lr(epoch, step, **self.sched)
            total += float(loss.item())
            nb += 1
        self.loss_history["lr"].append(lr)
        return total / nb if nb else float("inf")

    @torch.no_grad()
    def eval(self, epoch):
        if self.val_set is None:
            return float("inf")
        total, nb = 0.0, 0
        nval = max(1, len(self.val_set) // self.batch_size)
        for i, (images, answers, sfx) in enumerate(self._batches(self.val_set)):
            # the reference's validation loop noises its inputs too (maybe_add_noise, :209); its own Philox streams
            step = (1 << 30) + epoch * nval + i
            total += float(self.trainer.forward(images.to(self.engine.dev), answers, self.noise_level, seed=self.seed,
                                                step=step, suffix_ids=sfx, sample_offset=self.rank * self.batch_size).item())
            nb += 1
        return total / nb if nb else float("inf")

    def run(self):
        wait = 0
        for epoch in range(self.max_epoch):
            tl = self.train(epoch)
            vl = self.eval(epoch)
            self.loss_history["train_loss"].append(tl)
            self.loss_history["val_loss"].append(vl)
            if vl < self.best_val_loss:                             # :104-108 (checkpoint = the two trained tensors)
                self.best_val_loss, wait = vl, 0
                self.best_state = {"llama_proj.weight": self.trainer.Wp.clone(), "llama_proj.bias": self.trainer.bp.clone()}
            else:
                wait += 1
            if self.val_set is not None and wait >= self.patience:
                break
        return self.loss_history

    def finalize(self):
        torch.cuda.synchronize()
