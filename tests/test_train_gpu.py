"""GPU: the noise-augmented fine-tune step (certifiedgpt_b200/train.py + csrc/train_ops.cu).

* every backward / optimiser kernel against torch autograd / torch.optim on the same inputs;
* the whole step (forward with kept activations, backward through the frozen decoder, llama_proj gradient) against
  the CPU oracle's autograd over the restated reference forward (oracle.finetune_grads), bf16 tolerance;
* training on a fixed batch lowers the loss, and the generate path sees the updated llama_proj."""
import math

import pytest
import torch
import torch.nn.functional as F

from certifiedgpt_b200.config import LlmConfig, ModelConfig, QFormerConfig, VitConfig
from certifiedgpt_b200.weights import random_state_dict, round_to_bf16
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu

WIDE = ModelConfig(vit=VitConfig(img_size=56, depth=2), qf=QFormerConfig(layers=2),
                   llm=LlmConfig(hidden=512, layers=2, heads=4, inter=1024, vocab=512))


def _h():
    from certifiedgpt_b200 import _lib as L
    return L, L.load()


def test_swiglu_kernels_match_autograd():
    L, h = _h()
    g = torch.Generator(device="cuda").manual_seed(0)
    M, I = 37, 96
    gu = torch.randn(M, 2 * I, device="cuda", generator=g).bfloat16()
    dact = torch.randn(M, I, device="cuda", generator=g).bfloat16()
    act = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
    dgu = torch.empty_like(gu)
    L.check(h.cgpt_swiglu_fwd(L.ptr(gu), L.ptr(act), M, I, L.stream_ptr()))
    L.check(h.cgpt_swiglu_bwd(L.ptr(gu), L.ptr(dact), L.ptr(dgu), M, I, L.stream_ptr()))
    x = gu.float().requires_grad_(True)
    ref = F.silu(x[:, 0::2]) * x[:, 1::2]
    ref.backward(dact.float())
    assert (act.float() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    assert (dgu.float() - x.grad).abs().max().item() < 2e-2 * x.grad.abs().max().item()


def test_rmsnorm_bwd_matches_autograd_with_row_gather():
    L, h = _h()
    g = torch.Generator(device="cuda").manual_seed(1)
    B, T, D, na = 3, 9, 128, 2
    x = torch.randn(B * T, D, device="cuda", generator=g)
    gamma = torch.rand(D, device="cuda", generator=g) + 0.5
    for gather, rows in [((0, 0, 0), B * T), ((na, T, 4), B * na)]:
        dy = torch.randn(rows, D, device="cuda", generator=g)
        dx = torch.ones(B * T, D, device="cuda")
        L.check(h.cgpt_rmsnorm_bwd(L.ptr(x), D, L.ptr(gamma), L.ptr(dy), D, 1e-5, rows, D, L.ptr(dx), D, *gather, L.stream_ptr()))
        xr = x.clone().requires_grad_(True)
        idx = torch.arange(rows, device="cuda")
        src = idx if gather[0] == 0 else (idx // na) * T + 4 + idx % na
        xs = xr[src]
        hh = xs * torch.rsqrt(xs.pow(2).mean(-1, keepdim=True) + 1e-5) * gamma
        hh.backward(dy)
        assert (dx - (1.0 + xr.grad)).abs().max().item() < 1e-4


def test_rope_bwd_is_the_transpose_of_the_forward_rotation():
    L, h = _h()
    g = torch.Generator(device="cuda").manual_seed(2)
    B, T, H, hd, pos0 = 2, 5, 3, 32, 4
    D = H * hd
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.outer(torch.arange(64, dtype=torch.float32), inv)
    cos_t, sin_t = fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()
    dq = torch.randn(B * T, 3 * D, device="cuda", generator=g)
    out = torch.empty(B * T, 3 * D, device="cuda", dtype=torch.bfloat16)
    L.check(h.cgpt_rope_bwd_cast(L.ptr(dq), L.ptr(out), B * T, T, H, hd, pos0, L.ptr(cos_t), L.ptr(sin_t), L.stream_ptr()))
    x = torch.randn(B * T, 3 * D, device="cuda", generator=g).requires_grad_(True)
    pos = pos0 + torch.arange(B * T, device="cuda") % T
    c = torch.cat([cos_t[pos], cos_t[pos]], -1)[:, None, :]
    s = torch.cat([sin_t[pos], sin_t[pos]], -1)[:, None, :]
    rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., :hd // 2]], -1)
    parts = x.view(B * T, 3, H, hd)
    y = torch.stack([parts[:, 0] * c + rot(parts[:, 0]) * s, parts[:, 1] * c + rot(parts[:, 1]) * s, parts[:, 2]], 1)
    y.reshape(B * T, 3 * D).backward(dq)
    assert (out.float() - x.grad).abs().max().item() < 2e-2 * x.grad.abs().max().item()


@pytest.mark.parametrize("B,H,hd,Tq,P", [(2, 4, 32, 21, 3), (3, 4, 128, 60, 7), (1, 2, 128, 92, 7)])
def test_attention_bwd_matches_autograd(B, H, hd, Tq, P):
    L, h = _h()
    g = torch.Generator(device="cuda").manual_seed(Tq)
    D, Tk, rows = H * hd, P + Tq, P + Tq + 2
    q = torch.randn(B * Tq, D, device="cuda", generator=g).bfloat16()
    kc = torch.randn(B, rows, D, device="cuda", generator=g).bfloat16()
    vc = torch.randn(B, rows, D, device="cuda", generator=g).bfloat16()
    dout = torch.randn(B * Tq, D, device="cuda", generator=g).bfloat16()
    scale = hd ** -0.5
    qf = q.float().view(B, Tq, H, hd).transpose(1, 2).requires_grad_(True)
    kf = kc[:, :Tk].float().reshape(B, Tk, H, hd).transpose(1, 2).requires_grad_(True)
    vf = vc[:, :Tk].float().reshape(B, Tk, H, hd).transpose(1, 2).requires_grad_(True)
    sc = (qf @ kf.transpose(-1, -2)) * scale
    qi = torch.arange(Tq, device="cuda")[:, None] + P
    sc = sc.masked_fill(torch.arange(Tk, device="cuda")[None, :] > qi, float("-inf"))
    o = sc.softmax(-1) @ vf
    o.backward(dout.float().view(B, Tq, H, hd).transpose(1, 2))
    ob = o.detach().transpose(1, 2).reshape(B * Tq, D).bfloat16()
    dqkv = torch.full((B * Tq, 3 * D), float("nan"), device="cuda")
    L.check(h.cgpt_attention_bwd(L.ptr(q), D, L.ptr(kc), L.ptr(vc), D, rows, L.ptr(ob), D, L.ptr(dout), D, L.ptr(dqkv), B, H, hd,
                                 Tq, Tk, scale, L.stream_ptr()))
    torch.cuda.synchronize()
    assert not torch.isnan(dqkv).any()
    rq = qf.grad.transpose(1, 2).reshape(B * Tq, D)
    rk = kf.grad[:, :, P:].transpose(1, 2).reshape(B * Tq, D)
    rv = vf.grad[:, :, P:].transpose(1, 2).reshape(B * Tq, D)
    for got, ref in ((dqkv[:, :D], rq), (dqkv[:, D:2 * D], rk), (dqkv[:, 2 * D:], rv)):
        assert (got - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


def test_ce_grad_transpose_colsum_cast_adamw():
    L, h = _h()
    g = torch.Generator(device="cuda").manual_seed(3)
    R, V = 11, 512
    logits = torch.randn(R, V, device="cuda", generator=g) * 3
    tg = torch.randint(0, V, (R,), device="cuda", generator=g).int()
    tg[2] = -100
    tok, mc = torch.empty(R, device="cuda"), torch.empty(2, device="cuda")
    L.check(h.cgpt_ce_loss(L.ptr(logits), V, R, V, L.ptr(tg), L.ptr(tok), L.ptr(mc), L.stream_ptr()))
    dl = torch.empty(R, V, device="cuda", dtype=torch.bfloat16)
    L.check(h.cgpt_ce_grad(L.ptr(logits), V, R, V, L.ptr(tg), L.ptr(mc), L.ptr(dl), V, L.stream_ptr()))
    x = logits.clone().requires_grad_(True)
    F.cross_entropy(x, tg.long(), ignore_index=-100).backward()
    assert (dl.float() - x.grad).abs().max().item() < 1e-2 * x.grad.abs().max().item()
    assert (dl[2] == 0).all()
    # the reference's label-smoothed loss (CrossEntropyLoss(label_smoothing=0.1), modeling_llama.py:107) and its gradient
    L.check(h.cgpt_ce_loss_smooth(L.ptr(logits), V, R, V, L.ptr(tg), L.ptr(tok), L.ptr(mc), 0.1, L.stream_ptr()))
    L.check(h.cgpt_ce_grad_smooth(L.ptr(logits), V, R, V, L.ptr(tg), L.ptr(mc), L.ptr(dl), V, 0.1, L.stream_ptr()))
    x = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(x, tg.long(), ignore_index=-100, label_smoothing=0.1)
    ref.backward()
    assert abs(mc[0].item() - ref.item()) < 1e-4 and mc[1].item() == R - 1
    ref_tok = F.cross_entropy(logits, tg.long(), ignore_index=-100, label_smoothing=0.1, reduction="none")
    assert (tok - ref_tok).abs().max().item() < 1e-4
    assert (dl.float() - x.grad).abs().max().item() < 1e-2 * x.grad.abs().max().item()
    assert (dl[2] == 0).all()
    # transpose / column sum / gathered cast
    a = torch.randn(45, 70, device="cuda", generator=g).bfloat16()
    at = torch.zeros(70, 48, device="cuda", dtype=torch.bfloat16)
    L.check(h.cgpt_transpose_bf16(L.ptr(a), 70, L.ptr(at), 48, 45, 70, L.stream_ptr()))
    assert torch.equal(at[:, :45], a.t()) and (at[:, 45:] == 0).all()
    cs = torch.empty(70, device="cuda")
    L.check(h.cgpt_colsum_bf16(L.ptr(a), 70, 45, 70, L.ptr(cs), L.stream_ptr()))
    assert (cs - a.float().sum(0)).abs().max().item() < 1e-4
    src = torch.randn(12, 16, device="cuda", generator=g)
    dst = torch.empty(6, 16, device="cuda", dtype=torch.bfloat16)
    L.check(h.cgpt_cast_rows_f32_bf16(L.ptr(src), 16, L.ptr(dst), 16, 6, 16, 2, 4, 1, L.stream_ptr()))
    idx = torch.arange(6, device="cuda")
    assert torch.equal(dst, src[(idx // 2) * 4 + 1 + idx % 2].bfloat16())
    # AdamW against torch.optim.AdamW over three steps
    p0 = torch.randn(1000, device="cuda", generator=g)
    p, m, v = p0.clone(), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    pb = torch.empty(1000, device="cuda", dtype=torch.bfloat16)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    for step in range(1, 4):
        grad = torch.randn(1000, device="cuda", generator=g)
        L.check(h.cgpt_adamw_step(L.ptr(p), L.ptr(grad), L.ptr(m), L.ptr(v), L.ptr(pb), 1000, 1e-2, 0.9, 0.999, 1e-8, 0.05, step,
                                  1.0, L.stream_ptr()))
        ref.grad = grad.clone()
        opt.step()
    assert (p - ref.detach()).abs().max().item() < 1e-5
    assert torch.equal(pb, p.bfloat16())


def _trainer(cfg, seed, max_answer=4, **kw):
    from certifiedgpt_b200.engine import MiniGPT4Engine
    from certifiedgpt_b200.train import LlamaProjTrainer
    sd = round_to_bf16(random_state_dict(cfg, seed=seed))
    table = [((t,), t % 5) for t in range(3, cfg.llm.vocab)]
    eng = MiniGPT4Engine(cfg, sd, (1, 5, 6), (7, 8, 9, 10, 11), table, 6, max_new_tokens=4, use_graphs=False)
    return sd, eng, LlamaProjTrainer(eng, max_batch=4, max_answer=max_answer, **kw)


@pytest.mark.parametrize("name,cfg", [("tiny", ModelConfig.tiny()), ("wide", WIDE)])
def test_finetune_gradients_match_oracle_autograd(name, cfg):
    sd, eng, tr = _trainer(cfg, seed=41)
    S, V = cfg.vit.img_size, cfg.llm.vocab
    images = torch.randn(3, 3, S, S, generator=torch.Generator().manual_seed(6))
    answers = torch.tensor([[7, 9, 2, -100], [V - 1, 2, -100, -100], [12, 5, 6, 2]])
    loss = tr.forward(images.cuda(), answers, 0.0)
    gW, gb = tr.backward()
    torch.cuda.synchronize()
    ref_loss, rW, rb = mo.finetune_grads(sd, cfg, images, eng.prefix_ids, eng.suffix_ids, answers, label_smoothing=0.1)
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * max(1.0, ref_loss.item())
    for got, ref, nm in ((gW.cpu(), rW, "weight"), (gb.cpu(), rb, "bias")):
        cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        rel = (got - ref).abs().max().item() / ref.abs().max().item()
        assert cos > 0.995 and rel < 6e-2, (nm, cos, rel)
    # the inference-path loss (cgpt_lm_loss: fused SwiGLU epilogue, no activations kept) agrees with the training forward
    from certifiedgpt_b200.native import NativeMiniGPT4Engine
    nat = NativeMiniGPT4Engine.from_engine(eng)
    l2, _ = nat.lm_loss(images.cuda(), answers, 0.0)
    assert abs(l2.item() - loss.item()) < 1e-2 * max(1.0, loss.item())


def test_training_lowers_the_loss_and_updates_the_shared_projection():
    cfg = ModelConfig.tiny()
    sd, eng, tr = _trainer(cfg, seed=43, lr=2e-3, weight_decay=0.0)
    S = cfg.vit.img_size
    images = torch.rand(4, 3, S, S, generator=torch.Generator().manual_seed(9)).cuda()
    answers = torch.tensor([[20, 2, -100], [21, 22, 2], [23, 2, -100], [24, 25, 2]])
    w_before = eng.w["proj.w"].clone()
    losses = [tr.train_step(images, answers, 0.25, seed=1, step=s).item() for s in range(30)]
    assert losses[-1] < losses[0] - 0.3 and all(b < a + 0.05 for a, b in zip(losses[::5], losses[5::5])), losses[::5]
    assert not torch.equal(eng.w["proj.w"], w_before)                 # the engine's GEMM operand is the trained tensor
    assert torch.equal(eng.w["proj.w"], tr.Wp.bfloat16())
    # uniform noise is redrawn per step (Philox stream = step): same step -> same loss, different step -> different
    a = tr.forward(images, answers, 0.25, seed=1, step=3).item()
    b = tr.forward(images, answers, 0.25, seed=1, step=3).item()
    c = tr.forward(images, answers, 0.25, seed=1, step=4).item()
    assert a == b and a != c


def test_finetune_agent_loop_and_lr_schedule():
    from certifiedgpt_b200.agents import setup_agent
    from certifiedgpt_b200.agents.minigpt4_finetune_agent import linear_warmup_cosine_lr
    # the reference scheduler (optims.py:11-73) on its shipped settings: warm-up 1e-6 -> 1e-5 over 53 steps, then cosine
    kw = dict(max_epoch=4, iters_per_epoch=53, min_lr=1e-6, init_lr=1e-5, warmup_steps=53, warmup_start_lr=1e-6, warmup_max_lr=1e-5)
    assert linear_warmup_cosine_lr(0, 0, **kw) == pytest.approx(1e-6)
    assert linear_warmup_cosine_lr(0, 26, **kw) == pytest.approx(1e-6 + 9e-6 * 26 / 53)
    assert linear_warmup_cosine_lr(1, 0, **kw) == pytest.approx((1e-5 - 1e-6) * 0.5 * (1 + math.cos(math.pi * 53 / 212)) + 1e-6)
    assert linear_warmup_cosine_lr(3, 52, **kw) < 1.1e-6
    cfg = ModelConfig.tiny()
    sd, eng, _ = _trainer(cfg, seed=47)
    S = cfg.vit.img_size
    g = torch.Generator().manual_seed(3)
    train = [{"image": torch.rand(3, S, S, generator=g), "answer_ids": [20 + i % 3, 2]} for i in range(8)]
    val = [{"image": torch.rand(3, S, S, generator=g), "answer_ids": [20 + i % 3, 2]} for i in range(4)]
    agent = setup_agent("image_text_finetune", engine=eng, train_set=train, val_set=val, noise_level=0.25, batch_size=4,
                        max_epoch=6, init_lr=3e-3, min_lr=1e-3, warmup_steps=2, warmup_start_lr=1e-3, warmup_max_lr=3e-3,
                        weight_decay=0.0, max_answer=4)
    hist = agent.run()
    agent.finalize()
    assert len(hist["train_loss"]) >= 2 and hist["train_loss"][-1] < hist["train_loss"][0]
    assert agent.best_state is not None and agent.best_val_loss == min(hist["val_loss"])


def test_per_row_instructions_equal_one_question_at_a_time():
    """MiniGPTBase.forward trains every sample on ITS instruction (minigpt_base.py:323-362): a batch whose rows carry
    different suffix ids gives each row the token losses of a batch-of-one run under that question."""
    cfg = ModelConfig.tiny()
    sd, eng, tr = _trainer(cfg, seed=51)
    S = cfg.vit.img_size
    images = torch.rand(3, 3, S, S, generator=torch.Generator().manual_seed(2)).cuda()
    answers = torch.tensor([[20, 2, -100], [21, 22, 2], [23, 24, 2]])
    sfx = torch.tensor([[7, 8, 9], [30, 31, 32], [40, 41, 42]])
    tr.forward(images, answers, 0.25, seed=4, step=9, suffix_ids=sfx)
    tok = tr.last["tok"].clone()
    for b in range(3):
        eng.set_question(sfx[b].tolist())
        tr.forward(images[b:b + 1], answers[b:b + 1], 0.25, seed=4, step=9, sample_offset=b)
        assert torch.equal(tr.last["tok"][0], tok[b]), b
    assert not torch.equal(tok[0], tok[1])
    with pytest.raises(AssertionError):
        tr.forward(images, answers, 0.0, suffix_ids=torch.zeros(3, 9, dtype=torch.long))      # longer than the engine's


def test_finetune_agent_buckets_by_instruction_length_and_shards_by_rank():
    from certifiedgpt_b200.agents.minigpt4_finetune_agent import MiniGPT4FineTuneAgent, collate, linear_warmup_cosine_lr
    cfg = ModelConfig.tiny()
    sd, eng, _ = _trainer(cfg, seed=47)
    S = cfg.vit.img_size
    g = torch.Generator().manual_seed(3)
    lens = [3, 5, 3, 5, 5, 3, 3, 5, 4, 4]
    train = [{"image": torch.rand(3, S, S, generator=g), "answer_ids": [20 + i % 3, 2],
              "suffix_ids": list(range(7, 7 + lens[i]))} for i in range(10)]
    agent = MiniGPT4FineTuneAgent(eng, train, None, batch_size=2, max_epoch=2, init_lr=3e-3, min_lr=1e-3, warmup_steps=2,
                                  warmup_start_lr=1e-3, warmup_max_lr=3e-3, weight_decay=0.0, max_answer=4)
    b0 = agent._batch_indices(train, 0, shuffle=True)
    assert len(b0) == 5 and sorted(j for b in b0 for j in b) == list(range(10))
    assert all(len({lens[j] for j in b}) == 1 for b in b0)                       # one instruction length per batch
    assert b0 != agent._batch_indices(train, 1, shuffle=True)                    # reshuffled every epoch
    agent.rank, agent.world = 1, 2                                               # DistributedSampler-style sharding
    assert agent._batch_indices(train, 0, shuffle=True) == b0[:4][1::2]
    agent.rank, agent.world = 0, 1
    images, answers, sfx = collate([train[j] for j in b0[0]])
    assert sfx.shape == (2, lens[b0[0][0]]) and answers.shape == (2, 2)
    # the first step runs at AdamW's init_lr, step k at the rate the scheduler set after step k - 1 (:176-178)
    seen = []
    orig = agent.trainer.train_step
    agent.trainer.train_step = lambda *a, **kw: (seen.append(kw["lr"]), orig(*a, **kw))[1]
    agent.train(0)
    want = [3e-3] + [linear_warmup_cosine_lr(0, s, **agent.sched) for s in range(4)]
    assert seen == pytest.approx(want)
