"""Noise-augmented fine-tune step of MiniGPT-4.

Reference: MiniGPT4FineTuneAgent.train (agents/minigpt4_finetune_agent.py:142-195): uniform image noise
(`maybe_add_noise`), `loss = model(batch)["loss"]` (MiniGPTBase.forward, minigpt_base.py:323-362; shifted
cross-entropy with label_smoothing=0.1, modeling_llama.py:101-123), `loss.backward()`, gradient reduction across the data-parallel ranks
(`xm.reduce_gradients`) and `torch.optim.AdamW` (create_optimizer :338-347; lr 1e-5, weight decay 0.05, betas
(0.9, 0.999) in configs/train_configs/vqav2_finetuning_noise_*.yaml).  The ViT, the Q-Former and the Llama are
frozen (base_model.py:162-172,238-240; minigpt4.py:111-117): the ONLY trainable tensors are llama_proj.weight / .bias
(minigpt4.py:76-78), but their gradient needs the data gradient through all 32 frozen decoder layers.

Host orchestration only; every number is produced by libcgpt.so:
  forward   K1 (uniform noise, one image per row) -> frozen ViT / Q-Former (the engine's kernels) -> llama_proj GEMM ->
            [image | suffix | answer] rows through the decoder with the activations of every layer kept
            (fused rotary / KV-append QKV epilogue, persistent prefill attention, SwiGLU kernel) -> final norm + lm_head on
            the answer-predicting rows -> cgpt_ce_loss
  backward  cgpt_ce_grad -> lm_head dgrad GEMM -> per layer, reversed: down / gate-up / o_proj / qkv data-gradient GEMMs
            (the tcgen05 GEMM on transposed weight copies), cgpt_swiglu_bwd, cgpt_rmsnorm_bwd, cgpt_attention_bwd,
            cgpt_rope_bwd_cast -> llama_proj weight gradient = one GEMM over transposed operands, bias gradient = column sum
  step      (all-reduce of the 3.1 M gradient values over the ranks) -> cgpt_adamw_step on fp32 master weights
"""
import math

import torch

from . import _lib as L
from .engine import MiniGPT4Engine


def _lib():
    return L.load()


class LlamaProjTrainer:
    """One object = optimizer state + transposed frozen weights + activation storage for batches of up to
    `max_batch` images with up to `max_answer` answer tokens."""

    def __init__(self, engine: MiniGPT4Engine, *, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05,
                 max_batch=4, max_answer=8, process_group=None, label_smoothing=0.1):
        self.eng = engine
        self.label_smoothing = float(label_smoothing)   # CrossEntropyLoss(label_smoothing=0.1), modeling_llama.py:107
        self.cfg, self.dev, self.w = engine.cfg, engine.dev, engine.w
        self.lr, self.betas, self.eps, self.wd = float(lr), betas, float(eps), float(weight_decay)
        self.process_group = process_group
        self._comm = None
        self.step_count = 0
        l, q = self.cfg.llm, self.cfg.qf
        self.P, self.ns = engine.P, len(engine.suffix_ids)
        self.max_batch, self.max_answer = max_batch, max_answer
        dev, bf, f32 = self.dev, torch.bfloat16, torch.float32
        # trainable parameters: fp32 master copies (the bf16 GEMM operand engine.w["proj.w"] is refreshed by AdamW)
        self.Wp = self.w["proj.w"].float().contiguous()
        self.bp = self.w["proj.b"].float().contiguous()
        self.state = {n: (torch.zeros_like(p), torch.zeros_like(p)) for n, p in (("W", self.Wp), ("b", self.bp))}
        # transposed copies of the frozen decoder weights for the data-gradient GEMMs (dX = dY . W = dY . (W^T)^T)
        self.wt = {}
        for i in range(l.layers):
            o = f"llm.{i}."
            for n in ("qkv.w", "o.w", "gu.w", "down.w"):
                self.wt[o + n] = self.w[o + n].t().contiguous()
        self.wt["llm.head"] = self.w["llm.head"].t().contiguous()
        # rope tables must cover prefix + image + suffix + answers
        Tl = q.n_query + self.ns + max_answer
        need = self.P + Tl + 1
        if self.w["rope.cos"].shape[0] < need:
            inv = 1.0 / (l.rope_theta ** (torch.arange(0, l.head_dim, 2, dtype=f32) / l.head_dim))
            fr = torch.outer(torch.arange(need, dtype=f32), inv)
            self.cos, self.sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
        else:
            self.cos, self.sin = self.w["rope.cos"], self.w["rope.sin"]
        M = max_batch * Tl
        Hd, I = l.hidden, l.inter
        e = lambda *s, dtype=bf: torch.empty(*s, dtype=dtype, device=dev)
        self.cache_rows = self.P + Tl
        self.buf = {
            "res": e(M, Hd, dtype=f32), "xn": e(M, Hd), "qkv": e(M, 3 * Hd), "act": e(M, I),
            "kc": e(l.layers, max_batch, self.cache_rows, Hd), "vc": e(l.layers, max_batch, self.cache_rows, Hd),
            # saved per layer
            "x_in": e(l.layers, M, Hd, dtype=f32), "x_mid": e(l.layers, M, Hd, dtype=f32), "q": e(l.layers, M, Hd),
            "att": e(l.layers, M, Hd), "gu": e(l.layers, M, 2 * I),
            # backward scratch
            "dx": e(M, Hd, dtype=f32), "dxb": e(M, Hd), "dact": e(M, I), "dgu": e(M, 2 * I), "dh": e(M, Hd, dtype=f32),
            "datt": e(M, Hd), "dqkv": e(M, 3 * Hd, dtype=f32), "dqkvb": e(M, 3 * Hd),
        }
        if self.P > 0:
            self.buf["kc"][:, :, :self.P] = engine.kp[:, None, :self.P]
            self.buf["vc"][:, :, :self.P] = engine.vp[:, None, :self.P]
        self.last = {}

    # ------------------------------------------------------------------ thin kernel wrappers
    @staticmethod
    def _ck(rc):
        L.check(rc)

    def _rmsnorm_bwd(self, x, gamma, dy, eps, dx, rows, gather=(0, 0, 0)):
        self._ck(_lib().cgpt_rmsnorm_bwd(L.ptr(x), x.stride(0), L.ptr(gamma), L.ptr(dy), dy.stride(0), float(eps), rows,
                                         x.shape[1], L.ptr(dx), dx.stride(0), *gather, L.stream_ptr()))

    def _cast(self, src, dst, rows, gather=(0, 0, 0)):
        self._ck(_lib().cgpt_cast_rows_f32_bf16(L.ptr(src), src.stride(0), L.ptr(dst), dst.stride(0), rows, src.shape[1],
                                                *gather, L.stream_ptr()))

    # ------------------------------------------------------------------ forward (activations kept)
    @torch.no_grad()
    def forward(self, images, answers, noise_level=0.0, *, seed=0, step=0, noise_kind=L.NOISE_UNIFORM,
                noise_space=L.SPACE_NORMALIZED, suffix_ids=None, sample_offset=0):
        """images [B,3,S,S] fp32 CUDA, answers [B,na] int (-100 = padding) -> mean loss (device scalar).
        suffix_ids [B, ns] int (optional): every row's OWN instruction as the token ids after the image
        (MiniGPTBase.forward trains each sample on its instruction_input, minigpt_base.py:323-362); rows of one batch
        share the length ns <= the engine's (the agent buckets by length).  None: the engine's current question for
        every row.  sample_offset: first Philox sample index of this batch (data-parallel ranks draw different noise)."""
        eng, cfg, w, buf = self.eng, self.cfg, self.w, self.buf
        v, q, l = cfg.vit, cfg.qf, cfg.llm
        B, na = answers.shape
        assert B <= self.max_batch and na <= self.max_answer and images.is_cuda
        if suffix_ids is not None:
            sfx = torch.as_tensor(suffix_ids).to(device=self.dev, dtype=torch.int32).contiguous()
            assert sfx.dim() == 2 and sfx.shape[0] == B and sfx.shape[1] <= self.ns, "suffix_ids must be [B, ns <= engine's]"
            assert int(sfx.min()) >= 0 and int(sfx.max()) < l.vocab, "suffix ids outside the vocabulary"
            ns, sfx_ids, sfx_period = sfx.shape[1], sfx.view(-1), B * sfx.shape[1]
        else:
            ns, sfx_ids, sfx_period = len(eng.suffix_ids), eng.suffix_ids_dev, len(eng.suffix_ids)
        assert int(torch.as_tensor(answers).max()) < l.vocab, "answer ids outside the vocabulary"
        nq, P, Hd, I = q.n_query, self.P, l.hidden, l.inter
        Tl = nq + ns + na
        M = B * Tl
        lib = _lib()
        # frozen encoder (forward only): K1 per image -> ViT -> Q-Former
        eb = eng._encoder_buffers(B)
        G2 = v.grid * v.grid
        for b in range(B):
            L.noise_patchify(images[b].float().contiguous(), 1, float(noise_level), seed=seed, stream_id=step,
                             first_sample=sample_offset + b, noise_space=noise_space, noise_kind=noise_kind,
                             out=eb["patches"][b * G2:(b + 1) * G2])
        qo = eng.qformer(B, eb, eng.vit_from_patches(B, eb))              # [B*nq, qf.hidden] bf16
        self.last_q = qo[:B * nq].clone()
        res, xn, qkv, act = (buf[k][:M] for k in ("res", "xn", "qkv", "act"))
        # llama_proj (trainable) into the image rows; suffix / answer embeddings behind it
        L.gemm(self.last_q, w["proj.w"], bias=w["proj.b"], out=res, row_period=nq, remap_stride=Tl, remap_offset=0)
        if ns > 0:
            L.gather_rows(w["emb"], sfx_ids, B * ns, res, id_period=sfx_period, remap=(ns, Tl, nq))
        ans = answers.to(device=self.dev, dtype=torch.int32).contiguous()
        ans_in = torch.where(ans < 0, torch.full_like(ans, l.pad_id), ans).contiguous()
        L.gather_rows(w["emb"], ans_in.view(-1), B * na, res, id_period=B * na, remap=(na, Tl, nq + ns))
        scale = 1.0 / math.sqrt(l.head_dim)
        fused = l.head_dim == 128
        for i in range(l.layers):
            o = f"llm.{i}."
            kc, vc = buf["kc"][i, :B], buf["vc"][i, :B]
            buf["x_in"][i, :M].copy_(res)
            L.norm_rows(res, w[o + "n1"], None, l.rms_eps, xn, rms=True)
            if fused:
                L.gemm(xn, w[o + "qkv.w"], out=qkv, rope=dict(T=Tl, heads=l.heads, pos0=P, cos=self.cos, sin=self.sin, kcache=kc,
                                                              vcache=vc, cache_rows=self.cache_rows, cache_row0=P))
            else:
                L.gemm(xn, w[o + "qkv.w"], out=qkv)
                L.rope_split(qkv, Tl, l.heads, l.head_dim, P, self.cos, self.sin, kc, vc, self.cache_rows, P)
            buf["q"][i, :M].copy_(qkv[:, :Hd])
            att = buf["att"][i, :M]
            L.attention(qkv[:, :Hd], kc.reshape(-1, Hd), vc.reshape(-1, Hd), att, B=B, H=l.heads, Tq=Tl, Tk=P + Tl,
                        head_dim=l.head_dim, scale=scale, kv_rows_per_batch=self.cache_rows, causal=True)
            L.gemm(att, w[o + "o.w"], resid=res, out=res)
            buf["x_mid"][i, :M].copy_(res)
            L.norm_rows(res, w[o + "n2"], None, l.rms_eps, xn, rms=True)
            gu = buf["gu"][i, :M]
            L.gemm(xn, w[o + "gu.w"], out=gu)
            self._ck(lib.cgpt_swiglu_fwd(L.ptr(gu), L.ptr(act), M, I, L.stream_ptr()))
            L.gemm(act, w[o + "down.w"], resid=res, out=res)
        # answer-predicting rows: position nq + ns - 1 + j predicts answer token j
        R = B * na
        gat = (na, Tl, nq + ns - 1)
        xl = torch.empty(R, Hd, dtype=torch.bfloat16, device=self.dev)
        L.norm_rows(res, w["llm.norm"], None, l.rms_eps, xl, rms=True, gather=gat)
        logits = L.gemm(xl, w["llm.head"], out_dtype=torch.float32)
        tok = torch.empty(R, dtype=torch.float32, device=self.dev)
        mc = torch.empty(2, dtype=torch.float32, device=self.dev)
        self._ck(lib.cgpt_ce_loss_smooth(L.ptr(logits), logits.stride(0), R, l.vocab, L.ptr(ans), L.ptr(tok), L.ptr(mc),
                                         self.label_smoothing, L.stream_ptr()))
        self.last = dict(B=B, na=na, Tl=Tl, M=M, R=R, gat=gat, logits=logits, ans=ans, mc=mc, tok=tok.view(B, na))
        return mc[0]

    # ------------------------------------------------------------------ backward
    @torch.no_grad()
    def backward(self):
        """Gradients of the mean loss wrt llama_proj.weight [llm.hidden, qf.hidden] and .bias (fp32, device)."""
        cfg, w, wt, buf, st = self.cfg, self.w, self.wt, self.buf, self.last
        q, l = cfg.qf, cfg.llm
        B, na, Tl, M, R = st["B"], st["na"], st["Tl"], st["M"], st["R"]
        nq, P, Hd, I = q.n_query, self.P, l.hidden, l.inter
        lib = _lib()
        dx, dxb, dact, dgu, dh, datt, dqkv, dqkvb = (buf[k][:M] for k in ("dx", "dxb", "dact", "dgu", "dh", "datt", "dqkv", "dqkvb"))
        # loss -> logits -> last hidden state
        dlog = torch.empty(R, l.vocab, dtype=torch.bfloat16, device=self.dev)
        self._ck(lib.cgpt_ce_grad_smooth(L.ptr(st["logits"]), st["logits"].stride(0), R, l.vocab, L.ptr(st["ans"]), L.ptr(st["mc"]),
                                         L.ptr(dlog), dlog.stride(0), self.label_smoothing, L.stream_ptr()))
        dhl = L.gemm(dlog, wt["llm.head"], out_dtype=torch.float32)                       # [R, Hd]
        dx.zero_()
        self._rmsnorm_bwd(buf["res"][:M], w["llm.norm"], dhl, l.rms_eps, dx, R, st["gat"])    # res = output of the last layer
        scale = 1.0 / math.sqrt(l.head_dim)
        for i in reversed(range(l.layers)):
            o = f"llm.{i}."
            # MLP branch: x_out = x_mid + down(silu(g) * u), (g, u) = gate_up(rms2(x_mid))
            self._cast(dx, dxb, M)
            L.gemm(dxb, wt[o + "down.w"], out=dact)
            self._ck(lib.cgpt_swiglu_bwd(L.ptr(buf["gu"][i, :M]), L.ptr(dact), L.ptr(dgu), M, I, L.stream_ptr()))
            L.gemm(dgu, wt[o + "gu.w"], out=dh)
            self._rmsnorm_bwd(buf["x_mid"][i, :M], w[o + "n2"], dh, l.rms_eps, dx, M)
            # attention branch: x_mid = x_in + o_proj(attn(rope(q), rope(k), v)), (q, k, v) = qkv(rms1(x_in))
            self._cast(dx, dxb, M)
            L.gemm(dxb, wt[o + "o.w"], out=datt)
            kc, vc = buf["kc"][i, :B], buf["vc"][i, :B]
            qs, att = buf["q"][i, :M], buf["att"][i, :M]
            self._ck(lib.cgpt_attention_bwd(L.ptr(qs), qs.stride(0), L.ptr(kc), L.ptr(vc), Hd, self.cache_rows, L.ptr(att),
                                            att.stride(0), L.ptr(datt), datt.stride(0), L.ptr(dqkv), B, l.heads, l.head_dim,
                                            Tl, P + Tl, scale, L.stream_ptr()))
            self._ck(lib.cgpt_rope_bwd_cast(L.ptr(dqkv), L.ptr(dqkvb), M, Tl, l.heads, l.head_dim, P, L.ptr(self.cos),
                                            L.ptr(self.sin), L.stream_ptr()))
            L.gemm(dqkvb, wt[o + "qkv.w"], out=dh)
            self._rmsnorm_bwd(buf["x_in"][i, :M], w[o + "n1"], dh, l.rms_eps, dx, M)
        # dx rows [b*Tl, b*Tl + nq) = gradient wrt llama_proj's output
        Rq = B * nq
        dE = torch.empty(Rq, Hd, dtype=torch.bfloat16, device=self.dev)
        self._cast(dx, dE, Rq, (nq, Tl, 0))
        Rp = (Rq + 7) // 8 * 8                                   # GEMM K must be a multiple of 8
        dEt = torch.zeros(Hd, Rp, dtype=torch.bfloat16, device=self.dev)
        qt = torch.zeros(q.hidden, Rp, dtype=torch.bfloat16, device=self.dev)
        self._ck(lib.cgpt_transpose_bf16(L.ptr(dE), dE.stride(0), L.ptr(dEt), dEt.stride(0), Rq, Hd, L.stream_ptr()))
        self._ck(lib.cgpt_transpose_bf16(L.ptr(self.last_q), self.last_q.stride(0), L.ptr(qt), qt.stride(0), Rq, q.hidden,
                                         L.stream_ptr()))
        gW = L.gemm(dEt, qt, out_dtype=torch.float32)            # [Hd, qf.hidden] = dE^T . Q
        gb = torch.empty(Hd, dtype=torch.float32, device=self.dev)
        self._ck(lib.cgpt_colsum_bf16(L.ptr(dE), dE.stride(0), Rq, Hd, L.ptr(gb), L.stream_ptr()))
        self.grads = {"W": gW, "b": gb}
        return gW, gb

    # ------------------------------------------------------------------ optimiser step
    @torch.no_grad()
    def optimizer_step(self, lr=None):
        """xm.reduce_gradients (mean over the data-parallel ranks) + torch.optim.AdamW semantics."""
        gscale = 1.0
        if self.process_group is not None:
            import torch.distributed as dist
            g = None if self.process_group is True else self.process_group
            if dist.get_world_size(g) > 1:
                if self._comm is None:                                    # libcgpt's own NCCL communicator (NVLink)
                    from .native import CountsComm
                    self._comm = CountsComm(self.process_group)
                for t in self.grads.values():
                    self._comm.allreduce_f32(t)
                gscale = 1.0 / dist.get_world_size(g)                     # mean over the ranks
        self.step_count += 1
        lr = self.lr if lr is None else float(lr)
        lib = _lib()
        for name, p, pbf in (("W", self.Wp, self.w["proj.w"]), ("b", self.bp, None)):
            m, v = self.state[name]
            self._ck(lib.cgpt_adamw_step(L.ptr(p), L.ptr(self.grads[name]), L.ptr(m), L.ptr(v), L.ptr(pbf) if pbf is not None else None,
                                         p.numel(), lr, self.betas[0], self.betas[1], self.eps, self.wd, self.step_count,
                                         gscale, L.stream_ptr()))
        self.w["proj.b"].copy_(self.bp)                              # the GEMM epilogue reads the fp32 bias

    def train_step(self, images, answers, noise_level, *, seed=0, step=0, lr=None, suffix_ids=None, sample_offset=0):
        """maybe_add_noise -> forward -> backward -> reduce + AdamW (agents/minigpt4_finetune_agent.py:165-181)."""
        loss = self.forward(images, answers, noise_level, seed=seed, step=step, suffix_ids=suffix_ids,
                            sample_offset=sample_offset)
        self.backward()
        self.optimizer_step(lr)
        return loss
