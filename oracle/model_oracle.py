"""ORACLE (test infrastructure, never the product path).

fp32 CPU restatement of the MiniGPT-4 forward that `Smooth` classifies with, written
functionally over a reference-layout state dict.  Each function cites the reference lines it
follows (paths under /root/reference/graphs/models/minigpt4/models/ unless noted):

  vit_forward        eva_vit.py  PatchEmbed :204-210, cls/pos :337-340, Block :178-181,
                                 Attention :123-153, Mlp :59-66
  ln_vision          base_model.py:281-287 (fp32 LayerNorm), used at minigpt4.py:129
  qformer_forward    Qformer.py  BertEmbeddings :104-107, BertSelfAttention :169-275,
                                 BertSelfOutput :285-289, BertLayer :402-484 (query branch),
                                 BertIntermediate/Output :352-375; wiring minigpt4.py:133-141
  llama_*            third-party transformers (pinned 4.30.0 in docker/tpu-docker:32; the
                     reference's modeling_llama.py:65-84 keeps the stock logits path): published
                     Llama algorithm - RMSNorm, rotate_half RoPE, causal SDPA, SwiGLU - pinned in
                     tests against transformers.LlamaForCausalLM of this image
  generate_ids       minigpt_base.py:374-427 (greedy, min_length=1, EOS=2, pad after EOS) as
                     HF GenerationMixin greedy search executes it
  answer_label       minigpt_base.py:438-446 + agents/minigpt4_eval_agent.py:102 at token level
  lm_loss            minigpt_base.py:323-362 (training / validation forward) + modeling_llama.py:101-123 (shifted CE);
                     pinned against transformers.LlamaForCausalLM(inputs_embeds, labels).loss
  finetune_grads     loss.backward() of agents/minigpt4_finetune_agent.py:165-172 by torch autograd over lm_loss
                     (only llama_proj requires grad: base_model.py:162-172,238-240; minigpt4.py:111-117)

Pinning: tests/golden/ref_*.pt hold outputs of the reference's OWN eva_vit.py / Qformer.py
modules (imported by file path in the build container, script tests/golden/make_ref_fixtures.py);
tests/test_oracle_model_cpu.py checks this restatement against them; tests/golden/ref_encode_img.pt is a run of the
reference's OWN MiniGPT4.encode_img (minigpt4.py:121-149) over those modules (the whole tower and its wiring).
tests/golden/ref_generate.pt holds a run of the
reference's OWN MiniGPTBase.generate / get_context_emb / embed_tokens (minigpt_base.py, executed unmodified by
tests/golden/make_ref_generate_fixtures.py on a stub `self` with this image's transformers Llama): the embeddings handed to
llama_model.generate, its arguments, the generated ids and the post-processed answers; build_prompt_embeds, generate_ids and
canonical_answer reproduce them exactly.  tests/golden/ref_forward.pt is the same for the training forward
(MiniGPTBase.forward / preparing_embedding / prompt_wrap / concat_emb_input_output): lm_loss reproduces its inputs, labels and loss.
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ EVA ViT
def vit_forward(sd, cfg, images, prefix="visual_encoder.", collect=None):
    """VisionTransformer.forward_features (eva_vit.py:332-349). images [B,3,S,S] -> [B,T,D]."""
    v = cfg.vit
    x = F.conv2d(images, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"],
                 stride=v.patch)                                   # :202,209
    x = x.flatten(2).transpose(1, 2)
    B = x.shape[0]
    cls = sd[prefix + "cls_token"].expand(B, -1, -1)               # :337
    x = torch.cat((cls, x), dim=1) + sd[prefix + "pos_embed"]      # :338-340
    if collect is not None:
        collect["embed"] = x.clone()
    scale = v.head_dim ** -0.5                                     # :78
    for i in range(v.depth):
        p = f"{prefix}blocks.{i}."
        h = F.layer_norm(x, (v.dim,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], v.eps)
        bias = torch.cat((sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.v_bias"]),
                          sd[p + "attn.v_bias"]))                  # :127
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], bias)
        qkv = qkv.reshape(B, -1, 3, v.heads, v.head_dim).permute(2, 0, 3, 1, 4)
        q, k, vv = qkv[0] * scale, qkv[1], qkv[2]                   # :133
        attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)            # :134,147
        a = (attn @ vv).transpose(1, 2).reshape(B, -1, v.dim)       # :150
        x = x + F.linear(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])   # :151,180
        h = F.layer_norm(x, (v.dim,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], v.eps)
        h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))    # :60-61 exact GELU
        x = x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])        # :64,181
        if collect is not None:
            collect[f"block{i}"] = x.clone()
    return x


def ln_vision(sd, cfg, x):
    return F.layer_norm(x.float(), (cfg.vit.dim,), sd["ln_vision.weight"], sd["ln_vision.bias"],
                        cfg.ln_vision_eps)


# ------------------------------------------------------------------ Q-Former
def _bert_attention(sd, p, hidden, kv_src, heads, eps):
    """BertAttention = BertSelfAttention (Qformer.py:169-275, zero masks) + BertSelfOutput (:285-289)."""
    B, Tq, Hd = hidden.shape
    hd = Hd // heads

    def split(t):
        return t.view(B, -1, heads, hd).permute(0, 2, 1, 3)

    q = split(F.linear(hidden, sd[p + "self.query.weight"], sd[p + "self.query.bias"]))
    k = split(F.linear(kv_src, sd[p + "self.key.weight"], sd[p + "self.key.bias"]))
    v = split(F.linear(kv_src, sd[p + "self.value.weight"], sd[p + "self.value.bias"]))
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(hd)             # :197,240
    probs = scores.softmax(dim=-1)                                  # :246
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, Tq, Hd)        # :261-265
    out = F.linear(ctx, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return F.layer_norm(out + hidden, (Hd,), sd[p + "output.LayerNorm.weight"],
                        sd[p + "output.LayerNorm.bias"], eps)       # :288


def qformer_forward(sd, cfg, image_embeds, prefix="Qformer.bert.", collect=None):
    """Qformer.bert(query_embeds=query_tokens.expand(B), encoder_hidden_states=image_embeds)
    .last_hidden_state  (minigpt4.py:133-139 -> Qformer.py:804-965)."""
    q = cfg.qf
    B = image_embeds.shape[0]
    h = sd["query_tokens"].expand(B, -1, -1)
    h = F.layer_norm(h, (q.hidden,), sd[prefix + "embeddings.LayerNorm.weight"],
                     sd[prefix + "embeddings.LayerNorm.bias"], q.eps)   # :104-107
    for i in range(q.layers):
        p = f"{prefix}encoder.layer.{i}."
        h = _bert_attention(sd, p + "attention.", h, h, q.heads, q.eps)            # :417-423
        if i % q.cross_freq == 0:                                                    # :386-395
            h = _bert_attention(sd, p + "crossattention.", h, image_embeds, q.heads, q.eps)   # :432-444
        inter = F.gelu(F.linear(h, sd[p + "intermediate_query.dense.weight"],
                                sd[p + "intermediate_query.dense.bias"]))           # :481-482, :358-362
        out = F.linear(inter, sd[p + "output_query.dense.weight"], sd[p + "output_query.dense.bias"])
        h = F.layer_norm(out + h, (q.hidden,), sd[p + "output_query.LayerNorm.weight"],
                         sd[p + "output_query.LayerNorm.bias"], q.eps)              # :371-375
        if collect is not None:
            collect[f"layer{i}"] = h.clone()
    return h


def encode_img(sd, cfg, images, collect=None):
    """MiniGPT4.encode_img (minigpt4.py:121-149) -> inputs_llama [B, n_query, llm.hidden]."""
    feats = vit_forward(sd, cfg, images, collect=collect)
    image_embeds = ln_vision(sd, cfg, feats)
    if collect is not None:
        collect["image_embeds"] = image_embeds.clone()
    qout = qformer_forward(sd, cfg, image_embeds, collect=collect)
    if collect is not None:
        collect["qformer"] = qout.clone()
    return F.linear(qout, sd["llama_proj.weight"], sd["llama_proj.bias"])      # :141


# ------------------------------------------------------------------ Llama
def _rms(x, w, eps):
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def rope_tables(cfg, max_pos):
    l = cfg.llm
    inv = 1.0 / (l.rope_theta ** (torch.arange(0, l.head_dim, 2, dtype=torch.float32) / l.head_dim))
    fr = torch.outer(torch.arange(max_pos, dtype=torch.float32), inv)   # [pos, hd/2]
    return fr.cos(), fr.sin()


def _rope(x, cos, sin):
    # x [B,H,T,hd]; rotate_half convention: out = x*cos + cat(-x2, x1)*sin with duplicated freqs
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    c, s = cos[None, None], sin[None, None]
    return torch.cat((x1 * c - x2 * s, x2 * c + x1 * s), dim=-1)


def llama_forward(sd, cfg, embeds, past=None, pos0=0, prefix="llama_model."):
    """Decoder stack on input embeddings [B,T,hidden] with optional KV past (list of (k,v)).
    Returns (hidden after final norm [B,T,hidden], new past)."""
    l = cfg.llm
    B, T, _ = embeds.shape
    cos, sin = rope_tables(cfg, pos0 + T)
    cos, sin = cos[pos0:pos0 + T], sin[pos0:pos0 + T]
    x = embeds
    new_past = []
    for i in range(l.layers):
        p = f"{prefix}model.layers.{i}."
        h = _rms(x, sd[p + "input_layernorm.weight"], l.rms_eps)

        def split(t):
            return t.view(B, T, l.heads, l.head_dim).transpose(1, 2)

        q = _rope(split(F.linear(h, sd[p + "self_attn.q_proj.weight"])), cos, sin)
        k = _rope(split(F.linear(h, sd[p + "self_attn.k_proj.weight"])), cos, sin)
        v = split(F.linear(h, sd[p + "self_attn.v_proj.weight"]))
        if past is not None:
            k = torch.cat((past[i][0], k), dim=2)
            v = torch.cat((past[i][1], v), dim=2)
        new_past.append((k, v))
        Tk = k.shape[2]
        scores = (q @ k.transpose(-1, -2)) / math.sqrt(l.head_dim)
        qi = torch.arange(T)[:, None] + (Tk - T)
        mask = torch.arange(Tk)[None, :] > qi
        scores = scores.masked_fill(mask[None, None], float("-inf"))
        a = (scores.softmax(-1) @ v).transpose(1, 2).reshape(B, T, l.hidden)
        x = x + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        h = _rms(x, sd[p + "post_attention_layernorm.weight"], l.rms_eps)
        h = F.silu(F.linear(h, sd[p + "mlp.gate_proj.weight"])) * F.linear(h, sd[p + "mlp.up_proj.weight"])
        x = x + F.linear(h, sd[p + "mlp.down_proj.weight"])
    return _rms(x, sd[prefix + "model.norm.weight"], l.rms_eps), new_past


def build_prompt_embeds(sd, cfg, img_embeds, prefix_ids, suffix_ids, prefix="llama_model."):
    """get_context_emb (minigpt_base.py:75-89): [embed(prefix) | image (n_query) | embed(suffix)].
    All samples share the text, so no left padding occurs (:399-412)."""
    emb = sd[prefix + "model.embed_tokens.weight"]
    B = img_embeds.shape[0]
    pre = emb[torch.as_tensor(prefix_ids, dtype=torch.long)][None].expand(B, -1, -1)
    suf = emb[torch.as_tensor(suffix_ids, dtype=torch.long)][None].expand(B, -1, -1)
    return torch.cat((pre, img_embeds, suf), dim=1)


def generate_ids(sd, cfg, embeds, max_new_tokens=20, min_length=1, prefix="llama_model."):
    """HF greedy search on inputs_embeds (minigpt_base.py:414-427: do_sample=False, num_beams=1,
    min_length=1 -> EOS suppressed while fewer than min_length new tokens exist; finished rows
    emit pad).  Returns (ids [B,max_new], first-step logits [B,V], per-step top-2 margins [B,steps])."""
    l = cfg.llm
    B, S, _ = embeds.shape
    emb = sd[prefix + "model.embed_tokens.weight"]
    head = sd[prefix + "lm_head.weight"]
    hidden, past = llama_forward(sd, cfg, embeds, prefix=prefix)
    logits = F.linear(hidden[:, -1], head).float()                  # modeling_llama.py:83-84
    ids = torch.full((B, max_new_tokens), l.pad_id, dtype=torch.long)
    unfinished = torch.ones(B, dtype=torch.bool)
    first_logits = logits.clone()
    margins = []
    for t in range(max_new_tokens):
        scores = logits.clone()
        if t < min_length:
            scores[:, l.eos_id] = float("-inf")                     # MinLengthLogitsProcessor
        top2 = scores.topk(2, dim=-1).values
        margins.append(torch.where(unfinished, top2[:, 0] - top2[:, 1], torch.full((B,), float("inf"))))
        nxt = scores.argmax(-1)
        nxt = torch.where(unfinished, nxt, torch.full_like(nxt, l.pad_id))
        ids[:, t] = nxt
        unfinished = unfinished & (nxt != l.eos_id)
        if not unfinished.any() or t == max_new_tokens - 1:
            break
        hidden, past = llama_forward(sd, cfg, emb[nxt][:, None], past=past, pos0=S + t, prefix=prefix)
        logits = F.linear(hidden[:, -1], head).float()
    return ids, first_logits, torch.stack(margins, dim=1)


# ------------------------------------------------------------------ answer -> label adapter
def lm_loss(sd, cfg, images, prefix_ids, suffix_ids, answers, prefix="llama_model.", label_smoothing=0.0):
    """Training / validation forward of MiniGPTBase (minigpt_base.py:323-362): inputs = [bos+prompt | image | rest of
    the prompt | answer], targets = answer ids at the answer positions and -100 elsewhere, loss = the shifted
    CrossEntropyLoss(mean) of modeling_llama.py:101-123.  answers: LongTensor [B, na], -100 = right padding
    (pad embedding fed, target ignored).  Returns (mean loss, per-token losses [B, na]).
    label_smoothing: the reference's subclass builds CrossEntropyLoss(label_smoothing=0.1) (modeling_llama.py:107);
    0.1 reproduces a run of that file (tests/golden/ref_forward.pt), the default 0.0 is the stock transformers loss,
    which is what libcgpt's cgpt_ce_loss / cgpt_ce_grad compute today (known gap)."""
    l = cfg.llm
    emb = sd[prefix + "model.embed_tokens.weight"]
    img = encode_img(sd, cfg, images)
    cond = build_prompt_embeds(sd, cfg, img, prefix_ids, suffix_ids, prefix=prefix)
    ans_in = answers.clamp_min(0).masked_fill(answers < 0, l.pad_id)
    embeds = torch.cat((cond, emb[ans_in].float()), dim=1)
    hidden, _ = llama_forward(sd, cfg, embeds, prefix=prefix)
    logits = F.linear(hidden, sd[prefix + "lm_head.weight"]).float()
    Lc, na = cond.shape[1], answers.shape[1]
    targets = torch.full(embeds.shape[:2], -100, dtype=torch.long)
    targets[:, Lc:Lc + na] = answers
    shift_logits, shift_labels = logits[:, :-1].reshape(-1, l.vocab), targets[:, 1:].reshape(-1)
    loss = F.cross_entropy(shift_logits, shift_labels, ignore_index=-100, reduction="mean", label_smoothing=label_smoothing)
    tok = F.cross_entropy(shift_logits, shift_labels, ignore_index=-100, reduction="none",
                          label_smoothing=label_smoothing).view(embeds.shape[0], -1)
    return loss, tok[:, Lc - 1:Lc - 1 + na]


def finetune_grads(sd, cfg, images, prefix_ids, suffix_ids, answers, label_smoothing=0.0):
    """loss.backward() of the fine-tune step (agents/minigpt4_finetune_agent.py:165-172) by torch autograd over the
    restated forward: only llama_proj.weight / .bias require grad (everything else is frozen, base_model.py:162-172,
    238-240; minigpt4.py:111-117).  Returns (loss, dW, db)."""
    sd2 = dict(sd)
    W = sd["llama_proj.weight"].detach().float().clone().requires_grad_(True)
    b = sd["llama_proj.bias"].detach().float().clone().requires_grad_(True)
    sd2["llama_proj.weight"], sd2["llama_proj.bias"] = W, b
    with torch.enable_grad():
        loss, _ = lm_loss(sd2, cfg, images, prefix_ids, suffix_ids, answers, label_smoothing=label_smoothing)
        loss.backward()
    return loss.detach(), W.grad.detach(), b.grad.detach()


def canonical_answer(ids, eos_id=2):
    """Token-level restatement of minigpt_base.py:438-446: decode(skip_special_tokens=True) drops
    <unk>=0,<s>=1,</s>=2; everything after the first EOS is padding."""
    out = []
    for t in ids:
        t = int(t)
        if t == eos_id:
            break
        if t in (0, 1, 2):
            continue
        out.append(t)
    return tuple(out)


def answer_label(ids, table, other_label, eos_id=2):
    """table: dict canonical token tuple -> class id (fixed answer vocabulary); unknown -> other."""
    return table.get(canonical_answer(ids, eos_id), other_label)


class MiniGPT4ClassifierOracle(torch.nn.Module):
    """The VLM-as-classifier adapter `Smooth` needs (absent from the reference):
    images [B,3,S,S] -> one-hot logits [B,num_classes] of the normalised short answer."""

    def __init__(self, sd, cfg, prefix_ids, suffix_ids, answer_table, num_classes, max_new_tokens=20,
                 normalize=None):
        super().__init__()
        self.sd, self.cfg = sd, cfg
        self.prefix_ids, self.suffix_ids = list(prefix_ids), list(suffix_ids)
        self.table = {canonical_answer(k, cfg.llm.eos_id): v for k, v in answer_table}
        self.num_classes = num_classes
        self.max_new_tokens = max_new_tokens
        self.normalize = normalize       # (mean, std) when the noise lives in pixel space
        self.last = {}

    @torch.no_grad()
    def forward(self, images):
        images = images.float()
        if self.normalize is not None:
            m = torch.tensor(self.normalize[0]).view(1, 3, 1, 1)
            s = torch.tensor(self.normalize[1]).view(1, 3, 1, 1)
            images = (images - m) / s
        img = encode_img(self.sd, self.cfg, images)
        embeds = build_prompt_embeds(self.sd, self.cfg, img, self.prefix_ids, self.suffix_ids)
        ids, first_logits, margins = generate_ids(self.sd, self.cfg, embeds, self.max_new_tokens)
        labels = torch.tensor([answer_label(r.tolist(), self.table, self.num_classes - 1, self.cfg.llm.eos_id)
                               for r in ids])
        self.last = {"ids": ids, "first_logits": first_logits, "margins": margins, "labels": labels,
                     "img_embeds": img}
        onehot = torch.zeros(images.shape[0], self.num_classes)
        onehot[torch.arange(images.shape[0]), labels] = 1.0
        return onehot
