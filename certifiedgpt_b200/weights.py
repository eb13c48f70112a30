"""Weights of the MiniGPT-4 towers: reference state-dict layout, random init, bf16 packing.

Key names follow the reference checkpoints (minigpt4.py:193-197, eva_vit.py:444-454,
base_model.py:249-267, HF Llama), so a real MiniGPT-4 / Vicuna state dict can be bound the
same way as the synthetic one:
  visual_encoder.*  ln_vision.*  query_tokens  Qformer.bert.*  llama_proj.*  llama_model.*
"""
import math

import torch


def _normal(shape, std, gen, device):
    return torch.randn(*shape, generator=gen, device=device) * std


def random_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32, parts=("vit", "qf", "llm")):
    """Random init with the reference's own initialisers:
    ViT trunc_normal_(std=.02) + fix_init_weight rescale (eva_vit.py:293-323), LayerNorm 1/0;
    Q-Former normal_(0, 0.02) (Qformer.py:664-674), query_tokens normal_(0, 0.02)
    (minigpt4.py:99-102); llama_proj nn.Linear default; Llama HF default (std 0.02).
    Biases and norm parameters get small random values instead of 0/1 so parity tests exercise them."""
    gen = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    v, q, l = cfg.vit, cfg.qf, cfg.llm

    def put(name, t):
        sd[name] = t.to(dtype)

    def lin(name, out_f, in_f, std=0.02, bias=True):
        put(name + ".weight", _normal((out_f, in_f), std, gen, device))
        if bias:
            put(name + ".bias", _normal((out_f,), 0.02, gen, device))

    def ln(name, d):
        put(name + ".weight", 1.0 + _normal((d,), 0.05, gen, device))
        put(name + ".bias", _normal((d,), 0.05, gen, device))

    # ---- EVA ViT
    put("visual_encoder.patch_embed.proj.weight", _normal((v.dim, 3, v.patch, v.patch), 0.02, gen, device))
    put("visual_encoder.patch_embed.proj.bias", _normal((v.dim,), 0.02, gen, device))
    put("visual_encoder.cls_token", _normal((1, 1, v.dim), 0.02, gen, device))
    put("visual_encoder.pos_embed", _normal((1, v.tokens, v.dim), 0.02, gen, device))
    for i in range(v.depth):
        p = f"visual_encoder.blocks.{i}."
        ln(p + "norm1", v.dim)
        lin(p + "attn.qkv", 3 * v.dim, v.dim, bias=False)
        put(p + "attn.q_bias", _normal((v.dim,), 0.02, gen, device))
        put(p + "attn.v_bias", _normal((v.dim,), 0.02, gen, device))
        lin(p + "attn.proj", v.dim, v.dim)
        sd[p + "attn.proj.weight"] /= math.sqrt(2.0 * (i + 1))      # fix_init_weight
        ln(p + "norm2", v.dim)
        lin(p + "mlp.fc1", v.mlp, v.dim)
        lin(p + "mlp.fc2", v.dim, v.mlp)
        sd[p + "mlp.fc2.weight"] /= math.sqrt(2.0 * (i + 1))
    ln("ln_vision", v.dim)
    # ---- Q-Former
    put("query_tokens", _normal((1, q.n_query, q.hidden), 0.02, gen, device))
    ln("Qformer.bert.embeddings.LayerNorm", q.hidden)
    for i in range(q.layers):
        p = f"Qformer.bert.encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            lin(p + f"attention.self.{nm}", q.hidden, q.hidden)
        lin(p + "attention.output.dense", q.hidden, q.hidden)
        ln(p + "attention.output.LayerNorm", q.hidden)
        if i % q.cross_freq == 0:
            lin(p + "crossattention.self.query", q.hidden, q.hidden)
            lin(p + "crossattention.self.key", q.hidden, v.dim)
            lin(p + "crossattention.self.value", q.hidden, v.dim)
            lin(p + "crossattention.output.dense", q.hidden, q.hidden)
            ln(p + "crossattention.output.LayerNorm", q.hidden)
        lin(p + "intermediate_query.dense", q.inter, q.hidden)
        lin(p + "output_query.dense", q.hidden, q.inter)
        ln(p + "output_query.LayerNorm", q.hidden)
    if "llm" not in parts:      # (draw order is vit, qf, llm: skipping the tail keeps the head identical)
        return sd
    # ---- projection + Llama
    lin("llama_proj", l.hidden, q.hidden, std=1.0 / math.sqrt(q.hidden))
    put("llama_model.model.embed_tokens.weight", _normal((l.vocab, l.hidden), 0.02, gen, device))
    for i in range(l.layers):
        p = f"llama_model.model.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            lin(p + "self_attn." + nm, l.hidden, l.hidden, bias=False)
        lin(p + "mlp.gate_proj", l.inter, l.hidden, bias=False)
        lin(p + "mlp.up_proj", l.inter, l.hidden, bias=False)
        lin(p + "mlp.down_proj", l.hidden, l.inter, bias=False)
        put(p + "input_layernorm.weight", 1.0 + _normal((l.hidden,), 0.05, gen, device))
        put(p + "post_attention_layernorm.weight", 1.0 + _normal((l.hidden,), 0.05, gen, device))
    put("llama_model.model.norm.weight", 1.0 + _normal((l.hidden,), 0.05, gen, device))
    lin("llama_model.lm_head", l.vocab, l.hidden, bias=False)
    return sd


def aliased_state_dict(cfg, seed=0, round_bf16=False):
    """Full-depth state dict whose transformer layers ALIAS the tensors of the first layer of each tower (Q-Former: the
    first cross and the first plain layer): identical shapes, FLOPs and bytes per layer at a fraction of the host RAM
    and init time of 7 B distinct parameters.  For host-side work at BASELINE.json's full shape: the CPU timing sample
    of bench.py and the full-shape parity test, where the fp32 oracle and the engine must hold the SAME weights."""
    import copy
    c1 = copy.deepcopy(cfg)
    c1.vit.depth, c1.qf.layers, c1.llm.layers = 1, min(cfg.qf.layers, 2), 1
    sd = random_state_dict(c1, seed=seed)
    if round_bf16:
        sd = round_to_bf16(sd)
    for i in range(1, cfg.vit.depth):
        for k in [k for k in sd if k.startswith("visual_encoder.blocks.0.")]:
            sd[k.replace("blocks.0.", f"blocks.{i}.")] = sd[k]
    for i in range(2, cfg.qf.layers):
        src = i % 2
        for k in [k for k in sd if k.startswith(f"Qformer.bert.encoder.layer.{src}.")]:
            sd[k.replace(f"layer.{src}.", f"layer.{i}.")] = sd[k]
    for i in range(1, cfg.llm.layers):
        for k in [k for k in sd if k.startswith("llama_model.model.layers.0.")]:
            sd[k.replace("layers.0.", f"layers.{i}.")] = sd[k]
    return sd


def round_to_bf16(sd):
    """The same state dict with every GEMM weight rounded to bf16 (kept as fp32 tensors): the
    oracle then multiplies exactly the weights the GPU holds, so differences are activation
    rounding only."""
    out = {}
    for k, t in sd.items():
        if t.dim() >= 2 and not k.endswith(("pos_embed", "cls_token", "query_tokens")):
            out[k] = t.to(torch.bfloat16).to(torch.float32)
        else:
            out[k] = t.clone()
    return out


# ------------------------------------------------------------------------------- weight import
def interpolate_pos_embed(pos_embed, new_tokens, num_extra_tokens=1):
    """Bicubic resize of the ViT position grid to another image size (eva_vit.py:383-404);
    the class token's embedding is kept.  pos_embed: [1, 1+g*g, D] -> [1, new_tokens, D]."""
    pos_embed = pos_embed.float()
    D = pos_embed.shape[-1]
    orig = int((pos_embed.shape[-2] - num_extra_tokens) ** 0.5)
    new = int((new_tokens - num_extra_tokens) ** 0.5)
    if orig == new:
        return pos_embed
    extra = pos_embed[:, :num_extra_tokens]
    grid = pos_embed[:, num_extra_tokens:].reshape(-1, orig, orig, D).permute(0, 3, 1, 2)
    grid = torch.nn.functional.interpolate(grid, size=(new, new), mode="bicubic", align_corners=False)
    grid = grid.permute(0, 2, 3, 1).flatten(1, 2)
    return torch.cat((extra, grid), dim=1)


def import_state_dicts(cfg, vit_sd=None, qformer_sd=None, minigpt4_sd=None, llama_sd=None):
    """Assemble an engine state dict from the checkpoints the reference loads:
      vit_sd       eva_vit_g.pth keys (eva_vit.py:444-454) -> `visual_encoder.*`, pos-embed resized
      qformer_sd   BLIP-2 checkpoint `model` dict (base_model.py:249-267): `Qformer.*`, `query_tokens`,
                   `ln_vision.*`
      minigpt4_sd  MiniGPT-4 checkpoint `model_state_dict` (minigpt4.py:193-197): `llama_proj.*` and
                   any fine-tuned overrides
      llama_sd     HF LlamaForCausalLM state dict -> `llama_model.*`
    Later sources override earlier ones, as successive load_state_dict(strict=False) calls do."""
    out = {}
    if vit_sd is not None:
        for k, v in vit_sd.items():
            out["visual_encoder." + k] = v
        pe = "visual_encoder.pos_embed"
        if pe in out:
            out[pe] = interpolate_pos_embed(out[pe], cfg.vit.tokens)
    for src in (qformer_sd, minigpt4_sd):
        if src is not None:
            for k, v in src.items():
                if k.startswith(("Qformer.", "query_tokens", "ln_vision.", "llama_proj.", "visual_encoder.",
                                 "llama_model.")):
                    out[k] = v
    if llama_sd is not None:
        for k, v in llama_sd.items():
            out["llama_model." + k] = v
    return out
