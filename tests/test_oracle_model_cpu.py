"""CPU tests: the model oracle against (a) golden outputs of the reference's own eva_vit.py /
Qformer.py modules (tests/golden/ref_*.pt, made by tests/golden/make_ref_fixtures.py) and
(b) this image's transformers.LlamaForCausalLM (the reference's LLM arithmetic is third-party
transformers; its subclass only changes the loss, modeling_llama.py:65-84)."""
import os

import pytest
import torch

from certifiedgpt_b200.config import LlmConfig, ModelConfig, QFormerConfig, VitConfig
from certifiedgpt_b200.weights import random_state_dict
from oracle import model_oracle as mo

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {
    "tiny": ModelConfig.tiny(),
    "wide": ModelConfig(vit=VitConfig(img_size=56, depth=1), qf=QFormerConfig(layers=2)),
}


@pytest.mark.parametrize("name", ["tiny", "wide"])
def test_vit_matches_reference_module(name):
    ref = torch.load(os.path.join(GOLD, "ref_vit.pt"))[name]
    cfg = CASES[name]
    sd = random_state_dict(cfg, seed=11, parts=("vit", "qf"))
    out = mo.vit_forward(sd, cfg, ref["images"])
    assert out.shape == ref["features"].shape
    assert torch.allclose(out, ref["features"], atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("name", ["tiny", "wide"])
def test_qformer_matches_reference_module(name):
    ref = torch.load(os.path.join(GOLD, "ref_qformer.pt"))[name]
    cfg = CASES[name]
    sd = random_state_dict(cfg, seed=11, parts=("vit", "qf"))
    out = mo.qformer_forward(sd, cfg, ref["image_embeds"])
    assert torch.allclose(out, ref["last_hidden_state"], atol=2e-5, rtol=1e-5)


def _hf_llama(cfg, sd):
    from transformers import LlamaConfig, LlamaForCausalLM
    l = cfg.llm
    hc = LlamaConfig(hidden_size=l.hidden, intermediate_size=l.inter, num_hidden_layers=l.layers,
                     num_attention_heads=l.heads, num_key_value_heads=l.heads, vocab_size=l.vocab,
                     rms_norm_eps=l.rms_eps, rope_theta=l.rope_theta, max_position_embeddings=256,
                     bos_token_id=1, eos_token_id=l.eos_id, pad_token_id=l.pad_id,
                     attn_implementation="eager", tie_word_embeddings=False)
    m = LlamaForCausalLM(hc).eval()
    hsd = {k[len("llama_model."):]: v for k, v in sd.items() if k.startswith("llama_model.")}
    missing, unexpected = m.load_state_dict(hsd, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    return m


def test_llama_logits_and_greedy_generate_match_hf():
    cfg = ModelConfig.tiny()
    cfg.llm = LlmConfig(hidden=64, layers=2, heads=4, inter=128, vocab=96)
    sd = random_state_dict(cfg, seed=3)
    # make EOS reachable so the finished-row / padding semantics are exercised
    sd["llama_model.lm_head.weight"][cfg.llm.eos_id] *= 3.0
    hf = _hf_llama(cfg, sd)
    g = torch.Generator().manual_seed(0)
    embeds = torch.randn(5, 11, cfg.llm.hidden, generator=g) * 0.3
    with torch.no_grad():
        ref_logits = hf(inputs_embeds=embeds).logits[:, -1].float()
        ref_ids = hf.generate(inputs_embeds=embeds, attention_mask=torch.ones(5, 11, dtype=torch.long),
                              max_new_tokens=6, num_beams=1, do_sample=False, min_length=1,
                              top_p=0.9, repetition_penalty=1.0, length_penalty=1, temperature=1.0)
    ids, first_logits, margins = mo.generate_ids(sd, cfg, embeds, max_new_tokens=6)
    assert torch.allclose(first_logits, ref_logits, atol=1e-5, rtol=1e-5)
    assert margins.shape[0] == 5
    n = ref_ids.shape[1]
    assert torch.equal(ids[:, :n], ref_ids)
    assert (ids[:, n:] == cfg.llm.pad_id).all()
    assert (ids[:, 0] != cfg.llm.eos_id).all()           # min_length=1


def test_prompt_embedding_layout_and_adapter():
    cfg = ModelConfig.tiny()
    sd = random_state_dict(cfg, seed=4)
    img = torch.randn(3, cfg.qf.n_query, cfg.llm.hidden)
    e = mo.build_prompt_embeds(sd, cfg, img, [1, 5, 6], [7, 8, 9, 10])
    assert e.shape == (3, 3 + cfg.qf.n_query + 4, cfg.llm.hidden)
    emb = sd["llama_model.model.embed_tokens.weight"]
    assert torch.equal(e[1, 0], emb[1]) and torch.equal(e[2, 3:3 + cfg.qf.n_query], img[2])
    assert torch.equal(e[0, -1], emb[10])
    assert mo.canonical_answer([0, 5, 1, 6, 2, 9]) == (5, 6)
    assert mo.answer_label([5, 6, 2, 0], {(5, 6): 3}, other_label=9) == 3
    assert mo.answer_label([6, 5, 2, 0], {(5, 6): 3}, other_label=9) == 9


def test_classifier_oracle_end_to_end_tiny():
    cfg = ModelConfig.tiny()
    sd = random_state_dict(cfg, seed=5)
    table = [((t,), t % 7) for t in range(3, cfg.llm.vocab)]
    clf = mo.MiniGPT4ClassifierOracle(sd, cfg, [1, 4, 5], [6, 7, 8, 9], table, num_classes=8, max_new_tokens=1)
    x = torch.randn(4, 3, cfg.vit.img_size, cfg.vit.img_size)
    out = clf(x)
    assert out.shape == (4, 8) and torch.equal(out.sum(1), torch.ones(4))
    assert clf.last["ids"].shape == (4, 1)


def test_pos_embed_interpolation_matches_reference():
    from certifiedgpt_b200.weights import import_state_dicts, interpolate_pos_embed
    ref = torch.load(os.path.join(GOLD, "ref_pos_interp.pt"))
    assert torch.equal(interpolate_pos_embed(ref["in"], 65), ref["out"])
    assert interpolate_pos_embed(ref["in"], 17) is not None and interpolate_pos_embed(ref["in"], 17).shape == (1, 17, 16)
    # import path: eva_vit_g.pth-style keys are prefixed and the grid is resized for the configured image size
    cfg = ModelConfig.tiny()
    cfg.vit.img_size = 112                       # 8x8 patches + cls = 65 tokens
    sd = import_state_dicts(cfg, vit_sd={"pos_embed": ref["in"], "cls_token": torch.zeros(1, 1, 16)})
    assert torch.equal(sd["visual_encoder.pos_embed"], ref["out"]) and "visual_encoder.cls_token" in sd


def test_lm_loss_matches_hf_causal_lm_loss():
    """oracle.lm_loss (the restated training forward, minigpt_base.py:323-362 + modeling_llama.py:101-123) equals
    transformers' LlamaForCausalLM(inputs_embeds, labels).loss on the same embeddings, padding included."""
    cfg = ModelConfig.tiny()
    sd = random_state_dict(cfg, seed=9)
    hf = _hf_llama(cfg, sd)
    S = cfg.vit.img_size
    images = torch.randn(3, 3, S, S, generator=torch.Generator().manual_seed(4))
    prefix, suffix = (1, 5, 6), (7, 8, 9, 10)
    answers = torch.tensor([[11, 12, 2], [13, 2, -100], [14, 15, 16]])
    loss, tok = mo.lm_loss(sd, cfg, images, prefix, suffix, answers)
    with torch.no_grad():
        img = mo.encode_img(sd, cfg, images)
        cond = mo.build_prompt_embeds(sd, cfg, img, prefix, suffix)
        emb = sd["llama_model.model.embed_tokens.weight"]
        embeds = torch.cat((cond, emb[answers.clamp_min(0)]), dim=1)
        labels = torch.full(embeds.shape[:2], -100, dtype=torch.long)
        labels[:, cond.shape[1]:] = answers
        ref = hf(inputs_embeds=embeds, labels=labels).loss
    assert abs(loss.item() - ref.item()) < 1e-5
    assert tok.shape == (3, 3) and tok[1, 2].item() == 0.0 and (tok[answers >= 0] > 0).all()
    assert abs(tok.sum().item() / 8 - loss.item()) < 1e-5


def _ref_generate_cases():
    return torch.load(os.path.join(GOLD, "ref_generate.pt"))["cases"]


@pytest.mark.parametrize("rec", _ref_generate_cases(), ids=lambda r: r["case"]["name"])
def test_prompt_assembly_and_generate_equal_the_reference_generate_run(rec):
    """The oracle's glue (build_prompt_embeds / generate_ids / canonical_answer) and the product's split_prompt against a
    run of the reference's OWN MiniGPTBase.generate + get_context_emb + embed_tokens (tests/golden/
    make_ref_generate_fixtures.py): the embeddings handed to llama_model.generate, its arguments, the generated ids
    (min_length, EOS, pad) and the post-processed answer strings."""
    from certifiedgpt_b200.data import vqav2 as V
    from ref_generate_util import encode
    case = rec["case"]
    cfg = ModelConfig.tiny()
    cfg.llm = LlmConfig(hidden=64, layers=2, heads=4, inter=128, vocab=96)
    sd = random_state_dict(cfg, seed=case["seed"])
    if case["eos_boost"]:
        sd["llama_model.lm_head.weight"][cfg.llm.eos_id] *= case["eos_boost"]
    kw = rec["generate_kwargs"]
    assert kw["do_sample"] is False and kw["num_beams"] == 1 and kw["min_length"] == 1
    assert kw["max_new_tokens"] == case["max_new_tokens"]
    ref_ids, ref_emb, ref_mask = rec["outputs"], rec["inputs_embeds"], rec["attention_mask"]
    enc = lambda s: encode(s, cfg.llm.vocab)
    for b, text in enumerate(case["texts"]):
        prefix, suffix = V.split_prompt(text, enc, prompt_template="{}")     # the agents' prompt is already wrapped
        assert prefix[0] == 1 and 1 not in suffix                           # BOS on the first segment only (:79-82)
        emb = mo.build_prompt_embeds(sd, cfg, rec["img_embeds"][b:b + 1], prefix, suffix)
        S = emb.shape[1]
        assert int(ref_mask[b].sum()) == S and bool((ref_mask[b, -S:] == 1).all())      # left padding (:405-412)
        assert torch.equal(emb[0], ref_emb[b, -S:])
        assert bool((ref_emb[b, :-S] == 0).all()) if S < ref_emb.shape[1] else True
        ids = mo.generate_ids(sd, cfg, emb, max_new_tokens=case["max_new_tokens"])[0][0]
        n = ref_ids.shape[1]
        assert ids[:n].tolist() == ref_ids[b].tolist()
        assert (ids[n:] == cfg.llm.pad_id).all()
        assert " ".join(f"t{t}" for t in mo.canonical_answer(ids.tolist(), cfg.llm.eos_id)) == rec["answers"][b]


@pytest.mark.parametrize("name,cfg", [("tiny", ModelConfig.tiny()),
                                      # full ViT-g / Q-Former / llama_proj widths; the Llama tail is drawn after llama_proj
                                      # (weights.random_state_dict), so shrinking it leaves the tower's weights unchanged
                                      ("wide", ModelConfig(vit=VitConfig(img_size=56, depth=1), qf=QFormerConfig(layers=2),
                                                           llm=LlmConfig(layers=1, inter=128, vocab=96)))])
def test_encode_img_equals_the_reference_encode_img_run(name, cfg):
    """The whole image tower (A6-A10) against a run of the reference's OWN MiniGPT4.encode_img over its own
    VisionTransformer and Q-Former layers (tests/golden/make_ref_encode_img_fixture.py) on the same seeded weights."""
    ref = torch.load(os.path.join(GOLD, "ref_encode_img.pt"))[name]
    sd = random_state_dict(cfg, seed=ref["seed"])
    with torch.no_grad():
        out = mo.encode_img(sd, cfg, ref["images"])
    assert out.shape == (3, cfg.qf.n_query, cfg.llm.hidden)
    assert torch.allclose(out[..., ::ref["stride"]], ref["inputs_llama"], atol=3e-5, rtol=1e-5)
    assert float(out.double().sum()) == pytest.approx(ref["sum"], abs=2e-2)
    assert float(out.double().abs().sum()) == pytest.approx(ref["abs_sum"], rel=1e-5)


@pytest.mark.parametrize("rec", torch.load(os.path.join(GOLD, "ref_forward.pt"))["cases"], ids=lambda r: r["case"]["name"])
def test_training_forward_equals_the_reference_forward_run(rec):
    """oracle.lm_loss (and the product's split_prompt / answer tokenisation) against a run of the reference's OWN
    MiniGPTBase.forward -> preparing_embedding -> prompt_wrap -> concat_emb_input_output (tests/golden/
    make_ref_forward_fixture.py) over the reference's OWN LlamaForCausalLM subclass (modeling_llama.py): the embeddings
    and labels handed to the Llama and its label-smoothed mean cross-entropy."""
    from certifiedgpt_b200.data import vqav2 as V
    from ref_generate_util import encode_special
    case = rec["case"]
    cfg = ModelConfig.tiny()
    cfg.llm = LlmConfig(hidden=64, layers=2, heads=4, inter=128, vocab=96)
    sd = random_state_dict(cfg, seed=case["seed"])
    enc = lambda s: encode_special(s, cfg.llm.vocab)
    if case.get("style") == "shipped":      # the reference's own fine-tune configs: raw instruction, answers end with "###"
        template, end_sym = V.SHIPPED_PROMPT_TEMPLATE, V.SHIPPED_END_SYM
    else:                                   # chat style: "[INST] {} [/INST]", "</s>"
        template, end_sym = V.PROMPT_TEMPLATE, V.END_SYM
    prefix, suffix = V.split_prompt(case["instruction"], enc, prompt_template=template)
    rows = [enc(a + end_sym) for a in case["answers"]]
    na = max(len(r) for r in rows)
    answers = torch.tensor([r + [-100] * (na - len(r)) for r in rows])
    assert all(r[-1] == cfg.llm.eos_id for r in rows) or case.get("style") == "shipped"
    with torch.no_grad():
        loss, tok = mo.lm_loss(sd, cfg, rec["images"], prefix, suffix, answers)
        smooth, _ = mo.lm_loss(sd, cfg, rec["images"], prefix, suffix, answers, label_smoothing=0.1)
        cond = mo.build_prompt_embeds(sd, cfg, mo.encode_img(sd, cfg, rec["images"]), prefix, suffix)
    # the reference's own LlamaForCausalLM subclass: CrossEntropyLoss(label_smoothing=0.1) (modeling_llama.py:107)
    assert smooth.item() == pytest.approx(rec["loss"], abs=2e-5)
    # the stock transformers loss on the same call = plain CE = what cgpt_ce_loss computes (known gap)
    assert loss.item() == pytest.approx(rec["loss_stock_transformers"], abs=2e-5)
    assert abs(rec["loss"] - rec["loss_stock_transformers"]) > 1e-3
    Lc = cond.shape[1]
    labels = torch.full((len(rows), Lc + na), -100, dtype=torch.long)
    labels[:, Lc:] = answers
    assert torch.equal(labels, rec["labels"])                                # only the answer positions are scored
    assert torch.equal(rec["attention_mask"][:, :Lc].long(), torch.ones(len(rows), Lc, dtype=torch.long))
    assert torch.equal(rec["attention_mask"][:, Lc:].long(), (answers >= 0).long())
    assert torch.allclose(cond, rec["inputs_embeds"][:, :Lc], atol=1e-6)     # [bos + prompt | image | rest of the prompt]
    emb = sd["llama_model.model.embed_tokens.weight"]
    assert torch.equal(emb[answers.clamp_min(0).masked_fill(answers < 0, cfg.llm.pad_id)], rec["inputs_embeds"][:, Lc:])
    assert tok.shape == answers.shape and bool((tok[answers < 0] == 0).all())
