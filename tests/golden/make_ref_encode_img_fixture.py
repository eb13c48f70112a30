"""Generates tests/golden/ref_encode_img.pt by running the REFERENCE's own MiniGPT4.encode_img
(graphs/models/minigpt4/models/minigpt4.py:121-149, executed unmodified by file path) over the reference's own
VisionTransformer (eva_vit.py) and Q-Former layers (Qformer.py) on seeded weights.

Runs only in the build container (needs /root/reference, read-only).  Shims: the module imports that this image lacks
(torch_xla, registry, BaseModel / disabled_train, timm, transformers-5 renames - as in make_ref_fixtures.py and
make_ref_generate_fixtures.py).  encode_img is called on a stub `self` holding
  visual_encoder  the reference VisionTransformer
  ln_vision       LayerNorm(eps 1e-5) in fp32 (base_model.py:281-287 - that file cannot be imported: peft, torch_xla)
  query_tokens    the state dict's query_tokens
  Qformer.bert    the reference BertEmbeddings + BertEncoder behind the query-only path of BertModel.forward
                  (Qformer.py:804-965: all-ones masks become all-zero additive masks; BertModel.__init__ itself breaks
                  under transformers 5)
  llama_proj      nn.Linear with the state dict's weights
so the fixture pins the whole image tower A6-A10 and its wiring (ViT -> ln_vision -> expanded queries -> Q-Former ->
llama_proj, attention masks of ones) in one tensor: oracle.model_oracle.encode_img must reproduce inputs_llama.

    python tests/golden/make_ref_encode_img_fixture.py
"""
import contextlib
import os
import sys
import types
from functools import partial

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_fixtures as F0  # noqa: E402  (installs the transformers renames; provides _install_shims / _load)
import make_ref_generate_fixtures as G0  # noqa: E402

REF = F0.REF


def main():
    F0._install_shims()
    from certifiedgpt_b200.config import LlmConfig, ModelConfig, QFormerConfig, VitConfig
    from certifiedgpt_b200.weights import random_state_dict
    eva = F0._load("ref_eva_vit", os.path.join(REF, "eva_vit.py"))
    qf = F0._load("ref_qformer", os.path.join(REF, "Qformer.py"))
    base = G0.load_reference()                                     # minigpt_base.py under its shims
    sys.modules["graphs.models.minigpt4.models.minigpt_base"] = base
    sys.modules["graphs.models.minigpt4.models.Qformer"] = qf
    sys.modules["graphs.models.minigpt4.models.base_model"].disabled_train = lambda self, mode=True: self
    mg = F0._load("ref_minigpt4", os.path.join(REF, "minigpt4.py"))

    out = {}
    cases = {"tiny": ModelConfig.tiny(),
             "wide": ModelConfig(vit=VitConfig(img_size=56, depth=1), qf=QFormerConfig(layers=2),
                                 llm=LlmConfig(layers=1, inter=128, vocab=96))}      # llama_proj keeps its 4096-wide output
    for name, cfg in cases.items():
        sd = random_state_dict(cfg, seed=21)
        v, q = cfg.vit, cfg.qf
        vit = eva.VisionTransformer(img_size=v.img_size, patch_size=v.patch, use_mean_pooling=False, embed_dim=v.dim,
                                    depth=v.depth, num_heads=v.heads, mlp_ratio=v.mlp / v.dim, qkv_bias=True,
                                    norm_layer=partial(nn.LayerNorm, eps=v.eps)).eval()
        vsd = {k[len("visual_encoder."):]: t for k, t in sd.items() if k.startswith("visual_encoder.")}
        missing, unexpected = vit.load_state_dict(vsd, strict=False)
        assert not unexpected and all("relative_position" in m for m in missing), (missing, unexpected)
        ln = nn.LayerNorm(v.dim, eps=cfg.ln_vision_eps if hasattr(cfg, "ln_vision_eps") else 1e-5).eval()
        ln.load_state_dict({"weight": sd["ln_vision.weight"], "bias": sd["ln_vision.bias"]})
        bcfg = qf.BertConfig(hidden_size=q.hidden, num_hidden_layers=q.layers, num_attention_heads=q.heads,
                             intermediate_size=q.inter, layer_norm_eps=q.eps, hidden_dropout_prob=0.0,
                             attention_probs_dropout_prob=0.0)
        bcfg.encoder_width, bcfg.add_cross_attention = v.dim, True
        bcfg.cross_attention_freq, bcfg.query_length = q.cross_freq, q.n_query
        emb, enc = qf.BertEmbeddings(bcfg).eval(), qf.BertEncoder(bcfg).eval()
        emb.word_embeddings = emb.position_embeddings = None       # minigpt4.py:104-109
        for layer in enc.layer:
            layer.output = layer.intermediate = None
        m1 = emb.load_state_dict({k[len("Qformer.bert.embeddings."):]: t for k, t in sd.items()
                                  if k.startswith("Qformer.bert.embeddings.")}, strict=False)
        m2 = enc.load_state_dict({k[len("Qformer.bert.encoder."):]: t for k, t in sd.items()
                                  if k.startswith("Qformer.bert.encoder.")}, strict=False)
        assert not m1.unexpected_keys and not m2.unexpected_keys and not m2.missing_keys, (m1, m2)

        def bert(query_embeds, encoder_hidden_states, encoder_attention_mask, return_dict=True):
            # BertModel.forward for query-only input (Qformer.py:869-949): ones masks -> zero additive masks
            B, nq = query_embeds.shape[:2]
            assert bool((encoder_attention_mask == 1).all())
            h = emb(query_embeds=query_embeds)
            return enc(h, attention_mask=torch.zeros(B, 1, 1, nq), head_mask=[None] * q.layers,
                       encoder_hidden_states=encoder_hidden_states,
                       encoder_attention_mask=torch.zeros(B, 1, 1, encoder_hidden_states.shape[1]),
                       query_length=nq, return_dict=True)
        proj = nn.Linear(q.hidden, cfg.llm.hidden)
        proj.load_state_dict({"weight": sd["llama_proj.weight"], "bias": sd["llama_proj.bias"]})
        stub = types.SimpleNamespace(visual_encoder=vit, ln_vision=ln, has_qformer=True, query_tokens=sd["query_tokens"],
                                     Qformer=types.SimpleNamespace(bert=bert), llama_proj=proj.eval(),
                                     maybe_autocast=contextlib.nullcontext)
        images = torch.randn(3, 3, v.img_size, v.img_size, generator=torch.Generator().manual_seed(7))
        with torch.no_grad():
            inputs_llama, atts = mg.MiniGPT4.encode_img(stub, images)
        assert bool((atts == 1).all()) and atts.shape == inputs_llama.shape[:2]
        stride = 1 if inputs_llama.shape[-1] <= 256 else 8          # keep the committed fixture small
        out[name] = {"seed": 21, "images": images, "stride": stride, "inputs_llama": inputs_llama[..., ::stride].clone(),
                     "sum": float(inputs_llama.double().sum()), "abs_sum": float(inputs_llama.double().abs().sum())}
        print(name, tuple(inputs_llama.shape), float(inputs_llama.abs().mean()))
    torch.save(out, os.path.join(HERE, "ref_encode_img.pt"))


if __name__ == "__main__":
    main()
