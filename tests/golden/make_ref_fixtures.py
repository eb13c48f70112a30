"""Generates tests/golden/ref_vit.pt and ref_qformer.pt from the REFERENCE's own modules.

Runs only in the build container (needs /root/reference, read-only).  The reference's
graphs/models/minigpt4/models/eva_vit.py and Qformer.py are imported by file path behind small
shims for packages this image lacks (timm, omegaconf/registry) and for transformers-5 renames
(SURVEY.md 8c); they are executed unmodified on weights produced by
certifiedgpt_b200.weights.random_state_dict (seeded), and inputs + outputs are saved.
tests/test_oracle_model_cpu.py then checks oracle/model_oracle.py against these tensors.

    python tests/golden/make_ref_fixtures.py
"""
import importlib.util
import os
import sys
import types
from functools import partial

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/graphs/models/minigpt4/models"

import transformers  # noqa: E402  (must be imported before the fake timm is installed)
import transformers.modeling_utils as mu  # noqa: E402
import transformers.pytorch_utils as pu  # noqa: E402

for name in ("apply_chunking_to_forward", "prune_linear_layer", "find_pruneable_heads_and_indices"):
    if not hasattr(mu, name):
        setattr(mu, name, getattr(pu, name, lambda *a, **k: None))


def _install_shims():
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    registry_m = types.ModuleType("timm.models.registry")
    layers.drop_path = lambda x, p=0.0, training=False: x
    layers.to_2tuple = lambda v: (v, v) if not isinstance(v, tuple) else v
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    registry_m.register_model = lambda f: f
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers,
                        "timm.models.registry": registry_m})
    common = types.ModuleType("common")
    creg = types.ModuleType("common.registry")

    class _Reg:
        def __getattr__(self, k):
            return lambda *a, **kw: (lambda f: f)
    creg.registry = _Reg()
    sys.modules.update({"common": common, "common.registry": creg})
    for pkg in ("graphs", "graphs.models", "graphs.models.minigpt4", "graphs.models.minigpt4.common"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    du = types.ModuleType("graphs.models.minigpt4.common.dist_utils")
    du.download_cached_file = lambda *a, **k: None
    sys.modules["graphs.models.minigpt4.common.dist_utils"] = du


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    _install_shims()
    from certifiedgpt_b200.config import ModelConfig, VitConfig, QFormerConfig
    from certifiedgpt_b200.weights import random_state_dict

    eva = _load("ref_eva_vit", os.path.join(REF, "eva_vit.py"))
    qf = _load("ref_qformer", os.path.join(REF, "Qformer.py"))

    out_vit, out_qf = {}, {}
    cases = {
        "tiny": ModelConfig.tiny(),
        # full ViT-g / Q-Former widths, 1 layer each, small image: exercises 1408/88/6144 and 768/64/3072
        "wide": ModelConfig(vit=VitConfig(img_size=56, depth=1), qf=QFormerConfig(layers=2)),
    }
    for name, cfg in cases.items():
        sd = random_state_dict(cfg, seed=11, parts=("vit", "qf"))
        v = cfg.vit
        model = eva.VisionTransformer(img_size=v.img_size, patch_size=v.patch, use_mean_pooling=False,
                                      embed_dim=v.dim, depth=v.depth, num_heads=v.heads,
                                      mlp_ratio=v.mlp / v.dim, qkv_bias=True,
                                      norm_layer=partial(nn.LayerNorm, eps=v.eps)).eval()
        vsd = {k[len("visual_encoder."):]: t for k, t in sd.items() if k.startswith("visual_encoder.")}
        missing, unexpected = model.load_state_dict(vsd, strict=False)
        assert not unexpected and all("relative_position" in m for m in missing), (missing, unexpected)
        assert model.blocks[0].mlp.fc1.out_features == v.mlp
        g = torch.Generator().manual_seed(5)
        images = torch.randn(2, 3, v.img_size, v.img_size, generator=g)
        with torch.no_grad():
            feats = model.forward_features(images)
        out_vit[name] = {"images": images, "features": feats}

        # Q-Former: BertEmbeddings + BertEncoder built directly (BertModel.__init__ breaks under
        # transformers 5), text parts nulled as minigpt4.py:104-109 does
        q = cfg.qf
        bcfg = qf.BertConfig(hidden_size=q.hidden, num_hidden_layers=q.layers, num_attention_heads=q.heads,
                             intermediate_size=q.inter, layer_norm_eps=q.eps, hidden_dropout_prob=0.0,
                             attention_probs_dropout_prob=0.0)
        bcfg.encoder_width = v.dim
        bcfg.add_cross_attention = True
        bcfg.cross_attention_freq = q.cross_freq
        bcfg.query_length = q.n_query
        emb = qf.BertEmbeddings(bcfg).eval()
        enc = qf.BertEncoder(bcfg).eval()
        emb.word_embeddings = None
        emb.position_embeddings = None
        for layer in enc.layer:
            layer.output = None
            layer.intermediate = None
        esd = {k[len("Qformer.bert.embeddings."):]: t for k, t in sd.items()
               if k.startswith("Qformer.bert.embeddings.")}
        nsd = {k[len("Qformer.bert.encoder."):]: t for k, t in sd.items()
               if k.startswith("Qformer.bert.encoder.")}
        m1 = emb.load_state_dict(esd, strict=False)
        m2 = enc.load_state_dict(nsd, strict=False)
        assert not m1.unexpected_keys and not m2.unexpected_keys and not m2.missing_keys, (m1, m2)
        assert [l.has_cross_attention for l in enc.layer] == [i % q.cross_freq == 0 for i in range(q.layers)]
        g = torch.Generator().manual_seed(6)
        img_embeds = torch.randn(2, v.tokens, v.dim, generator=g)
        B = 2
        with torch.no_grad():
            h = emb(query_embeds=sd["query_tokens"].expand(B, -1, -1))
            res = enc(h, attention_mask=torch.zeros(B, 1, 1, q.n_query), head_mask=[None] * q.layers,
                      encoder_hidden_states=img_embeds,
                      encoder_attention_mask=torch.zeros(B, 1, 1, v.tokens), query_length=q.n_query,
                      return_dict=True)
        out_qf[name] = {"image_embeds": img_embeds, "last_hidden_state": res.last_hidden_state}
        print(name, "vit", tuple(feats.shape), "qformer", tuple(res.last_hidden_state.shape))

    torch.save(out_vit, os.path.join(HERE, "ref_vit.pt"))
    torch.save(out_qf, os.path.join(HERE, "ref_qformer.pt"))


def make_pos_interp_fixture():
    """ref_pos_interp.pt: the reference's interpolate_pos_embed (eva_vit.py:383-404) 4x4 -> 8x8."""
    _install_shims()
    eva = _load("ref_eva_vit", os.path.join(REF, "eva_vit.py"))

    class FakeModel:
        class patch_embed:
            num_patches = 64
        pos_embed = torch.zeros(1, 65, 16)
    inp = torch.randn(1, 17, 16, generator=torch.Generator().manual_seed(12))
    ck = {"pos_embed": inp.clone()}
    eva.interpolate_pos_embed(FakeModel, ck)
    torch.save({"in": inp, "out": ck["pos_embed"]}, os.path.join(HERE, "ref_pos_interp.pt"))


if __name__ == "__main__":
    main()
    make_pos_interp_fixture()
