"""Generates tests/golden/ref_generate.pt by running the REFERENCE's own MiniGPTBase.generate / get_context_emb /
embed_tokens (graphs/models/minigpt4/models/minigpt_base.py:75-89,366-448), executed unmodified by file path.

Runs only in the build container (needs /root/reference, read-only).  The module's imports are shimmed (torch_xla, the
registry, BaseModel, StoppingCriteriaSub - none of them is on the generate path) and the three methods are called on a
stub `self` that carries what they touch:
  * llama_model      this image's transformers.LlamaForCausalLM on seeded weights (the reference's LLM arithmetic is
                     third-party transformers; 4.30.0 pinned there, 5.5.0 here)
  * llama_tokenizer  a deterministic character tokenizer (no sentencepiece model ships with the reference): BOS = 1 only
                     when add_special_tokens, decode() prints "t<id>" words and honours skip_special_tokens
  * encode_img       returns seeded image embeddings [B, 32->n_query, hidden]
So the fixture pins the GLUE the oracle restates (oracle/model_oracle.py build_prompt_embeds / generate_ids /
canonical_answer): segment order around <ImageHere>, BOS on the first segment only, left padding and attention mask, the
generate() arguments, min_length, EOS / pad handling and the answer post-processing.

    python tests/golden/make_ref_generate_fixtures.py
"""
import contextlib
import functools
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import transformers  # noqa: E402  (before the shims)

from certifiedgpt_b200.config import LlmConfig, ModelConfig  # noqa: E402
from certifiedgpt_b200.weights import random_state_dict  # noqa: E402
from ref_generate_util import CASES, CharTokenizer, hf_llama  # noqa: E402

REF = "/root/reference/graphs/models/minigpt4/models/minigpt_base.py"


def load_reference():
    xla, amp = types.ModuleType("torch_xla"), types.ModuleType("torch_xla.amp")
    core, xm = types.ModuleType("torch_xla.core"), types.ModuleType("torch_xla.core.xla_model")
    amp.autocast = contextlib.nullcontext
    xm.master_print = lambda *a, **k: None
    xm.xla_device = lambda: torch.device("cpu")
    xla.amp, xla.core, core.xla_model = amp, core, xm
    sys.modules.update({"torch_xla": xla, "torch_xla.amp": amp, "torch_xla.core": core, "torch_xla.core.xla_model": xm})
    common, creg = types.ModuleType("common"), types.ModuleType("common.registry")

    class _Registry:
        def __getattr__(self, k):
            return lambda *a, **kw: (lambda f: f)
    creg.registry = _Registry()
    sys.modules.update({"common": common, "common.registry": creg})
    for pkg in ("graphs", "graphs.models", "graphs.models.minigpt4", "graphs.models.minigpt4.models",
                "graphs.models.minigpt4.conversation"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    bm = types.ModuleType("graphs.models.minigpt4.models.base_model")
    bm.BaseModel = type("BaseModel", (nn.Module,), {})
    conv = types.ModuleType("graphs.models.minigpt4.conversation.conversation")

    class StoppingCriteriaSub(transformers.StoppingCriteria):       # constructed by generate(), never passed on (:424)
        def __init__(self, stops=(), encounters=1):
            super().__init__()

        def __call__(self, input_ids, scores):
            return False
    conv.StoppingCriteriaSub = StoppingCriteriaSub
    sys.modules.update({"graphs.models.minigpt4.models.base_model": bm,
                        "graphs.models.minigpt4.conversation.conversation": conv})
    spec = importlib.util.spec_from_file_location("ref_minigpt_base", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference().MiniGPTBase
    out = {"reference": "MiniGPTBase.generate / get_context_emb / embed_tokens (minigpt_base.py, executed unmodified)",
           "cases": []}
    for case in CASES:
        cfg = ModelConfig.tiny()
        cfg.llm = LlmConfig(hidden=64, layers=2, heads=4, inter=128, vocab=96)
        sd = random_state_dict(cfg, seed=case["seed"])
        if case["eos_boost"]:
            sd["llama_model.lm_head.weight"][cfg.llm.eos_id] *= case["eos_boost"]
        hf = hf_llama(cfg, sd)
        g = torch.Generator().manual_seed(100 + case["seed"])
        img_embeds = torch.randn(case["B"], cfg.qf.n_query, cfg.llm.hidden, generator=g) * 0.3
        captured = {}
        real_generate = hf.generate

        def spy(**kw):
            captured["inputs_embeds"] = kw["inputs_embeds"].clone()
            captured["attention_mask"] = kw["attention_mask"].clone()
            captured["kwargs"] = {k: v for k, v in kw.items() if k not in ("inputs_embeds", "attention_mask")}
            captured["outputs"] = real_generate(**kw)
            return captured["outputs"]
        hf.generate = spy
        stub = types.SimpleNamespace(device=torch.device("cpu"), llama_model=hf, llama_tokenizer=CharTokenizer(cfg.llm.vocab),
                                     maybe_autocast=contextlib.nullcontext,
                                     encode_img=lambda images: (img_embeds, torch.ones(img_embeds.shape[:2], dtype=torch.long)))
        stub.embed_tokens = functools.partial(ref.embed_tokens, stub)
        stub.get_context_emb = functools.partial(ref.get_context_emb, stub)
        images = torch.zeros(case["B"], 3, 4, 4)                    # only handed to the stub encode_img
        answers = ref.generate(stub, images, case["texts"], max_new_tokens=case["max_new_tokens"])
        out["cases"].append({"case": case, "img_embeds": img_embeds, "inputs_embeds": captured["inputs_embeds"],
                             "attention_mask": captured["attention_mask"], "outputs": captured["outputs"],
                             "generate_kwargs": {k: (v if isinstance(v, (int, float, bool)) else str(v))
                                                 for k, v in captured["kwargs"].items()},
                             "answers": answers})
        print(case["name"], tuple(captured["inputs_embeds"].shape), captured["outputs"].tolist(), answers)
    torch.save(out, os.path.join(HERE, "ref_generate.pt"))


if __name__ == "__main__":
    main()
