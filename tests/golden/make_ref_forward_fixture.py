"""Generates tests/golden/ref_forward.pt by running the REFERENCE's own training / validation forward
MiniGPTBase.forward -> preparing_embedding -> prompt_wrap -> concat_emb_input_output (+ embed_tokens)
(graphs/models/minigpt4/models/minigpt_base.py:91-148,150-204,258-362, executed unmodified by file path).

Runs only in the build container (needs /root/reference, read-only).  Same shims and stub `self` as
make_ref_generate_fixtures.py; in addition
  * llama_model   the reference's OWN LlamaForCausalLM subclass (modeling_llama.py, executed unmodified by path over this
                  image's transformers 5.5 Llama; its loss is CrossEntropyLoss(label_smoothing=0.1), :107) on seeded
                  weights, behind a recorder of the one call forward() makes; the stock class's loss on the same call is
                  stored next to it
  * encode_img    oracle.model_oracle.encode_img on seeded images (the image tower is pinned separately,
                  make_ref_encode_img_fixture.py); the test regenerates the same embeddings
So the fixture pins the GLUE of the fine-tune step's forward: BOS embedding first, prompt segments around <ImageHere>
without special tokens, the answer + end_sym tokens right-padded, targets = -100 everywhere but the answer positions,
the attention mask, and the resulting mean cross-entropy.  oracle.model_oracle.lm_loss must reproduce loss and layout.

    python tests/golden/make_ref_forward_fixture.py
"""
import contextlib
import functools
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_generate_fixtures as G0  # noqa: E402

from certifiedgpt_b200.config import LlmConfig, ModelConfig  # noqa: E402
from certifiedgpt_b200.weights import random_state_dict  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402
from ref_generate_util import CharTokenizer, hf_llama  # noqa: E402

# style "v2_chat": chat_template set, prompt_template "[INST] {} [/INST]", end_sym "</s>" (the MiniGPT-v2 convention the
# evaluation agent's conversation template matches).  style "shipped": the reference's own fine-tune configs
# (configs/train_configs/vqav2_finetuning_noise_*.yaml: arch minigpt4, end_sym "###", prompt_template
# '###Human: {} ###Assistant: '): MiniGPT4 defines no chat_template, so forward() feeds the dataset's instruction RAW
# (minigpt_base.py:282-283) and the template only serves prompt_list, which instruction_input overrides (:274-279).
STYLES = {
    "v2_chat": dict(chat_template=True, prompt_template="[INST] {} [/INST]", end_sym="</s>"),
    "shipped": dict(prompt_template="###Human: {} ###Assistant: ", end_sym="###"),
}

CASES = [
    dict(name="shipped_config", seed=34, style="shipped", instruction="<Img><ImageHere></Img> [vqa] what color is the car ? ",
         answers=["red", "dark red"]),
    dict(name="equal_answers", seed=31, instruction="<Img><ImageHere></Img> [vqa] what color is the car ? ", answers=["red", "big"]),
    dict(name="ragged_answers", seed=32, instruction="<Img><ImageHere></Img> [vqa] is it raining ? ", answers=["no", "dark red", "2"]),
    dict(name="single", seed=33, instruction="<Img><ImageHere></Img> [vqa] Based on the image, respond to this question with a short answer: how many dogs ? ",
         answers=["three"]),
]


def load_reference_llama():
    """graphs/models/minigpt4/models/modeling_llama.py by path: the subclass whose forward adds `reduction` and builds
    CrossEntropyLoss(label_smoothing=0.1).  Two names it imports no longer exist in transformers 5 (docstring constants)."""
    import importlib.util
    import transformers.models.llama.modeling_llama as ml
    import transformers.utils as tu
    for name in ("LLAMA_INPUTS_DOCSTRING", "_CONFIG_FOR_DOC"):
        if not hasattr(ml, name):
            setattr(ml, name, "")
    for name in ("add_start_docstrings_to_model_forward", "replace_return_docstrings"):
        if not hasattr(tu, name):
            setattr(tu, name, lambda *a, **k: (lambda f: f))
    spec = importlib.util.spec_from_file_location("ref_modeling_llama",
                                                  "/root/reference/graphs/models/minigpt4/models/modeling_llama.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.LlamaForCausalLM


class LlamaSpy:
    """records the one call MiniGPTBase.forward makes and passes it to the reference's LlamaForCausalLM"""

    def __init__(self, model):
        self.model, self.base_model, self.calls = model, model.base_model, []

    def __call__(self, **kw):
        self.calls.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items()})
        return self.model(**kw)


def main():
    ref = G0.load_reference().MiniGPTBase
    RefLlama = load_reference_llama()
    out = {"reference": "MiniGPTBase.forward / preparing_embedding / prompt_wrap / concat_emb_input_output "
                        "(minigpt_base.py, executed unmodified)", "cases": []}
    for case in CASES:
        cfg = ModelConfig.tiny()
        cfg.llm = LlmConfig(hidden=64, layers=2, heads=4, inter=128, vocab=96)
        sd = random_state_dict(cfg, seed=case["seed"])
        hf = hf_llama(cfg, sd)                                   # stock transformers Llama on the seeded weights
        ref_llama = RefLlama(hf.config).eval()                   # the reference's subclass, same weights
        ref_llama.load_state_dict(hf.state_dict())
        spy = LlamaSpy(ref_llama)
        B = len(case["answers"])
        images = torch.randn(B, 3, cfg.vit.img_size, cfg.vit.img_size, generator=torch.Generator().manual_seed(case["seed"]))

        def encode_img(imgs):
            with torch.no_grad():
                e = mo.encode_img(sd, cfg, imgs)
            return e, torch.ones(e.shape[:2], dtype=torch.long)
        stub = types.SimpleNamespace(device=torch.device("cpu"), llama_model=spy, llama_tokenizer=CharTokenizer(cfg.llm.vocab),
                                     maybe_autocast=contextlib.nullcontext, encode_img=encode_img, prompt_list=[],
                                     max_txt_len=160, max_context_len=3800, **STYLES[case.get("style", "v2_chat")])
        for fn in ("embed_tokens", "prompt_wrap", "concat_emb_input_output", "preparing_embedding"):
            setattr(stub, fn, functools.partial(getattr(ref, fn), stub))
        samples = {"image": images, "instruction_input": [case["instruction"]] * B, "answer": case["answers"]}
        with torch.no_grad():
            res = ref.forward(stub, samples)
        assert res is not None and len(spy.calls) == 1, "the reference forward swallowed an exception"
        call = spy.calls[0]
        assert call["reduction"] == "mean"
        with torch.no_grad():                                    # the same call through the stock class: plain CE
            plain = hf(inputs_embeds=call["inputs_embeds"], attention_mask=call["attention_mask"], labels=call["labels"]).loss
        out["cases"].append({"case": case, "images": images, "loss": float(res["loss"]), "loss_stock_transformers": float(plain),
                             "attention_mask": call["attention_mask"], "labels": call["labels"],
                             "inputs_embeds": call["inputs_embeds"]})
        print(case["name"], tuple(call["inputs_embeds"].shape), float(res["loss"]), float(plain), call["labels"][-1].tolist()[-8:])
    torch.save(out, os.path.join(HERE, "ref_forward.pt"))


if __name__ == "__main__":
    main()
