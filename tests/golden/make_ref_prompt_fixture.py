"""Generates tests/golden/ref_prompt.json: the prompt the REFERENCE's evaluation call site hands to MiniGPTBase.generate.

The certify / predict agents are empty files in the reference; the call site of `generate` that does exist
is MiniGPT4EvalAgent.eval (agents/minigpt4_eval_agent.py:71-124): texts = prepare_texts(questions, CONV_VISION_minigptv2).
This script executes, unmodified,
  * graphs/models/minigpt4/conversation/conversation.py (by path; shim: common.registry) for CONV_VISION_minigptv2, and
  * the function prepare_texts of graphs/models/minigpt4/common/eval_utils.py:37-43 (its module imports nltk and the
    upstream minigpt4 package, so only this function's source is compiled, taken from the file's AST),
on questions formatted as VQAv2TestDataset.__getitem__ does (datasets/datasets/vqav2_dataset.py:201), and records the
texts plus the token ids MiniGPTBase.get_context_emb asks the tokenizer for (BOS from add_special_tokens AND the literal
"<s>" of the conversation role, i.e. two BOS ids).  Build container only (needs /root/reference).

    python tests/golden/make_ref_prompt_fixture.py
"""
import ast
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_generate_fixtures as G0  # noqa: E402
from ref_generate_util import CharTokenizer  # noqa: E402

REF = "/root/reference/graphs/models/minigpt4"
QUESTIONS = ["What color is the car?", "Is it raining!?", "How many (dogs)?"]


def main():
    base = G0.load_reference().MiniGPTBase                     # installs the common.registry shim as well
    spec = importlib.util.spec_from_file_location("ref_conversation", os.path.join(REF, "conversation", "conversation.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    tree = ast.parse(open(os.path.join(REF, "common", "eval_utils.py")).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "prepare_texts"][0]
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "eval_utils.py:prepare_texts", "exec"), ns)
    conv_temp = conv.CONV_VISION_minigptv2.copy()
    conv_temp.system = ""                                       # minigpt4_eval_agent.py:80-81
    dataset_questions = [f"[vqa] Based on the image, respond to this question with a short answer: {q}" for q in QUESTIONS]
    texts = ns["prepare_texts"](dataset_questions, conv_temp)

    vocab = 96
    calls = []

    class Recorder(CharTokenizer):
        def __call__(self, text, **kw):
            out = super().__call__(text, **kw)
            calls.append({"text": text, "add_special_tokens": kw.get("add_special_tokens", True), "ids": out.input_ids[0].tolist()})
            return out
    emb = torch.nn.Embedding(vocab, 8)
    stub = types.SimpleNamespace(llama_tokenizer=Recorder(vocab), embed_tokens=lambda ids: emb(ids))
    segs = []
    for t in texts:
        calls.clear()
        with torch.no_grad():
            mixed = base.get_context_emb(stub, t, [torch.zeros(1, 4, 8)])
        segs.append({"text": t, "segments": [dict(c) for c in calls], "rows": int(mixed.shape[1])})
    out = {"reference": "prepare_texts (eval_utils.py:37-43) + CONV_VISION_minigptv2 (conversation.py:130-137) + "
                        "MiniGPTBase.get_context_emb (minigpt_base.py:75-89), executed unmodified",
           "questions": QUESTIONS, "dataset_questions": dataset_questions, "vocab": vocab, "prompts": segs}
    json.dump(out, open(os.path.join(HERE, "ref_prompt.json"), "w"), indent=1)
    for s in segs:
        print(repr(s["text"]), [c["ids"][:4] for c in s["segments"]])


if __name__ == "__main__":
    main()
