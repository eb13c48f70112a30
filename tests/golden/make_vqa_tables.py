"""Generates (in the build container, from the reference's own VQAEval, imported by file path):
  certifiedgpt_b200/data/vqa_answer_tables.json  the DATA tables of the official VQA answer normaliser
        (contractions, number words, articles, punctuation list; common/vqa_tools/vqa_eval.py:29-191)
  tests/golden/vqa_normalize_kat.json             input -> output vectors of the reference pipeline:
        eval agent pre-step  answer.lower().replace('<unk>','').strip()   (agents/minigpt4_eval_agent.py:102)
        VQAEval.evaluate     replace('\\n',' '), replace('\\t',' '), strip, processPunctuation,
                             processDigitArticle                           (vqa_eval.py:213-218,249-274)
    python tests/golden/make_vqa_tables.py
"""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
spec = importlib.util.spec_from_file_location("ref_vqa_eval", "/root/reference/common/vqa_tools/vqa_eval.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
ev = mod.VQAEval(None, None)

tables = {"_source": "official VQA evaluation tables as shipped in leodesouza/certifiedGPT common/vqa_tools/vqa_eval.py:29-191 (data only)",
          "contractions": ev.contractions, "manualMap": dict(ev.manualMap), "articles": ev.articles, "punct": ev.punct}
with open(os.path.join(ROOT, "certifiedgpt_b200", "data", "vqa_answer_tables.json"), "w") as f:
    json.dump(tables, f, indent=0, sort_keys=True)


def reference_pipeline(answer):
    a = answer.lower().replace("<unk>", "").strip()          # eval agent :102
    a = a.replace("\n", " ").replace("\t", " ").strip()      # vqa_eval.py:213-215
    a = ev.processPunctuation(a)
    return ev.processDigitArticle(a)


cases = ["Yes", " yes ", "No.", "two", "Two dogs", "a cat", "the red car", "An apple", "dont know", "It's 3,000",
         "1,000", "3.5", "U.S.A.", "hello, world", "what?!", "black and white", "<unk>frisbee", "  Frisbee\n",
         "none", "ten", "t-shirt", "ice-cream cone", "on the table.", "isnt", "youre right", "2 people",
         "a", "", "...", "stop sign", "New York", "skate board", "skateboard", "o'clock", "10:30", "half & half",
         "blue/green", "50%", "it is a dog , not a cat", "tennis racket\t", "yes</s>", "Playing Wii"]
kat = [{"in": c, "out": reference_pipeline(c)} for c in cases]
with open(os.path.join(HERE, "vqa_normalize_kat.json"), "w") as f:
    json.dump(kat, f, indent=0)
print("tables:", {k: len(v) for k, v in tables.items() if k != "_source"}, "vectors:", len(kat))
for k in kat[:12]:
    print(repr(k["in"]), "->", repr(k["out"]))
