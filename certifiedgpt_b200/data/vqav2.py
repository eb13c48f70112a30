"""VQAv2 loader and BLIP-2 processors feeding the certify / predict / fine-tune agents.

Host-side data plumbing restated from the reference, without torch_xla / omegaconf / registry:
  * ImageProcessor      processors/base_processor.py:14-38 (Resize bicubic -> ToTensor -> Normalize with the CLIP/BLIP
                        constants; `normalize=False` leaves the image in [0,1] for Smooth(noise_space="pixel"), where
                        the fused noise kernel applies the Normalize itself)
  * pre_caption         processors/blip_processors.py:90-114 (BlipCaptionProcessor: the text processor of questions
                        and answers)
  * VQAv2Dataset        datasets/datasets/vqav2_dataset.py:19-166 + base_dataset.py:17-52: questions / annotations JSON,
                        COCO file naming, confidence-weighted answer sampling, "[vqa] ..." instruction templates,
                        "<Img><ImageHere></Img> {} " wrapping
  * split_prompt        MiniGPTBase.get_context_emb (minigpt_base.py:75-89): the prompt is split at <ImageHere>; the
                        part before it (with BOS) is the shared prefix, the part after it the per-sample suffix
"""
import collections
import json
import os
import random
import re

import numpy as np
import torch

MEAN = (0.48145466, 0.4578275, 0.40821073)     # base_processor.py:17
STD = (0.26862954, 0.26130258, 0.27577711)     # base_processor.py:19
INSTRUCTION_TEMPLATES = ("[vqa] {}",
                         "[vqa] Based on the image, respond to this question with a short answer: {}")   # vqav2_dataset.py:39-42
# Training-forward text conventions.  Two exist in the reference and both are pinned to runs of its forward
# (tests/golden/ref_forward.pt):
#   * chat style (defaults below): models that set `chat_template` wrap the instruction as "[INST] {} [/INST]" and end the
#     answer with "</s>" (minigpt_base.py:282-283,295) - the convention the evaluation agent's conversation template
#     (EVAL_PROMPT_TEMPLATE) matches;
#   * the SHIPPED fine-tune configs (configs/train_configs/vqav2_finetuning_noise_*.yaml: arch minigpt4, end_sym "###"):
#     MiniGPT4 defines no chat_template, so the dataset's instruction is fed RAW and the answer ends with "###"
#     -> split_prompt(..., prompt_template=SHIPPED_PROMPT_TEMPLATE), finetune_items(..., end_sym=SHIPPED_END_SYM).
PROMPT_TEMPLATE = "[INST] {} [/INST]"
END_SYM = "</s>"
SHIPPED_PROMPT_TEMPLATE = "{}"
SHIPPED_END_SYM = "###"
# The evaluation call site of MiniGPTBase.generate (agents/minigpt4_eval_agent.py:80-96): prepare_texts
# (graphs/models/minigpt4/common/eval_utils.py:37-43) over CONV_VISION_minigptv2 (conversation/conversation.py:130-137,
# roles "<s>[INST] " / " [/INST]", empty separator) on the questions of VQAv2TestDataset (vqav2_dataset.py:201).
EVAL_QUESTION_TEMPLATE = "[vqa] Based on the image, respond to this question with a short answer: {}"
EVAL_PROMPT_TEMPLATE = "<s>[INST] <Img><ImageHere></Img> {} [/INST]"


class ImageProcessor:
    """blip2_image_train / blip2_image_val (identical transforms in the reference)."""

    def __init__(self, image_size=448, mean=None, std=None, normalize=True):
        self.image_size = image_size
        self.mean = torch.tensor(mean or MEAN, dtype=torch.float32).view(3, 1, 1)
        self.std = torch.tensor(std or STD, dtype=torch.float32).view(3, 1, 1)
        self.normalize = normalize

    def __call__(self, image):
        from PIL import Image
        img = image.convert("RGB").resize((self.image_size, self.image_size), Image.BICUBIC)   # transforms.Resize on PIL
        x = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div_(255.0)   # ToTensor
        return (x - self.mean) / self.std if self.normalize else x


def pre_caption(caption, max_words=50):
    caption = re.sub(r"([.!\"()*#:;~])", " ", caption.lower())
    caption = re.sub(r"\s{2,}", " ", caption)
    caption = caption.rstrip("\n").strip(" ")
    words = caption.split(" ")
    if len(words) > max_words:
        caption = " ".join(words[:max_words])
    return caption


class VQAv2Dataset:
    def __init__(self, questions_paths, annotation_paths, vis_paths, split="train", vis_processor=None,
                 text_processor=pre_caption, seed=0):
        self.vis_paths, self.split = vis_paths, split
        self.vis_processor = vis_processor or ImageProcessor()
        self.text_processor = text_processor
        self.rng = random.Random(seed)            # the reference draws from the global `random`; seeded here
        questions, self.annotations = [], []
        for p in questions_paths:
            q = json.load(open(p))
            if isinstance(q, dict):
                questions.extend(q["questions"])
        for p in annotation_paths:
            a = json.load(open(p))
            if isinstance(a, dict):
                self.annotations.extend(a["annotations"])
        by_id = {q["question_id"]: q for q in questions}
        # keep the questions that have an annotation (vqav2_dataset.py:57-75)
        self.questions = [by_id[a["question_id"]] for a in self.annotations
                          if a.get("question_id") is not None and a["question_id"] in by_id]
        self.questions_dict = {q["question_id"]: q for q in self.questions}

    def __len__(self):
        return len(self.questions)

    def image_path(self, image_id):
        return os.path.join(self.vis_paths, f"COCO_{self.split}2014_{image_id:012d}.jpg")       # vqav2_dataset.py:103-104

    @staticmethod
    def answer_weights(annotation):
        """confidence-weighted answer distribution (vqav2_dataset.py:114-135): yes = 2, maybe = 1, else 0."""
        w = collections.defaultdict(float)
        for a in annotation["answers"]:
            ans = a.get("answer")
            if not ans:
                continue
            w[ans] += {"yes": 2, "maybe": 1}.get(a.get("answer_confidence"), 0)
        total = sum(w.values())
        if total > 0:
            for k in w:
                w[k] /= total
        return dict(w)

    def get_data(self, index):
        from PIL import Image
        ann = self.annotations[index]
        if "image_id" not in ann or "question_id" not in ann or "answers" not in ann:
            raise ValueError(f"Invalid annotation at index {index}: {ann}")
        if len(ann["answers"]) == 0:
            raise ValueError(f"No answers found for question_id {ann['question_id']}")
        question = self.text_processor(self.questions_dict[ann["question_id"]]["question"])
        image = self.vis_processor(Image.open(self.image_path(ann["image_id"])))
        w = self.answer_weights(ann)
        answers, weights = list(w.keys()), list(w.values())
        answer = self.rng.choices(answers, weights=weights if sum(weights) > 0 else None, k=1)[0]
        return {"image": image, "question": question, "question_id": ann["question_id"],
                "answer": self.text_processor(answer), "answer_weights": w}

    def __getitem__(self, index):
        d = self.get_data(index)
        instruction = self.rng.choice(INSTRUCTION_TEMPLATES).format(d["question"])
        return {"image": d["image"], "question_id": d["question_id"],
                "instruction_input": "<Img><ImageHere></Img> {} ".format(instruction), "answer": d["answer"],
                "answer_weights": d["answer_weights"]}


class VQAv2TestDataset:
    """Questions-only evaluation split (vqav2_dataset.py:173-205): COCO_<split>2015 file names, the RAW question inside
    the short-answer instruction (no text processor), no answers."""

    def __init__(self, questions_paths, vis_paths, split="test", vis_processor=None):
        self.vis_paths, self.split = vis_paths, split
        self.vis_processor = vis_processor or ImageProcessor()
        self.questions = []
        for p in questions_paths:
            q = json.load(open(p))
            if isinstance(q, dict):
                self.questions.extend(q["questions"])

    def __len__(self):
        return len(self.questions)

    def image_path(self, image_id):
        return os.path.join(self.vis_paths, f"COCO_{self.split}2015_{image_id:012d}.jpg")

    def __getitem__(self, idx):
        from PIL import Image
        d = self.questions[idx]
        return {"image": self.vis_processor(Image.open(self.image_path(d["image_id"]))),
                "question": EVAL_QUESTION_TEMPLATE.format(d["question"]), "question_id": d["question_id"],
                "img_id": d["image_id"]}


def eval_prompt(question):
    """The text the reference's evaluation loop hands to generate() for one dataset question (see EVAL_PROMPT_TEMPLATE).
    It starts with the literal "<s>" of the conversation role AND get_context_emb adds the tokenizer's BOS to the first
    segment (minigpt_base.py:79-82), so with a Llama tokenizer the prompt opens with two BOS ids:
    `split_prompt(eval_prompt(q), encode, prompt_template="{}")` reproduces exactly that when `encode` maps the literal
    "<s>" to the BOS id, as LlamaTokenizer does."""
    return EVAL_PROMPT_TEMPLATE.format(question)


def split_prompt(instruction_input, encode, bos_id=1, prompt_template=PROMPT_TEMPLATE):
    """(prefix_ids, suffix_ids) around <ImageHere> for the engines: get_context_emb (minigpt_base.py:75-89) tokenises
    the segments separately, BOS only in front of the first one."""
    before, after = prompt_template.format(instruction_input).split("<ImageHere>")
    return [bos_id] + [int(t) for t in encode(before)], [int(t) for t in encode(after)]


def certify_items(dataset, vocabulary, indices=None, encode=None):
    """{"image", "label"[, "prefix_ids", "suffix_ids"]} items for MiniGPT4CertifyAgent / MiniGPT4PredictAgent: label =
    class of the most probable ground-truth answer under the answer vocabulary
    (certifiedgpt_b200.answers.AnswerVocabulary).  With `encode` (the Llama tokenizer's text -> ids) every item carries
    ITS question as token ids around <ImageHere>, in the evaluation call site's wording (`eval_prompt`): the agents hand
    `suffix_ids` to engine.set_question before certifying the item, and the engine is built with the longest one.
    Without `encode` the engine's fixed prompt is used - only meaningful when every item asks the same question."""
    out = []
    for i in (range(len(dataset)) if indices is None else indices):
        d = dataset.get_data(i)
        best = max(d["answer_weights"].items(), key=lambda kv: kv[1])[0] if d["answer_weights"] else d["answer"]
        item = {"image": d["image"], "label": vocabulary.label_of_text(best), "question_id": d["question_id"]}
        if encode is not None:
            question = EVAL_QUESTION_TEMPLATE.format(d["question"])
            item["prefix_ids"], item["suffix_ids"] = split_prompt(eval_prompt(question), encode, prompt_template="{}")
        out.append(item)
    return out


def finetune_items(dataset, encode, indices=None, end_sym=END_SYM, max_txt_len=160, prompt_template=PROMPT_TEMPLATE):
    """{"image", "answer_ids", "suffix_ids"} items for MiniGPT4FineTuneAgent: answer + end_sym tokenised without special
    tokens (minigpt_base.py:297-311); suffix_ids = the item's own instruction as the token ids after <ImageHere>
    (MiniGPTBase.forward trains every sample on ITS instruction_input, minigpt_base.py:323-362)."""
    out = []
    for i in (range(len(dataset)) if indices is None else indices):
        d = dataset[i]
        prefix, suffix = split_prompt(d["instruction_input"], encode, prompt_template=prompt_template)
        out.append({"image": d["image"], "answer_ids": [int(t) for t in encode(d["answer"] + end_sym)][:max_txt_len],
                    "question_id": d["question_id"], "instruction_input": d["instruction_input"],
                    "prefix_ids": prefix, "suffix_ids": suffix})
    return out
