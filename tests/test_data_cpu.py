"""CPU: VQAv2 loader + BLIP-2 processors (certifiedgpt_b200/data/vqav2.py) against the reference's definitions:
torchvision's Resize(bicubic) -> ToTensor -> Normalize (processors/base_processor.py:14-38) on a synthetic
VQAv2-shaped directory, the caption processor's regexes, the COCO file naming and the confidence weighting."""
import json
import os

import numpy as np
import pytest
import torch

from certifiedgpt_b200.data import vqav2 as V


import ref_data_util as U  # noqa: E402

REF_DATA = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_data.json")))


@pytest.fixture()
def fake_vqav2(tmp_path):
    return U.build(str(tmp_path))


def test_image_processor_equals_torchvision_pipeline(fake_vqav2):
    from PIL import Image
    from torchvision import transforms
    from torchvision.transforms.functional import InterpolationMode
    _, _, img_dir = fake_vqav2
    img = Image.open(os.path.join(img_dir, "COCO_train2014_000000000042.jpg"))
    for size in (224, 448):
        ref = transforms.Compose([transforms.Resize((size, size), interpolation=InterpolationMode.BICUBIC),
                                  transforms.ToTensor(), transforms.Normalize(V.MEAN, V.STD)])(img.convert("RGB"))
        got = V.ImageProcessor(size)(img)
        assert got.shape == (3, size, size) and torch.allclose(got, ref, atol=1e-6)
        px = V.ImageProcessor(size, normalize=False)(img)
        assert px.min() >= 0 and px.max() <= 1
        assert torch.allclose((px - torch.tensor(V.MEAN).view(3, 1, 1)) / torch.tensor(V.STD).view(3, 1, 1), ref, atol=1e-6)


def test_pre_caption_regexes():
    assert V.pre_caption('What  color is the "car"?') == "what color is the car ?"
    assert V.pre_caption("Is it raining!?") == "is it raining ?"
    assert V.pre_caption("a b c d", max_words=2) == "a b"
    assert V.pre_caption("How many (dogs)?\n") == "how many dogs ?"


def test_dataset_items_follow_the_reference_loader(fake_vqav2):
    qp, ap, img_dir = fake_vqav2
    ds = V.VQAv2Dataset([qp], [ap], img_dir, split="train", vis_processor=V.ImageProcessor(56), seed=1)
    assert len(ds) == 3                                      # the question without annotation is dropped
    assert ds.image_path(42).endswith("COCO_train2014_000000000042.jpg")
    w = ds.answer_weights(ds.annotations[0])
    assert w["red"] == pytest.approx(14 / 16) and w["dark red"] == pytest.approx(2 / 16) and w["blue"] == 0.0
    assert ds.answer_weights(ds.annotations[2]) == {"2": 1.0}            # empty answers skipped
    item = ds[1]
    assert item["image"].shape == (3, 56, 56) and item["question_id"] == 1001 and item["answer"] == "no"
    assert item["instruction_input"].startswith("<Img><ImageHere></Img> [vqa] ") and item["instruction_input"].endswith("is it raining ? ")
    picks = {ds.get_data(0)["answer"] for _ in range(60)}
    assert picks <= {"red", "dark red"} and "red" in picks                # zero-weight answers are never drawn


def test_prompt_split_and_agent_items(fake_vqav2):
    from certifiedgpt_b200.answers import AnswerVocabulary
    qp, ap, img_dir = fake_vqav2
    ds = V.VQAv2Dataset([qp], [ap], img_dir, split="train", vis_processor=V.ImageProcessor(56), seed=2)
    enc = lambda s: [3 + (ord(c) % 90) for c in s]
    prefix, suffix = V.split_prompt(ds[0]["instruction_input"], enc)
    assert prefix[0] == 1 and prefix[1:] == enc("[INST] <Img>")
    assert suffix[:len(enc("</Img> [vqa]"))] == enc("</Img> [vqa]") and suffix[-len(enc(" [/INST]")):] == enc(" [/INST]")
    vocab = AnswerVocabulary(["red", "no", "2"])
    items = V.certify_items(ds, vocab)
    assert [it["label"] for it in items] == [0, 1, 2]                    # "No." normalises to "no"
    ft = V.finetune_items(ds, enc)
    assert ft[1]["answer_ids"] == enc("no</s>") and ft[1]["image"].shape == (3, 56, 56)
    # every item carries ITS question as the token ids around <ImageHere> (what the agents hand to set_question)
    assert ft[1]["suffix_ids"] == V.split_prompt(ft[1]["instruction_input"], enc)[1] and ft[1]["prefix_ids"] == prefix
    assert "is it raining" in ft[1]["instruction_input"] and ft[0]["suffix_ids"] != ft[1]["suffix_ids"]
    with_q = V.certify_items(ds, vocab, encode=enc)
    assert [it["label"] for it in with_q] == [0, 1, 2]
    for it, d in zip(with_q, (ds.get_data(i) for i in range(len(ds)))):
        want = V.split_prompt(V.eval_prompt(V.EVAL_QUESTION_TEMPLATE.format(d["question"])), enc, prompt_template="{}")
        assert (it["prefix_ids"], it["suffix_ids"]) == want
    assert len({tuple(it["suffix_ids"]) for it in with_q}) == len(with_q)    # different questions, different ids
    assert len({tuple(it["prefix_ids"]) for it in with_q}) == 1             # one shared prefix (the engine's prefix KV)


@pytest.mark.parametrize("size", [56, 224])
def test_items_equal_the_reference_loader_run(fake_vqav2, size):
    """Every item of the reference's OWN VQAv2Dataset + Blip2ImageTrainProcessor + BlipCaptionProcessor (executed
    unmodified by tests/golden/make_ref_data_fixtures.py on the same files, global `random` seeded): question ids,
    instruction strings, sampled answers and the processed image."""
    qp, ap, img_dir = fake_vqav2
    rec = REF_DATA["image_sizes"][str(size)]
    for seed in (1, 2):
        ds = V.VQAv2Dataset([qp], [ap], img_dir, split="train", vis_processor=V.ImageProcessor(size), seed=seed)
        assert len(ds) == rec["len"]
        for want in [r for r in rec["items"] if r["seed"] == seed]:
            got = ds[want["index"]]
            assert got["question_id"] == want["question_id"]
            assert got["instruction_input"] == want["instruction_input"]
            assert got["answer"] == want["answer"]
            assert list(got["image"].shape) == want["shape"]
            dg = U.image_digest(got["image"])
            assert dg["sum"] == pytest.approx(want["image"]["sum"], rel=1e-6, abs=1e-3)
            assert dg["sumsq"] == pytest.approx(want["image"]["sumsq"], rel=1e-6)
            assert np.allclose(dg["pooled"], want["image"]["pooled"], atol=1e-5)


def test_text_processor_and_answer_sampling_equal_the_reference_run(fake_vqav2):
    for row in REF_DATA["pre_caption"]:
        assert V.pre_caption(row["in"]) == row["out"]
    for row in REF_DATA["pre_caption_max3"]:
        assert V.pre_caption(row["in"], max_words=3) == row["out"]
    qp, ap, img_dir = fake_vqav2
    ds = V.VQAv2Dataset([qp], [ap], img_dir, split="train", vis_processor=V.ImageProcessor(56), seed=3)
    assert [ds.get_data(0)["answer"] for _ in range(400)] == REF_DATA["answer_draws_seed3"]


def test_eval_split_equals_the_reference_test_dataset_run(fake_vqav2):
    """VQAv2TestDataset (questions only) against a run of the reference's own class on the same files."""
    qp, _, img_dir = fake_vqav2
    rec = REF_DATA["test_dataset"]
    ds = V.VQAv2TestDataset([qp], os.path.join(os.path.dirname(img_dir), "test2015"), split="test",
                            vis_processor=V.ImageProcessor(56))
    assert len(ds) == rec["len"] == 4                       # no annotation filter on the evaluation split
    for want in rec["items"]:
        got = ds[want["index"]]
        assert (got["question"], got["question_id"], got["img_id"]) == (want["question"], want["question_id"], want["img_id"])
        dg = U.image_digest(got["image"])
        assert dg["sumsq"] == pytest.approx(want["image"]["sumsq"], rel=1e-6)
        assert np.allclose(dg["pooled"], want["image"]["pooled"], atol=1e-5)


def test_eval_prompt_equals_the_reference_eval_call_site():
    """eval_prompt + split_prompt against tests/golden/ref_prompt.json: the texts the reference's prepare_texts +
    CONV_VISION_minigptv2 produce for generate(), and the token ids its get_context_emb requests (two BOS ids: the
    tokenizer's and the literal "<s>" of the conversation role)."""
    from ref_generate_util import encode_special
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_prompt.json")))
    enc = lambda s: encode_special(s, ref["vocab"])
    for q, dq, rec in zip(ref["questions"], ref["dataset_questions"], ref["prompts"]):
        assert V.EVAL_QUESTION_TEMPLATE.format(q) == dq
        text = V.eval_prompt(dq)
        assert text == rec["text"]
        prefix, suffix = V.split_prompt(text, enc, prompt_template="{}")
        seg0, seg1 = rec["segments"]
        assert seg0["add_special_tokens"] is True and seg1["add_special_tokens"] is False
        assert prefix == seg0["ids"] and prefix[:2] == [1, 1]
        assert suffix == seg1["ids"]
        assert rec["rows"] == len(prefix) + 4 + len(suffix)
