"""CPU: VQAv2 loader + BLIP-2 processors (certifiedgpt_b200/data/vqav2.py) against the reference's definitions:
torchvision's Resize(bicubic) -> ToTensor -> Normalize (processors/base_processor.py:14-38) on a synthetic
VQAv2-shaped directory, the caption processor's regexes, the COCO file naming and the confidence weighting."""
import json
import os

import numpy as np
import pytest
import torch

from certifiedgpt_b200.data import vqav2 as V


@pytest.fixture()
def fake_vqav2(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    img_dir = tmp_path / "train2014"
    img_dir.mkdir()
    ids = [42, 9, 123456]
    for i, iid in enumerate(ids):
        arr = rng.integers(0, 256, size=(60 + 10 * i, 80, 3), dtype=np.uint8)
        Image.fromarray(arr).save(img_dir / f"COCO_train2014_{iid:012d}.jpg", quality=95)
    questions = {"questions": [{"question_id": 1000 + i, "image_id": iid, "question": q}
                               for i, (iid, q) in enumerate(zip(ids, ["What color is the car?", "Is it raining!?", "How many (dogs)?"]))]}
    questions["questions"].append({"question_id": 7, "image_id": 9, "question": "unused"})
    ann = {"annotations": [
        {"question_id": 1000, "image_id": 42, "answers": [{"answer": "red", "answer_confidence": "yes"}] * 7
         + [{"answer": "dark red", "answer_confidence": "maybe"}] * 2 + [{"answer": "blue", "answer_confidence": "no"}]},
        {"question_id": 1001, "image_id": 9, "answers": [{"answer": "No.", "answer_confidence": "yes"}] * 10},
        {"question_id": 1002, "image_id": 123456, "answers": [{"answer": "2", "answer_confidence": "maybe"}] * 3
         + [{"answer": "", "answer_confidence": "yes"}]},
    ]}
    qp, ap = tmp_path / "q.json", tmp_path / "a.json"
    qp.write_text(json.dumps(questions))
    ap.write_text(json.dumps(ann))
    return str(qp), str(ap), str(img_dir)


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
