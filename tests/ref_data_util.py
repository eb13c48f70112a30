"""Synthetic VQAv2-shaped directory shared by tests/golden/make_ref_data_fixtures.py (which runs the REFERENCE's loader and
processors on it) and tests/test_data_cpu.py (which runs ours on the same files).  Images are lossless PNG bytes under the
COCO .jpg names the reference hard-codes (PIL detects the format from the header), so both sides decode identical pixels."""
import json
import os

import numpy as np

IDS = [42, 9, 123456]
QUESTIONS = ["What color is the car?", "Is it raining!?", "How many (dogs)?"]
CAPTIONS = ['What  color is the "car"?', "Is it raining!?", "How many (dogs)?\n", "A:B;C~D  *E#", " leading and trailing  ",
            "one two three four five six seven eight nine ten eleven twelve"]


def build(root):
    """-> (questions.json, annotations.json, image dir)"""
    from PIL import Image
    rng = np.random.default_rng(0)
    img_dir = os.path.join(root, "train2014")
    os.makedirs(img_dir, exist_ok=True)
    for i, iid in enumerate(IDS):
        arr = rng.integers(0, 256, size=(60 + 10 * i, 80, 3), dtype=np.uint8)
        with open(os.path.join(img_dir, f"COCO_train2014_{iid:012d}.jpg"), "wb") as f:
            Image.fromarray(arr).save(f, format="PNG")
    test_dir = os.path.join(root, "test2015")
    os.makedirs(test_dir, exist_ok=True)
    for iid in IDS:                                        # same pixels under the evaluation split's file names
        src = os.path.join(img_dir, f"COCO_train2014_{iid:012d}.jpg")
        with open(src, "rb") as f, open(os.path.join(test_dir, f"COCO_test2015_{iid:012d}.jpg"), "wb") as g:
            g.write(f.read())
    questions = {"questions": [{"question_id": 1000 + i, "image_id": iid, "question": q}
                               for i, (iid, q) in enumerate(zip(IDS, QUESTIONS))]}
    questions["questions"].append({"question_id": 7, "image_id": 9, "question": "unused"})
    ann = {"annotations": [
        {"question_id": 1000, "image_id": 42, "answers": [{"answer": "red", "answer_confidence": "yes"}] * 7
         + [{"answer": "dark red", "answer_confidence": "maybe"}] * 2 + [{"answer": "blue", "answer_confidence": "no"}]},
        {"question_id": 1001, "image_id": 9, "answers": [{"answer": "No.", "answer_confidence": "yes"}] * 10},
        {"question_id": 1002, "image_id": 123456, "answers": [{"answer": "2", "answer_confidence": "maybe"}] * 3
         + [{"answer": "", "answer_confidence": "yes"}]},
    ]}
    qp, ap = os.path.join(root, "q.json"), os.path.join(root, "a.json")
    json.dump(questions, open(qp, "w"))
    json.dump(ann, open(ap, "w"))
    return qp, ap, img_dir


def image_digest(t):
    """Small, version-robust summary of a [3,S,S] image tensor: fp64 sum, sum of squares and an 8x8 average pooling."""
    import torch
    d = t.double()
    pooled = torch.nn.functional.adaptive_avg_pool2d(d[None], 8)[0]
    return {"sum": float(d.sum()), "sumsq": float((d * d).sum()), "pooled": [round(float(v), 9) for v in pooled.flatten()]}
