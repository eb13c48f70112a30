"""CPU: the answer normaliser against golden vectors produced by the reference's own VQAEval
(tests/golden/vqa_normalize_kat.json, made by tests/golden/make_vqa_tables.py), and the answer vocabulary."""
import json
import os

import pytest

from certifiedgpt_b200 import answers as A

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLD, "vqa_normalize_kat.json")))


@pytest.mark.parametrize("row", KAT, ids=[repr(r["in"])[:20] for r in KAT])
def test_normalize_matches_reference(row):
    assert A.normalize_answer(row["in"]) == row["out"]


def test_tables_are_the_official_ones():
    t = A._tables()
    assert len(t["contractions"]) == 120 and t["manualMap"]["ten"] == "10" and t["articles"] == ["a", "an", "the"]


def test_vocabulary_and_table_entries():
    vocab = A.AnswerVocabulary(["yes", "no", "2", "red car", "don't know"])
    assert vocab.num_classes == 6 and vocab.other == 5
    assert vocab.label_of_text(" Yes. ") == 0 and vocab.label_of_text("Two") == 2
    assert vocab.label_of_text("the red car") == 3 and vocab.label_of_text("dont know") == 4
    assert vocab.label_of_text("purple") == 5
    # toy tokenizer: one id per character (offset past the special ids)
    enc = lambda s: [3 + (ord(c) % 90) for c in s]
    entries = vocab.table_entries(enc)
    assert entries and all(isinstance(k, list) and 0 <= v < 5 for k, v in entries)
    keys = [tuple(k) for k, _ in entries]
    assert len(keys) == len(set(keys))                       # one class per token sequence
    assert (enc("yes"), 0) in entries and (enc(" Yes"), 0) in entries
