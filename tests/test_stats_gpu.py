"""GPU parity: histogram, label hash and the fp64 statistics tail vs the oracle / golden KATs."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import smoothing_oracle as so

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "smoothing_kat.json")))
# The device tail evaluates the Clopper-Pearson quantile by bisection on the fp64 binomial tail
# (lgamma/exp), SciPy uses Boost's ibeta_inv: agreement is ~1e-12 relative, not bit-level.
RTOL = 1e-9


@pytest.mark.parametrize("row", KAT["certify"])
def test_certify_tail_known_answers(lib, row):
    sel = torch.tensor([1, 5, 2], dtype=torch.int64, device="cuda")
    est = torch.tensor([0, row["nA"], row["n"] - row["nA"]], dtype=torch.int64, device="cuda")
    lab, st = lib.certify_tail(sel, est, row["n"], row["alpha"], row["sigma"])
    lab, st = lab.cpu().tolist(), st.cpu().tolist()
    assert (lab[0] == -1) == row["abstain"]
    assert lab[1] == 1
    assert st[1] == pytest.approx(row["pABar"], rel=RTOL, abs=1e-300)
    assert st[0] == pytest.approx(row["radius"], rel=1e-7 if row["radius"] < 1e-2 else RTOL, abs=1e-300)


@pytest.mark.parametrize("row", KAT["predict"])
def test_predict_tail_known_answers(lib, row):
    counts = torch.zeros(6, dtype=torch.int64, device="cuda")
    counts[4] = row["count1"]
    counts[1] = row["count2"]
    lab, st = lib.predict_tail(counts, row["alpha"])
    lab, st = lab.cpu().tolist(), st.cpu().tolist()
    assert st[0] == pytest.approx(row["pvalue"], rel=RTOL)
    assert (lab[0] == -1) == row["abstain"]
    if not row["abstain"]:
        assert lab[0] == (4 if row["count1"] >= row["count2"] else 1)


def test_abstain_boundary_is_exact_for_every_count(lib):
    """pABar >= 0.5 must flip at exactly the same nA as SciPy for n = 100 and 1000."""
    for n in (100, 1000):
        for nA in range(n // 2 - 2, int(n * 0.7)):
            sel = torch.tensor([9, 1], dtype=torch.int64, device="cuda")
            est = torch.tensor([nA, n - nA], dtype=torch.int64, device="cuda")
            lab, _ = lib.certify_tail(sel, est, n, 0.001, 0.25)
            assert (lab[0].item() == -1) == (so.lower_confidence_bound(nA, n, 0.001) < 0.5), (n, nA)


def test_radius_sweep_matches_scipy(lib):
    n = 1000
    for nA in list(range(550, 1001, 25)) + [999, 1000]:
        est = torch.tensor([nA, n - nA], dtype=torch.int64, device="cuda")
        sel = torch.tensor([2, 1], dtype=torch.int64, device="cuda")
        lab, st = lib.certify_tail(sel, est, n, 0.001, 0.5)
        ref = so.certify_tail([2, 1], [nA, n - nA], n, 0.001, 0.5)
        assert lab[0].item() == ref[0]
        assert st[0].item() == pytest.approx(ref[1], rel=1e-8)


def test_argmax_ties_take_lowest_index(lib):
    sel = torch.tensor([3, 7, 7, 1], dtype=torch.int64, device="cuda")
    est = torch.tensor([0, 1000, 0, 0], dtype=torch.int64, device="cuda")
    lab, _ = lib.certify_tail(sel, est, 1000, 0.001, 0.5)
    assert lab.cpu().tolist() == [1, 1]
    big = torch.zeros(3130, dtype=torch.int64, device="cuda")
    big[3000] = 5
    big[17] = 5
    lab, _ = lib.certify_tail(big, big, 10, 0.001, 0.5)
    assert lab[1].item() == 17


@pytest.mark.parametrize("B,classes", [(1, 3), (1000, 10), (4097, 3130), (100000, 7)])
def test_label_hist_matches_count_arr(lib, B, classes):
    g = torch.Generator().manual_seed(B)
    labels = torch.randint(0, classes, (B,), generator=g, dtype=torch.int32)
    counts = torch.zeros(classes, dtype=torch.int64, device="cuda")
    lib.label_hist(labels.cuda(), counts)
    ref = so.SmoothOracle(None, classes, 0.0)._count_arr(labels.numpy(), classes)
    assert np.array_equal(counts.cpu().numpy(), ref)
    lib.label_hist(labels.cuda(), counts)  # accumulates across batches
    assert np.array_equal(counts.cpu().numpy(), 2 * ref)


def test_label_hist_counts_invalid(lib):
    labels = torch.tensor([0, 5, -1, 2, 99], dtype=torch.int32, device="cuda")
    counts = torch.zeros(5, dtype=torch.int64, device="cuda")
    invalid = torch.zeros(1, dtype=torch.int32, device="cuda")
    lib.label_hist(labels, counts, invalid)
    assert counts.cpu().tolist() == [1, 0, 1, 0, 0] and invalid.item() == 3


def test_argmax_rows(lib):
    torch.manual_seed(0)
    logits = torch.randn(300, 32000, device="cuda")
    logits[5, 100] = logits[5, 7] = 50.0  # tie -> lowest index
    idx, margin = lib.argmax_rows(logits, want_margin=True)
    assert torch.equal(idx.long(), logits.argmax(1))
    assert idx[5].item() == 7 and margin[5].item() == 0.0
    top2 = logits.topk(2, dim=1).values
    assert torch.allclose(margin, top2[:, 0] - top2[:, 1])
    idx2 = lib.argmax_rows(logits, suppress_col=7)
    assert idx2[5].item() == 100


def test_answer_labels(lib):
    entries = [([5, 6], 0), ([7], 1), ([5], 2), ([31999, 4, 9], 3)]
    keys, vals = lib.build_answer_table(entries)
    ids = torch.tensor([[5, 6, 2, 0], [7, 2, 0, 0], [5, 2, 8, 8], [31999, 4, 9, 11], [0, 5, 6, 2],
                        [6, 5, 2, 0], [2, 5, 6, 0], [31999, 4, 9, 2]], dtype=torch.int32, device="cuda")
    lab = lib.answer_labels(ids, keys, vals, other_label=4)
    assert lab.cpu().tolist() == [0, 1, 2, 4, 0, 4, 4, 3]


def test_answer_vocabulary_to_device_labels(lib):
    """text vocabulary -> token-sequence table -> device hash lookup agrees with the text-level normaliser."""
    from certifiedgpt_b200.answers import AnswerVocabulary
    vocab = AnswerVocabulary(["yes", "no", "2", "red car"])
    enc = lambda s: [3 + (ord(c) % 90) for c in s]           # toy tokenizer
    keys, vals = lib.build_answer_table(vocab.table_entries(enc))
    gen = ["yes", " Yes", "No.", "red car", "purple", " 2", "Red Car"]
    width = max(len(enc(g)) for g in gen) + 2
    rows = []
    for g in gen:
        ids = enc(g) + [2]                                    # EOS terminates the answer
        rows.append(ids + [0] * (width - len(ids)))
    ids = torch.tensor(rows, dtype=torch.int32, device="cuda")
    lab = lib.answer_labels(ids, keys, vals, other_label=vocab.other).cpu().tolist()
    assert lab == [vocab.label_of_text(g) for g in gen]
