"""CPU tests: the oracle against known-answer vectors (no GPU, no libcgpt compute)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import philox_oracle as po
from oracle import smoothing_oracle as so

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLD, "smoothing_kat.json")))


def test_philox_random123_known_answers():
    # Salmon et al., Random123 kat_vectors: philox4x32 10 rounds
    f = 0xFFFFFFFF
    assert [int(v) for v in po.philox4x32_10(0, 0, 0, 0, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(v) for v in po.philox4x32_10(f, f, f, f, f, f)] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert [int(v) for v in po.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344,
                                             0xA4093822, 0x299F31D0)] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_gaussian_moments_and_independence_of_batching():
    z = po.draws(150528, np.arange(4), seed=42, stream_id=3)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    z2 = po.draws(150528, [2], seed=42, stream_id=3)
    assert np.array_equal(z[2], z2[0])  # a sample's draw depends only on its global index
    u = po.draws(1024, [0], seed=1, kind="uniform")
    assert u.min() >= 0.0 and u.max() < 1.0


@pytest.mark.parametrize("row", KAT["certify"])
def test_certify_tail_known_answers(row):
    counts_sel = np.array([1, 5, 2])
    counts_est = np.array([0, row["nA"], row["n"] - row["nA"]])
    label, radius = so.certify_tail(counts_sel, counts_est, row["n"], row["alpha"], row["sigma"])
    assert so.lower_confidence_bound(row["nA"], row["n"], row["alpha"]) == pytest.approx(row["pABar"], rel=1e-13, abs=0)
    assert (label == so.SmoothOracle.ABSTAIN) == row["abstain"]
    assert radius == pytest.approx(row["radius"], rel=1e-13, abs=0)
    if not row["abstain"]:
        assert label == 1


def test_min_count_that_certifies():
    for n, k in KAT["min_nA_certifying"].items():
        n = int(n)
        assert so.lower_confidence_bound(k, n, 0.001) >= 0.5
        assert so.lower_confidence_bound(k - 1, n, 0.001) < 0.5


@pytest.mark.parametrize("row", [r for r in KAT["predict"] if r["count1"] + r["count2"] > 0])
def test_predict_tail_known_answers(row):
    counts = np.zeros(6, dtype=int)
    counts[4] = row["count1"]
    counts[1] = row["count2"]
    assert so.binom_pvalue(row["count1"], row["count2"]) == pytest.approx(row["pvalue"], rel=1e-12)
    got = so.predict_tail(counts, row["alpha"])
    if row["abstain"]:
        assert got == so.SmoothOracle.ABSTAIN
    else:
        assert got == (4 if row["count1"] >= row["count2"] else 1)


def test_certify_argmax_tie_takes_lowest_index():
    # smoothing.py:46 ndarray.argmax -> first maximum
    label, _ = so.certify_tail(np.array([3, 7, 7, 1]), np.array([0, 1000, 0, 0]), 1000, 0.001, 0.5)
    assert label == 1


class _Toy(torch.nn.Module):
    def __init__(self, classes=5):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.w = torch.nn.Parameter(torch.randn(classes, 3 * 8 * 8, generator=g))

    def forward(self, x):
        return x.flatten(1) @ self.w.t()


def test_smooth_oracle_end_to_end_cpu():
    torch.manual_seed(0)
    model = _Toy()
    x = torch.rand(3, 8, 8)
    g = torch.Generator().manual_seed(1234)
    eps = torch.randn(140, 3, 8, 8, generator=g)
    cur = {"base": 0}

    def noise_fn(drawn, count, batch):
        return eps[cur["base"] + drawn: cur["base"] + drawn + count]

    s = so.SmoothOracle(model, 5, 0.25, noise_fn=noise_fn)
    c = s._sample_noise(x, 40, 16)
    assert c.sum() == 40 and len(s.last_margins) == 40
    cur["base"] = 0
    c2 = s._sample_noise(x, 40, 7)  # batch size must not matter under injected noise
    assert np.array_equal(c, c2)
    label = s.predict(x, 40, 0.001, 16)
    assert label in (-1, int(c.argmax()))


def test_smooth_oracle_edge_cases_of_the_reference_loop():
    """smoothing.py:91-99: num = 0 runs no batch; a batch size above num runs ONE short batch; one draw can never
    certify (alpha ** 1 < 0.5) nor pass the binomial test (p = 1)."""
    model = _Toy()
    x = torch.rand(3, 8, 8, generator=torch.Generator().manual_seed(2))
    seen = []
    s = so.SmoothOracle(model, 5, 0.25, noise_fn=lambda d, c, b: (seen.append(c), torch.zeros_like(b))[1])
    z = s._sample_noise(x, 0, 16)
    assert z.tolist() == [0] * 5 and seen == []
    c = s._sample_noise(x, 5, 1000)
    assert seen == [5] and c.sum() == 5 and c.max() == 5           # zero noise: every draw votes the same class
    seen.clear()
    s._sample_noise(x, 37, 16)
    assert seen == [16, 16, 5]
    assert s.certify(x, 1, 1, 0.001, 8) == (-1, 0.0)
    assert s.predict(x, 1, 0.001, 8) == -1
    assert so.lower_confidence_bound(1, 1, 0.001) == pytest.approx(0.001, rel=1e-12)   # alpha ** (1 / n)


from ref_smooth_util import REF_SMOOTH, case_id, ref_smooth_inputs  # noqa: E402


@pytest.mark.parametrize("case", REF_SMOOTH["cases"], ids=case_id)
def test_oracle_equals_the_reference_smooth_run(case):
    """The oracle restatement against outputs of the reference's OWN smoothing.py (executed unmodified in the build
    container on the same seeded classifier, image and draws): count vectors, certify and predict results."""
    model, x, eps = ref_smooth_inputs(case)
    cur = {"base": 0}
    s = so.SmoothOracle(model, case["classes"], case["sigma"],
                        noise_fn=lambda d, c, b: eps[cur["base"] + d: cur["base"] + d + c])
    n0, n, bs, alpha = case["n0"], case["n"], case["batch_size"], case["alpha"]
    sel = s._sample_noise(x, n0, bs)
    cur["base"] = n0
    est = s._sample_noise(x, n, bs)
    assert sel.tolist() == case["counts_selection"] and est.tolist() == case["counts_estimation"]
    label, radius = so.certify_tail(sel, est, n, alpha, case["sigma"])
    assert label == case["certify"][0] and radius == pytest.approx(case["certify"][1], rel=1e-12, abs=0)
    cur["base"] = 0
    assert s.predict(x, n, alpha, bs) == case["predict"]
    cur["base"] = 0
    assert s._sample_noise(x, n, bs).tolist() == case["predict_counts"]


def test_oracle_helpers_equal_the_reference_run():
    for row in REF_SMOOTH["lower_confidence_bound"]:
        assert so.lower_confidence_bound(row["NA"], row["N"], row["alpha"]) == pytest.approx(row["value"], rel=1e-13, abs=0)
    ca = REF_SMOOTH["count_arr"]
    s = so.SmoothOracle(None, ca["length"], 0.25)
    assert s._count_arr(np.array(ca["arr"]), ca["length"]).tolist() == ca["counts"]
    assert so.SmoothOracle.ABSTAIN == REF_SMOOTH["ABSTAIN"] == -1
