"""GPU parity: drop-in Smooth (generic nn.Module path) vs the oracle on identical injected noise."""
import numpy as np
import pytest
import torch

from oracle import smoothing_oracle as so

pytestmark = pytest.mark.gpu


class Toy(torch.nn.Module):
    """fp32 linear classifier: both sides compute the same logits up to summation order."""

    def __init__(self, classes, shape, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w = torch.nn.Parameter(torch.randn(classes, int(np.prod(shape)), generator=g) / 10)

    def forward(self, x):
        return x.flatten(1).double() @ self.w.t().double()


def _pair(classes=6, shape=(3, 16, 16), sigma=0.5, n_total=1200, seed=0):
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    model = Toy(classes, shape, seed)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(seed + 1))
    eps = torch.randn(n_total, *shape, generator=torch.Generator().manual_seed(1234))
    cur = {"base": 0}

    def noise_fn(drawn, count, batch):
        return eps[cur["base"] + drawn: cur["base"] + drawn + count]

    import copy
    oracle = so.SmoothOracle(model, classes, sigma, noise_fn=noise_fn)
    ours = Smooth(copy.deepcopy(model).cuda(), classes, sigma)
    ours.inject_noise(eps.cuda())
    return oracle, ours, x, cur, model


def test_sample_noise_counts_match_oracle():
    oracle, ours, x, cur, model = _pair()
    ref = oracle._sample_noise(x, 500, 64)
    ours._cursor = 0
    got = ours._sample_noise(x.cuda(), 500, 64)
    safe = np.array(oracle.last_margins) > 1e-6
    assert safe.all()
    assert got.dtype.kind == "i" and np.array_equal(got, ref)
    ours._cursor = 0
    assert np.array_equal(ours._sample_noise(x.cuda(), 500, 500), ref)  # batch size irrelevant


@pytest.mark.parametrize("sigma,seed", [(0.25, 0), (0.5, 1), (1.0, 2)])
def test_certify_matches_oracle(sigma, seed):
    oracle, ours, x, cur, model = _pair(sigma=sigma, seed=seed)
    n0, n = 100, 1000
    cur["base"] = 0
    sel = oracle._sample_noise(x, n0, 128)
    cur["base"] = n0
    est = oracle._sample_noise(x, n, 128)
    ref_label, ref_radius = so.certify_tail(sel, est, n, 0.001, sigma)
    label, radius = ours.certify(x.cuda(), n0, n, 0.001, 128)
    assert np.array_equal(ours.last_counts_selection.cpu().numpy(), sel)
    assert np.array_equal(ours.last_counts_estimation.cpu().numpy(), est)
    assert label == ref_label
    assert radius == ref_radius          # bit-exact: per-(n, alpha) SciPy table, IEEE fp64 product on the device
    assert ours.last_pABar == so.lower_confidence_bound(int(est[ref_label if ref_label >= 0 else ours.last_cAHat]), n, 0.001)
    assert isinstance(label, int) and isinstance(radius, float)
    # the table-free device tail (fp64 bisection + AS241) stays as the cross-check
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    dev = Smooth(ours.base_classifier, ours.num_classes, sigma, exact_tail=False)
    dev.inject_noise(ours._injected)
    l2, r2 = dev.certify(x.cuda(), n0, n, 0.001, 128)
    assert l2 == label and r2 == pytest.approx(radius, rel=1e-9)


def test_predict_matches_oracle():
    oracle, ours, x, cur, model = _pair(sigma=0.5, seed=3)
    cur["base"] = 0
    ref = oracle.predict(x, 100, 0.001, 32)
    got = ours.predict(x.cuda(), 100, 0.001, 32)
    assert got == ref


def test_philox_path_runs_and_is_reproducible():
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    model = Toy(4, (3, 16, 16)).cuda()
    x = torch.rand(3, 16, 16).cuda()
    a = Smooth(model, 4, 0.5, seed=7)
    b = Smooth(model, 4, 0.5, seed=7)
    ca = a._sample_noise(x, 300, 64)
    cb = b._sample_noise(x, 300, 300)
    assert ca.sum() == 300 and np.array_equal(ca, cb)
    c = Smooth(model, 4, 0.5, seed=8)._sample_noise(x, 300, 64)
    assert c.sum() == 300


def test_helpers_match_reference_semantics():
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    s = Smooth(torch.nn.Identity(), 5, 0.25)
    assert s._count_arr(np.array([0, 4, 4, 2]), 5).tolist() == [1, 0, 1, 0, 2]
    assert s._lower_confidence_bound(990, 1000, 0.001) == 0.9760361871553114      # SURVEY 8c known answer, exact
    assert Smooth(torch.nn.Identity(), 5, 0.25, exact_tail=False)._lower_confidence_bound(990, 1000, 0.001) == \
        pytest.approx(0.9760361871553114, rel=1e-9)
    assert s._lower_confidence_bound(0, 1000, 0.001) == 0.0
    assert Smooth.ABSTAIN == -1


def test_edge_cases_follow_the_reference_loop():
    """Empty and ragged draws as smoothing.py:91-99 handles them: num = 0 runs no batch (all-zero counts), a batch
    size larger than num runs one short batch, batch size 1 walks sample by sample; a one-draw certify abstains."""
    oracle, ours, x, cur, model = _pair(n_total=64)
    xc = x.cuda()
    ours._cursor = 0
    zero = ours._sample_noise(xc, 0, 32)
    assert zero.shape == (6,) and zero.sum() == 0
    assert np.array_equal(oracle._sample_noise(x, 0, 32), zero)
    cur["base"] = 0
    ref = oracle._sample_noise(x, 37, 1000)
    for bs in (1000, 37, 36, 5, 1):
        ours._cursor = 0
        assert np.array_equal(ours._sample_noise(xc, 37, bs), ref), bs
    # n0 = n = 1: pABar = alpha ** 1 < 0.5 -> ABSTAIN, radius exactly 0.0 (smoothing.py:53-54)
    assert ours.certify(xc, 1, 1, 0.001, 1000) == (-1, 0.0)
    cur["base"] = 0
    assert oracle.certify(x, 1, 1, 0.001, 1000) == (-1, 0.0)
    # predict on one draw: binomial test of 1 vs 0 has p = 1 > alpha -> ABSTAIN (smoothing.py:76-77)
    assert ours.predict(xc, 1, 0.001, 8) == -1


def test_rejects_host_tensor_on_the_generic_path():
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    s = Smooth(Toy(4, (3, 16, 16)).cuda(), 4, 0.5)
    with pytest.raises(L.CgptError):
        s._sample_noise(torch.rand(3, 16, 16), 8, 8)      # no CPU path exists


from ref_smooth_util import REF_SMOOTH, case_id, ref_smooth_inputs  # noqa: E402


@pytest.mark.parametrize("case", REF_SMOOTH["cases"], ids=case_id)
def test_cuda_smooth_equals_the_reference_smooth_run(case):
    """The CUDA path against outputs of the reference's OWN smoothing.py (tests/golden/make_ref_smooth_fixtures.py ran it
    unmodified on the same seeded classifier, image and draws): count vectors, (label, radius) and predict, exactly."""
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    model, x, eps = ref_smooth_inputs(case)
    n0, n, bs, alpha = case["n0"], case["n"], case["batch_size"], case["alpha"]
    ours = Smooth(model.cuda(), case["classes"], case["sigma"])
    ours.inject_noise(eps.cuda())
    label, radius = ours.certify(x.cuda(), n0, n, alpha, bs)
    assert ours.last_counts_selection.cpu().tolist() == case["counts_selection"]
    assert ours.last_counts_estimation.cpu().tolist() == case["counts_estimation"]
    assert label == case["certify"][0] and radius == case["certify"][1]       # bit-exact radius
    assert ours.last_pABar == so.lower_confidence_bound(case["counts_estimation"][ours.last_cAHat], n, alpha)
    # the reference's two separate _sample_noise calls (no fused selection pass) give the same result
    seq = Smooth(model.cuda(), case["classes"], case["sigma"], fuse_selection=False)
    seq.inject_noise(eps.cuda())
    assert seq.certify(x.cuda(), n0, n, alpha, bs) == (label, radius)
    assert ours.predict(x.cuda(), n, alpha, bs) == case["predict"]
    assert ours.last_counts.cpu().tolist() == case["predict_counts"]


def test_lower_confidence_bound_equals_the_reference_run_exactly():
    """Smooth._lower_confidence_bound against the values the reference's own smoothing.py produced: equal bit for bit."""
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    s = Smooth(torch.nn.Identity(), 5, 0.25)
    for row in REF_SMOOTH["lower_confidence_bound"]:
        assert s._lower_confidence_bound(row["NA"], row["N"], row["alpha"]) == row["value"], row
