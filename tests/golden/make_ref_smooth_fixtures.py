"""Generates tests/golden/ref_smooth.json by running the REFERENCE's own randomized_smoothing/smoothing.py.

Runs only in the build container (needs /root/reference, read-only).  The file is executed unmodified, by path,
behind three shims for what this image lacks:

* `scipy.stats.binom_test` (removed from SciPy 1.12)  -> `scipy.stats.binomtest(k, n, p).pvalue`, the same exact
  two-sided binomial test;
* `statsmodels.stats.proportion.proportion_confint(count, nobs, alpha, method="beta")` (statsmodels is not installed
  and the reference pins no version: docker/tpu-docker has no statsmodels line) -> its published Clopper-Pearson
  definition: (beta.ppf(alpha/2, count, nobs-count+1), beta.isf(alpha/2, count+1, nobs-count)), lower bound 0 when
  count == 0, upper bound 1 when count == nobs;
* `torch.randn_like(batch, device='cuda')` (smoothing.py:96 hard-codes the device) -> serves the next rows of a seeded
  standard-normal tensor on the CPU, so the draws are identical on every side of the comparison.

Everything else - the selection / estimation order, the batch loop, argmax, _count_arr, the abstain rule, norm.ppf -
is the reference's code.  Both the oracle (tests/test_oracle_cpu.py) and the CUDA path (tests/test_smooth_gpu.py)
are then checked against the counts, labels and radii it produced.

    python tests/golden/make_ref_smooth_fixtures.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import scipy.stats
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/randomized_smoothing/smoothing.py"


def toy_weight(classes, shape, seed):
    """The toy classifier of tests/test_smooth_gpu.py::Toy (fp64 linear map of the flattened image)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(classes, int(np.prod(shape)), generator=g) / 10


class Toy(torch.nn.Module):
    def __init__(self, classes, shape, seed):
        super().__init__()
        self.w = torch.nn.Parameter(toy_weight(classes, shape, seed))

    def forward(self, x):
        return x.flatten(1).double() @ self.w.t().double()


def load_reference():
    if not hasattr(scipy.stats, "binom_test"):
        scipy.stats.binom_test = lambda k, n=None, p=0.5: float(scipy.stats.binomtest(int(k), int(n), p).pvalue)
    sm = types.ModuleType("statsmodels")
    sms = types.ModuleType("statsmodels.stats")
    smp = types.ModuleType("statsmodels.stats.proportion")

    def proportion_confint(count, nobs, alpha=0.05, method="normal"):
        assert method == "beta"
        lo = 0.0 if count == 0 else float(scipy.stats.beta.ppf(alpha / 2, count, nobs - count + 1))
        hi = 1.0 if count == nobs else float(scipy.stats.beta.isf(alpha / 2, count + 1, nobs - count))
        return lo, hi
    smp.proportion_confint = proportion_confint
    sys.modules.update({"statsmodels": sm, "statsmodels.stats": sms, "statsmodels.stats.proportion": smp})
    spec = importlib.util.spec_from_file_location("ref_smoothing", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Draws:
    """Stands in for torch.randn_like: consecutive rows of one seeded tensor, whatever the device argument."""

    def __init__(self, eps):
        self.eps, self.pos = eps, 0

    def __call__(self, batch, **kwargs):
        n = batch.shape[0]
        out = self.eps[self.pos:self.pos + n]
        assert out.shape == batch.shape, "injected draws exhausted"
        self.pos += n
        return out


CASES = [  # classes, shape, sigma, model seed, n0, n, batch_size, alpha
    (6, (3, 16, 16), 0.25, 0, 100, 1000, 128, 0.001),
    (6, (3, 16, 16), 0.5, 1, 100, 1000, 128, 0.001),
    (6, (3, 16, 16), 1.0, 2, 100, 1000, 128, 0.001),
    (6, (3, 16, 16), 0.5, 3, 10, 37, 16, 0.001),          # ragged last batch
    (6, (3, 16, 16), 2.0, 4, 100, 1000, 1000, 0.001),     # heavy noise: low top-class share
    (3, (3, 8, 8), 0.5, 5, 1, 1, 1000, 0.001),            # single draws: must abstain
    (10, (3, 8, 8), 4.0, 6, 50, 400, 64, 0.05),           # looser alpha, many classes
    (6, (3, 16, 16), 0.12, 7, 100, 1000, 100, 0.001),     # light noise: every draw votes one class
    (4, (3, 16, 16), 0.35, 8, 64, 512, 50, 0.01),
]


def main():
    ref = load_reference()
    real_randn_like = torch.randn_like
    out = {"reference": "randomized_smoothing/smoothing.py (executed unmodified; shims: see this script's docstring)",
           "eps_seed": 1234, "cases": [], "lower_confidence_bound": []}
    try:
        for classes, shape, sigma, seed, n0, n, bs, alpha in CASES:
            model = Toy(classes, shape, seed)
            x = torch.rand(*shape, generator=torch.Generator().manual_seed(seed + 1))
            eps = torch.randn(n0 + n, *shape, generator=torch.Generator().manual_seed(1234))
            smooth = ref.Smooth(model, classes, sigma)
            # certify: selection consumes draws [0, n0), estimation [n0, n0 + n)  (smoothing.py:44,48)
            draws = Draws(eps)
            torch.randn_like = draws
            label, radius = smooth.certify(x, n0, n, alpha, bs)
            assert draws.pos == n0 + n
            # the two count vectors, through the reference's own _sample_noise on the same draws
            draws = Draws(eps)
            torch.randn_like = draws
            sel = smooth._sample_noise(x, n0, bs)
            est = smooth._sample_noise(x, n, bs)
            # predict on draws [0, n)
            draws = Draws(eps)
            torch.randn_like = draws
            pred = smooth.predict(x, n, alpha, bs)
            pred_counts = None
            draws = Draws(eps)
            torch.randn_like = draws
            pred_counts = smooth._sample_noise(x, n, bs)
            out["cases"].append({
                "classes": classes, "shape": list(shape), "sigma": sigma, "model_seed": seed, "x_seed": seed + 1,
                "n0": n0, "n": n, "batch_size": bs, "alpha": alpha,
                "eps_checksum": float(eps.double().sum()), "w_checksum": float(model.w.detach().double().sum()),
                "counts_selection": [int(v) for v in sel], "counts_estimation": [int(v) for v in est],
                "certify": [int(label), float(radius)],
                "predict": int(pred), "predict_counts": [int(v) for v in pred_counts]})
        smooth = ref.Smooth(Toy(2, (1, 1, 1), 0), 2, 0.25)
        for NA, N, alpha in [(0, 100, 0.001), (1, 1, 0.001), (66, 100, 0.001), (550, 1000, 0.001), (990, 1000, 0.001),
                             (1000, 1000, 0.001), (5156, 10000, 0.001), (37, 50, 0.05), (400, 400, 0.05)]:
            out["lower_confidence_bound"].append({"NA": NA, "N": N, "alpha": alpha,
                                                  "value": float(smooth._lower_confidence_bound(NA, N, alpha))})
        # _count_arr on a fixed vector
        arr = np.array([0, 4, 4, 2, 0, 4])
        out["count_arr"] = {"arr": arr.tolist(), "length": 6, "counts": [int(v) for v in smooth._count_arr(arr, 6)]}
        out["ABSTAIN"] = int(ref.Smooth.ABSTAIN)
    finally:
        torch.randn_like = real_randn_like
    with open(os.path.join(HERE, "ref_smooth.json"), "w") as f:
        json.dump(out, f, indent=1)
    for c in out["cases"]:
        print(c["sigma"], c["n0"], c["n"], "certify", c["certify"], "predict", c["predict"], c["counts_estimation"])


if __name__ == "__main__":
    main()
