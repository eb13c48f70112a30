"""ORACLE (test infrastructure, never the product path).

CPU restatement of the reference's smoothed classifier, following
/root/reference/randomized_smoothing/smoothing.py line by line.  The reference file itself
cannot be imported in this image (`scipy.stats.binom_test` was removed in SciPy >= 1.12,
`statsmodels` is absent, and `:96` hard-codes device='cuda'), so the two library calls are
replaced by their exact SciPy equivalents:

  binom_test(k, n, p)                              -> scipy.stats.binomtest(k, n, p).pvalue
  proportion_confint(NA, N, 2*alpha, "beta")[0]    -> 0.0 if NA == 0 else beta.ppf(alpha, NA, N-NA+1)
     (statsmodels' method="beta" lower limit is exactly this Clopper-Pearson quantile)

Parity status: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md F3), so nothing of the reference's own test suite pins it.  It is pinned by
(a) OUTPUTS OF THE REFERENCE FILE ITSELF: tests/golden/make_ref_smooth_fixtures.py executes
smoothing.py unmodified in the build container (behind shims for the two library calls above and
for the hard-coded noise device) on seeded classifiers, images and draws, and
tests/golden/ref_smooth.json holds its count vectors, certify / predict results and
_lower_confidence_bound values - this module reproduces them exactly
(tests/test_oracle_cpu.py::test_oracle_equals_the_reference_smooth_run), and so does the CUDA path
(tests/test_smooth_gpu.py::test_cuda_smooth_equals_the_reference_smooth_run);
(b) the known-answer table of SURVEY.md 8(c) (tests/golden/smoothing_kat.json, generated with the
SciPy calls above).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from math import ceil

import numpy as np
import torch
from scipy.stats import beta as _beta
from scipy.stats import binomtest as _binomtest
from scipy.stats import norm as _norm


class SmoothOracle(object):
    """Restatement of `Smooth` (smoothing.py:13-117).  `noise_fn(batch) -> eps` lets a test
    inject the same standard-normal draws on both sides; default is torch.randn_like."""

    ABSTAIN = -1  # smoothing.py:17

    def __init__(self, base_classifier, num_classes: int, sigma: float, noise_fn=None):
        # smoothing.py:19-27
        self.base_classifier = base_classifier
        self.num_classes = num_classes
        self.sigma = sigma
        self.noise_fn = noise_fn
        self.last_margins = []  # per-sample top-2 logit margins of the last _sample_noise call

    def certify(self, x, n0: int, n: int, alpha: float, batch_size: int):
        # smoothing.py:29-56
        self.base_classifier.eval()
        counts_selection = self._sample_noise(x, n0, batch_size)
        cAHat = counts_selection.argmax().item()
        counts_estimation = self._sample_noise(x, n, batch_size)
        nA = counts_estimation[cAHat].item()
        pABar = self._lower_confidence_bound(nA, n, alpha)
        if pABar < 0.5:
            return SmoothOracle.ABSTAIN, 0.0
        else:
            radius = self.sigma * _norm.ppf(pABar)
            return cAHat, radius

    def predict(self, x, n: int, alpha: float, batch_size: int):
        # smoothing.py:58-79
        self.base_classifier.eval()
        counts = self._sample_noise(x, n, batch_size)
        return predict_tail(counts, alpha)

    def _sample_noise(self, x, num: int, batch_size):
        # smoothing.py:81-99 (noise device = x.device instead of the hard-coded 'cuda')
        with torch.no_grad():
            counts = np.zeros(self.num_classes, dtype=int)
            self.last_margins = []
            drawn = 0
            for _ in range(ceil(num / batch_size)):
                this_batch_size = min(batch_size, num)
                num -= this_batch_size
                batch = x.repeat((this_batch_size, 1, 1, 1))
                if self.noise_fn is not None:
                    eps = self.noise_fn(drawn, this_batch_size, batch)
                else:
                    eps = torch.randn_like(batch)
                drawn += this_batch_size
                noise = eps * self.sigma
                logits = self.base_classifier(batch + noise)
                predictions = logits.argmax(1)
                if logits.shape[1] >= 2:
                    top2 = logits.float().topk(2, dim=1).values
                    self.last_margins.extend((top2[:, 0] - top2[:, 1]).tolist())
                counts += self._count_arr(predictions.cpu().numpy(), self.num_classes)
            return counts

    def _count_arr(self, arr, length: int):
        # smoothing.py:101-105
        counts = np.zeros(length, dtype=int)
        for idx in arr:
            counts[idx] += 1
        return counts

    def _lower_confidence_bound(self, NA: int, N: int, alpha: float) -> float:
        # smoothing.py:107-117
        return lower_confidence_bound(NA, N, alpha)


def lower_confidence_bound(NA: int, N: int, alpha: float) -> float:
    """Clopper-Pearson one-sided lower bound (smoothing.py:117)."""
    if NA == 0:
        return 0.0
    return float(_beta.ppf(alpha, NA, N - NA + 1))


def certify_tail(counts_selection, counts_estimation, n: int, alpha: float, sigma: float):
    """smoothing.py:46-56 on given count vectors."""
    cAHat = int(np.asarray(counts_selection).argmax())
    nA = int(np.asarray(counts_estimation)[cAHat])
    pABar = lower_confidence_bound(nA, n, alpha)
    if pABar < 0.5:
        return SmoothOracle.ABSTAIN, 0.0
    return cAHat, float(sigma * _norm.ppf(pABar))


def binom_pvalue(count1: int, count2: int) -> float:
    return float(_binomtest(int(count1), int(count1 + count2), p=0.5).pvalue)


def predict_tail(counts, alpha: float):
    """smoothing.py:73-79 on a given count vector."""
    counts = np.asarray(counts)
    top2 = counts.argsort()[::-1][:2]
    count1 = counts[top2[0]]
    count2 = counts[top2[1]]
    if binom_pvalue(count1, count2) > alpha:
        return SmoothOracle.ABSTAIN
    else:
        return int(top2[0])
