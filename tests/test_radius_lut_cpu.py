"""The per-(n, alpha) pABar / Phi^-1 table behind the bit-exact radius (certifiedgpt_b200/_lib.py::radius_lut_host):
built with vectorised SciPy calls, it must equal the reference's scalar per-image calls (smoothing.py:55,117, restated in
oracle/smoothing_oracle.py) entry by entry, and the reference-run fixtures."""
import json
import os

import numpy as np
import pytest
from scipy.stats import norm

from certifiedgpt_b200._lib import radius_lut_host
from oracle import smoothing_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("n,alpha", [(1, 0.001), (32, 0.001), (100, 0.001), (1000, 0.001), (1000, 0.01), (2000, 0.05)])
def test_table_equals_scalar_reference_calls(n, alpha):
    lut = radius_lut_host(n, alpha)
    assert lut.dtype == np.float64 and lut.shape == (2 * (n + 1),)
    for na in range(n + 1):
        p = so.lower_confidence_bound(na, n, alpha)
        assert lut[na] == p, (na, lut[na], p)                        # bit-exact, not approx
        if p >= 0.5:
            assert lut[n + 1 + na] == float(norm.ppf(p)), na
        else:
            assert lut[n + 1 + na] == 0.0                            # the reference abstains there (smoothing.py:53-54)
    assert lut[0] == 0.0
    assert np.all(np.diff(lut[: n + 1]) > 0)                         # pABar strictly increasing in nA


def test_table_reproduces_certify_tail_and_reference_run():
    for sigma in (0.25, 0.5, 1.0):
        lut = radius_lut_host(1000, 0.001)
        for na in (0, 1, 540, 549, 550, 600, 900, 990, 1000):
            counts = np.array([na, 1000 - na], dtype=np.int64)
            label, radius = so.certify_tail(np.array([1, 0]), counts, 1000, 0.001, sigma)
            got = (0, sigma * lut[1001 + na]) if lut[na] >= 0.5 else (-1, 0.0)
            assert got == (label, radius), (sigma, na)
    ref = json.load(open(os.path.join(HERE, "golden", "ref_smooth.json")))
    for row in ref["lower_confidence_bound"]:
        assert radius_lut_host(row["N"], row["alpha"])[row["NA"]] == row["value"], row
    for case in ref["cases"]:
        lut = radius_lut_host(case["n"], case["alpha"])
        sel, est = case["counts_selection"], case["counts_estimation"]
        ca = int(np.argmax(sel))
        na = est[ca]
        want = tuple(case["certify"])
        got = (ca, case["sigma"] * lut[case["n"] + 1 + na]) if lut[na] >= 0.5 else (-1, 0.0)
        assert got == want, case
