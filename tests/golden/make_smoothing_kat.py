"""Generates tests/golden/smoothing_kat.json: known-answer values for the statistics tail of
Smooth.certify / Smooth.predict, computed with the SciPy calls that are the
exact equivalents of the reference's statsmodels/scipy calls (smoothing.py:51-56,76,117).
Run:  python tests/golden/make_smoothing_kat.py
"""
import json
import os

from scipy.stats import beta, binomtest, norm

ALPHA = 0.001
certify = []
for nA, n, sigma in [(1000, 1000, 0.25), (1000, 1000, 0.5), (990, 1000, 0.25), (900, 1000, 0.25),
                     (600, 1000, 0.5), (550, 1000, 0.5), (549, 1000, 0.5), (540, 1000, 0.5), (0, 1000, 0.25),
                     (100, 100, 0.25), (66, 100, 0.25), (65, 100, 0.25), (5156, 10000, 1.0),
                     (5155, 10000, 1.0), (9999, 10000, 0.12), (31, 32, 0.25), (1, 1000, 0.25)]:
    p = 0.0 if nA == 0 else float(beta.ppf(ALPHA, nA, n - nA + 1))
    certify.append({"nA": nA, "n": n, "sigma": sigma, "alpha": ALPHA, "pABar": p,
                    "abstain": bool(p < 0.5), "radius": 0.0 if p < 0.5 else float(sigma * norm.ppf(p))})
predict = []
for c1, c2 in [(32, 0), (27, 5), (24, 8), (16, 16), (11, 0), (10, 0), (70, 30), (60, 40), (0, 0),
               (1, 0), (600, 400), (540, 460), (5, 27)]:
    pv = 1.0 if c1 + c2 == 0 else float(binomtest(c1, c1 + c2, p=0.5).pvalue)
    predict.append({"count1": c1, "count2": c2, "alpha": ALPHA, "pvalue": pv, "abstain": bool(pv > ALPHA)})
out = {"alpha": ALPHA, "certify": certify, "predict": predict,
       "min_nA_certifying": {"100": 66, "1000": 550, "10000": 5156}}
with open(os.path.join(os.path.dirname(__file__), "smoothing_kat.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(certify), "certify and", len(predict), "predict vectors")
