"""Inputs of the tests/golden/ref_smooth.json cases (produced by the reference's own smoothing.py, see
tests/golden/make_ref_smooth_fixtures.py), regenerated from their seeds and checked against the stored checksums."""
import json
import os

import numpy as np
import pytest
import torch

REF_SMOOTH = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_smooth.json")))


class RefToy(torch.nn.Module):
    """The fp64 linear classifier the fixture script ran under the reference's Smooth."""

    def __init__(self, classes, shape, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w = torch.nn.Parameter(torch.randn(classes, int(np.prod(shape)), generator=g) / 10)

    def forward(self, x):
        return x.flatten(1).double() @ self.w.t().double()


def case_id(c):
    return f"sigma{c['sigma']}-n{c['n']}-bs{c['batch_size']}"


def ref_smooth_inputs(case):
    """(model, x, eps) of one fixture case."""
    shape = tuple(case["shape"])
    model = RefToy(case["classes"], shape, case["model_seed"])
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(case["x_seed"]))
    eps = torch.randn(case["n0"] + case["n"], *shape, generator=torch.Generator().manual_seed(REF_SMOOTH["eps_seed"]))
    assert float(eps.double().sum()) == pytest.approx(case["eps_checksum"], rel=1e-12, abs=1e-9)
    assert float(model.w.detach().double().sum()) == pytest.approx(case["w_checksum"], rel=1e-12, abs=1e-9)
    return model, x, eps
