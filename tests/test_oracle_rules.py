"""Oracle quadrature rules: exactness on all monomials up to the stated degree, positivity,
agreement with the mpmath-generated fixture (tests/golden/rules_sym.json)."""
import itertools
import json
import os
from math import factorial

import numpy as np
import pytest

from oracle import xoracle as xo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def moment(e):
    d = len(e)
    num = 1
    for k in e:
        num *= factorial(k)
    return num / factorial(sum(e) + d)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("order", list(range(0, 17)))
def test_rule_exact(dim, order):
    pts, wts = xo.select_rule(dim, order)
    assert len(wts) == xo.lib().xo_rule_size(dim, order)
    assert np.all(wts > 0)
    lam_last = 1 - pts[:, :dim].sum(axis=1)
    assert np.all(pts[:, :dim] >= -1e-15) and np.all(lam_last >= -1e-15)
    for e in itertools.product(range(order + 1), repeat=dim):
        if sum(e) > order:
            continue
        val = np.sum(wts * np.prod(pts[:, :dim] ** np.array(e), axis=1))
        assert abs(val - moment(e)) < 2e-16 * (1 + order) + 1e-17, (e, val, moment(e))


def test_sym_tables_match_fixture():
    with open(os.path.join(GOLD, "rules_sym.json")) as f:
        js = json.load(f)
    for name, dim, order in [("trig1", 2, 1), ("trig3", 2, 2), ("trig6", 2, 4), ("trig12", 2, 6),
                             ("tet1", 3, 0), ("tet4", 3, 2), ("tet14", 3, 5)]:
        pts, wts = xo.select_rule(dim, order)
        assert np.array_equal(pts[:, :dim], np.array(js[name]["points"]))
        assert np.array_equal(wts, np.array(js[name]["weights"]))


@pytest.mark.parametrize("alpha", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 12])
def test_gauss_jacobi(n, alpha):
    x, w = xo.gauss_jacobi01(n, alpha)
    assert np.all(np.diff(x) > 0) and np.all(w > 0)
    for k in range(2 * n):
        # int_0^1 (1-x)^alpha x^k dx = B(k+1, alpha+1)
        ex = factorial(k) * factorial(alpha) / factorial(k + alpha + 1)
        assert abs(np.sum(w * x ** k) - ex) < 1e-15
