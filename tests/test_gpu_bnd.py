"""Boundary (VorB = BND) elements: CutInfo classification of the boundary facets (xfem/cutinfo.cpp:82-125 loops over VOL
and BND) against the oracle, and the scalar known answers of the reference's tests/pytests/test_bndcut.py:30-86
(Integrate(1 * dCut(lset, NEG/POS, vb=BND)))."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from oracle import xoracle as xo
from util import unpack_bits

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("dim,n", [(2, 9), (3, 5)])
def test_bnd_classification_matches_oracle(dim, n):
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV) if dim == 3 else xf.MakeStructured2DMesh(nx=n, ny=n, device=DEV)
    c = (.1, .2, .15)
    lset = xf.InterpolateToP1(lambda *x: torch.sqrt(sum((x[i] - c[i]) ** 2 for i in range(dim))) - 0.6, mesh)
    ci = xf.CutInfo(mesh, lset)
    bv, ids = mesh.boundary_elements()
    nb = bv.shape[0]
    assert nb == (4 * n if dim == 2 else 12 * n * n)
    bn, bp, bi, cr, err, _ = xo.cutinfo_update(dim - 1, bv.cpu().numpy(), lset.vec.cpu().numpy())
    assert err == 0
    for dt, ref in ((xf.NEG, bn), (xf.POS, bp), (xf.IF, bi)):
        got = ci.GetElementsOfType(dt, xf.BND)
        assert len(got) == nb
        assert np.array_equal(got.to_numpy(), unpack_bits(ref, nb))
    assert 0 < ci.GetElementsOfType(xf.IF, xf.BND).NumSet() < nb
    assert np.allclose(ci.GetCutRatios(xf.BND).cpu().numpy(), cr, rtol=1e-13, atol=1e-15)
    has = ci.GetElementsOfType(xf.HASNEG, xf.BND).to_numpy()
    assert np.array_equal(has, unpack_bits(bn, nb) | unpack_bits(bi, nb))


@pytest.mark.parametrize("cut", [0.3, 0.5, 0.7])
def test_integrate_bnd_2d(cut):
    # test_bndcut.py:30-47: lset = x - cut on the unit square; NEG boundary length 2 cut + 1
    mesh = xf.MakeStructured2DMesh(nx=7, ny=6, device=DEV)
    lset = xf.InterpolateToP1(lambda x, y: x - cut, mesh)
    neg = xf.Integrate({"levelset": lset, "domain_type": xf.NEG, "vb": xf.BND}, mesh=mesh, cf=1.0, order=0)
    pos = xf.Integrate({"levelset": lset, "domain_type": xf.POS, "vb": xf.BND}, mesh=mesh, cf=1.0, order=0)
    assert abs(neg - (2 * cut + 1.0)) < 1e-10
    assert abs(neg + pos - 4.0) < 1e-10          # test_bndcut.py:60-70: NEG + POS = the whole boundary
    # first moment of y over the NEG boundary: bottom 0, top cut, left 1/2
    my = xf.Integrate({"levelset": lset, "domain_type": xf.NEG, "vb": xf.BND}, mesh=mesh, cf=xf.y, order=1)
    assert abs(my - (cut + 0.5)) < 1e-10


def test_integrate_bnd_3d():
    # test_bndcut.py:73-86: unit cube, lset = x - 0.3: NEG boundary area 1 + 4 cut
    cut = 0.3
    mesh = xf.MakeStructured3DMesh(nx=4, device=DEV)
    lset = xf.InterpolateToP1(lambda x, y, z: x - cut, mesh)
    val = xf.Integrate({"levelset": lset, "domain_type": xf.NEG, "vb": xf.BND}, mesh=mesh, cf=1.0, order=0)
    assert abs(val - (1.0 + 4.0 * cut)) < 1e-10
    mx = xf.Integrate({"levelset": lset, "domain_type": xf.NEG, "vb": xf.BND}, mesh=mesh, cf=xf.x, order=1)
    assert abs(mx - 4.0 * cut * cut / 2) < 1e-10
