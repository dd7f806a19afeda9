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


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (3, 2)])
def test_bnd_bilinear_and_linear_forms(dim, p):
    """tests/pytests/test_bndcut.py:91-150 (BFI / LFI over dCut(lset, NEG, vb=BND)) with basis-independent answers:
    for g = 1 + 2x - y (in every H1 space): c^T M c = int_{boundary, x < cut} g^2, f^T c = int g * (x y)"""
    cut = 0.4
    mesh = xf.MakeStructured3DMesh(nx=3, device=DEV) if dim == 3 else xf.MakeStructured2DMesh(nx=5, ny=4, device=DEV)
    lset = xf.InterpolateToP1((lambda x, y, z: x - cut) if dim == 3 else (lambda x, y: x - cut), mesh)
    V = xf.H1(mesh, order=p)
    u, v = V.TnT()
    a = xf.BilinearForm(V)
    a += u * v * xf.dCut(lset, xf.NEG, vb=xf.BND)
    a.Assemble()
    f = xf.LinearForm(V)
    f += (xf.x * xf.y) * v * xf.dCut(lset, xf.NEG, vb=xf.BND, order=2 * p + 2)
    f.Assemble()
    pts = mesh.points
    c = torch.zeros(V.ndof, dtype=torch.float64, device=DEV)
    c[:mesh.nv] = 1.0 + 2.0 * pts[:, 0] - pts[:, 1]                     # P1 function: higher-order dofs vanish
    quad = float((c * (a.mat * c)).sum())
    lin = float((c * f.vec).sum())
    X = cut
    if dim == 2:
        # edges: bottom y=0 (x in [0,X]), top y=1, left x=0 (y in [0,1]); g = 1 + 2x - y
        ex_quad = ((1 + 2 * X) ** 3 - 1) / 6 + (2 * X) ** 3 / 6 + ((1 - 0) ** 3 - (1 - 1) ** 3) / 3
        ex_lin = 0.0 + (X ** 2 / 2 * 0 + 2 * X ** 3 / 3) + 0.0            # int_0^X x * 1 * (2x) dx on the top edge
    else:
        import scipy.integrate as si
        g = lambda x, y: 1 + 2 * x - y
        faces = [(lambda s, t: g(s, 0.0), lambda s, t: s * 0.0, (0, X)), (lambda s, t: g(s, 1.0), lambda s, t: s * 1.0, (0, X)),
                 (lambda s, t: g(s, t), lambda s, t: s * t, (0, X)), (lambda s, t: g(s, t), lambda s, t: s * t, (0, X))]
        # y = 0, y = 1 faces (params s = x, t = z); z = 0, z = 1 faces (params s = x, t = y); x = 0 face (params y, z)
        ex_quad = ex_lin = 0.0
        for gg, xy, (lo, hi) in faces:
            ex_quad += si.dblquad(lambda t, s_: gg(s_, t) ** 2, lo, hi, 0, 1)[0]
            ex_lin += si.dblquad(lambda t, s_: gg(s_, t) * xy(s_, t), lo, hi, 0, 1)[0]
        ex_quad += si.dblquad(lambda t, s_: g(0.0, s_) ** 2, 0, 1, 0, 1)[0]
    assert abs(quad - ex_quad) < 1e-10 * abs(ex_quad)
    assert abs(lin - ex_lin) < 1e-10 * max(1.0, abs(ex_lin))


# ---- dCut(lset, IF, vb=BND): the codim-2 point measure on the boundary segments of a 2D mesh
# (cutint/straightcutrule.cpp:646-656; known answers of tests/pytests/test_bndcut.py:166-217)
@pytest.mark.parametrize("cut", [0.3, 0.5, 0.7])
def test_integrate_if_bnd_2d_counts_the_intersections(cut):
    mesh = xf.MakeStructured2DMesh(nx=7, ny=6, device=DEV)
    lset = xf.InterpolateToP1(lambda x, y: x - cut, mesh)
    ld = {"levelset": lset, "domain_type": xf.IF, "vb": xf.BND}
    assert abs(xf.Integrate(ld, mesh=mesh, cf=1.0, order=0) - 2.0) < 1e-10          # (cut, 0) and (cut, 1)
    assert abs(xf.Integrate(ld, mesh=mesh, cf=xf.x, order=1) - 2.0 * cut) < 1e-10
    assert abs(xf.Integrate(ld, mesh=mesh, cf=xf.y, order=1) - 1.0) < 1e-10
    # definedonelements over the boundary elements (the reference's definedon=mesh.Boundaries("bottom"), :203-217)
    bv, _ = mesh.boundary_elements()
    ymid = mesh.points[bv.long()][:, :, 1].mean(dim=1)
    xmid = mesh.points[bv.long()][:, :, 0].mean(dim=1)
    for mask, want in ((ymid < 1e-12, 1.0), (xmid < 1e-12, 0.0)):
        ba = xf.BitArray.from_numpy(mask.cpu().numpy(), DEV)
        got = xf.Integrate({**ld, "definedonelements": ba}, mesh=mesh, cf=1.0, order=0)
        assert abs(got - want) < 1e-10


def test_if_bnd_2d_against_the_cut_points_in_closed_form():
    """circle through the boundary of the square: every cut boundary segment contributes f at the zero of the P1 level set"""
    mesh = xf.MakeStructured2DMesh(nx=9, ny=8, device=DEV)
    lset = xf.InterpolateToP1(lambda x, y: torch.sqrt((x - .1) ** 2 + (y - .2) ** 2) - 0.62, mesh)
    bv, _ = mesh.boundary_elements()
    P, phi, b = mesh.points.cpu().numpy(), lset.vec.cpu().numpy(), bv.cpu().numpy()
    want1 = wantf = 0.0
    for s in b:
        a, c = phi[s[0]], phi[s[1]]
        if (a < 0) != (c < 0):
            t = a / (a - c)
            xs = P[s[0]] + t * (P[s[1]] - P[s[0]])
            want1 += 1.0
            wantf += xs[0] ** 2 * xs[1]
    assert want1 >= 2
    ld = {"levelset": lset, "domain_type": xf.IF, "vb": xf.BND}
    assert abs(xf.Integrate(ld, mesh=mesh, cf=1.0, order=0) - want1) < 1e-12
    assert abs(xf.Integrate(ld, mesh=mesh, cf=xf.x ** 2 * xf.y, order=3) - wantf) < 1e-12
    # linear / bilinear forms with a P1 function g = 1 + 2x - y (higher-order dofs vanish): f^T c = sum g(x*),
    # c^T M c = sum g(x*)^2; the vertex functions sum to 1 at every cut point
    V = xf.H1(mesh, order=2)
    u, v = V.TnT()
    f = xf.LinearForm(V)
    f += 1.0 * v * xf.dCut(lset, xf.IF, vb=xf.BND)
    f.Assemble()
    assert abs(float(f.vec[:mesh.nv].sum()) - want1) < 1e-12
    a = xf.BilinearForm(V)
    a += u * v * xf.dCut(lset, xf.IF, vb=xf.BND)
    a.Assemble()
    c = torch.zeros(V.ndof, dtype=torch.float64, device=DEV)
    c[:mesh.nv] = 1.0 + 2.0 * mesh.points[:, 0] - mesh.points[:, 1]
    wantg = wantg2 = 0.0
    for s in b:
        p0, p1 = phi[s[0]], phi[s[1]]
        if (p0 < 0) != (p1 < 0):
            xs = P[s[0]] + p0 / (p0 - p1) * (P[s[1]] - P[s[0]])
            wantg += 1.0 + 2.0 * xs[0] - xs[1]
            wantg2 += (1.0 + 2.0 * xs[0] - xs[1]) ** 2
    assert abs(float((f.vec * c).sum()) - wantg) < 1e-11
    assert abs(float((c * (a.mat * c)).sum()) - wantg2) < 1e-11


def test_if_bnd_3d_is_declined():
    mesh = xf.MakeStructured3DMesh(nx=3, device=DEV)
    lset = xf.InterpolateToP1(lambda x, y, z: x - 0.4, mesh)
    with pytest.raises(xf.XgpuError):
        xf.Integrate({"levelset": lset, "domain_type": xf.IF, "vb": xf.BND}, mesh=mesh, cf=1.0, order=0)
