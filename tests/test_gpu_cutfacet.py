"""SURVEY.md §8 row f4 on the GPU: cut integrals over interior facets (dCut(..., skeleton=True)) and element boundaries
(dCut(..., element_boundary=True)) — xgpu_facet_elmat through the C ABI against the oracle restatement of
xfem/symboliccutbfi.cpp:886-1290 / :488-883 facet by facet, and assembled operators against properties that do not depend on
the basis (jumps of globally smooth functions vanish; cut skeleton measure)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from oracle import xoracle as xo
from util import interpolate_to_p1, mesh_np

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NEG, POS = 0, 1

DG = lambda lam: [(0, 0, 0, 0, lam, -1), (0, 1, 0, 0, -lam, -1), (0, 0, 0, 1, -lam, -1), (0, 1, 0, 1, lam, -1),
                  (1, 0, 0, 0, -0.5, 0), (1, 0, 0, 1, 0.5, 0), (1, 1, 0, 0, -0.5, 0), (1, 1, 0, 1, 0.5, 0),
                  (0, 0, 1, 0, -0.5, 0), (0, 0, 1, 1, -0.5, 0), (0, 1, 1, 0, 0.5, 0), (0, 1, 1, 1, 0.5, 0)]


def make(dim, n):
    return xf.MakeStructured2DMesh(nx=n, ny=n, device=DEV) if dim == 2 else xf.MakeStructured3DMesh(nx=n, device=DEV)


@pytest.mark.parametrize("dim,n,p", [(2, 6, 1), (2, 5, 2), (2, 4, 3), (2, 3, 4), (3, 3, 1), (3, 3, 2), (3, 2, 3)])
@pytest.mark.parametrize("dt", [NEG, POS])
def test_skeleton_facet_matrices_match_oracle(dim, n, p, dt):
    m = make(dim, n)
    pts, els = mesh_np(m)
    ents, el_f, fe = (t.cpu().numpy() for t in m.facets())
    phi = interpolate_to_p1(lambda *x: np.sqrt(sum((xi - 0.47) ** 2 for xi in x)) - 0.31, pts)
    inner = np.flatnonzero(fe[:, 1] >= 0)
    fv = torch.from_numpy(ents[inner].astype(np.int32)).to(DEV)
    fel = torch.from_numpy(fe[inner].astype(np.int32)).to(DEV)
    g = torch.from_numpy(phi).to(DEV)
    terms = DG(10.0 * p * p)
    off, q, w = ops.cutquad(dim - 1, None, fv, m.points, g, dt, 2 * p)
    assert ops.default_ctx().status()[0] == 0
    mats = ops.facet_elmat(dim, p, fv, fel, m.elements, m.points, (off, q, w), terms, 2).cpu().numpy()
    ncut = 0
    for k, f in enumerate(inner):
        e1, e2 = fe[f]
        T1, T2 = xo.make_trafo(pts[els[e1]], els[e1]), xo.make_trafo(pts[els[e2]], els[e2])
        ref, npt = xo.cut_facet_elmat(T1, T2, p, ents[f], phi[ents[f]], dt, 2 * p, terms)
        assert max(npt, 0) == int(off[k + 1] - off[k])
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(mats[k] - ref).max() <= 1e-12 * scale + 1e-300
        ncut += phi[ents[f]].min() < 0 < phi[ents[f]].max()
    assert ncut > 0


@pytest.mark.parametrize("dim,n,p", [(2, 5, 2), (3, 3, 2), (2, 4, 3)])
def test_element_boundary_matrices_match_oracle(dim, n, p):
    m = make(dim, n)
    pts, els = mesh_np(m)
    ents, el_f, fe = (t.cpu().numpy() for t in m.facets())
    phi = interpolate_to_p1(lambda *x: x[0] + 0.25 * x[1] - 0.52, pts)
    terms = [(0, 0, 0, 0, 2.0, -1), (1, 0, 0, 0, -1.0, 0), (0, 0, 1, 0, -1.0, 0), (1, 0, 1, 0, 0.3, 1)]
    nl = dim + 1
    fids = el_f.reshape(-1)
    fv = torch.from_numpy(ents[fids].astype(np.int32)).to(DEV)
    fel = torch.stack([torch.arange(len(els), dtype=torch.int32).repeat_interleave(nl), torch.full((len(els) * nl,), -1, dtype=torch.int32)], 1).to(DEV)
    g = torch.from_numpy(phi).to(DEV)
    off, q, w = ops.cutquad(dim - 1, None, fv, m.points, g, NEG, 2 * p)
    mats = ops.facet_elmat(dim, p, fv, fel, m.elements, m.points, (off, q, w), terms, 1).cpu().numpy()
    nd = ops.h1_ndof(dim, p)
    for e in range(len(els)):
        T = xo.make_trafo(pts[els[e]], els[e])
        for k in range(nl):
            f = el_f[e, k]
            ref, _ = xo.cut_facet_elmat(T, None, p, ents[f], phi[ents[f]], NEG, 2 * p, terms)
            assert np.abs(mats[e * nl + k] - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-300) + 1e-300


@pytest.mark.parametrize("dim,n,p", [(2, 8, 2), (3, 4, 2), (2, 6, 1)])
def test_assembled_derivative_jump_penalty_kills_global_polynomials(dim, n, p):
    """gamma h (dn(u) - dn(u.Other())) (dn(v) - dn(v.Other())) dCut(lset, NEG, skeleton=True) assembled on Restrict(H1): the
    operator vanishes on the interpolant of a global polynomial of degree <= p (its gradient is continuous), is symmetric
    positive semidefinite, and does not vanish on a function with a kink"""
    mesh = make(dim, n)
    lset = xf.InterpolateToP1(lambda *x: torch.sqrt(sum((xi - 0.5) ** 2 for xi in x)) - 0.33, mesh)
    ci = xf.CutInfo(mesh, lset)
    hasneg = ci.GetElementsOfType(xf.HASNEG)
    facets = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasneg)
    V = xf.H1(mesh, order=p, dgjumps=True)
    u, v = V.TnT()
    h = xf.specialcf.mesh_size
    a = xf.RestrictedBilinearForm(V, element_restriction=hasneg, facet_restriction=facets, check_unused=False)
    a += 0.1 * h * (xf.dn(u, 1) - xf.dn(u.Other(), 1)) * (xf.dn(v, 1) - xf.dn(v.Other(), 1)) * xf.dCut(lset, xf.NEG, skeleton=True, definedonelements=facets)
    a.Assemble()
    A = a.mat.to_scipy()
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    # nodal interpolant of a polynomial: vertex dofs = values, P2 edge dofs from the hierarchical basis (-2 lam_a lam_b)
    pts = mesh.points.cpu().numpy()
    if p == 1:
        f = lambda X: 1 + 2 * X[:, 0] - X[:, 1]
        uvec = f(pts)
    else:
        f = lambda X: 1 + X[:, 0] - X[:, 1] + X[:, 0] * X[:, 1] + 0.5 * X[:, -1] ** 2
        edges = mesh.edges()[0].cpu().numpy()
        mid = 0.5 * (pts[edges[:, 0]] + pts[edges[:, 1]])
        # f(mid) = (f_a + f_b)/2 + c_e * (-2 * 1/4)  =>  c_e = (f_a + f_b) - 2 f(mid)
        ce = f(pts[edges[:, 0]]) + f(pts[edges[:, 1]]) - 2 * f(mid)
        uvec = np.concatenate([f(pts), ce])
    y = A @ uvec
    assert np.abs(y).max() < 1e-10 * abs(A).max() * np.abs(uvec).max()
    kink = np.abs(pts[:, 0] - 0.5)
    w = np.concatenate([kink, np.zeros(len(uvec) - len(kink))])
    assert w @ (A @ w) > 1e-6


def test_assembled_cut_skeleton_measure():
    """u*v on the cut skeleton with P1 functions: 1^T A 1 = measure of the interior facets inside the level-set domain"""
    n = 8
    mesh = make(2, n)
    c = 0.41
    lset = xf.InterpolateToP1(lambda x, y: x - c, mesh)
    V = xf.H1(mesh, order=1, dgjumps=True)
    u, v = V.TnT()
    allf = xf.BitArray(mesh.nfacet, DEV).Set()
    a = xf.BilinearForm(V, check_unused=False)
    a += u * v * xf.dCut(lset, xf.NEG, skeleton=True, definedonelements=allf)
    a.Assemble()
    one = torch.ones(V.ndof, dtype=torch.float64, device=DEV)
    total = float((one * (a.mat * one)).sum())
    vert = sum(1.0 for i in range(1, n) if i / n < c)
    ref = (n - 1) * c + n * np.sqrt(2) * c + vert
    assert abs(total - ref) < 1e-12 * ref


def test_unsupported_facet_requests_raise():
    mesh = make(2, 3)
    lset = xf.InterpolateToP1(lambda x, y: x - 0.4, mesh)
    V = xf.H1(mesh, order=1, dgjumps=True)
    u, v = V.TnT()
    els = xf.BitArray(mesh.ne, DEV).Set()
    a = xf.BilinearForm(V, check_unused=False)
    a += (u - u.Other()) * v * xf.dCut(lset, xf.NEG, element_boundary=True, definedonelements=els)
    with pytest.raises(xf.XgpuError):
        a.Assemble()                      # .Other() on an element boundary
    with pytest.raises(xf.XgpuError):
        xf.dCut(lset, xf.NEG, skeleton=True, definedonelements=els, deformation=object())


def test_coefficient_functions_on_cut_facet_integrands():
    """cf(x, y) * (u*v) * dCut(lset, NEG, skeleton=True | element_boundary=True): the coefficient is tabulated at the physical
    points of the cut facet rules and folded into their weights.  lset = x - c, cf = x, P1: 1^T A 1 = int_{facets, x < c} x ds
    in closed form; a constant callable reproduces the constant factor bit for bit"""
    n, c = 8, 0.41
    mesh = make(2, n)
    lset = xf.InterpolateToP1(lambda x, y: x - c, mesh)
    V = xf.H1(mesh, order=1, dgjumps=True)
    u, v = V.TnT()
    allf = xf.BitArray(mesh.nfacet, DEV).Set()
    alle = xf.BitArray(mesh.ne, DEV).Set()
    cf = xf.CoefficientFunction(lambda x, y: x)
    one = torch.ones(V.ndof, dtype=torch.float64, device=DEV)
    a = xf.BilinearForm(V, check_unused=False)
    a += cf * (u * v) * xf.dCut(lset, xf.NEG, skeleton=True, definedonelements=allf)
    a.Assemble()
    vert = sum(i / n for i in range(1, n) if i / n < c)
    ref = (n - 1) * c * c / 2 + n * np.sqrt(2) * c * c / 2 + vert
    assert abs(float((one * (a.mat * one)).sum()) - ref) < 1e-12 * ref
    b = xf.BilinearForm(V, check_unused=False)
    b += cf * (u * v) * xf.dCut(lset, xf.NEG, element_boundary=True, definedonelements=alle)
    b.Assemble()
    ref_eb = 2 * ref + c * c           # interior facets from both sides + bottom and top edge (the left edge has x = 0)
    assert abs(float((one * (b.mat * one)).sum()) - ref_eb) < 1e-12 * ref_eb
    # DG jump term with a constant written as a callable = the constant factor
    k1 = xf.BilinearForm(V, check_unused=False)
    k1 += xf.CoefficientFunction(lambda x, y: torch.full_like(x, 2.5)) * ((u - u.Other()) * (v - v.Other())) \
        * xf.dCut(lset, xf.NEG, skeleton=True, definedonelements=allf)
    k1.Assemble()
    k2 = xf.BilinearForm(V, check_unused=False)
    k2 += 2.5 * ((u - u.Other()) * (v - v.Other())) * xf.dCut(lset, xf.NEG, skeleton=True, definedonelements=allf)
    k2.Assemble()
    assert torch.equal(k1.mat.col, k2.mat.col)
    assert float((k1.mat.vals - k2.mat.vals).abs().max()) <= 1e-15 * float(k2.mat.vals.abs().max())
    # 3D: cf = z on the cut skeleton of lset = x - c against the facet rules themselves (sum_q w_q |F| z_q)
    m3 = make(3, 3)
    l3 = xf.InterpolateToP1(lambda x, y, z: x - 0.37, m3)
    V3 = xf.H1(m3, order=1, dgjumps=True)
    u3, v3 = V3.TnT()
    f3 = xf.BitArray(m3.nfacet, DEV).Set()
    o3 = torch.ones(V3.ndof, dtype=torch.float64, device=DEV)
    az = xf.BilinearForm(V3, check_unused=False)
    az += xf.CoefficientFunction(lambda x, y, z: z) * (u3 * v3) * xf.dCut(l3, xf.NEG, skeleton=True, definedonelements=f3)
    az.Assemble()
    a1 = xf.BilinearForm(V3, check_unused=False)
    a1 += u3 * v3 * xf.dCut(l3, xf.NEG, skeleton=True, definedonelements=f3)
    a1.Assemble()
    tz, t1 = float((o3 * (az.mat * o3)).sum()), float((o3 * (a1.mat * o3)).sum())
    ents, _, fe = m3.facets()
    inner = torch.nonzero(fe[:, 1] >= 0).reshape(-1)
    fv = ents[inner].to(torch.int32).contiguous()
    off, q, w = ops.cutquad(2, None, fv, m3.points, l3.vec, NEG, 2)
    X = m3.points[fv.long()]                                            # (nf, 3, 3)
    meas = torch.linalg.cross(X[:, 0] - X[:, 2], X[:, 1] - X[:, 2]).norm(dim=1)
    f_of = torch.repeat_interleave(torch.arange(fv.shape[0], device=DEV), off[1:] - off[:-1])
    lam = torch.stack([q[:, 0], q[:, 1], 1.0 - q[:, 0] - q[:, 1]], dim=1)
    zq = (lam * X[f_of][:, :, 2]).sum(dim=1)
    want1, wantz = float((w * meas[f_of]).sum()), float((w * meas[f_of] * zq).sum())
    assert t1 > 1.0 and abs(t1 - want1) < 1e-12 * t1 and abs(tz - wantz) < 1e-12 * t1
