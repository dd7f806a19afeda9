"""Space-time element matrices / vectors and their assembly on the GPU (xgpu_elmat_st, xgpu_elvec_st, BilinearForm /
LinearForm on a SpaceTimeFESpace; BASELINE config 5; SURVEY.md §8 a9 + a10) — against the oracle restatement of
SpaceTimeFE (1e-12 relative, Frobenius), against the closed-form answers of the reference's own space-time test
(tests/pytests/test_spacetimecutrule.py:44-69: level set 1-2x-2t on the unit square, NEG 7/8, POS 1/8, IF 1/2) and
against the weak form of a heat equation whose solution lies in the discrete space (the DG-in-time formulation of
tests/pytests/test_spacetimecutrule.py:100-104)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import _lib, ops
from oracle import xoracle as xo

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEG, POS, IF = 0, 1, 2
ID, GRAD, DT = 0, 1, 3


def moving_lset(mesh, tfe):
    if mesh.dim == 3:
        f = lambda x, y, z, t: torch.sqrt((x - .45 - .1 * t) ** 2 + (y - .5) ** 2 + (z - .55) ** 2) - 0.37
    elif mesh.dim == 2:
        f = lambda x, y, t: torch.sqrt((x - .5) ** 2 + (y - .45 - .15 * t) ** 2) - 0.31
    else:
        f = lambda x, t: x - 0.43 - 0.2 * t
    return xf.SpaceTimeInterpolateToP1(f, mesh, tfe)


def make_mesh(dim, n):
    if dim == 1:
        return xf.MakeStructured1DMesh(nx=n, device=DEV)
    if dim == 2:
        return xf.MakeStructured2DMesh(nx=n, ny=n, device=DEV)
    return xf.MakeStructured3DMesh(nx=n, device=DEV)


@pytest.mark.parametrize("dim,n,p,kt,dt", [(3, 3, 1, 2, NEG), (3, 3, 2, 1, IF), (3, 2, 2, 2, NEG), (2, 5, 1, 1, POS), (2, 5, 2, 2, NEG),
                                            (2, 4, 3, 3, IF), (1, 7, 2, 2, NEG), (3, 2, 3, 1, NEG), (3, 3, 1, 1, IF), (2, 5, 1, 2, NEG),
                                            (2, 5, 2, 1, IF), (3, 2, 1, 0, NEG)])
def test_spacetime_element_matrices_match_oracle(dim, n, p, kt, dt):
    mesh = make_mesh(dim, n)
    ls = moving_lset(mesh, xf.ScalarTimeFE(2))
    terms = [(ID, ID, 1.25, 0), (GRAD, GRAD, 0.5, 1), (DT, ID, -0.75, 0), (ID, DT, 0.3, -1), (DT, DT, 0.1, 0)]
    off, q, tv, w = ops.cutquad_st(dim, None, mesh.elements, mesh.points, ls.vec, 2, dt, 2 * kt, 2 * p)
    assert w.numel() > 0
    m = ops.elmat_st(dim, p, kt, None, mesh.elements, mesh.points, off, q, tv, w, terms)
    P, E = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy()
    offn, qn, tvn, wn = off.cpu().numpy(), q.cpu().numpy(), tv.cpu().numpy(), w.cpu().numpy()
    nd = ops.h1_ndof(dim, p) * (kt + 1)
    mo = np.zeros((mesh.ne, nd, nd))
    for el in range(mesh.ne):
        a, b = offn[el], offn[el + 1]
        if b > a:
            T = xo.make_trafo(P[E[el]], E[el])
            xo.elmat_st_from_rule(T, p, kt, qn[a:b], tvn[a:b], wn[a:b], terms, out=mo[el])
    assert np.linalg.norm(mo) > 0
    assert np.linalg.norm(m.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)
    # accumulate + slots
    perm = torch.randperm(mesh.ne, generator=torch.Generator().manual_seed(2)).to(torch.int32).to(DEV)
    buf = torch.zeros_like(m)
    for _ in range(2):
        ops.elmat_st(dim, p, kt, None, mesh.elements, mesh.points, off, q, tv, w, terms, out=buf, accumulate=True, slots=perm)
    assert torch.allclose(buf[perm.long()], 2 * m, rtol=1e-14, atol=0)
    # run-to-run bit reproducibility
    m2 = ops.elmat_st(dim, p, kt, None, mesh.elements, mesh.points, off, q, tv, w, terms)
    assert torch.equal(m, m2)
    # the generic kernel (k_elmat_st) and the tensor-structured one (k_elmat_st_tp, default where instantiated) agree
    ctx = _lib.default_ctx()
    ctx.set_option("elmat_st_tp", 0)
    try:
        mg = ops.elmat_st(dim, p, kt, None, mesh.elements, mesh.points, off, q, tv, w, terms)
    finally:
        ctx.set_option("elmat_st_tp", 1)
    assert np.linalg.norm(mg.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)


@pytest.mark.parametrize("dim,n,p,kt,dt", [(3, 3, 1, 2, NEG), (3, 2, 2, 1, IF), (2, 5, 2, 2, NEG), (2, 4, 4, 3, POS), (1, 7, 3, 1, NEG)])
def test_spacetime_element_vectors_match_oracle(dim, n, p, kt, dt):
    mesh = make_mesh(dim, n)
    ls = moving_lset(mesh, xf.ScalarTimeFE(1))
    ex = lambda e: tuple(v if k < dim or k == 3 else 0 for k, v in enumerate(e))
    terms = [(ID, 1.5, 0, ex((2, 1, 0, 1))), (DT, -0.5, 1, ex((0, 0, 1, 2))), (ID, 0.25, -1, ex((1, 0, 0, 0)))]
    off, q, tv, w = ops.cutquad_st(dim, None, mesh.elements, mesh.points, ls.vec, 1, dt, 2 * kt, 2 * p)
    v = ops.elvec_st(dim, p, kt, None, mesh.elements, mesh.points, off, q, tv, w, terms)
    P, E = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy()
    offn, qn, tvn, wn = off.cpu().numpy(), q.cpu().numpy(), tv.cpu().numpy(), w.cpu().numpy()
    nd = ops.h1_ndof(dim, p) * (kt + 1)
    vo = np.zeros((mesh.ne, nd))
    for el in range(mesh.ne):
        a, b = offn[el], offn[el + 1]
        if b > a:
            T = xo.make_trafo(P[E[el]], E[el])
            xo.elvec_st_from_rule(T, p, kt, qn[a:b], tvn[a:b], wn[a:b], terms, out=vo[el])
    assert np.linalg.norm(vo) > 0
    assert np.linalg.norm(v.cpu().numpy() - vo) <= 1e-12 * np.linalg.norm(vo)


def test_unsupported_requests_fail_loudly():
    mesh = make_mesh(2, 3)
    ls = moving_lset(mesh, xf.ScalarTimeFE(1))
    off, q, tv, w = ops.cutquad_st(2, None, mesh.elements, mesh.points, ls.vec, 1, NEG, 2, 2)
    with pytest.raises(xf.XgpuError):   # GRAD paired with ID
        ops.elmat_st(2, 1, 1, None, mesh.elements, mesh.points, off, q, tv, w, [(GRAD, ID, 1.0, 0)])
    with pytest.raises(xf.XgpuError):   # DN is not a space-time operator
        ops.elmat_st(2, 1, 1, None, mesh.elements, mesh.points, off, q, tv, w, [(2, 2, 1.0, 0)])
    with pytest.raises(xf.XgpuError):   # time order
        ops.elmat_st(2, 1, 4, None, mesh.elements, mesh.points, off, q, tv, w, [(ID, ID, 1.0, 0)])
    # a space-time proxy on a merely spatial rule: the reference throws in SpaceTimeFE::CalcShape
    V = xf.SpaceTimeFESpace(xf.H1(mesh, order=1), xf.ScalarTimeFE(1))
    u, v = V.TnT()
    a = xf.BilinearForm(V)
    a += u * v * xf.dCut(xf.InterpolateToP1(lambda x, y: x - 0.4, mesh), NEG)
    with pytest.raises(xf.XgpuError, match="mere space IR"):
        a.Assemble()
    # and a space-time symbol on a spatial space
    W = xf.H1(mesh, order=1)
    uw, vw = W.TnT()
    b = xf.BilinearForm(W)
    b += uw * vw * xf.dCut(ls, NEG, time_order=2)
    with pytest.raises(xf.XgpuError):
        b.Assemble()


def st_coeffs(V, fn):
    """coefficients of a function that is P1/P2-representable in space and nodal in time: fn(points, t) -> values;
    hierarchical P2 edge coefficient = f(a) + f(b) - 2 f(mid) (fes.Deformation.interpolate_p2)"""
    mesh = V.mesh
    p = mesh.points
    out = []
    for t in V.TimeFE_nodes():
        c = [fn(p, t)]
        if V.order == 2:
            edges, _ = mesh.edges()
            a, b = p[edges[:, 0].long()], p[edges[:, 1].long()]
            c.append(fn(a, t) + fn(b, t) - 2.0 * fn(0.5 * (a + b), t))
        out.append(torch.cat(c))
    return torch.cat(out)


@pytest.mark.parametrize("p,kt", [(1, 1), (2, 2)])
def test_moving_plane_known_answers(p, kt):
    """tests/pytests/test_spacetimecutrule.py:44-69: lset(t) = 1 - 2x - 2t on the unit square, space-time measures
    NEG 7/8, POS 1/8, IF 1/2 — here through ASSEMBLED operators applied to polynomial coefficient vectors"""
    mesh = xf.MakeStructured2DMesh(nx=4, ny=4, device=DEV)
    tfe_l = xf.ScalarTimeFE(1)
    ls = xf.SpaceTimeInterpolateToP1(lambda x, y, t: 1 - 2 * x - 2 * t, mesh, tfe_l)
    V = xf.SpaceTimeFESpace(xf.H1(mesh, order=p), xf.ScalarTimeFE(kt))
    u, v = V.TnT()
    one = st_coeffs(V, lambda P, t: torch.ones_like(P[:, 0]))
    tt = st_coeffs(V, lambda P, t: torch.full_like(P[:, 0], t))
    xx = st_coeffs(V, lambda P, t: P[:, 0])
    xt = st_coeffs(V, lambda P, t: P[:, 0] * t)
    ref = {NEG: 7 / 8, POS: 1 / 8, IF: 0.5}
    for dom in (NEG, POS, IF):
        m = xf.BilinearForm(V)
        m += u * v * xf.dCut(ls, dom, time_order=2 * kt)
        m.Assemble()
        assert abs(float(one @ (m.mat * one)) - ref[dom]) < 1e-13
    # dt(u) v with u = t, v = 1: the NEG measure again; u = 1: zero
    a = xf.BilinearForm(V)
    a += xf.dt(u) * v * xf.dCut(ls, NEG, time_order=2 * kt)
    a.Assemble()
    assert abs(float(one @ (a.mat * tt)) - 7 / 8) < 1e-13
    assert abs(float(one @ (a.mat * one))) < 1e-13
    # grad(u).grad(v) with u = x t, v = x: int_NEG t = 23/48
    k = xf.BilinearForm(V)
    k += xf.grad(u) * xf.grad(v) * xf.dCut(ls, NEG, time_order=2 * kt)
    k.Assemble()
    assert abs(float(xx @ (k.mat * xt)) - 23 / 48) < 1e-13
    # linear form: int_NEG x^0 tref v, summed over a partition of unity = int_NEG t = 23/48
    f = xf.LinearForm(V)
    f += xf.tref * v * xf.dCut(ls, NEG, time_order=2 * kt)
    f.Assemble()
    assert abs(float(one @ f.vec) - 23 / 48) < 1e-13
    # fixed reference time: the NEG domain at tref = 0.25 is x > 0.25 (area 3/4), at tref = 1 everything
    for t0, area in ((0.25, 0.75), (1.0, 1.0), (0.0, 0.5)):
        for lset0 in (ls, xf.InterpolateToP1(lambda x, y: 1 - 2 * x - 2 * t0, mesh)):
            g = xf.BilinearForm(V)
            g += xf.fix_tref(u, t0) * xf.fix_tref(v, t0) * xf.dCut(lset0, NEG)
            g.Assemble()
            assert abs(float(one @ (g.mat * one)) - area) < 1e-13
        g = xf.BilinearForm(V)
        g += u * v * xf.dCut(xf.InterpolateToP1(lambda x, y: 1 - 2 * x - 2 * t0, mesh), NEG, tref=t0)
        g.Assemble()
        assert abs(float(one @ (g.mat * one)) - area) < 1e-13


def test_heat_equation_weak_form_is_satisfied_by_a_solution_in_the_space():
    """DG-in-time slab formulation of tests/pytests/test_spacetimecutrule.py:100-104 on [0,1]^2:
         a = delta_t grad(u).grad(v) dxtref + dt(u) v dxtref + fix_tref(u,0) fix_tref(v,0) dx
         f = delta_t f v dxtref + u_old fix_tref(v,0) dx
    u = x y (1 + t_old + delta_t tref) solves u_t - Laplace u = x y and lies in P2 x P1, so (A u - f) vanishes on every
    test function that vanishes on the boundary of the square."""
    n, told, dlt = 4, 0.3, 0.1
    mesh = xf.MakeStructured2DMesh(nx=n, ny=n, device=DEV)
    Vs = xf.H1(mesh, order=2)
    V = xf.SpaceTimeFESpace(Vs, xf.ScalarTimeFE(1))
    u, v = V.TnT()
    a = xf.BilinearForm(V)
    a += dlt * xf.grad(u) * xf.grad(v) * xf.dxtref(mesh, time_order=2)
    a += xf.dt(u) * v * xf.dxtref(mesh, time_order=2)
    a += xf.fix_tref(u, 0) * xf.fix_tref(v, 0) * xf.dmesh(mesh)
    a.Assemble()
    f = xf.LinearForm(V)
    f += dlt * xf.x * xf.y * v * xf.dxtref(mesh, time_order=2)
    f += (1 + told) * xf.x * xf.y * xf.fix_tref(v, 0) * xf.dmesh(mesh)
    f.Assemble()
    uex = st_coeffs(V, lambda P, t: P[:, 0] * P[:, 1] * (1 + told + dlt * t))
    r = a.mat * uex - f.vec
    # interior space dofs: vertices / edges not on the boundary
    P = mesh.points
    onb = lambda X: ((X[:, 0] < 1e-12) | (X[:, 0] > 1 - 1e-12) | (X[:, 1] < 1e-12) | (X[:, 1] > 1 - 1e-12))
    edges, _ = mesh.edges()
    mid = 0.5 * (P[edges[:, 0].long()] + P[edges[:, 1].long()])
    interior = torch.cat([~onb(P), ~onb(mid)])
    interior = torch.cat([interior, interior])
    assert int(interior.sum()) > 0
    assert float(r[interior].abs().max()) < 1e-13
    assert float(r[~interior].abs().max()) > 1e-6   # boundary rows carry the flux: the check is not vacuous
    # sparsity: time-major blocks share the spatial pattern
    ms = xf.BilinearForm(Vs)
    us, vs = Vs.TnT()
    ms += us * vs * xf.dCut(xf.GridFunction(mesh, -torch.ones(mesh.nv, dtype=torch.float64)), NEG)
    ms.Assemble()
    assert a.mat.nze == 4 * ms.mat.nze


def test_config5_shape_assembles_and_integrates_to_the_cylinder_volume():
    """config 5 shape at small size: moving sphere, P1 space x P2 time level set and FE space; 1^T M 1 = space-time
    volume of the discrete NEG domain = sum of the rule weights x |det F| (basis-independent)"""
    mesh = xf.MakeStructured3DMesh(nx=8, device=DEV)
    tfe = xf.ScalarTimeFE(2)
    ls = xf.SpaceTimeInterpolateToP1(
        lambda x, y, z, t: torch.sqrt((x - .5) ** 2 + (y - .5 - .1 * np.sin(2 * np.pi * 0.1 * t)) ** 2 + (z - .5) ** 2) - 0.3, mesh, tfe)
    V = xf.SpaceTimeFESpace(xf.H1(mesh, order=1), tfe)
    u, v = V.TnT()
    m = xf.BilinearForm(V)
    m += (u * v + xf.dt(u) * v) * xf.dCut(ls, NEG, time_order=4)
    m += 0.01 * xf.grad(u) * xf.grad(v) * xf.dCut(ls, NEG, time_order=4)
    m.Assemble()
    one = torch.ones(V.ndof, dtype=torch.float64, device=DEV)
    off, q, tv, w = ops.cutquad_st(3, None, mesh.elements, mesh.points, ls.vec, 2, NEG, 4, 2)
    cnt = (off[1:] - off[:-1])
    vol = float((w * torch.repeat_interleave(torch.full((mesh.ne,), 1.0 / 8 ** 3, dtype=torch.float64, device=DEV), cnt)).sum())
    got = float(one @ (m.mat * one))
    assert abs(got - vol) < 1e-13
    assert abs(vol - 4 / 3 * np.pi * 0.3 ** 3) < 0.02   # polyhedral approximation of the ball, constant in time


# ---------------------------------------------------------------- Integrate on space-time rules: the reference's own tests
@pytest.mark.parametrize("cf,ref,so,to", [(lambda: xf.tref, 0.5, 0, 1), (lambda: xf.tref ** 3, 0.25, 0, 3),
                                          (lambda: xf.x, 0.5, 1, 0), (lambda: xf.tref * xf.tref * xf.x * xf.x, 1 / 9, 2, 2),
                                          (lambda: xf.tref * xf.tref * xf.y * xf.y, 1 / 9, 2, 2)])
def test_spacetime_integrate_no_cut(cf, ref, so, to):
    """tests/pytests/test_spacetimecutrule.py:14-37 (tref^2 (x^2+y^2) = 2/9 split into its two monomials)"""
    mesh = xf.MakeStructured2DMesh(nx=1, ny=1, device=DEV)
    ls = xf.SpaceTimeGridFunction(mesh, xf.ScalarTimeFE(1), -torch.ones((2, mesh.nv), dtype=torch.float64))
    val = xf.Integrate(levelset_domain={"levelset": ls, "domain_type": NEG}, cf=cf(), mesh=mesh, order=so, time_order=to)
    assert abs(val - ref) < 5e-15


@pytest.mark.parametrize("dom,ref", [(POS, 1 / 8), (NEG, 7 / 8), (IF, 0.5)])
def test_spacetime_integrate_moving_plane(dom, ref):
    """tests/pytests/test_spacetimecutrule.py:40-69"""
    mesh = xf.MakeStructured2DMesh(nx=1, ny=1, device=DEV)
    ls = xf.SpaceTimeInterpolateToP1(lambda x, y, t: 1 - 2 * x - 2 * t, mesh, xf.ScalarTimeFE(1))
    val = xf.Integrate(levelset_domain={"levelset": ls, "domain_type": dom}, cf=1.0, mesh=mesh, order=0, time_order=0)
    assert abs(val - ref) < 5e-15
    # a spatial level set under a space-time rule is constant in time (cutint/xintegration.cpp:60-70)
    l0 = xf.InterpolateToP1(lambda x, y: x - 0.25, mesh)
    v0 = xf.Integrate(levelset_domain={"levelset": l0, "domain_type": NEG, "time_order": 2}, cf=xf.tref, mesh=mesh, order=1)
    assert abs(v0 - 0.25 * 0.5) < 5e-15
    # fixed reference time (dmesh(tref=...) style domain)
    v1 = xf.Integrate(levelset_domain={"levelset": l0, "domain_type": POS, "tref": 0.5}, cf=xf.tref * xf.x, mesh=mesh, order=1)
    assert abs(v1 - 0.5 * (0.5 - 0.25 ** 2 / 2)) < 5e-15


# ---------------------------------------------------------------- CutInfo on a time slab
@pytest.mark.parametrize("dim,n,kt,to", [(2, 6, 1, 0), (2, 5, 2, 2), (3, 3, 2, 0), (1, 9, 1, 1)])
def test_spacetime_cutinfo_matches_oracle(dim, n, kt, to):
    """CutInformation::Update (xfem/cutinfo.cpp:90-124) with a space-time level set: part volumes from the order-0
    space-time rules; bit-exact BitArrays against the oracle's rules, cut ratios to 1e-13"""
    mesh = make_mesh(dim, n)
    ls = moving_lset(mesh, xf.ScalarTimeFE(kt))
    ci = xf.CutInfo(mesh, ls, time_order=to)
    P, E, L = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy(), ls.vec.cpu().numpy()
    neg = np.zeros(mesh.ne, bool); pos = np.zeros(mesh.ne, bool); cut = np.zeros(mesh.ne, bool)
    ratio = np.zeros(mesh.ne)
    for el in range(mesh.ne):
        T = xo.make_trafo(P[E[el]], E[el])
        pv = {}
        for np_ in (POS, NEG):
            _, _, w = xo.spacetime_cut_rule(T, L[:, E[el]], kt, np_, to, 0)
            s = 0.0
            for x in w:
                s += x
            pv[np_] = s
        ratio[el] = pv[NEG] / (pv[NEG] + pv[POS])
        if pv[NEG] > 0:
            (cut if pv[POS] > 0 else neg)[el] = True
        else:
            pos[el] = True
    got = {k: ci.GetElementsOfType(k).to_numpy() for k in (xf.NEG, xf.POS, xf.IF)}
    assert (got[xf.NEG] == neg).all() and (got[xf.POS] == pos).all() and (got[xf.IF] == cut).all()
    assert cut.any() and neg.any() and pos.any()
    assert np.allclose(ci.GetCutRatios().cpu().numpy(), ratio, rtol=1e-13, atol=1e-15)
    hasneg = ci.GetElementsOfType(xf.HASNEG).to_numpy()
    assert (hasneg == (neg | cut)).all()
    # an element that is cut at ANY time of the slab is IF: superset of the spatial IF sets at the time nodes
    for j in range(kt + 1):
        cj = xf.CutInfo(mesh, xf.GridFunction(mesh, ls.vec[j]))
        assert not (cj.GetElementsOfType(xf.IF).to_numpy() & ~cut).any()


def test_spacetime_cutinfo_of_a_spatial_level_set_equals_the_spatial_one():
    mesh = make_mesh(3, 4)
    l0 = xf.InterpolateToP1(lambda x, y, z: torch.sqrt((x - .5) ** 2 + (y - .5) ** 2 + (z - .5) ** 2) - 0.3, mesh)
    a, b = xf.CutInfo(mesh, l0), xf.CutInfo(mesh, l0, time_order=0)
    for k in (xf.NEG, xf.POS, xf.IF, xf.HASNEG, xf.HASPOS, xf.UNCUT):
        assert (a.GetElementsOfType(k).to_numpy() == b.GetElementsOfType(k).to_numpy()).all()
    assert torch.allclose(a.GetCutRatios(), b.GetCutRatios(), rtol=1e-13, atol=1e-15)


# ---------------------------------------------------------------- space-time ghost penalty
def time_mass(kt, time_order=None, tref=None):
    """time matrix of the tensor-product patch rule (xfem/symboliccutbfi.cpp:2095-2177) from numpy's Gauss-Legendre"""
    nodes = xf.ScalarTimeFE(kt).nodes()

    def L(t):
        return np.array([np.prod([(t - nj) / (ni - nj) for j, nj in enumerate(nodes) if j != i]) for i, ni in enumerate(nodes)])
    if tref is not None:
        return np.outer(L(tref), L(tref))
    x, w = np.polynomial.legendre.leggauss(time_order // 2 + 1)
    return sum(0.5 * wi * np.outer(L(0.5 * (xi + 1)), L(0.5 * (xi + 1))) for xi, wi in zip(x, w))


@pytest.mark.parametrize("dim,n,p,kt,to,tref", [(2, 4, 1, 1, 2, None), (2, 3, 2, 2, 4, None), (3, 2, 1, 2, 4, None), (3, 2, 2, 1, -1, 0.25),
                                                (2, 3, 3, 1, 2, None)])
def test_spacetime_facet_patch_matches_oracle(dim, n, p, kt, to, tref):
    mesh = make_mesh(dim, n)
    _, _, fe = mesh.facets()
    inner = torch.nonzero(fe[:, 1] >= 0).reshape(-1).to(torch.int32)
    m = ops.facetpatch_st(dim, p, kt, inner, fe, mesh.elements, mesh.points, -1, to, tref, 0.7, -2)
    F, nfail = xo.batch_facet_patch(dim, p, inner.cpu().numpy(), fe.cpu().numpy(), mesh.elements.cpu().numpy(),
                                    mesh.points.cpu().numpy(), -1, 0.7, -2)
    assert nfail == 0
    T = time_mass(kt, to, tref)
    nds, nt = ops.h1_ndof(dim, p), kt + 1
    nf = F.shape[0]
    # out[(h,j,i),(g,l,k)] = T[j,l] F[(h,i),(g,k)]
    F4 = F.reshape(nf, 2, nds, 2, nds)
    mo = np.einsum("jl,fhigk->fhjiglk", T, F4).reshape(nf, 2 * nds * nt, 2 * nds * nt)
    assert np.linalg.norm(mo) > 0
    assert np.linalg.norm(m.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)
    # downscale: the spatial patch matrix carries the scaling, the time matrix does not
    ms = ops.facetpatch_st(dim, p, kt, inner, fe, mesh.elements, mesh.points, -1, to, tref, 0.7, -2, downscale=0.5)
    Fs, nfail = xo.batch_facet_patch(dim, p, inner.cpu().numpy(), fe.cpu().numpy(), mesh.elements.cpu().numpy(),
                                     mesh.points.cpu().numpy(), -1, 0.7, -2, downscale=0.5)
    mos = np.einsum("jl,fhigk->fhjiglk", T, Fs.reshape(nf, 2, nds, 2, nds)).reshape(nf, 2 * nds * nt, 2 * nds * nt)
    assert nfail == 0 and np.linalg.norm(ms.cpu().numpy() - mos) <= 1e-12 * np.linalg.norm(mos)


def test_spacetime_ghost_penalty_assembled_kernel_and_definiteness():
    """gamma h^-2 (u - u.Other())(v - v.Other()) dFacetPatch(time_order): vanishes on functions that are ONE polynomial of
    degree <= p in space (any nodal time dependence), positive on others; fixed-tref variant likewise"""
    mesh = xf.MakeStructured2DMesh(nx=4, ny=4, device=DEV)
    ls = xf.SpaceTimeInterpolateToP1(lambda x, y, t: torch.sqrt((x - .5) ** 2 + (y - .45 - .1 * t) ** 2) - 0.31, mesh, xf.ScalarTimeFE(1))
    ci = xf.CutInfo(mesh, ls, time_order=0)
    hasneg = ci.GetElementsOfType(xf.HASNEG)
    ba_facets = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=ci.GetElementsOfType(xf.IF))
    assert ba_facets.NumSet() > 0
    V = xf.SpaceTimeFESpace(xf.H1(mesh, order=2, dgjumps=True), xf.ScalarTimeFE(1))
    u, v = V.TnT()
    h = xf.specialcf.mesh_size
    for dx in (xf.dFacetPatch(definedonelements=ba_facets, time_order=2), xf.dFacetPatch(definedonelements=ba_facets, tref=1.0)):
        g = xf.BilinearForm(V)
        g += 0.1 / h ** 2 * (u - u.Other()) * (v - v.Other()) * dx
        g += u * v * xf.dCut(ls, NEG, time_order=2, definedonelements=hasneg)
        g.Assemble()
        m = xf.BilinearForm(V)
        m += u * v * xf.dCut(ls, NEG, time_order=2, definedonelements=hasneg)
        m.Assemble()
        poly = st_coeffs(V, lambda P, t: (1 + 2 * t) * P[:, 0] * P[:, 1] + P[:, 0] ** 2 - t)
        gp = g.mat * poly - m.mat * poly     # same space, same dof numbering: the ghost-penalty part alone
        assert float(gp.abs().max()) < 1e-12
        rough = torch.rand(V.ndof, dtype=torch.float64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
        e = float(rough @ (g.mat * rough)) - float(rough @ (m.mat * rough))
        assert e > 1e-3
    assert g.mat.nze > m.mat.nze


@pytest.mark.parametrize("dim,n,kt,to", [(3, 5, 2, 0), (2, 9, 1, 2), (1, 17, 2, 1), (3, 4, 3, 0)])
def test_one_pass_slab_classification_equals_the_rule_based_one_bit_for_bit(dim, n, kt, to):
    """xgpu_classify_st (no rule storage) == xgpu_cutquad_st_* + xgpu_rule_element_sums + xgpu_classify_volumes"""
    mesh = make_mesh(dim, n)
    ls = moving_lset(mesh, xf.ScalarTimeFE(kt))
    vols = {}
    for np_ in (POS, NEG):
        off, _, _, w = ops.cutquad_st(dim, None, mesh.elements, mesh.points, ls.vec, kt, np_, to, 0)
        vols[np_] = ops.rule_element_sums(off, w)
    ref = (ops.new_bits(mesh.ne, DEV), ops.new_bits(mesh.ne, DEV), ops.new_bits(mesh.ne, DEV),
           torch.empty(mesh.ne, dtype=torch.float64, device=DEV))
    ops.classify_volumes(vols[NEG], vols[POS], ref)
    got = ops.classify_st(dim, mesh.elements, ls.vec, kt, to)
    for a, b in zip(got[:3], ref[:3]):
        assert torch.equal(a, b)
    assert torch.equal(got[3], ref[3])
    assert int(ops.bitset_count(got[2], mesh.ne)) > 0


# ---------------------------------------------------------------- edge cases
def test_spacetime_edge_cases():
    mesh = make_mesh(2, 3)                       # 18 triangles: not a multiple of 32
    tfe = xf.ScalarTimeFE(1)
    allpos = xf.SpaceTimeGridFunction(mesh, tfe, torch.ones((2, mesh.nv), dtype=torch.float64))
    allneg = xf.SpaceTimeGridFunction(mesh, tfe, -torch.ones((2, mesh.nv), dtype=torch.float64))
    for ls, want in ((allpos, xf.POS), (allneg, xf.NEG)):
        ci = xf.CutInfo(mesh, ls, time_order=0)
        assert ci.GetElementsOfType(want).NumSet() == mesh.ne
        assert ci.GetElementsOfType(xf.IF).NumSet() == 0
        assert ci.GetElementsOfType(xf.UNCUT).NumSet() == mesh.ne
    assert float(xf.CutInfo(mesh, allneg, time_order=0).GetCutRatios().min()) == 1.0
    # empty element list: nothing launched, nothing written
    empty = torch.zeros(0, dtype=torch.int32, device=DEV)
    off, q, tv, w = ops.cutquad_st(2, empty, mesh.elements, mesh.points, allneg.vec, 1, NEG, 2, 2)
    assert off.numel() == 1 and w.numel() == 0
    m = ops.elmat_st(2, 1, 1, empty, mesh.elements, mesh.points, off, q, tv, w, [(ID, ID, 1.0, 0)])
    assert m.shape == (0, 6, 6)
    # a NEG integral over an all-POS slab: the pattern of the space, all values zero
    V = xf.SpaceTimeFESpace(xf.H1(mesh, order=1), tfe)
    u, v = V.TnT()
    a = xf.BilinearForm(V)
    a += u * v * xf.dCut(allpos, NEG, time_order=2)
    a.Assemble()
    assert a.mat.nze > 0 and float(a.mat.vals.abs().max()) == 0.0
    # nothing active at all: the empty matrix
    none = xf.BitArray(mesh.ne, DEV)
    Vr = xf.SpaceTimeFESpace(xf.Restrict(xf.H1(mesh, order=1), none), tfe)
    ur, vr = Vr.TnT()
    b = xf.BilinearForm(Vr)
    b += ur * vr * xf.dCut(allneg, NEG, time_order=2, definedonelements=none)
    b.Assemble()
    assert Vr.ndof == 0 and b.mat.nze == 0
    f = xf.LinearForm(Vr)
    f += xf.tref * vr * xf.dCut(allneg, NEG, time_order=2)
    f.Assemble()
    assert f.vec.numel() == 0
    # integrals of empty domains
    assert xf.Integrate({"levelset": allpos, "domain_type": NEG}, mesh=mesh, cf=1.0, order=1, time_order=1) == 0.0
    assert xf.Integrate({"levelset": allpos, "domain_type": IF}, mesh=mesh, cf=xf.tref, order=1, time_order=1) == 0.0


# ---------------------------------------------------------------- solver level: DG-in-time heat equation
def _heat_run(nx, nsteps, tend):
    """tests/pytests/test_spacetimecutrule.py:72-140 restated with a polynomial solution (the accelerated path takes
    monomial coefficients): u = t x(1-x) y(1-y) on the unit square, homogeneous Dirichlet data, slabs of length
    delta_t, P1 x P1;  a = delta_t grad(u).grad(v) + dt(u) v on the slab + u(0) v(0),  f = delta_t f v + u_old v(0)."""
    import scipy.sparse.linalg as spla
    mesh = xf.MakeStructured2DMesh(nx=nx, ny=nx, device=DEV)
    dlt = tend / nsteps
    Vs = xf.H1(mesh, order=1)
    V = xf.SpaceTimeFESpace(Vs, xf.ScalarTimeFE(1))
    u, v = V.TnT()
    dxt = xf.dxtref(mesh, time_order=2)
    a = xf.BilinearForm(V)
    a += dlt * xf.grad(u) * xf.grad(v) * dxt
    a += xf.dt(u) * v * dxt
    a += xf.fix_tref(u, 0) * xf.fix_tref(v, 0) * xf.dmesh(mesh)
    a.Assemble()
    ic = xf.BilinearForm(V)
    ic += xf.fix_tref(u, 0) * xf.fix_tref(v, 0) * xf.dmesh(mesh)
    ic.Assemble()
    A, IC = a.mat.to_scipy().tocsr(), ic.mat.to_scipy().tocsr()
    P = mesh.points.cpu().numpy()
    inner = ~((P[:, 0] < 1e-12) | (P[:, 0] > 1 - 1e-12) | (P[:, 1] < 1e-12) | (P[:, 1] > 1 - 1e-12))
    free = np.flatnonzero(np.concatenate([inner, inner]))
    solve = spla.factorized(A[free][:, free].tocsc())
    nv = mesh.nv
    g = [(1.0, (1, 1)), (-1.0, (2, 1)), (-1.0, (1, 2)), (1.0, (2, 2))]       # x(1-x) y(1-y)
    s = [(1.0, (1, 0)), (-1.0, (2, 0)), (1.0, (0, 1)), (-1.0, (0, 2))]       # x(1-x) + y(1-y)
    mono = lambda c, e, te=0: c * xf.x ** e[0] * xf.y ** e[1] * (xf.tref ** te if te else 1.0)
    uold = np.zeros(nv)
    told = 0.0
    for _ in range(nsteps):
        f1 = xf.LinearForm(V)     # delta_t (g + 2 t_old s) v
        for c, e in g:
            f1 += mono(dlt * c, e) * v * dxt
        for c, e in s:
            f1 += mono(dlt * 2 * told * c, e) * v * dxt
        for c, e in s:            # delta_t * 2 delta_t tref s v
            f1 += mono(dlt * 2 * dlt * c, e, 1) * v * dxt
        f1.Assemble()
        rhs = f1.vec.cpu().numpy() + IC @ np.concatenate([uold, np.zeros(nv)])
        sol = np.zeros(2 * nv)
        sol[free] = solve(rhs[free])
        uold = sol[nv:]           # nodal in time: the values at tref = 1
        told += dlt
    exact = told * P[:, 0] * (1 - P[:, 0]) * P[:, 1] * (1 - P[:, 1])
    return np.sqrt(np.mean((uold - exact) ** 2)), np.abs(exact).max()


def test_spacetime_heat_equation_converges():
    e8, umax = _heat_run(8, 4, 0.5)
    e16, _ = _heat_run(16, 4, 0.5)
    assert e16 < 2e-4 * 1.0 and e16 < 0.02 * umax        # the reference's own bound for its sin solution is 2e-4
    assert e8 / e16 > 3.0                                 # second order in h (the solution is linear in time: no time error)


# ---------------------------------------------------------------- coefficient functions on space-time integrands
def test_spacetime_coefficient_functions_fold_into_the_rule_weights():
    """cf(x, y, tref) * v * dCut(ls, NEG, time_order) and cf * (u*v + grad(u)*grad(v)) * dCut(...): a callable of the physical
    coordinates and the reference time (xfem/symboliccutlfi.cpp:96-118, symboliccutbfi.cpp:174-193 on space-time rules).
    A monomial written as a callable gives the vector of the monomial path; M(cf) applied to the dofs of g gives the vector
    of cf * g; a non-polynomial cf is checked against the rule points themselves."""
    mesh = xf.MakeStructured2DMesh(nx=5, ny=4, device=DEV)
    tfe = xf.ScalarTimeFE(1)
    ls = xf.SpaceTimeInterpolateToP1(lambda x, y, t: torch.sqrt((x - .45 - .1 * t) ** 2 + (y - .5) ** 2) - 0.31, mesh, tfe)
    V = xf.SpaceTimeFESpace(xf.H1(mesh, order=2), tfe)
    u, v = V.TnT()
    dx = xf.dCut(ls, NEG, time_order=2, order=6)
    f1 = xf.LinearForm(V)
    f1 += (2.0 * xf.x * xf.y ** 2 * xf.tref) * v * dx
    f1.Assemble()
    cf = xf.CoefficientFunction(lambda x, y, t: 2.0 * x * y ** 2 * t)
    f2 = xf.LinearForm(V)
    f2 += cf * v * dx
    f2.Assemble()
    assert float(f1.vec.abs().max()) > 1e-4
    assert float((f1.vec - f2.vec).abs().max()) < 1e-14
    # bilinear: M(cf) g = f(cf g) for g = 1 (vertex dofs 1 at both time nodes)
    a = xf.BilinearForm(V)
    a += cf * (u * v) * dx
    a.Assemble()
    g = st_coeffs(V, lambda P, t: torch.ones_like(P[:, 0]))
    assert float((a.mat * g - f1.vec).abs().max()) < 1e-14
    # a non-polynomial coefficient: sum_j f_j g_j = sum_q w_q |det F| cf(x_q, t_q) g(x_q) with g = 1
    cfe = xf.CoefficientFunction(lambda x, y, t: torch.exp(x - 0.3 * t) * torch.cos(y))
    f3 = xf.LinearForm(V)
    f3 += cfe * v * dx
    f3.Assemble()
    off, q, tv, w = ops.cutquad_st(2, None, mesh.elements, mesh.points, ls.vec, 1, NEG, 2, 6)
    xyz, meas = ops.rule_phys(2, None, mesh.elements, mesh.points, off, q, want_meas=True)
    want = float((w * meas * torch.exp(xyz[:, 0] - 0.3 * tv) * torch.cos(xyz[:, 1])).sum())
    assert abs(float((f3.vec * g).sum()) - want) < 1e-13 * max(1.0, abs(want))
    # fixed reference time: the function sees tref = the fixed time
    f4 = xf.LinearForm(V)
    f4 += xf.CoefficientFunction(lambda x, y, t: x * t) * xf.fix_tref(v, 1.0) * xf.dmesh(mesh)
    f4.Assemble()
    f5 = xf.LinearForm(V)
    f5 += xf.x * xf.fix_tref(v, 1.0) * xf.dmesh(mesh)
    f5.Assemble()
    assert float((f4.vec - f5.vec).abs().max()) < 1e-14
    with pytest.raises(xf.XgpuError):
        cf * u                                      # a coefficient multiplies a whole integrand
