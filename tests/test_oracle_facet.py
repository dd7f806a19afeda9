"""Pins of the oracle's cut-facet restatement (oracle/xo_facet.c — SymbolicCutFacetBilinearFormIntegrator::T_CalcFacetMatrix,
xfem/symboliccutbfi.cpp:886-1290, and the element-boundary variant :488-883) against closed forms: cut facet lengths / areas of a
planar level set, the divergence theorem on element boundaries, vanishing jumps of globally smooth functions; and of the symbolic
layer that expands the DG forms of demos/fictdom_dg.py:108-125 into the kernel's term list."""
import numpy as np
import pytest

import ngsxfem_b200 as xf
from ngsxfem_b200 import meshes
from oracle import xoracle as xo
from util import interpolate_to_p1, mesh_np

NEG, POS = 0, 1
VAL, DN = 0, 1


def facet_tables(m):
    ents, el_f, fe = m.facets()
    return ents.cpu().numpy(), el_f.cpu().numpy(), fe.cpu().numpy()


@pytest.mark.parametrize("dim,n", [(2, 5), (3, 3)])
@pytest.mark.parametrize("dt", [NEG, POS])
def test_cut_measure_of_the_interior_facets(dim, n, dt):
    """level set x - c: the part of the interior skeleton in x < c (NEG) / x > c (POS) has a closed-form measure"""
    c = 0.37
    m = meshes.MakeStructured2DMesh(nx=n, ny=n) if dim == 2 else meshes.MakeStructured3DMesh(nx=n)
    pts, els = mesh_np(m)
    ents, el_f, fe = facet_tables(m)
    phi = interpolate_to_p1(lambda *x: x[0] - c, pts)
    total = 0.0
    full = 0.0
    for f in range(len(ents)):
        if fe[f, 1] < 0:
            continue
        e1 = fe[f, 0]
        T1 = xo.make_trafo(pts[els[e1]], els[e1])
        fv = ents[f]
        mat, _ = xo.cut_facet_elmat(T1, None, 1, fv, phi[fv], dt, 2, [(VAL, 0, VAL, 0, 1.0, 0)])
        total += mat.sum()            # sum_ij int N_i N_j = int 1
        fp = pts[fv]
        full += np.linalg.norm(fp[0] - fp[1]) if dim == 2 else 0.5 * np.linalg.norm(np.cross(fp[0] - fp[2], fp[1] - fp[2]))
    frac = c if dt == NEG else 1 - c
    if dim == 2:
        # horizontal lines: (n-1) * frac ; diagonals: n rows * sqrt(2) * frac ; vertical lines x = i/n strictly inside the part
        vert = sum(1.0 for i in range(1, n) if ((i / n < c) if dt == NEG else (i / n > c)))
        ref = (n - 1) * frac + n * np.sqrt(2) * frac + vert
        assert abs(total - ref) < 1e-12 * ref
    else:
        # every facet family of the Kuhn mesh is translation invariant in x or a plane x = const: check against the sum of
        # uncut parts by linearity instead — NEG + POS = full skeleton
        other, _ = 0.0, None
        for f in range(len(ents)):
            if fe[f, 1] < 0:
                continue
            e1 = fe[f, 0]
            T1 = xo.make_trafo(pts[els[e1]], els[e1])
            fv = ents[f]
            mat, _ = xo.cut_facet_elmat(T1, None, 1, fv, phi[fv], 1 - dt, 2, [(VAL, 0, VAL, 0, 1.0, 0)])
            other += mat.sum()
        assert abs(total + other - full) < 1e-12 * full
        assert 0 < total < full


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (3, 2), (2, 3)])
def test_element_boundary_divergence_theorem(dim, p):
    """one-sided integrals over ALL facets of an uncut element: sum_j M[i][j] c_j with M = int_{dT} (nF.grad N_i) N_j and c the
    coefficients of the constant 1 equals int_T laplace(N_i) — zero for the vertex hats, -4 |T| grad(lam_a).grad(lam_b) for the
    P2 edge functions -2 lam_a lam_b"""
    m = meshes.MakeStructured2DMesh(nx=2, ny=2) if dim == 2 else meshes.MakeStructured3DMesh(nx=2)
    pts, els = mesh_np(m)
    ents, el_f, fe = facet_tables(m)
    phi = -np.ones(len(pts))
    nd = xo.h1_ndof(dim, p)
    one = np.zeros(nd); one[:dim + 1] = 1.0       # hierarchical basis: the constant is the sum of the vertex hats
    for e in (0, len(els) - 1):
        T = xo.make_trafo(pts[els[e]], els[e])
        M = np.zeros((nd, nd))
        for k in range(dim + 1):
            fv = ents[el_f[e, k]]
            mat, _ = xo.cut_facet_elmat(T, None, p, fv, phi[fv], NEG, 2 * p, [(VAL, 0, DN, 0, 1.0, 0)])
            M += mat
        lap = M @ one
        assert np.abs(lap[:dim + 1]).max() < 1e-12
        if p == 2:
            X = pts[els[e]]
            A = np.hstack([X, np.ones((dim + 1, 1))])
            G = np.linalg.inv(A)[:dim, :].T              # rows: grad lam_k
            vol = abs(np.linalg.det(A)) / (2 if dim == 2 else 6)
            edges = [(2, 0), (1, 2), (0, 1)] if dim == 2 else [(3, 0), (3, 1), (3, 2), (0, 1), (0, 2), (1, 2)]
            for i, (a, b) in enumerate(edges):
                assert abs(lap[dim + 1 + i] - (-4 * vol * G[a] @ G[b])) < 1e-12


@pytest.mark.parametrize("dim,p", [(2, 2), (3, 2), (2, 3)])
def test_jumps_of_a_global_polynomial_vanish(dim, p):
    """u = a global polynomial of degree p in the FE space: value jump and normal-derivative jump across a cut facet vanish,
    so the jump forms applied to its coefficient vectors give zero"""
    m = meshes.MakeStructured2DMesh(nx=2, ny=2) if dim == 2 else meshes.MakeStructured3DMesh(nx=2)
    pts, els = mesh_np(m)
    ents, el_f, fe = facet_tables(m)
    phi = interpolate_to_p1(lambda *x: x[0] + 0.3 * x[1] - 0.55, pts)
    nd = xo.h1_ndof(dim, p)
    poly = lambda X: 1 + X[..., 0] - 2 * X[..., 1] + 0.5 * X[..., 0] * X[..., 1] + X[..., -1] ** 2

    def coefs(e):   # element-local coefficients of the polynomial: solve at sample points
        rng = np.random.default_rng(7)
        xh = rng.random((3 * nd, dim)) / dim
        T = xo.make_trafo(pts[els[e]], els[e])
        B = np.array([xo.h1_shapes(dim, p, els[e], x)[0] for x in xh])
        X = np.array([xo.trafo_eval(T, x)[1] for x in xh])
        return np.linalg.lstsq(B, poly(X), rcond=None)[0]

    jj = [(VAL, 0, VAL, 0, 1, 0), (VAL, 1, VAL, 0, -1, 0), (VAL, 0, VAL, 1, -1, 0), (VAL, 1, VAL, 1, 1, 0)]
    dd = [(DN, 0, DN, 0, 1, 1), (DN, 1, DN, 0, -1, 1), (DN, 0, DN, 1, -1, 1), (DN, 1, DN, 1, 1, 1)]
    checked = 0
    for f in range(len(ents)):
        if fe[f, 1] < 0:
            continue
        e1, e2 = fe[f]
        fv = ents[f]
        if not (phi[fv].min() < 0 < phi[fv].max()):
            continue
        T1, T2 = xo.make_trafo(pts[els[e1]], els[e1]), xo.make_trafo(pts[els[e2]], els[e2])
        c = np.concatenate([coefs(e1), coefs(e2)])
        for terms in (jj, dd):
            M, npt = xo.cut_facet_elmat(T1, T2, p, fv, phi[fv], NEG, 2 * p, terms)
            assert npt > 0 and np.abs(M).max() > 0
            assert np.abs(M @ c).max() < 1e-10 * np.abs(M).max() * np.abs(c).max()
        checked += 1
    assert checked >= 2


def test_dg_forms_expand_into_the_term_list():
    mesh = xf.MakeStructured2DMesh(nx=2, ny=2)
    V = xf.H1(mesh, order=2)
    u, v = V.TnT()
    h = xf.specialcf.mesh_size
    nF = xf.specialcf.normal(2)
    flux_u = -0.5 * (xf.grad(u) + xf.grad(u.Other())) * nF
    jump_u, jump_v = u - u.Other(), v - v.Other()
    form = 10 / h * jump_u * jump_v + flux_u * jump_v
    assert sorted(form.terms) == sorted([(0, 0, 0, 0, 10.0, -1), (0, 1, 0, 0, -10.0, -1), (0, 0, 0, 1, -10.0, -1), (0, 1, 0, 1, 10.0, -1),
                                         (1, 0, 0, 0, -0.5, 0), (1, 0, 0, 1, 0.5, 0), (1, 1, 0, 0, -0.5, 0), (1, 1, 0, 1, 0.5, 0)])
    d = xf.dn(u, 1) - xf.dn(u.Other(), 1)
    assert d.items == [(1, 0, 1.0, 0), (1, 1, -1.0, 0)]
    facets = xf.BitArray(mesh.nfacet, "cpu")
    lset = xf.GridFunction(mesh)
    integ = form * xf.dCut(lset, xf.NEG, skeleton=True, definedonelements=facets)
    assert integ.dx.skeleton and len(integ.form.terms) == 8
    with pytest.raises(xf.XgpuError):
        xf.grad(u.Other()) * xf.grad(v)                       # a gradient of .Other() needs the facet normal
    with pytest.raises(xf.XgpuError):
        (xf.grad(u) + xf.grad(u.Other())) * jump_v             # ... before it meets the test function
    with pytest.raises(xf.XgpuError):
        form * xf.dCut(lset, xf.NEG, definedonelements=facets)  # facet integrand on a volume symbol
    with pytest.raises(xf.XgpuError):
        xf.dCut(lset, xf.IF, skeleton=True, definedonelements=facets)     # codim-2 case: not provided
    with pytest.raises(xf.XgpuError):
        xf.dn(u, 2)


# ---- dFacetPatch(downscale = s): ScaleAndMapIntegrationRule (symboliccutbfi.cpp:1849-1885) pinned by closed forms
@pytest.mark.parametrize("dim,n", [(1, 5), (2, 3), (3, 2)])
@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("s", [1.0, 0.5, 0.2])
def test_downscaled_patch_rule_in_closed_form(dim, n, p, s):
    """the scaled rule of an element is the rule of its image under the homothety (centre: midpoint of the common facet,
    ratio s) with weights s^2 |T| w_q.  With u = a polynomial on element 1 and 0 on element 2 the jump is that polynomial
    on the whole patch: u = 1 gives s^2 (|T1| + |T2|), u = x gives s^(2 - D) times the second moments of the two images"""
    import math
    m = {1: lambda: meshes.MakeStructured1DMesh(n), 2: lambda: meshes.MakeStructured2DMesh(nx=n, ny=n + 1),
         3: lambda: meshes.MakeStructured3DMesh(nx=n)}[dim]()
    pts, els = mesh_np(m)
    rng = np.random.default_rng(dim * 10 + p)
    pts = pts + 0.15 / n * rng.random(pts.shape)
    ents, _, fe = facet_tables(m)
    inner = np.nonzero(fe[:, 1] >= 0)[0].astype(np.int32)
    M, nfail = xo.batch_facet_patch(dim, p, inner, fe, els, pts, -1, 1.0, 0, downscale=s)
    assert nfail == 0
    nv, nd = dim + 1, M.shape[1] // 2

    def vol(X):
        return abs(np.linalg.det(X[:dim] - X[dim])) / math.factorial(dim)

    def moment_x2(X):   # int_T x^2 = |T| / ((D+1)(D+2)) (sum x_i^2 + (sum x_i)^2)
        x = X[:, 0]
        return vol(X) / ((dim + 1) * (dim + 2)) * ((x ** 2).sum() + x.sum() ** 2)

    for k, f in enumerate(inner):
        e1, e2 = fe[f]
        X1, X2 = pts[els[e1]], pts[els[e2]]
        mid = pts[np.atleast_1d(ents[f])].reshape(-1, dim).mean(axis=0)
        one = np.zeros(2 * nd); one[:nv] = 1.0
        assert abs(one @ M[k] @ one - s * s * (vol(X1) + vol(X2))) < 1e-13
        cx = np.zeros(2 * nd); cx[:nv] = X1[:, 0]
        want = s ** (2 - dim) * (moment_x2(mid + s * (X1 - mid)) + moment_x2(mid + s * (X2 - mid)))
        assert abs(cx @ M[k] @ cx - want) < 1e-13 * max(1.0, abs(want))
        # the other side: u = x on element 2, 0 on element 1 -> the same moments
        cy = np.zeros(2 * nd); cy[nd:nd + nv] = X2[:, 0]
        assert abs(cy @ M[k] @ cy - want) < 1e-13 * max(1.0, abs(want))
    with pytest.raises(ValueError):
        xo.batch_facet_patch(dim, p, inner, fe, els, pts, -1, 1.0, 0, downscale=1.5)
    with pytest.raises(ValueError):
        xo.batch_facet_patch(dim, p, inner, fe, els, pts, -1, 1.0, 0, downscale=0.0)
