"""Isoparametric deformation on the GPU path (xgpu_set_deformation; config 3 of BASELINE.json: curved elements +
ghost-penalty facet patches) against the oracle and against the closed-form answers of test_oracle_deformation.py.
Tolerances: interface weights 1e-13, element / facet-patch matrices 1e-12 (Frobenius), integrals 1e-12."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from oracle import xoracle as xo

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEG, POS, IF = 0, 1, 2
A, B, Cc, X = 0.3, -0.25, 0.2, 0.55


def random_deformation(V, scale, seed):
    g = torch.Generator().manual_seed(seed)
    vals = (torch.rand((V.ndof, V.mesh.dim), generator=g, dtype=torch.float64) - 0.5) * 2 * scale
    return xf.Deformation(V, vals.to(DEV))


def sphere(mesh, r=0.37, c=(.45, .5, .55)):
    return xf.InterpolateToP1(lambda *x: torch.sqrt(sum((x[i] - c[i]) ** 2 for i in range(len(x)))) - r, mesh)


def test_closed_form_integrals_on_the_deformed_mesh():
    mesh = xf.MakeStructured3DMesh(nx=5, device=DEV)
    V = xf.H1(mesh, order=2)
    defo = xf.Deformation.interpolate_p2(V, lambda p: torch.stack([A * p[:, 1] ** 2, B * p[:, 2] ** 2, Cc * p[:, 0] ** 2], dim=1))
    lset = xf.GridFunction(mesh, mesh.points[:, 0] - X)
    abc8 = 8 * A * B * Cc
    vol = xf.Integrate({"levelset": lset, "domain_type": NEG}, mesh=mesh, cf=1.0, order=3, deformation=defo)
    assert abs(vol - (X + abc8 * X ** 2 / 8)) < 1e-13
    mx = xf.Integrate({"levelset": lset, "domain_type": NEG}, mesh=mesh, cf=xf.x, order=5, deformation=defo)
    assert abs(mx - (X ** 2 / 2 + A * X / 3 + abc8 * (X ** 3 / 12 + A * X ** 2 / 16))) < 1e-13
    g, wg = np.polynomial.legendre.leggauss(40)
    y = 0.5 * (g + 1); w = 0.5 * wg
    Y, Z = np.meshgrid(y, y, indexing="ij")
    exact = float((np.sqrt(1 + 4 * A ** 2 * Y ** 2 + 16 * A ** 2 * B ** 2 * Y ** 2 * Z ** 2) * np.outer(w, w)).sum())
    area = xf.Integrate({"levelset": lset, "domain_type": IF}, mesh=mesh, cf=1.0, order=8, deformation=defo)
    assert abs(area - exact) < 1e-9
    # without the deformation the plain answers come back (the state is unset after the call)
    assert abs(xf.Integrate({"levelset": lset, "domain_type": NEG}, mesh=mesh, cf=1.0, order=3) - X) < 1e-13


@pytest.mark.parametrize("dim,n,pdef", [(3, 4, 2), (3, 3, 3), (2, 7, 2), (2, 5, 4)])
def test_interface_weights_on_deformed_elements_match_oracle(dim, n, pdef):
    mesh = (xf.MakeStructured3DMesh if dim == 3 else xf.MakeStructured2DMesh)(nx=n, **({} if dim == 3 else {"ny": n}), device=DEV)
    Vd = xf.H1(mesh, order=pdef)
    defo = random_deformation(Vd, 0.04 / n, 5)
    lset = sphere(mesh) if dim == 3 else xf.InterpolateToP1(lambda x, y: torch.sqrt((x - .5) ** 2 + (y - .45) ** 2) - 0.31, mesh)
    ctx = ops.default_ctx()
    with ctx.deformation(defo):
        off, q, w = ops.cutquad(dim, None, mesh.elements, mesh.points, lset.vec, IF, 4)
    pts, els, phi = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy(), lset.vec.cpu().numpy()
    loc = defo.local().cpu().numpy()
    off = off.cpu().numpy(); q = q.cpu().numpy(); w = w.cpu().numpy()
    ncut = 0
    for el in range(els.shape[0]):
        T = xo.make_trafo(pts[els[el]], els[el], def_order=pdef, deform=loc[el])
        po, wo, _ = xo.straight_cut_rule(T, phi[els[el]], IF, 4)
        assert off[el + 1] - off[el] == len(wo)
        if len(wo):
            ncut += 1
            assert np.allclose(q[off[el]:off[el + 1]], po, rtol=1e-13, atol=1e-15)
            assert np.allclose(w[off[el]:off[el + 1]], wo, rtol=1e-13, atol=1e-13 * np.abs(wo).max())
    assert ncut > 10


@pytest.mark.parametrize("p,pdef,n,dt,dmma", [(2, 2, 4, NEG, 0), (2, 2, 4, IF, 0), (3, 3, 3, NEG, 0), (3, 3, 3, NEG, -1),
                                              (4, 4, 2, POS, 0), (4, 4, 2, POS, -1), (3, 2, 3, IF, 0), (1, 2, 4, NEG, 0)])
def test_element_matrices_on_deformed_elements_match_oracle(p, pdef, n, dt, dmma):
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    Vd = xf.H1(mesh, order=pdef)
    defo = random_deformation(Vd, 0.05 / n, 11)
    lset = sphere(mesh)
    ctx = ops.default_ctx()
    ctx.set_option("elmat_dmma", dmma)
    terms = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 0.7, -1)]
    try:
        with ctx.deformation(defo):
            off, q, w = ops.cutquad(3, None, mesh.elements, mesh.points, lset.vec, dt, 2 * p)
            m = ops.elmat_rule(3, p, None, mesh.elements, mesh.points, lset.vec, off, q, w, terms)
    finally:
        ctx.set_option("elmat_dmma", 0)
    pts, els, phi = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy(), lset.vec.cpu().numpy()
    mo = xo.batch_cut_elmats(3, p, None, els, pts, phi, dt, 2 * p, terms, def_order=pdef, deform=defo.local().cpu().numpy())
    assert np.linalg.norm(mo) > 0
    assert np.linalg.norm(m.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)
    # general kernel (normal-derivative term) on the same rules
    if p <= 3:
        terms2 = terms + [(ops.OP_DN, ops.OP_DN, 0.3, 1)]
        with ctx.deformation(defo):
            m2 = ops.elmat_rule(3, p, None, mesh.elements, mesh.points, lset.vec, off, q, w, terms2)
        mo2 = xo.batch_cut_elmats(3, p, None, els, pts, phi, dt, 2 * p, terms2, def_order=pdef, deform=defo.local().cpu().numpy())
        assert np.linalg.norm(m2.cpu().numpy() - mo2) <= 1e-12 * np.linalg.norm(mo2)


@pytest.mark.parametrize("p", [1, 2])
def test_uncut_elements_on_a_deformed_mesh_leave_the_affine_fast_path(p):
    mesh = xf.MakeStructured3DMesh(nx=3, device=DEV)
    defo = random_deformation(xf.H1(mesh, order=2), 0.02, 3)
    terms = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 1.0, 0)]
    ctx = ops.default_ctx()
    phi = torch.full((mesh.nv,), -1.0, dtype=torch.float64, device=DEV)
    with ctx.deformation(defo):
        m = ops.elmat_std(3, p, None, mesh.elements, mesh.points, phi, 2 * p, terms)
    m0 = ops.elmat_std(3, p, None, mesh.elements, mesh.points, phi, 2 * p, terms)
    mo = xo.batch_cut_elmats(3, p, None, mesh.elements.cpu().numpy(), mesh.points.cpu().numpy(), phi.cpu().numpy(), NEG,
                             2 * p, terms, def_order=2, deform=defo.local().cpu().numpy())
    assert np.linalg.norm(m.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)
    assert np.linalg.norm((m - m0).cpu().numpy()) > 1e-4 * np.linalg.norm(mo)


@pytest.mark.parametrize("dim,p,n", [(3, 2, 3), (3, 3, 2), (2, 2, 5), (2, 4, 3)])
def test_facet_patch_on_deformed_elements_matches_oracle(dim, p, n):
    mesh = (xf.MakeStructured3DMesh if dim == 3 else xf.MakeStructured2DMesh)(nx=n, **({} if dim == 3 else {"ny": n}), device=DEV)
    # higher-order bubbles are extrapolated into the neighbour element by the patch mapping: keep them small enough
    # for the Newton iteration of MapPatchIntegrationPoint to converge everywhere (the oracle counts failures)
    defo = random_deformation(xf.H1(mesh, order=p), (0.04 if p <= 2 else 0.004) / n, 17)
    _, _, fe = mesh.facets()
    inner = torch.nonzero(fe[:, 1] >= 0).reshape(-1).to(torch.int32)
    ctx = ops.default_ctx()
    with ctx.deformation(defo):
        m = ops.facetpatch(dim, p, inner, fe, mesh.elements, mesh.points, -1, 0.1, -2)
    mo, nfail = xo.batch_facet_patch(dim, p, inner.cpu().numpy(), fe.cpu().numpy(), mesh.elements.cpu().numpy(),
                                     mesh.points.cpu().numpy(), -1, 0.1, -2, def_order=p, deform=defo.local().cpu().numpy())
    assert nfail == 0
    assert np.linalg.norm(m.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)
    # dFacetPatch(downscale=) on curved elements: the scaled points go through the same Newton map
    with ctx.deformation(defo):
        ms = ops.facetpatch(dim, p, inner, fe, mesh.elements, mesh.points, -1, 0.1, -2, downscale=0.4)
    mos, nfail = xo.batch_facet_patch(dim, p, inner.cpu().numpy(), fe.cpu().numpy(), mesh.elements.cpu().numpy(),
                                      mesh.points.cpu().numpy(), -1, 0.1, -2, def_order=p, deform=defo.local().cpu().numpy(),
                                      downscale=0.4)
    assert nfail == 0
    assert np.linalg.norm(ms.cpu().numpy() - mos) <= 1e-12 * np.linalg.norm(mos)


def test_isoparametric_assembly_with_ghost_penalty_matches_oracle():
    """config-3 shape at test size: curved cut elements (NEG + IF integrators) + facet-patch ghost penalty, order 3"""
    p, n = 3, 3
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    lset = sphere(mesh)
    V0 = xf.H1(mesh, order=p, dgjumps=True)
    defo = random_deformation(xf.H1(mesh, order=p), 0.004 / n, 23)
    ci = xf.CutInfo(mesh, lset)
    hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
    gpf = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif, use_and=True)
    V = xf.Restrict(V0, hasneg)
    u, v = V.TnT()
    h = xf.specialcf.mesh_size
    a = xf.RestrictedBilinearForm(V, element_restriction=hasneg, facet_restriction=gpf)
    a += (xf.grad(u) * xf.grad(v)) * xf.dCut(lset, NEG, definedonelements=hasneg, deformation=defo)
    a += (4.0 / h) * u * v * xf.dCut(lset, IF, definedonelements=hasif, deformation=defo)
    a += (0.1 / h ** 2) * (u - u.Other()) * (v - v.Other()) * xf.dFacetPatch(definedonelements=gpf, deformation=defo)
    a.Assemble()
    pts, els, phi = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy(), lset.vec.cpu().numpy()
    loc = defo.local().cpu().numpy()
    l_act = hasneg.indices().cpu().numpy()
    m = xo.batch_cut_elmats(3, p, l_act, els, pts, phi, NEG, -1, [(1, 1, 1.0, 0)], def_order=p, deform=loc)
    l_if = hasif.indices().cpu().numpy()
    m_if = xo.batch_cut_elmats(3, p, l_if, els, pts, phi, IF, -1, [(0, 0, 4.0, -1)], def_order=p, deform=loc)
    pos = {e: i for i, e in enumerate(l_act)}
    for k, e in enumerate(l_if):
        m[pos[e]] += m_if[k]
    _, _, fe = mesh.facets()
    l_f = gpf.indices()
    l_f = l_f[fe[l_f.long(), 1] >= 0].contiguous()  # boundary facets carry no patch
    fel = fe[l_f.long()]
    l_f = l_f.cpu().numpy()
    mf, nfail = xo.batch_facet_patch(3, p, l_f, fe.cpu().numpy(), els, pts, -1, 0.1, -2, def_order=p, deform=loc)
    assert nfail == 0 and len(l_f) > 0
    t_el = V.rows(hasneg.indices()).cpu().numpy()
    t_f = torch.cat([V.rows(fel[:, 0].contiguous()), V.rows(fel[:, 1].contiguous())], dim=1).cpu().numpy()
    av = a.mat.to_scipy()
    # dense comparison through scipy (two batches of different width)
    import scipy.sparse as sp
    ref = sp.lil_matrix((V.ndof, V.ndof))
    for tab, mats in ((t_el, m), (t_f, mf)):
        for r in range(tab.shape[0]):
            d = tab[r]
            for i in range(len(d)):
                for j in range(len(d)):
                    ref[d[i], d[j]] += mats[r, i, j]
    diff = (av - ref.tocsr())
    assert abs(diff).max() <= 1e-12 * abs(ref).max() * 10
    assert np.sqrt((diff.multiply(diff)).sum()) <= 1e-12 * np.sqrt((ref.multiply(ref)).sum())
