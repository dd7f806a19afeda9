"""BASELINE.json configs 3, 4 and 5 at (single-GPU) scale, checked through size-independent properties — the oracle
would need minutes at these sizes (config 2 at full size: tests/test_pipeline.py::test_full_size_invariants).

config 3: order-4 isoparametric stiffness + ghost penalty on a torus, 384 000 tets: constants are in the kernel of both
          operators on ANY deformed mesh (grad 1 = 0, jump of 1 = 0): |A 1| = 0 up to rounding; A symmetric.
config 4: multi-level-set domain, order 3, 1.57 M tets: 1^T M 1 = |domain| (order-0 multi-level-set rule), K 1 = 0.
config 5: space-time cut rules P1 x P2 on 6 367 248 tets (the size BASELINE names): NEG + POS measure = 1 exactly,
          NEG measure = ball volume up to the P1/P2 interface error."""
import math

import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEG, POS, IF = 0, 1, 2




def test_config3_isoparametric_ghost_penalty_kernel_of_constants():
    n, k = 40, 4
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    # the deformation is PRODUCED (LevelSetMeshAdaptation, row f3), as BASELINE.json's config 3 names it
    torus = lambda x, y, z: (torch.sqrt((x - .5) ** 2 + (y - .5) ** 2) - 0.28) ** 2 + (z - .5) ** 2 - 0.12 ** 2   # noqa: E731
    ad = xf.LevelSetMeshAdaptation(mesh, order=k, threshold=0.1, discontinuous_qn=True)
    defo = ad.CalcDeformation(torus)
    lset = ad.lset_p1
    assert ad.nfail == 0 and float(defo.vals.abs().max()) > 0
    exact = 2 * math.pi ** 2 * 0.28 * 0.12 ** 2
    vol = xf.Integrate(levelset_domain={"levelset": lset, "domain_type": xf.NEG}, cf=1.0, mesh=mesh, order=k, deformation=defo)
    vol0 = xf.Integrate(levelset_domain={"levelset": lset, "domain_type": xf.NEG}, cf=1.0, mesh=mesh, order=k)
    print("config 3: torus volume error curved", vol - exact, "uncurved", vol0 - exact)
    assert abs(vol - exact) < 1e-5 and abs(vol0 - exact) > 1e-4
    V0 = xf.H1(mesh, order=k, dgjumps=True)
    ci = xf.CutInfo(mesh, lset)
    hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
    gpf = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif, use_and=True)
    V = xf.Restrict(V0, hasneg)
    u, v = V.TnT()
    h = xf.specialcf.mesh_size
    a = xf.RestrictedBilinearForm(V, element_restriction=hasneg, facet_restriction=gpf)
    a += (xf.grad(u) * xf.grad(v)) * xf.dCut(lset, xf.NEG, definedonelements=hasneg, deformation=defo)
    a += (0.1 / h ** 2) * (u - u.Other()) * (v - v.Other()) * xf.dFacetPatch(definedonelements=gpf, deformation=defo)
    a.Assemble()
    assert hasif.NumSet() > 10000 and gpf.NumSet() > 20000
    rows = V.rows(hasneg.indices())
    one = torch.zeros(V.ndof, dtype=torch.float64, device=DEV)
    one[rows[:, :4].reshape(-1).long()] = 1.0          # u = 1: vertex dofs 1, all higher-order dofs 0
    y = a.mat * one
    scale = float(a.mat.vals.abs().max())
    assert float(y.abs().max()) < 1e-10 * scale
    x = torch.rand(V.ndof, generator=torch.Generator().manual_seed(1), dtype=torch.float64).to(DEV)
    z = torch.rand(V.ndof, generator=torch.Generator().manual_seed(2), dtype=torch.float64).to(DEV)
    s1, s2 = float((z * (a.mat * x)).sum()), float((x * (a.mat * z)).sum())
    assert abs(s1 - s2) < 1e-11 * abs(s1)              # symmetric operator


def test_config4_mlset_order3_volume_and_constants():
    n = 64
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    x, y, z = (mesh.points[:, i] for i in range(3))
    fns = [x + 0.31 * y - 0.83, -0.27 * x + y + 0.11 * z - 0.79, 0.13 * x - 0.21 * y + z - 0.81,
           torch.sqrt((x - .4) ** 2 + (y - .45) ** 2 + (z - .4) ** 2) - 0.37]
    lsets = [xf.GridFunction(mesh, torch.where(f.abs() < 1e-14, torch.full_like(f, 1e-14), f).contiguous()) for f in fns]
    dts = (NEG, NEG, NEG, NEG)
    active = xf.MultiLevelsetCutInfo(mesh, lsets).GetElementsWithContribution(dts)
    V = xf.Restrict(xf.H1(mesh, order=3), active)
    u, v = V.TnT()

    def assemble(form):
        a = xf.RestrictedBilinearForm(V, element_restriction=active)
        a += form * xf.dCut(lsets, dts, definedonelements=active)
        return a.Assemble()

    m, kk = assemble(u * v), assemble(xf.grad(u) * xf.grad(v))
    rows = V.rows(active.indices())
    one = torch.zeros(V.ndof, dtype=torch.float64, device=DEV)
    one[rows[:, :4].reshape(-1).long()] = 1.0
    vol = xf.Integrate({"levelset": lsets, "domain_type": dts}, mesh=mesh, cf=1.0, order=0)
    tot = float((one * (m.mat * one)).sum())
    assert active.NumSet() > 300000
    assert abs(tot - vol) < 1e-11 * vol
    assert float((kk.mat * one).abs().max()) < 1e-10 * float(kk.mat.vals.abs().max())


def test_config5_spacetime_rules_full_size():
    n, r = 102, 0.3
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    assert mesh.ne == 6367248
    phis = []
    for tn in (0.0, 0.5, 1.0):
        cy = 0.5 + 0.1 * math.sin(2 * math.pi * tn)
        f = torch.sqrt((mesh.points[:, 0] - .5) ** 2 + (mesh.points[:, 1] - cy) ** 2 + (mesh.points[:, 2] - .5) ** 2) - r
        phis.append(torch.where(f.abs() < 1e-14, torch.full_like(f, 1e-14), f))
    phi_st = torch.stack(phis).contiguous()
    res = {}
    for dt in (NEG, POS):
        off, q, tv, w = ops.cutquad_st(3, None, mesh.elements, mesh.points, phi_st, 2, dt, 4, 2)
        assert ops.default_ctx().status()[0] == 0
        assert float(tv.min()) >= 0.0 and float(tv.max()) <= 1.0
        res[dt] = ops.integrate_monomial(3, None, mesh.elements, mesh.points, off, q, w, (0, 0, 0))
        del off, q, tv, w
    assert abs(res[NEG] + res[POS] - 1.0) < 1e-11
    ball = 4.0 / 3.0 * math.pi * r ** 3
    assert abs(res[NEG] - ball) < 2e-3 * ball


# ------------------------------------------------------------------ sampled oracle parity AT SCALE
# The invariants above hold for any consistent kernel; these compare the CUDA kernels with the oracle element by element on a
# random SAMPLE of the cut elements of the full-size inputs (element numbers in the millions, the produced order-4 deformation,
# the real dof / vertex numbering of the big mesh) — the oracle does a few hundred elements in seconds.
from oracle import xoracle as xo  # noqa: E402


def _sample(lst, k, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(lst.numel(), generator=g)[:k].sort().values
    return lst[idx.to(lst.device)].contiguous()


def test_config3_sampled_element_and_facet_matrices_match_oracle_at_scale():
    n, k = 40, 4
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    torus = lambda x, y, z: (torch.sqrt((x - .5) ** 2 + (y - .5) ** 2) - 0.28) ** 2 + (z - .5) ** 2 - 0.12 ** 2   # noqa: E731
    ad = xf.LevelSetMeshAdaptation(mesh, order=k, threshold=0.1, discontinuous_qn=True)
    defo = ad.CalcDeformation(torus)
    phi = ad.lset_p1.vec
    ci = xf.CutInfo(mesh, ad.lset_p1)
    hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
    l_s = _sample(hasif.indices(), 150, 11)
    ctx = ops.default_ctx()
    terms = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 0.5, 0)]
    with ctx.deformation(defo):
        off, q, w = ops.cutquad(3, l_s, mesh.elements, mesh.points, phi, NEG, 2 * k)
        m = ops.elmat_rule(3, k, l_s, mesh.elements, mesh.points, phi, off, q, w, terms)
        off2, q2, w2 = ops.cutquad(3, l_s, mesh.elements, mesh.points, phi, IF, 2 * k)
        mi = ops.elmat_rule(3, k, l_s, mesh.elements, mesh.points, phi, off2, q2, w2, [(ops.OP_ID, ops.OP_ID, 4.0, -1)])
    pts, els = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy()
    ls_np = l_s.cpu().numpy()
    # the oracle takes the deformation as per-element coefficient blocks: hand it the sampled elements as a small mesh
    sub_els = els[ls_np]
    loc = defo.vals[defo.el_dofs[l_s.long()].long()].cpu().numpy()
    mo = xo.batch_cut_elmats(3, k, None, sub_els, pts, phi.cpu().numpy(), xo.NEG, 2 * k, terms, def_order=k, deform=loc)
    moi = xo.batch_cut_elmats(3, k, None, sub_els, pts, phi.cpu().numpy(), xo.IF, 2 * k, [(0, 0, 4.0, -1)], def_order=k, deform=loc)
    assert np.linalg.norm(m.cpu().numpy() - mo) <= 1e-11 * np.linalg.norm(mo)
    assert np.linalg.norm(mi.cpu().numpy() - moi) <= 1e-11 * np.linalg.norm(moi)


def test_config4_sampled_mlset_element_matrices_match_oracle_at_scale():
    n, p = 64, 3
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    x, y, z = (mesh.points[:, i] for i in range(3))
    fns = [x + 0.31 * y - 0.83, -0.27 * x + y + 0.11 * z - 0.79, 0.13 * x - 0.21 * y + z - 0.81,
           torch.sqrt((x - .4) ** 2 + (y - .45) ** 2 + (z - .4) ** 2) - 0.37]
    lsets = [xf.GridFunction(mesh, torch.where(f.abs() < 1e-14, torch.full_like(f, 1e-14), f).contiguous()) for f in fns]
    dts = [(NEG, NEG, NEG, NEG)]
    mlci = xf.MultiLevelsetCutInfo(mesh, lsets)
    active = mlci.GetElementsWithContribution(dts[0])
    inner = mlci.GetElementsOfType(dts[0])                           # uncut w.r.t. every level set
    cut = xf.BitArray(mesh.ne, words=ops.bitset_op(ops.BIT_ANDNOT, active.bits, inner.bits, mesh.ne))
    l_s = _sample(cut.indices(), 200, 5)
    terms = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 2.5, 0)]
    phis = torch.stack([g.vec for g in lsets]).contiguous()
    off, q, w = ops.cutquad_ml(3, l_s, mesh.elements, mesh.points, phis, dts, 2 * p)
    assert w.numel() > 0
    m = ops.elmat_rule(3, p, l_s, mesh.elements, mesh.points, phis[0].contiguous(), off, q, w, terms)
    pts, els = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy()
    ph = [g.vec.cpu().numpy() for g in lsets]
    nd = xo.h1_ndof(3, p)
    mo = np.zeros((l_s.numel(), nd, nd))
    for kk, el in enumerate(l_s.cpu().numpy()):
        T = xo.make_trafo(pts[els[el]], els[el])
        ls = np.stack([a[els[el]] for a in ph], axis=1)
        qq, ww = xo.mlset_cut_rule(T, ls, dts, 2 * p)
        if len(ww):
            xo.elmat_from_rule(T, p, ls[:, 0].copy(), qq, ww, terms, out=mo[kk])
    assert np.linalg.norm(mo) > 0 and np.linalg.norm(m.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)


def test_config5_sampled_spacetime_element_matrices_match_oracle_at_scale():
    n, r, p, kt = 102, 0.3, 1, 2
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    phis = []
    for tn in (0.0, 0.5, 1.0):
        cy = 0.5 + 0.1 * math.sin(2 * math.pi * tn)
        f = torch.sqrt((mesh.points[:, 0] - .5) ** 2 + (mesh.points[:, 1] - cy) ** 2 + (mesh.points[:, 2] - .5) ** 2) - r
        phis.append(torch.where(f.abs() < 1e-14, torch.full_like(f, 1e-14), f))
    phi_st = torch.stack(phis).contiguous()
    bn, bp, bi = ops.classify_st(3, mesh.elements, phi_st, 2, 2, want_ratio=False)[:3]
    l_s = _sample(ops.bitset_to_list(bi, mesh.ne), 200, 9)
    terms = [(0, 0, 1.0, 0), (1, 1, 0.5, 0), (3, 0, 1.0, 0)]       # u v + grad u . grad v + dt(u) v
    off, q, tv, w = ops.cutquad_st(3, l_s, mesh.elements, mesh.points, phi_st, 2, NEG, 2 * kt, 2 * p)
    m = ops.elmat_st(3, p, kt, l_s, mesh.elements, mesh.points, off, q, tv, w, terms)
    P, E = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy()
    offn, qn, tvn, wn = off.cpu().numpy(), q.cpu().numpy(), tv.cpu().numpy(), w.cpu().numpy()
    L = phi_st.cpu().numpy()
    nd = ops.h1_ndof(3, p) * (kt + 1)
    mo = np.zeros((l_s.numel(), nd, nd))
    for kk, el in enumerate(l_s.cpu().numpy()):
        T = xo.make_trafo(P[E[el]], E[el])
        # rule AND matrix from the oracle (not the GPU rule fed back)
        qo, to, wo = xo.spacetime_cut_rule(T, L[:, E[el]], 2, xo.NEG, 2 * kt, 2 * p)
        a, b = offn[kk], offn[kk + 1]
        assert b - a == len(wo)
        assert np.allclose(qn[a:b], qo, rtol=0, atol=1e-13) and np.allclose(wn[a:b], wo, rtol=1e-12, atol=1e-18)
        xo.elmat_st_from_rule(T, p, kt, qo, to, wo, terms, out=mo[kk])
    assert np.linalg.norm(mo) > 0 and np.linalg.norm(m.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)
