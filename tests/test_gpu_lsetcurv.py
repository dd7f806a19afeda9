"""Row f3 — the isoparametric deformation producer on the GPU (csrc/lsetcurv.cu behind ngsxfem_b200/lsetcurv.py):
parity with the oracle (oracle/xo_shift.c) entry by entry, and the reference's own known answer
tests/pytests/test_intcurved.py:17-91 assembled end to end on the device (LevelSetMeshAdaptation -> Integrate on the
deformed mesh)."""
from math import log, pi

import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from oracle import xoracle as xo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NEG, POS, IF = xf.NEG, xf.POS, xf.IF


def circle(x, y):
    return torch.sqrt(x * x + y * y) - 0.5


def sphere(x, y, z):
    return torch.sqrt((x - 0.5) ** 2 + (y - 0.48) ** 2 + (z - 0.52) ** 2) - 0.33


def oracle_adaptation(mesh, V, fn, order, threshold, qn_mode, lower=0.0, upper=0.0):
    dim = mesh.dim
    pts, els, eld = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy(), V.el_dofs.cpu().numpy()
    xq = xo.rule_points_phys(dim, 2 * order, els, pts)
    f = fn(*[torch.from_numpy(xq[..., c]) for c in range(dim)]).numpy()
    lset_ho = xo.set_from_points(dim, order, 2 * order, els, eld, V.ndof, f)
    lset_p1 = lset_ho[:mesh.nv].copy()
    lset_p1[np.abs(lset_p1) < 1e-14] = 1e-14
    qn = xo.set_gradient(dim, order, els, pts, eld, V.ndof, lset_ho) if qn_mode == 2 else None
    deform, nfail = xo.project_shift(dim, order, els, pts, eld, V.ndof, lset_ho, lset_p1, qn_mode=qn_mode, qn=qn,
                                     lower=lower, upper=upper, threshold=threshold)
    return lset_ho, lset_p1, qn, deform, nfail


@pytest.mark.parametrize("dim,n,order,disc", [(2, 12, 2, True), (2, 9, 3, True), (2, 10, 2, False), (2, 7, 4, False),
                                              (3, 5, 2, True), (3, 4, 3, False), (2, 8, 1, True)])
def test_calc_deformation_matches_oracle(dim, n, order, disc):
    if dim == 2:
        mesh = xf.MakeStructured2DMesh(nx=n, ny=n + 1, mapping=lambda x, y: (2 * x - 1, 2 * y - 1), device=DEV)
        fn = circle
    else:
        mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
        fn = sphere
    ad = xf.LevelSetMeshAdaptation(mesh, order=order, threshold=0.2, discontinuous_qn=disc)
    d = ad.CalcDeformation(fn)
    lset_ho, lset_p1, qn, deform, nfail = oracle_adaptation(mesh, ad.v_ho, fn, order, 0.2, 1 if disc else 2)
    assert ad.nfail == nfail      # (3D, order 3 on the 4^3 mesh: one search does not converge in 20 steps — in both)
    # element-local Cholesky solves in another summation order: 1e-10 of the largest coefficient (orders 3, 4: the mass matrix
    # of the hierarchical basis has a condition number of 1e3-1e4)
    assert np.abs(ad.lset_ho.cpu().numpy() - lset_ho).max() <= 1e-10 * np.abs(lset_ho).max()
    assert np.abs(ad.lset_p1.vec.cpu().numpy() - lset_p1).max() <= 1e-10 * np.abs(lset_p1).max()
    if not disc:
        assert np.abs(ad.qn.cpu().numpy() - qn).max() <= 1e-9 * np.abs(qn).max()
    scale = np.abs(deform).max()
    assert scale > 0 or order == 1
    assert np.abs(d.vals.cpu().numpy() - deform).max() <= 1e-8 * max(scale, 1e-30) + 1e-14
    # run to run reproducible
    d0 = d.vals.clone()
    ad.CalcDeformation(fn)
    assert torch.equal(ad.deform.vals, d0) and torch.equal(ad.deform_last.vals, d0)


def test_band_elements_and_marker_argument():
    mesh = xf.MakeStructured2DMesh(nx=16, ny=16, mapping=lambda x, y: (2 * x - 1, 2 * y - 1), device=DEV)
    ad = xf.LevelSetMeshAdaptation(mesh, order=2, threshold=0.2, discontinuous_qn=True, lset_lower_bound=-0.1, lset_upper_bound=0.15)
    ad.CalcDeformation(circle)
    l1 = ad.lset_p1.vec.cpu().numpy()
    els = mesh.elements.cpu().numpy()
    bits = ops.band_elements(2, mesh.elements, ad.lset_p1.vec, -0.1, 0.15)
    got = ops.bitset_to_list(bits, mesh.ne).cpu().numpy()
    want = np.array([e for e in range(mesh.ne) if xo.element_in_band(l1[els[e]], -0.1, 0.15)])
    assert np.array_equal(got, want)
    # the wider band moves more dofs than the cut elements alone; ba= restricts to a marked set
    _, _, _, deform, _ = oracle_adaptation(mesh, ad.v_ho, circle, 2, 0.2, 1, lower=-0.1, upper=0.15)
    assert np.abs(ad.deform.vals.cpu().numpy() - deform).max() <= 1e-9 * np.abs(deform).max()
    ci = xf.CutInfo(mesh, ad.lset_p1)
    ad.CalcDeformation(circle, ba=ci.GetElementsOfType(IF))
    _, _, _, deform0, _ = oracle_adaptation(mesh, ad.v_ho, circle, 2, 0.2, 1)
    assert np.abs(ad.deform.vals.cpu().numpy() - deform0).max() <= 1e-9 * np.abs(deform0).max()


@pytest.mark.parametrize("order", [1, 2, 3])
def test_intcurved_known_answers_on_the_device(order):
    """tests/pytests/test_intcurved.py:17-91 (triangles): error < 1e-5 on the finest mesh, average order > order + 0.75"""
    ref = {NEG: 0.25 * pi, POS: 4.0 - 0.25 * pi, IF: pi}
    levels = 5 if order > 1 else 7
    err = {k: [] for k in ref}
    for lvl in range(levels):
        n = 4 * 2 ** lvl
        mesh = xf.MakeStructured2DMesh(nx=n, ny=n, mapping=lambda x, y: (2 * x - 1, 2 * y - 1), device=DEV)
        ad = xf.LevelSetMeshAdaptation(mesh, order=order, threshold=0.2, discontinuous_qn=True)
        deformation = ad.CalcDeformation(circle)
        for k in ref:
            mesh.SetDeformation(deformation)
            v = xf.Integrate(levelset_domain={"levelset": ad.lset_p1, "domain_type": k}, cf=1.0, mesh=mesh, order=order)
            mesh.UnsetDeformation()
            err[k].append(abs(v - ref[k]))
    for k in ref:
        eoc = [log(a / b) / log(2) for a, b in zip(err[k][:-1], err[k][1:])]
        assert err[k][-1] < (1e-5 if order > 1 else 1e-4)
        s = 2 if order == 1 else 0
        assert sum(eoc[s:]) / len(eoc[s:]) > order + 0.75, (k, err[k], eoc)


def test_context_manager_and_blending():
    mesh = xf.MakeStructured2DMesh(nx=24, ny=24, mapping=lambda x, y: (2 * x - 1, 2 * y - 1), device=DEV)
    ad = xf.LevelSetMeshAdaptation(mesh, order=2, threshold=0.2, levelset=circle)      # continuous qn, computed right away
    with ad as lset_p1:
        a = xf.Integrate(levelset_domain={"levelset": lset_p1, "domain_type": IF}, cf=1.0, mesh=mesh, order=4)
    b = xf.Integrate(levelset_domain={"levelset": ad.lset_p1, "domain_type": IF}, cf=1.0, mesh=mesh, order=4)
    assert mesh.deformation is None
    assert abs(a - pi) < 1e-4 < abs(b - pi)
    # blending pulls the goal value towards lset_ho away from the interface: the interface itself is still mapped
    for bl in ("quadratic", "quartic"):
        d = ad.CalcDeformation(circle, blending=bl)
        v = xf.Integrate(levelset_domain={"levelset": ad.lset_p1, "domain_type": IF}, cf=1.0, mesh=mesh, order=4, deformation=d)
        assert abs(v - pi) < 5e-4 < abs(b - pi)
    with pytest.raises(xf.XgpuError):
        xf.LevelSetMeshAdaptation(mesh, order=2, boundary_tangential=True)


def test_deformed_cut_assembly_on_a_produced_deformation():
    """config 3 of BASELINE.json names "LevelSetMeshAdaptation deformation": the stiffness + mass form on the cut elements of a
    sphere, assembled on the deformation produced here; 1^T A 1 is the deformed volume + area"""
    n, order = 16, 2
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    ad = xf.LevelSetMeshAdaptation(mesh, order=order, threshold=0.1, discontinuous_qn=True)
    d = ad.CalcDeformation(sphere)
    vol = xf.Integrate(levelset_domain={"levelset": ad.lset_p1, "domain_type": NEG}, cf=1.0, mesh=mesh, order=4, deformation=d)
    area = xf.Integrate(levelset_domain={"levelset": ad.lset_p1, "domain_type": IF}, cf=1.0, mesh=mesh, order=4, deformation=d)
    assert abs(vol - 4 / 3 * pi * 0.33 ** 3) < 5e-5 and abs(area - 4 * pi * 0.33 ** 2) < 3e-4
    vol0 = xf.Integrate(levelset_domain={"levelset": ad.lset_p1, "domain_type": NEG}, cf=1.0, mesh=mesh, order=4)
    assert abs(vol0 - 4 / 3 * pi * 0.33 ** 3) > 1e-3
    # config-3 shape on the PRODUCED deformation: stiffness + Nitsche-type mass on the interface + ghost penalty, order 2
    ci = xf.CutInfo(mesh, ad.lset_p1)
    hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
    gpf = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif, use_and=True)
    V = xf.Restrict(xf.H1(mesh, order=order, dgjumps=True), hasneg)
    u, v = V.TnT()
    h = xf.specialcf.mesh_size
    a = xf.RestrictedBilinearForm(V, element_restriction=hasneg, facet_restriction=gpf)
    a += (xf.grad(u) * xf.grad(v) + u * v) * xf.dCut(ad.lset_p1, NEG, definedonelements=hasneg, deformation=d)
    a += u * v * xf.dCut(ad.lset_p1, IF, definedonelements=hasif, deformation=d)
    a += (0.1 / h ** 2) * (u - u.Other()) * (v - v.Other()) * xf.dFacetPatch(definedonelements=gpf, deformation=d)
    a.Assemble()
    # u = 1 in the hierarchical basis: vertex dofs 1, edge dofs 0; stiffness and jump terms vanish on it:
    # 1^T A 1 = |Omega^-_h| + |Gamma_h| on the deformed geometry
    rows = V.rows(hasneg.indices())
    one = torch.zeros(V.ndof, dtype=torch.float64, device=DEV)
    one[rows[:, :4].reshape(-1).long()] = 1.0
    y = ops.csr_spmv(a.mat.rowptr, a.mat.col, a.mat.vals, one)
    assert abs(float((one * y).sum()) - (vol + area)) < 1e-10 * (vol + area)
