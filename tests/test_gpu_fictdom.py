"""Config 1 of BASELINE.json (demos/fictdom.py restated on a structured triangle mesh): fictitious-domain Poisson with
Nitsche boundary terms on the P1 level-set interface and ghost-penalty facet patches, the WHOLE problem on the accelerated
path (CutInfo, GetFacetsWithNeighborTypes, Restrict, dCut NEG / IF with normal-derivative terms, dFacetPatch, LinearForm,
CSR), solved on the host with scipy.  Basis-independent known answer (the kind of pin the reference's own tests use,
tests/pytests/test_nxfem.py style EOC checks): u = R^2 - |x - c|^2 on the disk, f = -Laplace u = 4.  The geometry error of
the P1 interface is O(h^2), so the L2 error must decay with second order; for order 2 the exact solution is in the
discrete space and the error is the geometry error alone."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf

pytestmark = pytest.mark.gpu
DEV = "cuda"
R, CX, CY = 0.3, 0.5, 0.5


def solve(n, p, gamma=0.1):
    import scipy.sparse.linalg as spla
    mesh = xf.MakeStructured2DMesh(nx=n, ny=n, device=DEV)
    lset = xf.InterpolateToP1(lambda x, y: torch.sqrt((x - CX) ** 2 + (y - CY) ** 2) - R, mesh)
    ci = xf.CutInfo(mesh, lset)
    hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
    ba_facets = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif)
    Vb = xf.H1(mesh, order=p, dgjumps=True)
    Vh = xf.Restrict(Vb, hasneg)
    u, v = Vh.TnT()
    h = xf.specialcf.mesh_size
    nrm = xf.Normalize(xf.grad(lset))
    dx = xf.dCut(lset, xf.NEG, definedonelements=hasneg)
    ds = xf.dCut(lset, xf.IF, definedonelements=hasif)
    dw = xf.dFacetPatch(definedonelements=ba_facets)
    a = xf.RestrictedBilinearForm(Vh, element_restriction=hasneg, facet_restriction=ba_facets)
    a += xf.grad(u) * xf.grad(v) * dx
    a += -1.0 * (xf.grad(u) * nrm) * v * ds
    a += -1.0 * (xf.grad(v) * nrm) * u * ds
    a += (10.0 * p * p / h) * u * v * ds
    a += (gamma / h ** 2) * (u - u.Other()) * (v - v.Other()) * dw
    f = xf.LinearForm(Vh)
    f += 4.0 * v * dx
    a.Assemble()
    f.Assemble()
    m = xf.RestrictedBilinearForm(Vh, element_restriction=hasneg)
    m += u * v * dx
    m.Assemble()
    A = a.mat.to_scipy().tocsc()
    uh = spla.spsolve(A, f.vec.cpu().numpy())
    # coefficients of the exact solution in the hierarchical basis (vertex values; order-2 edge dof = u(a)+u(b)-2u(mid))
    ex = lambda q: R * R - (q[:, 0] - CX) ** 2 - (q[:, 1] - CY) ** 2
    pts = mesh.points
    full = [ex(pts)]
    if p == 2:
        edges, _ = mesh.edges()
        pa, pb = pts[edges[:, 0].long()], pts[edges[:, 1].long()]
        full.append(ex(pa) + ex(pb) - 2.0 * ex(0.5 * (pa + pb)))
    full = torch.cat(full)
    act = Vh.dofmap >= 0
    ci_ = np.zeros(Vh.ndof)
    ci_[Vh.dofmap[act].cpu().numpy()] = full[act].cpu().numpy()
    e = uh - ci_
    M = m.mat.to_scipy()
    return float(np.sqrt(e @ (M @ e))), float(np.sqrt(ci_ @ (M @ ci_))), Vh.ndof


@pytest.mark.parametrize("p", [1, 2])
def test_fictitious_domain_poisson_converges_with_second_order(p):
    errs = []
    for n in (8, 16, 32):
        err, nrm, ndof = solve(n, p)
        errs.append(err / nrm)
    assert errs[0] < 0.1 and errs[2] < 6e-3
    assert errs[1] / errs[0] < 0.4 and errs[2] / errs[1] < 0.4, errs   # second order: ratio 1/4 (+ pre-asymptotic slack)


def solve_demo(n, order=3, gamma_stab=0.1):
    """demos/fictdom.py:60-165 as shipped: ring 1/4 < r < 3/4 in [-1,1]^2, order 3, LevelSetMeshAdaptation(order, threshold=0.1,
    discontinuous_qn=True) deformation on every dCut / dFacetPatch, Nitsche with lambda = 10 order^2, ghost penalty 0.1 / h^2,
    right-hand side coeff_f = -Laplace(exact) as a coefficient function, error sqrt(Integrate((gfu - exact)^2 dx)).
    (structured triangles with h = 2/n instead of the netgen mesh with maxh = 0.1)"""
    import scipy.sparse.linalg as spla
    mesh = xf.MakeStructured2DMesh(nx=n, ny=n, mapping=lambda x, y: (2 * x - 1, 2 * y - 1), device=DEV)
    r2, r1 = 3 / 4, 1 / 4
    rc, rr = (r1 + r2) / 2.0, (r2 - r1) / 2.0
    rad = lambda x, y: torch.sqrt(x ** 2 + y ** 2)                                              # noqa: E731
    levelset = lambda x, y: torch.where(rad(x, y) - rc > 0, rad(x, y) - rc - rr, rc - rad(x, y) - rr)   # IfPos(r - rc, ...)   # noqa: E731
    exact = lambda x, y: 20 * (r2 - rad(x, y)) * (rad(x, y) - r1)                               # noqa: E731
    coeff_f = xf.CoefficientFunction(lambda x, y: 80.0 - 20.0 / rad(x, y))                      # -(u'' + u'/r), r1 + r2 = 1
    lsetmeshadap = xf.LevelSetMeshAdaptation(mesh, order=order, threshold=0.1, discontinuous_qn=True)
    deformation = lsetmeshadap.CalcDeformation(levelset)
    lsetp1 = lsetmeshadap.lset_p1
    ci = xf.CutInfo(mesh, lsetp1)
    hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
    ba_facets = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif)
    Vh = xf.Restrict(xf.H1(mesh, order=order, dgjumps=True), hasneg)
    u, v = Vh.TrialFunction(), Vh.TestFunction()
    h = xf.specialcf.mesh_size
    nrm = xf.Normalize(xf.grad(lsetp1))
    dx = xf.dCut(lsetp1, xf.NEG, definedonelements=hasneg, deformation=deformation)
    ds = xf.dCut(lsetp1, xf.IF, definedonelements=hasif, deformation=deformation)
    dw = xf.dFacetPatch(definedonelements=ba_facets, deformation=deformation)
    a = xf.RestrictedBilinearForm(Vh, element_restriction=hasneg, facet_restriction=ba_facets)
    a += xf.grad(u) * xf.grad(v) * dx
    a += -1.0 * (xf.grad(u) * nrm) * v * ds
    a += -1.0 * (xf.grad(v) * nrm) * u * ds
    a += (10.0 * order * order / h) * u * v * ds
    a += gamma_stab / h ** 2 * (u - u.Other()) * (v - v.Other()) * dw
    f = xf.LinearForm(Vh)
    f += coeff_f * v * dx
    a.Assemble()
    f.Assemble()
    uh = spla.spsolve(a.mat.to_scipy().tocsc(), f.vec.cpu().numpy())
    gfu = torch.from_numpy(uh).to(DEV)
    err2 = xf.Integrate(levelset_domain={"levelset": lsetp1, "domain_type": xf.NEG}, mesh=mesh, order=2 * order, deformation=deformation,
                        cf=xf.CoefficientFunction(lambda x, y, uh_: (uh_ - exact(x, y)) ** 2, gfs=[(Vh, gfu)]))
    area = xf.Integrate(levelset_domain={"levelset": lsetp1, "domain_type": xf.NEG}, mesh=mesh, order=2 * order, deformation=deformation)
    return float(np.sqrt(err2)), area, lsetmeshadap.nfail


def test_fictdom_demo_as_shipped_order3_deformation_and_its_right_hand_side():
    ring = np.pi * (0.75 ** 2 - 0.25 ** 2)
    err20, area20, nfail = solve_demo(20)      # h = 0.1: the demo's maxh
    assert nfail == 0
    err40, area40, _ = solve_demo(40)
    print("fictdom demo: L2 errors", err20, err40, "area errors", area20 - ring, area40 - ring)
    assert abs(area20 - ring) < 5e-4 and abs(area40 - ring) < abs(area20 - ring) / 8.0     # order-3 geometry: fourth order
    assert err20 < 5e-3 and err40 < err20 / 8.0                           # order 3: fourth-order L2 convergence (ratio 1/16, slack 2)


def test_variable_diffusion_coefficient_in_the_bilinear_form():
    """cf * grad(u)*grad(v): the operator applied to a linear function u = x tested with v = x gives int_Omega D(x) dx;
    mass form with a coefficient: 1^T M 1 = int D — against Integrate with the same coefficient (same rule, same points)"""
    mesh = xf.MakeStructured2DMesh(nx=12, ny=12, device=DEV)
    lset = xf.InterpolateToP1(lambda x, y: torch.sqrt((x - CX) ** 2 + (y - CY) ** 2) - R, mesh)
    ci = xf.CutInfo(mesh, lset)
    hasneg = ci.GetElementsOfType(xf.HASNEG)
    D = xf.CoefficientFunction(lambda x, y: 1.0 + x * y + torch.sin(3 * x))
    for p in (1, 2):
        Vh = xf.Restrict(xf.H1(mesh, order=p), hasneg)
        u, v = Vh.TnT()
        dx = xf.dCut(lset, xf.NEG, definedonelements=hasneg, order=6)
        k = xf.RestrictedBilinearForm(Vh, element_restriction=hasneg)
        k += D * (xf.grad(u) * xf.grad(v)) * dx
        k.Assemble()
        m = xf.RestrictedBilinearForm(Vh, element_restriction=hasneg)
        m += (u * v) * D * dx
        m.Assemble()
        ref = xf.Integrate(levelset_domain={"levelset": lset, "domain_type": xf.NEG, "order": 6}, mesh=mesh, cf=D)
        rows = Vh.rows(hasneg.indices())
        nvert = rows[:, :3].reshape(-1).long()
        one = torch.zeros(Vh.ndof, dtype=torch.float64, device=DEV); one[nvert] = 1.0
        xs = torch.zeros(Vh.ndof, dtype=torch.float64, device=DEV)
        act = Vh.dofmap[:mesh.nv] >= 0
        xs[Vh.dofmap[:mesh.nv][act].long()] = mesh.points[:, 0][act]        # u = x: vertex dofs x_v, higher dofs 0
        assert abs(float((one * (m.mat * one)).sum()) - ref) < 1e-12 * abs(ref)
        assert abs(float((xs * (k.mat * xs)).sum()) - ref) < 1e-11 * abs(ref)
        assert float((k.mat * one).abs().max()) < 1e-11 * float(k.mat.vals.abs().max())
    with pytest.raises(xf.XgpuError):
        Vh = xf.Restrict(xf.H1(mesh, order=1), hasneg)
        u, v = Vh.TnT()
        (D * (u * v) + xf.CoefficientFunction(lambda x, y: x) * (u * v))
