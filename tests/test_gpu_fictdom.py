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
