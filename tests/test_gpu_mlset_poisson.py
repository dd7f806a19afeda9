"""The reference's tests/pytests/test_mlset_poisson.py:148-160 (`test_solution_in_space`) restated on a structured mesh:
Poisson on the unit square described by FOUR level sets (-y, x-1, y-1, -x) inside the background mesh (-0.2, 1.2)^2,
Nitsche on each of the four boundary pieces (multi-level-set tuples with one IF entry, normal of that level set), ghost
penalty on the facet patches, right-hand side 32 (y(1-y) + x(1-x)).  The exact solution 16 x(1-x) y(1-y) lies in the order-4
space, so the discrete solution must reproduce it: L2 error < 1e-10 (the reference asserts 1e-12 with a direct solver;
here the matrix comes from the GPU path, the solve is scipy's SuperLU).  Everything assembled on the accelerated path:
MultiLevelsetCutInfo, multi-level-set volume and codim-1 rules, element matrices with normal-derivative terms,
facet patches, LinearForm, CSR."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEG, POS, IF = 0, 1, 2


def test_solution_in_space():
    import scipy.sparse.linalg as spla
    k, n = 4, 6
    mesh = xf.MakeStructured2DMesh(nx=n, ny=n, mapping=lambda x, y: (1.4 * x - 0.2, 1.4 * y - 0.2), device=DEV)
    lsets = [xf.InterpolateToP1(f, mesh) for f in (lambda x, y: -y, lambda x, y: x - 1.0, lambda x, y: y - 1.0, lambda x, y: -x)]
    inner = (NEG, NEG, NEG, NEG)
    bnd = [tuple(IF if j == i else NEG for j in range(4)) for i in range(4)]
    mlci = xf.MultiLevelsetCutInfo(mesh, lsets)
    els_hasneg = mlci.GetElementsWithContribution(inner)
    els_if = els_hasneg & ~mlci.GetElementsOfType(inner)
    els_bnd = [mlci.GetElementsWithContribution(t) for t in bnd]
    facets_gp = xf.GetFacetsWithNeighborTypes(mesh, a=els_hasneg, b=els_if, use_and=True)
    assert 0 < els_if.NumSet() < els_hasneg.NumSet() < mesh.ne
    Vb = xf.H1(mesh, order=k, dgjumps=True)
    V = xf.Restrict(Vb, els_hasneg)
    u, v = V.TnT()
    h = xf.specialcf.mesh_size
    a = xf.RestrictedBilinearForm(V, element_restriction=els_hasneg, facet_restriction=facets_gp)
    a += xf.grad(u) * xf.grad(v) * xf.dCut(lsets, inner, definedonelements=els_hasneg)
    for i, t in enumerate(bnd):
        nrm = xf.Normalize(xf.grad(lsets[i]))       # square.GetOuterNormals: the normal of the level set that is IF
        form = -1.0 * (xf.grad(u) * nrm) * v + -1.0 * (xf.grad(v) * nrm) * u + (10.0 * k * k / h) * u * v
        a += form * xf.dCut(lsets, t, definedonelements=els_bnd[i])
    a += (0.1 / h ** 2) * (u - u.Other()) * (v - v.Other()) * xf.dFacetPatch(definedonelements=facets_gp)
    f = xf.LinearForm(V)
    f += (32.0 * xf.y) * v * xf.dCut(lsets, inner, definedonelements=els_hasneg, order=2 * k)
    f += (-32.0 * xf.y ** 2) * v * xf.dCut(lsets, inner, definedonelements=els_hasneg, order=2 * k)
    f += (32.0 * xf.x) * v * xf.dCut(lsets, inner, definedonelements=els_hasneg, order=2 * k)
    f += (-32.0 * xf.x ** 2) * v * xf.dCut(lsets, inner, definedonelements=els_hasneg, order=2 * k)
    a.Assemble()
    f.Assemble()
    uh = spla.spsolve(a.mat.to_scipy().tocsc(), f.vec.cpu().numpy())
    # L2 error over the multi-level-set domain with the order-2k rule (test_mlset_poisson.py:128-131), on the host
    l_act = els_hasneg.indices()
    phis = torch.stack([g.vec for g in lsets]).contiguous()
    off, q, w = ops.cutquad_ml(2, l_act, mesh.elements, mesh.points, phis, [inner], 2 * k)
    off, q, w = off.cpu().numpy(), q.cpu().numpy(), w.cpu().numpy()
    rows = V.rows(l_act).cpu().numpy()
    els, pts = mesh.elements.cpu().numpy(), mesh.points.cpu().numpy()
    err2 = nrm2 = 0.0
    for e, el in enumerate(l_act.cpu().numpy()):
        vt = pts[els[el]]
        det = abs(np.linalg.det(np.stack([vt[0] - vt[2], vt[1] - vt[2]], axis=1)))
        for j in range(off[e], off[e + 1]):
            sh, _ = ops.h1_shapes_host(2, k, els[el], q[j, :2])
            xq = q[j, 0] * vt[0] + q[j, 1] * vt[1] + (1 - q[j, 0] - q[j, 1]) * vt[2]
            ue = 16 * xq[0] * (1 - xq[0]) * xq[1] * (1 - xq[1])
            err2 += w[j] * det * (float(sh.numpy() @ uh[rows[e]]) - ue) ** 2
            nrm2 += w[j] * det * ue ** 2
    assert abs(nrm2 - 256.0 / 900.0) < 1e-12        # int u_ex^2 over the unit square = 256 (1/30)^2: the rule is exact
    assert np.sqrt(err2) < 1e-10, np.sqrt(err2)
