"""Host-side space-time surface (ngsxfem_b200/spacetime.py) without a GPU: the symbolic layer builds the term lists the
kernels take, the dof tables are time-major (spacetime/SpaceTimeFE.hpp:185-218), unsupported requests raise."""
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import spacetime as st

ID, GRAD, DT = 0, 1, 3


def space(p=1, kt=2, n=2):
    mesh = xf.MakeStructured2DMesh(nx=n, ny=n)
    return mesh, xf.SpaceTimeFESpace(xf.H1(mesh, order=p), xf.ScalarTimeFE(kt))


def test_dof_table_is_time_major():
    mesh, V = space(p=2, kt=2)
    Vs = V.spaceFES
    assert V.ndof == 3 * Vs.ndof and V.nd == 3 * Vs.nd
    rows = V.rows(torch.arange(mesh.ne, dtype=torch.int32)) if hasattr(Vs, "rows") else None
    if rows is None:
        r = Vs.el_dofs
        rows = torch.cat([r + j * Vs.ndof for j in range(3)], dim=1)
    for j in range(3):
        assert torch.equal(rows[:, j * Vs.nd:(j + 1) * Vs.nd].long(), Vs.el_dofs.long() + j * Vs.ndof)
    assert xf.ScalarTimeFE(2).nodes() == [0.0, 0.5, 1.0] and xf.ScalarTimeFE(1).nodes() == [0.0, 1.0]
    n3 = xf.ScalarTimeFE(3).nodes()
    assert abs(n3[1] + n3[2] - 1.0) < 1e-15 and abs(n3[1] - 0.5 * (1 - 5 ** -0.5)) < 1e-15   # Gauss-Lobatto
    for t in (0.0, 0.3, 1.0):
        assert abs(sum(xf.ScalarTimeFE(3).shapes(t)) - 1.0) < 1e-14


def test_forms_become_term_lists():
    mesh, V = space()
    u, v = V.TnT()
    h = xf.specialcf.mesh_size
    f = 2.0 * u * v + xf.dt(u) * v - 0.5 / h * xf.grad(u) * xf.grad(v) + u * xf.dtref(v)
    assert f.terms == [(ID, ID, 2.0, 0), (DT, ID, 1.0, 0), (GRAD, GRAD, -0.5, -1), (ID, DT, 1.0, 0)] and f.fixt is None
    g = xf.fix_tref(u, 0) * xf.fix_tref(v, 0)
    assert g.terms == [(ID, ID, 1.0, 0)] and g.fixt == 0.0
    j = 0.1 / h ** 2 * (u - u.Other()) * (v - v.Other())
    assert j.terms == [(st.OP_JUMP, st.OP_JUMP, 0.1, -2)]
    lf = 3.0 * xf.x * xf.y ** 2 * xf.tref * v + xf.tref ** 2 * xf.dt(v)
    assert lf.terms == [(ID, 3.0, 0, (1, 2, 0, 1)), (DT, 1.0, 0, (0, 0, 0, 2))]
    ls = xf.SpaceTimeGridFunction(mesh, xf.ScalarTimeFE(1))
    integ = f * xf.dCut(ls, xf.NEG, time_order=2)
    assert integ.dx.time_order == 2 and integ.dx.tref is None and integ.form is f
    dx = xf.dxtref(mesh, time_order=3)
    assert int(dx.domain_type) == xf.POS and dx.time_order == 3 and torch.equal(dx.levelset.vec, torch.ones(2, mesh.nv, dtype=torch.float64))
    dm = xf.dmesh(mesh, tref=1.0)
    assert dm.tref == 1.0 and dm.time_order == -1
    fp = xf.dFacetPatch(definedonelements=xf.BitArray(mesh.nfacet, "cpu"), time_order=2)
    assert fp.time_order == 2 and fp.tref is None


def test_unsupported_requests_raise():
    mesh, V = space()
    u, v = V.TnT()
    with pytest.raises(xf.XgpuError):
        xf.grad(u) * v                                   # grad must meet grad
    with pytest.raises(xf.XgpuError):
        xf.fix_tref(u, 0) * xf.fix_tref(v, 1)           # one reference time per term
    with pytest.raises(xf.XgpuError):
        xf.dt(xf.dt(u))
    with pytest.raises(xf.XgpuError):
        (u - u.Other()) * v                              # jump meets jump
    with pytest.raises(xf.XgpuError):
        u - 2 * u.Other()                                # not a scaled jump: both sides must carry the same factor
    with pytest.raises(xf.XgpuError):
        (2 * u) - u.Other()
    u0, _ = xf.H1(mesh, order=1).TnT()                   # the same rule on a plain space
    with pytest.raises(xf.XgpuError):
        u0 - 2 * u0.Other()
    with pytest.raises(xf.XgpuError):
        u * v + xf.fix_tref(u, 0) * xf.fix_tref(v, 0)   # one fixed time per integrand
    with pytest.raises(xf.XgpuError):
        xf.ScalarTimeFE(4)
    with pytest.raises(xf.XgpuError):
        xf.dxtref(mesh, time_order=1.5)
    with pytest.raises(xf.XgpuError):
        xf.dmesh(None, tref=0.0)
    with pytest.raises(xf.XgpuError):
        xf.dFacetPatch(definedonelements=xf.BitArray(mesh.nfacet, "cpu"), downscale=1.5)   # only 0 < downscale <= 1
    with pytest.raises(xf.XgpuError):
        xf.dCut([xf.GridFunction(mesh), xf.GridFunction(mesh)], (xf.NEG, xf.NEG), time_order=1)   # no multi-level-set slabs
    a = xf.BilinearForm(V)
    with pytest.raises(xf.XgpuError):
        a += 3.0                                          # not an integral


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a host without a GPU")
def test_space_time_assembly_has_no_cpu_fallback():
    mesh, V = space()
    u, v = V.TnT()
    a = xf.BilinearForm(V)
    a += u * v * xf.dxtref(mesh, time_order=2)
    with pytest.raises(xf.XgpuError):
        a.Assemble()
    with pytest.raises(xf.XgpuError):
        xf.CutInfo(mesh, xf.SpaceTimeGridFunction(mesh, xf.ScalarTimeFE(1)), time_order=0)
