"""Cut LINEAR forms on the GPU (xgpu_elvec_*, xgpu_vec_scatter, LinearForm) — SymbolicCutLinearFormIntegrator,
xfem/symboliccutlfi.cpp:51-140 — against the oracle (1e-12) and against closed-form integrals."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from oracle import xoracle as xo

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEG, POS, IF = 0, 1, 2


def lset_of(mesh):
    if mesh.dim == 3:
        return xf.InterpolateToP1(lambda x, y, z: torch.sqrt((x - .45) ** 2 + (y - .5) ** 2 + (z - .55) ** 2) - 0.37, mesh)
    return xf.InterpolateToP1(lambda x, y: torch.sqrt((x - .5) ** 2 + (y - .45) ** 2) - 0.31, mesh)


@pytest.mark.parametrize("dim,n,p,dt", [(3, 4, 1, NEG), (3, 4, 2, IF), (3, 3, 3, POS), (3, 2, 4, NEG), (2, 7, 2, NEG), (2, 6, 4, IF),
                                        (1, 9, 3, NEG)])
def test_element_vectors_match_oracle(dim, n, p, dt):
    mk = {1: xf.MakeStructured1DMesh, 2: xf.MakeStructured2DMesh, 3: xf.MakeStructured3DMesh}[dim]
    mesh = mk(nx=n, device=DEV) if dim != 2 else mk(nx=n, ny=n, device=DEV)
    lset = lset_of(mesh) if dim > 1 else xf.InterpolateToP1(lambda x: x - 0.43, mesh)
    ex = [(2, 1, 0), (0, 0, 1), (1, 0, 0)]
    terms = [(ops.OP_ID, 1.5, 0, tuple(e if k < dim else 0 for k, e in enumerate(ex[0]))),
             (ops.OP_ID, -0.5, 1, tuple(e if k < dim else 0 for k, e in enumerate(ex[1]))),
             (ops.OP_DN, 0.25, -1, tuple(e if k < dim else 0 for k, e in enumerate(ex[2])))]
    off, q, w = ops.cutquad(dim, None, mesh.elements, mesh.points, lset.vec, dt, 2 * p)
    v = ops.elvec_rule(dim, p, None, mesh.elements, mesh.points, lset.vec, off, q, w, terms)
    vo = xo.batch_cut_elvecs(dim, p, None, mesh.elements.cpu().numpy(), mesh.points.cpu().numpy(), lset.vec.cpu().numpy(),
                             dt, -1, terms)
    assert np.linalg.norm(vo) > 0
    assert np.linalg.norm(v.cpu().numpy() - vo) <= 1e-12 * np.linalg.norm(vo)
    # accumulate + slots: add the same vectors into a permuted buffer
    perm = torch.randperm(mesh.ne, generator=torch.Generator().manual_seed(1)).to(torch.int32).to(DEV)
    buf = torch.zeros_like(v)
    ops.elvec_rule(dim, p, None, mesh.elements, mesh.points, lset.vec, off, q, w, terms, out=buf, accumulate=True, slots=perm)
    ops.elvec_rule(dim, p, None, mesh.elements, mesh.points, lset.vec, off, q, w, terms, out=buf, accumulate=True, slots=perm)
    assert torch.allclose(buf[perm.long()], 2 * v, rtol=1e-14, atol=0)


def test_uncut_element_vectors_with_the_standard_rule():
    mesh = xf.MakeStructured3DMesh(nx=3, device=DEV)
    phi = torch.full((mesh.nv,), -1.0, dtype=torch.float64, device=DEV)
    terms = [(ops.OP_ID, 1.0, 0, (1, 1, 0))]
    v = ops.elvec_std(3, 3, None, mesh.elements, mesh.points, phi, -1, terms)
    vo = xo.batch_cut_elvecs(3, 3, None, mesh.elements.cpu().numpy(), mesh.points.cpu().numpy(), phi.cpu().numpy(), NEG, -1, terms)
    assert np.linalg.norm(v.cpu().numpy() - vo) <= 1e-12 * np.linalg.norm(vo)


def test_element_vectors_on_a_deformed_mesh_match_oracle():
    mesh = xf.MakeStructured3DMesh(nx=3, device=DEV)
    lset = lset_of(mesh)
    Vd = xf.H1(mesh, order=2)
    g = torch.Generator().manual_seed(4)
    defo = xf.Deformation(Vd, ((torch.rand((Vd.ndof, 3), generator=g, dtype=torch.float64) - 0.5) * 0.02).to(DEV))
    terms = [(ops.OP_ID, 1.0, 0, (1, 0, 2)), (ops.OP_DN, 2.0, 1, (0, 0, 0))]
    ctx = ops.default_ctx()
    for dt in (NEG, IF):
        with ctx.deformation(defo):
            off, q, w = ops.cutquad(3, None, mesh.elements, mesh.points, lset.vec, dt, 6)
            v = ops.elvec_rule(3, 3, None, mesh.elements, mesh.points, lset.vec, off, q, w, terms)
        vo = xo.batch_cut_elvecs(3, 3, None, mesh.elements.cpu().numpy(), mesh.points.cpu().numpy(), lset.vec.cpu().numpy(),
                                 dt, 6, terms, def_order=2, deform=defo.local().cpu().numpy())
        assert np.linalg.norm(v.cpu().numpy() - vo) <= 1e-12 * np.linalg.norm(vo)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_linear_form_assembly_matches_oracle_and_closed_form(p):
    """f = (3 x^2 y + 2 z) v dCut(NEG) on {x + 0.2 y < 0.57}: sum_i f_i c_i with c = coefficients of u = 1 equals
    int_{Omega^-} 3 x^2 y + 2 z (closed form below); entries vs the oracle scatter (fixed order -> also bit reproducible)."""
    mesh = xf.MakeStructured3DMesh(nx=4, device=DEV)
    lset = xf.GridFunction(mesh, mesh.points[:, 0] + 0.2 * mesh.points[:, 1] - 0.57)
    ci = xf.CutInfo(mesh, lset)
    hasneg = ci.GetElementsOfType(xf.HASNEG)
    V = xf.Restrict(xf.H1(mesh, order=p), hasneg)
    v = V.TestFunction()
    f = xf.LinearForm(V)
    f += (3.0 * xf.x ** 2 * xf.y) * v * xf.dCut(lset, NEG, definedonelements=hasneg)
    f += 2.0 * xf.z * v * xf.dCut(lset, NEG, definedonelements=hasneg, order=2 * p)
    f.Assemble()
    terms = [(ops.OP_ID, 3.0, 0, (2, 1, 0)), (ops.OP_ID, 2.0, 0, (0, 0, 1))]
    l_act = hasneg.indices().cpu().numpy()
    vo = xo.batch_cut_elvecs(3, p, l_act, mesh.elements.cpu().numpy(), mesh.points.cpu().numpy(), lset.vec.cpu().numpy(),
                             NEG, -1, terms)
    table = V.rows(hasneg.indices()).cpu().numpy()
    ref = np.zeros(V.ndof)
    np.add.at(ref, table.reshape(-1), vo.reshape(-1))
    got = f.vec.cpu().numpy()
    assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)
    f2 = xf.LinearForm(V)
    f2 += (3.0 * xf.x ** 2 * xf.y) * v * xf.dCut(lset, NEG, definedonelements=hasneg)
    f2 += 2.0 * xf.z * v * xf.dCut(lset, NEG, definedonelements=hasneg, order=2 * p)
    assert torch.equal(f2.Assemble().vec, f.vec)
    if p >= 2:  # rule order 2p >= 3 integrates the cubic exactly
        one = np.zeros(V.ndof); one[table[:, :4].reshape(-1)] = 1.0
        # int_0^1 int_0^1 int_0^{0.57-0.2y} (3 x^2 y + 2 z) dx dz dy = int_0^1 y X^3 + X dy,  X = 0.57 - 0.2 y
        ys, wy = np.polynomial.legendre.leggauss(8)
        ys = 0.5 * (ys + 1); wy = 0.5 * wy
        X = 0.57 - 0.2 * ys
        exact = float((wy * (ys * X ** 3 + X)).sum())
        assert abs(got @ one - exact) < 1e-13
