"""Edge cases of the hot path through the public surface: no cut elements, no active elements at all, a single element,
empty lists on every batched entry point (the reference returns nullptr rules / empty BitArrays in these cases and
assembles an empty or uncut-only matrix)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from ngsxfem_b200.pipeline import CutAssembly
from oracle import xoracle as xo
from oracle.pipeline_ref import oracle_step

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _step(mesh, phi, p=2):
    V0 = xf.H1(mesh, order=p)
    r = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, p).step(phi)
    o, _ = oracle_step(3, mesh.elements.cpu().numpy(), mesh.points.cpu().numpy(), V0.el_dofs.cpu().numpy(), V0.ndof, p,
                       phi.cpu().numpy())
    return r, o


def test_all_negative_no_cut_elements():
    mesh = xf.MakeStructured3DMesh(nx=3, device=DEV)
    phi = torch.full((mesh.nv,), -0.7, dtype=torch.float64, device=DEV)
    r, o = _step(mesh, phi)
    assert r["n_if"] == 0 and r["n_neg"] == mesh.ne and r["ndof"] == o["ndof"] and r["nnz"] == o["nnz"]
    assert np.array_equal(r["col"].cpu().numpy(), o["col"])
    assert np.linalg.norm(r["vals"].cpu().numpy() - o["vals"]) <= 1e-12 * np.linalg.norm(o["vals"])


def test_all_positive_nothing_active():
    mesh = xf.MakeStructured3DMesh(nx=3, device=DEV)
    phi = torch.full((mesh.nv,), 0.3, dtype=torch.float64, device=DEV)
    V0 = xf.H1(mesh, order=2)
    r = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2).step(phi)
    assert r["units"] == 0 and r["ndof"] == 0 and r["nnz"] == 0
    assert r["rowptr"].numel() == 1 and r["col"].numel() == 0 and r["vals"].numel() == 0
    ci = xf.CutInfo(mesh, xf.GridFunction(mesh, phi))
    assert ci.GetElementsOfType(xf.HASNEG).NumSet() == 0 and ci.GetElementsOfType(xf.POS).NumSet() == mesh.ne
    assert float(ci.GetCutRatios().abs().max()) == 0.0


def test_single_cut_element():
    pts = torch.tensor([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]], dtype=torch.float64, device=DEV)
    mesh = xf.Mesh(pts, torch.tensor([[1, 2, 3, 0]], dtype=torch.int32, device=DEV))
    phi = torch.tensor([-0.3, 0.5, 0.2, 0.4], dtype=torch.float64, device=DEV)
    r, o = _step(mesh, phi, p=2)
    assert r["n_if"] == 1 and r["nnz"] == 100
    assert np.array_equal(r["col"].cpu().numpy(), o["col"])
    assert np.linalg.norm(r["vals"].cpu().numpy() - o["vals"]) <= 1e-12 * np.linalg.norm(o["vals"])


def test_empty_lists_on_every_batched_entry_point():
    mesh = xf.MakeStructured3DMesh(nx=2, device=DEV)
    phi = xf.InterpolateToP1(lambda x, y, z: x - 0.3, mesh).vec
    empty = torch.zeros(0, dtype=torch.int32, device=DEV)
    off, q, w = ops.cutquad(3, empty, mesh.elements, mesh.points, phi, ops.NEG, 4)
    assert off.tolist() == [0] and q.shape == (0, 3) and w.numel() == 0
    T = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0)]
    assert ops.elmat_rule(3, 2, empty, mesh.elements, mesh.points, phi, off, q, w, T).shape == (0, 10, 10)
    assert ops.elmat_std(3, 2, empty, mesh.elements, mesh.points, phi, -1, T).shape == (0, 10, 10)
    assert ops.elmat_rule_pair(3, 2, empty, mesh.elements, mesh.points, phi, (off, q, w), T, (off, q, w), T).shape == (0, 10, 10)
    assert ops.elvec_rule(3, 2, empty, mesh.elements, mesh.points, phi, off, q, w, [(ops.OP_ID, 1.0, 0, (0, 0, 0))]).shape == (0, 10)
    _, _, fe = mesh.facets()
    assert ops.facetpatch(3, 2, empty, fe, mesh.elements, mesh.points, -1, 1.0, 0).shape == (0, 20, 20)
    assert ops.integrate_monomial(3, empty, mesh.elements, mesh.points, off, q, w, (0, 0, 0)) == 0.0
    assert ops.bitset_to_list(ops.new_bits(100, DEV), 100).numel() == 0
    # a plan whose batches are all empty
    t0 = torch.zeros((0, 10), dtype=torch.int32, device=DEV)
    plan = ops.csr_plan(7, [t0])
    rp, col, vals = ops.csr_assemble(plan, [(t0, torch.zeros((0, 10, 10), dtype=torch.float64, device=DEV))])
    assert plan.nnz == 0 and rp.tolist() == [0] * 8 and col.numel() == 0 and vals.numel() == 0
    # forms on a space without active elements
    none = xf.CutInfo(mesh, xf.GridFunction(mesh, torch.full((mesh.nv,), 1.0, dtype=torch.float64, device=DEV))).GetElementsOfType(xf.HASNEG)
    V = xf.Restrict(xf.H1(mesh, order=1), none)
    u, v = V.TnT()
    a = xf.RestrictedBilinearForm(V, element_restriction=none)
    a += u * v * xf.dCut(xf.GridFunction(mesh, phi), xf.NEG, definedonelements=none)
    a.Assemble()
    assert V.ndof == 0 and a.mat.nze == 0
    f = xf.LinearForm(V)
    f += 1.0 * v * xf.dCut(xf.GridFunction(mesh, phi), xf.NEG, definedonelements=none)
    assert f.Assemble().vec.numel() == 0


def test_context_stats_count_calls_launches_and_carry_the_status():
    """xgpu_get_stats: NVTX-wrapped entry points and kernel launches are counted, the sticky geometry status comes along"""
    import ngsxfem_b200 as xf
    from ngsxfem_b200 import ops
    ctx = ops.default_ctx()
    s0 = ctx.stats()
    mesh = xf.MakeStructured2DMesh(nx=4, ny=4, device="cuda:0")
    phi = (mesh.points[:, 0] - 0.3).contiguous()
    ops.classify(2, mesh.elements, phi, ctx=ctx)
    s1 = ctx.stats()
    assert s1["api_calls"] > s0["api_calls"] and s1["kernel_launches"] > s0["kernel_launches"]
    assert s1["status_code"] == 0 and s1["status_element"] == -1 and s1["scratch_bytes"] >= 0
