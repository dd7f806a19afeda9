"""Config-4 style assembly (BASELINE.json configs[3]): multi-level-set domain cube ∩ planes ∩ sphere through
BilinearForm.Assemble (dCut with a LIST of level sets -> cutint/mlsetintegration.cpp:6-186 rules ->
xfem/symboliccutbfi.cpp:76-486 element matrices -> CSR) against the oracle: bit-exact pattern, values 1e-12 (Frobenius)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from oracle import xoracle as xo

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEG, POS, IF = 0, 1, 2


def _lsets(mesh):
    x, y, z = (mesh.points[:, i] for i in range(3))
    fns = [x + 0.31 * y - 0.83, -0.27 * x + y + 0.11 * z - 0.79, 0.13 * x - 0.21 * y + z - 0.81,
           torch.sqrt((x - .4) ** 2 + (y - .45) ** 2 + (z - .4) ** 2) - 0.37]
    out = []
    for f in fns:
        f = torch.where(f.abs() < 1e-14, torch.full_like(f, 1e-14), f)
        out.append(xf.GridFunction(mesh, f.contiguous()))
    return out


def _oracle_mats(mesh, elems, phis, dts, p, order, terms):
    pts, els = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy()
    nd = xo.h1_ndof(3, p)
    mats = np.zeros((len(elems), nd, nd))
    for k, el in enumerate(elems):
        T = xo.make_trafo(pts[els[el]], els[el])
        ls = np.stack([ph[els[el]] for ph in phis], axis=1)
        q, w = xo.mlset_cut_rule(T, ls, dts, order)
        if len(w):
            xo.elmat_from_rule(T, p, ls[:, 0].copy(), q, w, terms, out=mats[k])
    return mats


@pytest.mark.parametrize("n,p,dts", [(4, 3, [(NEG, NEG, NEG, NEG)]), (5, 2, [(NEG, NEG, NEG, NEG), (NEG, NEG, NEG, IF)]),
                                     (4, 1, [(IF, NEG, NEG, NEG)])])
def test_mlset_domain_assembly_matches_oracle(n, p, dts):
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    lsets = _lsets(mesh)
    mlci = xf.MultiLevelsetCutInfo(mesh, lsets)
    active = mlci.GetElementsWithContribution(dts if len(dts) > 1 else dts[0])
    assert 0 < active.NumSet() < mesh.ne
    V = xf.Restrict(xf.H1(mesh, order=p), active)
    u, v = V.TnT()
    a = xf.RestrictedBilinearForm(V, element_restriction=active)
    a += (xf.grad(u) * xf.grad(v) + 2.5 * u * v) * xf.dCut(lsets, dts if len(dts) > 1 else dts[0], definedonelements=active)
    a.Assemble()
    terms = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 2.5, 0)]
    l_act = active.indices().cpu().numpy()
    phis = [g.vec.cpu().numpy() for g in lsets]
    m = _oracle_mats(mesh, l_act, phis, dts, p, 2 * p, terms)
    table = V.rows(active.indices()).cpu().numpy()
    rp, col = xo.graph(V.ndof, table)
    vals = xo.scatter(V.ndof, table, m, rp, col)
    assert np.array_equal(a.mat.rowptr.cpu().numpy(), rp) and np.array_equal(a.mat.col.cpu().numpy(), col)
    gv = a.mat.vals.cpu().numpy()
    assert np.linalg.norm(gv - vals) <= 1e-12 * np.linalg.norm(vals)
    # basis independent check: the mass part applied to u = 1 gives 2.5 |domain| (volume tuples only)
    if all(IF not in t for t in dts):
        one = torch.zeros(V.ndof, dtype=torch.float64, device=DEV)
        one[torch.from_numpy(table[:, :4].reshape(-1)).long().to(DEV)] = 1.0
        tot = float((one * (a.mat * one)).sum())
        vol = xf.Integrate({"levelset": lsets, "domain_type": dts if len(dts) > 1 else dts[0]}, mesh=mesh, cf=1.0, order=0)
        assert abs(tot - 2.5 * vol) < 1e-11 * abs(tot)


def test_two_different_levelsets_in_one_form():
    """two integrators with different single level sets take the general path as well"""
    mesh = xf.MakeStructured3DMesh(nx=4, device=DEV)
    lsets = _lsets(mesh)
    V = xf.H1(mesh, order=2)
    u, v = V.TnT()
    a = xf.BilinearForm(V)
    a += u * v * xf.dCut(lsets[3], NEG)
    a += xf.grad(u) * xf.grad(v) * xf.dCut(lsets[0], POS)
    a.Assemble()
    pts, els = mesh.points.cpu().numpy(), mesh.elements.cpu().numpy()
    allel = np.arange(mesh.ne, dtype=np.int32)
    m = xo.batch_cut_elmats(3, 2, allel, els, pts, lsets[3].vec.cpu().numpy(), NEG, -1, [(0, 0, 1.0, 0)])
    m += xo.batch_cut_elmats(3, 2, allel, els, pts, lsets[0].vec.cpu().numpy(), POS, -1, [(1, 1, 1.0, 0)])
    table = V.el_dofs.cpu().numpy()
    rp, col = xo.graph(V.ndof, table)
    vals = xo.scatter(V.ndof, table, m, rp, col)
    assert np.array_equal(a.mat.col.cpu().numpy(), col)
    assert np.linalg.norm(a.mat.vals.cpu().numpy() - vals) <= 1e-12 * np.linalg.norm(vals)
