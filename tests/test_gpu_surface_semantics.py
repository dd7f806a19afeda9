"""Semantics of the Python surface that user code relies on: in-place identity of the CutInfo BitArrays
across Update (§9.13), the sparsity pattern of a plain BilinearForm on a dgjumps space vs a facet-restricted one (§9.14),
rejection of the legacy paths (§9.15), marker defaults (§9.8)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import XgpuError
from oracle import xoracle as xo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def sphere(mesh, r):
    return xf.InterpolateToP1(lambda x, y, z: torch.sqrt((x - .5) ** 2 + (y - .5) ** 2 + (z - .5) ** 2) - r, mesh)


def test_cutinfo_update_refills_the_same_bitarrays():
    mesh = xf.MakeStructured3DMesh(nx=6, device=DEV)
    ci = xf.CutInfo(mesh, sphere(mesh, 0.3))
    hasneg, iff, ratios = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF), ci.GetCutRatios()
    n0, i0 = hasneg.NumSet(), iff.NumSet()
    dx = xf.dCut(sphere(mesh, 0.3), xf.NEG, definedonelements=hasneg)    # captured BEFORE the update (time-loop idiom)
    ci.Update(sphere(mesh, 0.42))
    assert ci.GetElementsOfType(xf.HASNEG) is hasneg and ci.GetElementsOfType(xf.IF) is iff and ci.GetCutRatios() is ratios
    assert hasneg.NumSet() > n0 and iff.NumSet() != i0
    assert dx.definedonelements.NumSet() == hasneg.NumSet()
    # values equal a fresh CutInfo
    fresh = xf.CutInfo(mesh, sphere(mesh, 0.42))
    for dt in (xf.NEG, xf.POS, xf.IF, xf.HASNEG, xf.HASPOS, xf.UNCUT, xf.ANY, xf.NO):
        assert np.array_equal(ci.GetElementsOfType(dt).to_numpy(), fresh.GetElementsOfType(dt).to_numpy())
    assert fresh.GetElementsOfType(xf.ANY).NumSet() == mesh.ne and fresh.GetElementsOfType(xf.NO).NumSet() == 0
    thr = ci.GetElementsWithThresholdContribution(xf.NEG, 0.5).to_numpy()
    assert np.array_equal(thr, ratios.cpu().numpy() >= 0.5)


def test_mlset_cutinfo_update_in_place():
    mesh = xf.MakeStructured2DMesh(nx=8, ny=8, device=DEV)
    mk = lambda s: [xf.InterpolateToP1(lambda x, y: x - 0.5 - s, mesh), xf.InterpolateToP1(lambda x, y: y - 0.37, mesh)]
    ls = mk(0.0)
    mlci = xf.MultiLevelsetCutInfo(mesh, ls)
    a = mlci.GetElementsOfType((xf.NEG, xf.NEG))
    b = mlci.GetElementsWithContribution((xf.NEG, xf.NEG))
    na, nb = a.NumSet(), b.NumSet()
    assert mlci.GetElementsOfType([(xf.NEG, xf.NEG)]) is a        # cached per requested tuple list
    mlci.Update(mk(0.2))
    assert mlci.GetElementsOfType((xf.NEG, xf.NEG)) is a and a.NumSet() > na and b.NumSet() > nb
    with pytest.raises(XgpuError):
        mlci.GetElementsOfType((xf.NEG,))                            # tuple length must match the number of level sets


@pytest.mark.parametrize("restricted", [False, True])
def test_dgjumps_sparsity_pattern(restricted):
    """plain BilinearForm on Restrict(H1(dgjumps=True), hasneg): couples across EVERY facet between active elements;
    RestrictedBilinearForm(facet_restriction=ba): only across the marked facets (utils/restrictedblf.cpp:83-134)"""
    mesh = xf.MakeStructured3DMesh(nx=4, device=DEV)
    lset = sphere(mesh, 0.33)
    ci = xf.CutInfo(mesh, lset)
    hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
    ba = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif)
    V = xf.Restrict(xf.H1(mesh, order=2, dgjumps=True), hasneg)
    u, v = V.TnT()
    a = xf.RestrictedBilinearForm(V, element_restriction=hasneg, facet_restriction=ba) if restricted else xf.BilinearForm(V)
    a += u * v * xf.dCut(lset, xf.NEG, definedonelements=hasneg)
    a.Assemble()
    _, _, fe = mesh.facets()
    fe_np, act = fe.cpu().numpy(), hasneg.to_numpy()
    if restricted:
        fsel = np.nonzero(ba.to_numpy() & (fe_np[:, 1] >= 0))[0]
    else:
        fsel = np.nonzero((fe_np[:, 1] >= 0) & act[fe_np[:, 0]] & act[np.maximum(fe_np[:, 1], 0)])[0]
    t_el = V.rows(hasneg.indices()).cpu().numpy()
    fel = fe[torch.from_numpy(fsel).to(DEV)]
    t_f = torch.cat([V.rows(fel[:, 0].contiguous()), V.rows(fel[:, 1].contiguous())], dim=1).cpu().numpy()
    W = t_f.shape[1]
    tab = np.full((t_el.shape[0] + t_f.shape[0], W), -1, np.int32)
    tab[:t_el.shape[0], :t_el.shape[1]] = t_el
    tab[t_el.shape[0]:] = t_f
    rp, col = xo.graph(V.ndof, tab)
    assert np.array_equal(a.mat.rowptr.cpu().numpy(), rp) and np.array_equal(a.mat.col.cpu().numpy(), col)


def test_legacy_paths_are_rejected_not_emulated():
    mesh = xf.MakeStructured2DMesh(nx=4, ny=4, device=DEV)
    lset = xf.InterpolateToP1(lambda x, y: x - 0.4, mesh)
    with pytest.raises(XgpuError):
        xf.CutInfo(mesh, lset, subdivlvl=1)
    with pytest.raises(XgpuError):
        xf.dCut(lset, xf.NEG, subdivlvl=2)
    with pytest.raises(XgpuError):
        xf.CutInfo(mesh, lambda x, y: x - 0.4)            # not a P1 GridFunction
    with pytest.raises(NotImplementedError):
        xf.MakeStructured2DMesh(quads=True)
    with pytest.raises(XgpuError):
        xf.BilinearForm(xf.H1(mesh, order=1), symmetric=True)
    # marker defaults (python_xfem.cpp:346-369): b = a, boundary values True -> boundary facets of marked elements count
    ci = xf.CutInfo(mesh, lset)
    f_def = xf.GetFacetsWithNeighborTypes(mesh, a=ci.GetElementsOfType(xf.HASNEG))
    f_nob = xf.GetFacetsWithNeighborTypes(mesh, a=ci.GetElementsOfType(xf.HASNEG), bnd_val_a=False, bnd_val_b=False)
    assert f_def.NumSet() > f_nob.NumSet() > 0


def test_dfacetpatch_downscale_assembles_the_scaled_patch_matrices():
    """dFacetPatch(definedonelements=..., downscale=s) (python/__init__.py:749-778 -> SetIRScaling): the assembled ghost
    penalty is the sum of the oracle's scaled patch matrices; the pattern is the one of the unscaled form"""
    import scipy.sparse as sp
    mesh = xf.MakeStructured3DMesh(nx=4, device=DEV)
    lset = sphere(mesh, 0.34)
    ci = xf.CutInfo(mesh, lset)
    hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
    gpf = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif, use_and=True)
    V = xf.Restrict(xf.H1(mesh, order=2, dgjumps=True), hasneg)
    u, v = V.TnT()
    h = xf.specialcf.mesh_size

    def assemble(s):
        a = xf.RestrictedBilinearForm(V, element_restriction=hasneg, facet_restriction=gpf)
        a += (0.1 / h ** 2) * (u - u.Other()) * (v - v.Other()) * xf.dFacetPatch(definedonelements=gpf, downscale=s)
        a.Assemble()
        return a.mat

    A1, As = assemble(None), assemble(0.5)
    assert torch.equal(A1.rowptr, As.rowptr) and torch.equal(A1.col, As.col)
    assert torch.equal(assemble(1.0).vals, A1.vals)
    _, _, fe = mesh.facets()
    l_f = gpf.indices()
    l_f = l_f[fe[l_f.long(), 1] >= 0].contiguous()
    fel = fe[l_f.long()]
    mf, nfail = xo.batch_facet_patch(3, 2, l_f.cpu().numpy(), fe.cpu().numpy(), mesh.elements.cpu().numpy(),
                                     mesh.points.cpu().numpy(), -1, 0.1, -2, downscale=0.5)
    assert nfail == 0 and len(mf) > 0
    t_f = torch.cat([V.rows(fel[:, 0].contiguous()), V.rows(fel[:, 1].contiguous())], dim=1).cpu().numpy()
    ref = sp.lil_matrix((V.ndof, V.ndof))
    for r in range(t_f.shape[0]):
        d = t_f[r]                      # the dofs on the common facet appear in both halves: add entry by entry
        for i in range(len(d)):
            for j in range(len(d)):
                ref[d[i], d[j]] += mf[r, i, j]
    diff = As.to_scipy() - ref.tocsr()
    assert np.sqrt(diff.multiply(diff).sum()) <= 1e-12 * np.sqrt(ref.multiply(ref).sum())
    with pytest.raises(XgpuError):
        xf.dFacetPatch(definedonelements=gpf, downscale=0.0)
