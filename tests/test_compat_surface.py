"""The older spellings of the reference's Python surface (ngsxfem_b200/compat.py; python/__init__.py:156-337, 455-500, 574-579,
654-690): the levelset_domain dictionary maps onto the same differential symbol as dCut, NoDeformation interpolates and
carries no deformation, kappa / IsCut are functions of the cut ratios.  Host logic only (no GPU needed)."""
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import compat, meshes
from ngsxfem_b200.forms import Integral

NEG, POS, IF = xf.NEG, xf.POS, xf.IF


def setup():
    mesh = meshes.MakeStructured2DMesh(nx=4, ny=4)
    lset = xf.InterpolateToP1(lambda x, y: torch.sqrt((x - .5) ** 2 + (y - .5) ** 2) - 0.3, mesh)
    V = xf.H1(mesh, order=2)
    return mesh, lset, V


def same_symbol(a, b):
    keys = ("levelset", "domain_type", "order", "definedonelements", "deformation", "time_order", "tref", "vb", "skeleton",
            "element_boundary")
    return all(getattr(a, k) is getattr(b, k) or getattr(a, k) == getattr(b, k) for k in keys)


def test_symbolic_bfi_lfi_are_the_cut_differential_symbol():
    mesh, lset, V = setup()
    u, v = V.TnT()
    ba = xf.BitArray(mesh.ne, "cpu")
    form = xf.grad(u) * xf.grad(v) + u * v
    i1 = xf.SymbolicBFI(levelset_domain={"levelset": lset, "domain_type": NEG, "subdivlvl": 0}, form=form, definedonelements=ba)
    i2 = form * xf.dCut(lset, NEG, definedonelements=ba)
    assert isinstance(i1, Integral) and i1.form is form and same_symbol(i1.dx, i2.dx) and i1.dx.order == -1
    # an `order` keyword only fills a missing dictionary entry (python/__init__.py:226-229)
    assert xf.SymbolicBFI(levelset_domain={"levelset": lset, "domain_type": IF}, form=u * v, order=4).dx.order == 4
    assert xf.SymbolicBFI(levelset_domain={"levelset": lset, "domain_type": IF, "order": 2}, form=u * v, order=4).dx.order == 2
    assert xf.SymbolicBFI(levelset_domain={"levelset": lset, "domain_type": IF, "order": -1}, form=u * v).dx.order == -1
    # time_order: the keyword wins over a dictionary entry of -1 (:230-233)
    s = xf.SymbolicBFI(levelset_domain={"levelset": lset, "domain_type": NEG, "time_order": -1}, form=u * v, time_order=2).dx
    assert s.time_order == 2
    # boundary elements, facets
    assert xf.SymbolicLFI(levelset_domain={"levelset": lset, "domain_type": NEG}, form=1.0 * v, VOL_or_BND=xf.BND).dx.vb == xf.BND
    fac = xf.BitArray(mesh.nfacet, "cpu")
    sk = xf.SymbolicBFI(levelset_domain={"levelset": lset, "domain_type": POS}, form=(u - u.Other()) * (v - v.Other()),
                        skeleton=True, definedonelements=fac)
    assert sk.dx.skeleton and not sk.dx.element_boundary and sk.dx.definedonelements is fac
    lf = xf.SymbolicLFI(levelset_domain={"levelset": lset, "domain_type": NEG}, form=(xf.x * xf.y) * v)
    assert same_symbol(lf.dx, xf.dCut(lset, NEG))
    # multi-level-set domains go through unchanged
    dta = xf.DomainTypeArray((NEG, POS))
    ml = xf.SymbolicBFI(levelset_domain={"levelset": (lset, lset), "domain_type": dta}, form=u * v)
    assert ml.dx.domain_type is dta
    for bad in (dict(levelset_domain=None, form=u * v), dict(levelset_domain={"levelset": lset}, form=u * v),
                dict(levelset_domain={"levelset": lset, "domain_type": NEG, "subdivlvl": 2}, form=u * v),
                dict(levelset_domain={"levelset": lset, "domain_type": NEG, "colour": 1}, form=u * v),
                dict(levelset_domain={"levelset": lset, "domain_type": NEG}, form=None),
                dict(levelset_domain={"levelset": lset, "domain_type": NEG}, form=u * v, definedon="bottom"),
                dict(levelset_domain={"levelset": lset, "domain_type": NEG}, form=u * v, bonus=3)):
        with pytest.raises(xf.XgpuError):
            xf.SymbolicBFI(**bad)
    a = xf.BilinearForm(V)
    a += i1                                                  # accepted like form * dCut(...)
    assert a.integrals[-1] is i1


def test_no_deformation():
    mesh, lset, V = setup()
    fn = lambda x, y: torch.sqrt((x - .5) ** 2 + (y - .5) ** 2) - 0.3
    nd = xf.NoDeformation(mesh, fn, warn=False)
    assert nd.deform is None and torch.equal(nd.lset_p1.vec, lset.vec)
    with nd as l1:
        assert l1 is nd.lset_p1
    old = nd.lset_p1
    assert nd.CalcDeformation(lambda x, y: x - 0.4) is None and nd.lset_p1 is old     # re-interpolated in place
    assert torch.allclose(nd.lset_p1.vec, mesh.points[:, 0] - 0.4)
    with pytest.raises(xf.XgpuError):
        xf.NoDeformation(None, fn)
    late = xf.NoDeformation(mesh, warn=False)
    late.CalcDeformation(fn)
    assert torch.equal(late.lset_p1.vec, lset.vec)
    # dCut takes its deform like any other `deformation=` argument
    assert xf.dCut(nd.lset_p1, NEG, deformation=nd.deform).deformation is None


def test_kappa_and_iscut_from_the_cut_ratios(monkeypatch):
    ratios = torch.tensor([0.0, 1.0, 0.25, 1e-17, 1.0 - 1e-17, 0.999], dtype=torch.float64)

    class FakeCutInfo:
        def __init__(self, mesh, lset):
            self.args = (mesh, lset)

        def GetCutRatios(self, vb=xf.VOL):
            assert vb == xf.VOL
            return ratios

    monkeypatch.setattr(compat, "CutInfo", FakeCutInfo)
    k1, k2 = xf.kappa("mesh", "lset")
    assert torch.equal(k1, ratios) and torch.equal(k2, 1.0 - ratios)
    assert xf.IsCut("mesh", "lset").tolist() == [0.0, 0.0, 1.0, 0.0, 0.0, 1.0]
    assert xf.CutRatioGF(FakeCutInfo(None, None)) is ratios
    with pytest.raises(xf.XgpuError):
        xf.kappa("mesh", "lset", subdivlvl=1)


def test_fix_t_is_fix_tref(capsys):
    mesh, lset, Vs = setup()
    V = xf.SpaceTimeFESpace(Vs, xf.ScalarTimeFE(1))
    u, v = V.TnT()
    a, b = xf.fix_t(u, 0.0), xf.fix_tref(u, 0.0)
    assert a.fixt == b.fixt == 0.0 and a.op == b.op
    assert "deprecated" in capsys.readouterr().out


def test_element_wise_coefficients():
    """ElementCF / IndicatorCF: the element's value at each of its rule points (offsets of the listed elements), folded into
    integrands like any coefficient; pure index arithmetic, no kernel"""
    mesh, lset, V = setup()
    u, v = V.TnT()
    mask = torch.zeros(mesh.ne, dtype=torch.bool)
    mask[[1, 4, 5, 30]] = True
    ba = xf.BitArray.from_numpy(mask.numpy(), "cpu")
    ind = xf.IndicatorCF(mesh, ba)
    assert isinstance(ind, xf.ElementCF) and isinstance(ind, xf.CoefficientFunction) and torch.equal(ind.values, mask.double())
    # three listed elements with 2, 0 and 3 rule points
    elems = torch.tensor([4, 7, 30], dtype=torch.int32)
    off = torch.tensor([0, 2, 2, 5], dtype=torch.int64)
    pts = torch.zeros((5, 3), dtype=torch.float64)
    assert ind.evaluate(2, mesh, elems, off, pts, None).tolist() == [1.0, 1.0, 1.0, 1.0, 1.0]
    assert ind.evaluate(2, mesh, torch.tensor([0, 4, 2], dtype=torch.int32), off, pts, None).tolist() == [0.0, 0.0, 0.0, 0.0, 0.0]
    cr = xf.ElementCF(torch.arange(mesh.ne, dtype=torch.float64))
    assert (2.0 * cr).evaluate(2, mesh, elems, off, pts, None).tolist() == [8.0, 8.0, 60.0, 60.0, 60.0]
    assert (cr * ind * 0.5).evaluate_st(2, mesh, elems, off, pts, torch.zeros(5), None).tolist() == [2.0, 2.0, 15.0, 15.0, 15.0]
    all_off = torch.arange(mesh.ne + 1, dtype=torch.int64)                        # elems=None: one point per element
    assert torch.equal(ind.evaluate(2, mesh, None, all_off, None, None), mask.double())
    # integrands carry it like a pointwise coefficient
    f = ind * (u * v)
    assert f.cf is ind and (ind * v).cf is ind and ((u * v) * ind).cf is ind
    itg = f * xf.dCut(lset, NEG)
    assert itg.form.cf is ind
    pointwise = xf.CoefficientFunction(lambda x, y: x)
    for bad in (lambda: ind * pointwise, lambda: pointwise * ind, lambda: cr * xf.ElementCF(torch.zeros(3)),
                lambda: xf.ElementCF(torch.zeros(3)).evaluate(2, mesh, None, all_off, None, None),
                lambda: ind.evaluate_facets(mesh, None, off, pts), lambda: xf.IndicatorCF(mesh, ba, facets=True),
                lambda: xf.IndicatorCF(mesh, xf.BitArray(mesh.ne + 1, "cpu")), lambda: xf.ElementCF(torch.zeros((2, 2)))):
        with pytest.raises(xf.XgpuError):
            bad()
