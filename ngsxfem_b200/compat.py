"""The older spellings of ngsxfem's Python surface for the accelerated path (python/__init__.py of the reference): integrators
given by a `levelset_domain` dictionary (`SymbolicBFI` / `SymbolicLFI`, :156-337), `NoDeformation` (:654-690), `CutRatioGF`,
`kappa`, `IsCut` (:455-500) and the deprecated `fix_t` (:574-579).  Host-side glue only: everything maps onto dCut / CutInfo /
fix_tref of this package (and raises where they raise)."""
import torch

from ._lib import XgpuError
from .cutinfo import CutInfo, GridFunction, InterpolateToP1, NEG, VOL
from .forms import dCut
from .spacetime import fix_tref

_DOMAIN_KEYS = {"levelset", "domain_type", "subdivlvl", "order", "time_order", "quad_dir_policy", "tref"}


def _cut_symbol(levelset_domain, kwargs):
    """the differential symbol of a levelset_domain dictionary plus the integrator keywords (python/__init__.py:225-238: an
    `order` / `time_order` keyword only fills the dictionary entry that is missing; python_xfem.cpp:722-931 for the rest)"""
    if not isinstance(levelset_domain, dict):
        raise XgpuError("SymbolicBFI / SymbolicLFI without a levelset_domain dictionary are NGSolve's own integrators "
                        "(not part of the accelerated path)")
    unknown = set(levelset_domain) - _DOMAIN_KEYS
    if unknown:
        raise XgpuError(f"levelset_domain: unknown entries {sorted(unknown)}")
    if "levelset" not in levelset_domain or "domain_type" not in levelset_domain:
        raise XgpuError("levelset_domain needs the entries 'levelset' and 'domain_type'")
    ld = dict(levelset_domain)
    kw = dict(kwargs)
    if "order" in kw:
        o = kw.pop("order")
        if "order" not in ld:
            ld["order"] = o
    if "time_order" in kw:
        t = kw.pop("time_order")
        if "time_order" not in ld or ld["time_order"] == -1:
            ld["time_order"] = t
    vb = kw.pop("VOL_or_BND", VOL)
    if kw.pop("definedon", None) is not None:
        raise XgpuError("definedon=Region: give definedonelements (a BitArray) instead")
    order = ld.get("order", None)
    return dCut(ld["levelset"], ld["domain_type"], order=None if order is None or order < 0 else order,
                subdivlvl=ld.get("subdivlvl", None), time_order=ld.get("time_order", -1), tref=ld.get("tref", None),
                definedonelements=kw.pop("definedonelements", None), deformation=kw.pop("deformation", None), vb=vb,
                element_boundary=bool(kw.pop("element_boundary", False)), skeleton=bool(kw.pop("skeleton", False)), **kw)


def SymbolicBFI(levelset_domain=None, form=None, **kwargs):
    """a.+= SymbolicBFI(levelset_domain={"levelset": lsetp1, "domain_type": NEG, ...}, form=u*v, definedonelements=ba)
    == a += u*v * dCut(lsetp1, NEG, definedonelements=ba)   (python/__init__.py:156-246)"""
    if form is None:
        raise XgpuError("SymbolicBFI needs form=")
    return form * _cut_symbol(levelset_domain, kwargs)


def SymbolicLFI(levelset_domain=None, form=None, **kwargs):
    """f += SymbolicLFI(levelset_domain=..., form=coef*v, ...) == f += coef*v * dCut(...)   (python/__init__.py:248-337)"""
    if form is None:
        raise XgpuError("SymbolicLFI needs form=")
    return form * _cut_symbol(levelset_domain, kwargs)


class NoDeformation:
    """the stand-in for LevelSetMeshAdaptation at order 1 (python/__init__.py:654-690): `lset_p1` is the P1 interpolant,
    `deform` is no deformation (None: every `deformation=` argument of this package takes it), `CalcDeformation` only
    re-interpolates.  Usable as a context manager like LevelSetMeshAdaptation (yields lset_p1)."""

    def __init__(self, mesh=None, levelset=None, warn=True):
        self.mesh, self.deform, self.warn = mesh, None, warn
        if levelset is not None:
            if mesh is None:
                raise XgpuError("need mesh")
            self.lset_p1 = InterpolateToP1(levelset, GridFunction(mesh))

    def __enter__(self):
        return self.lset_p1

    def __exit__(self, typ, value, tb):
        return False

    def CalcDeformation(self, levelset):
        if self.warn is True:
            print("WARNING: CalcDeformation called on instance of NoDeformation. No deformation computed.")
        if self.mesh is None:
            raise XgpuError("need mesh")
        if not hasattr(self, "lset_p1"):
            self.lset_p1 = GridFunction(self.mesh)
        InterpolateToP1(levelset, self.lset_p1)
        return self.deform


def IndicatorCF(mesh, ba, facets=False):
    """1 on the elements marked in the BitArray, 0 on the others (python/__init__.py:441-453; a snapshot of `ba`, as the
    reference's BitArrayCF(BitArray(ba))).  The facet variant is a facet-wise function and is not provided."""
    if facets:
        raise XgpuError("IndicatorCF(facets=True): facet-wise functions are not on the accelerated path")
    if len(ba) != mesh.ne:
        raise XgpuError("IndicatorCF: the BitArray is not over the elements of the mesh")
    from .forms import ElementCF
    return ElementCF(torch.from_numpy(ba.to_numpy().astype("float64")).to(mesh.device))


def CutRatioGF(cutinfo):
    """the ratio |element ∩ NEG| / |element| of every element as a piecewise constant function: one value per element
    (python/__init__.py:455-462; an L2(order=0) GridFunction there, the device vector of CutInfo here)"""
    return cutinfo.GetCutRatios(VOL)


def kappa(mesh, lset_approx, subdivlvl=0):
    """(kappa_neg, kappa_pos) per element (deprecated in the reference, python/__init__.py:464-478: there through an L2
    projection of the NEG indicator, which IS the cut ratio)"""
    if subdivlvl != 0:
        raise XgpuError("subdivlvl > 0 selects ngsxfem's legacy subdivision path: not provided")
    k1 = CutRatioGF(CutInfo(mesh, lset_approx))
    return k1, 1.0 - k1


def IsCut(mesh, lset_approx, subdivlvl=0):
    """1 on cut elements, 0 elsewhere (deprecated in the reference, python/__init__.py:480-498: 1e-16 < kappa < 1 - 1e-16)"""
    k1, _ = kappa(mesh, lset_approx, subdivlvl)
    return ((k1 > 1e-16) & (k1 < 1.0 - 1e-16)).to(torch.float64)


def fix_t(obj, time, *args, **kwargs):
    """deprecated name of fix_tref (python/__init__.py:574-579)"""
    print('WARNING: fix_t is deprecated. Use "fix_tref" instead. \n         Note that operators act w.r.t. the reference time intervals.')
    return fix_tref(obj, time, *args, **kwargs)
