"""dCut / dFacetPatch / SymbolicCutBFI-style integrators and (Restricted)BilinearForm.Assemble —
the Python mirror of ngsxfem's form surface (python/__init__.py:749-864, 973-1037;
cutint/python_cutint.cpp:359-392, 447-483; cutint/cutintegral.cpp:12-66, 233-265;
utils/restrictedblf.cpp) driving the batched CUDA path.

NGSolve interprets arbitrary CoefficientFunction trees per quadrature point
(xfem/symboliccutbfi.cpp:174-193).  The accelerated path supports a closed menu and raises for
anything else (no silent fallback):  sums of  c * h^k * A(u) * B(v)  with A, B in
{u, grad(u)., grad(u)*n}, n = Normalize(grad(lsetp1)), h = specialcf.mesh_size, and on facet
patches  c * h^k * (u-u.Other())*(v-v.Other()).
"""
import numbers

import torch

from . import ops
from ._lib import XgpuError, default_ctx
from .cutinfo import BitArray, GridFunction, NEG, POS, IF, _lset_vec, _normalize_dtlist

OP_ID, OP_GRAD, OP_DN, OP_JUMP = 0, 1, 2, 3


# ------------------------------------------------------------------ tiny symbolic layer
class MeshSize:
    """specialcf.mesh_size and its integer powers"""

    def __init__(self, power=1):
        self.power = power

    def __pow__(self, k):
        if not isinstance(k, int):
            raise XgpuError("only integer powers of the mesh size are supported")
        return MeshSize(self.power * k)

    def __rtruediv__(self, c):
        return Scalar(float(c), -self.power)

    def __mul__(self, o):
        return Scalar(1.0, self.power) * o

    __rmul__ = __mul__


class specialcf:
    mesh_size = MeshSize()

    @staticmethod
    def normal(dim=None):
        """the facet normal nF of skeleton / element-boundary integrals (outward unit normal of the first element)"""
        return FacetNormal()


class Scalar:
    """c * h^k"""

    def __init__(self, c, hpow=0):
        self.c, self.hpow = float(c), int(hpow)

    def __mul__(self, o):
        if isinstance(o, numbers.Number):
            return Scalar(self.c * o, self.hpow)
        if isinstance(o, Scalar):
            return Scalar(self.c * o.c, self.hpow + o.hpow)
        if isinstance(o, MeshSize):
            return Scalar(self.c, self.hpow + o.power)
        return o.__rmul__(self)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, MeshSize):
            return Scalar(self.c, self.hpow - o.power)
        return Scalar(self.c / float(o), self.hpow)


class Normal:
    """n = Normalize(grad(lsetp1)), or its negative (the outward normals of mlset.DomainTypeArray.GetOuterNormals,
    python/mlset.py:436-440): the sign goes into the coefficient of grad(u) * n"""

    def __init__(self, lset, sign=1.0):
        self.lset, self.sign = lset, float(sign)

    def __neg__(self):
        return Normal(self.lset, -self.sign)


class GradLset:
    def __init__(self, lset):
        self.lset = lset


def Normalize(g):
    if isinstance(g, GradLset):
        return Normal(g.lset)
    raise XgpuError("Normalize: only Normalize(grad(lsetp1)) is supported")


class Proxy:
    """trial / test function with an operator"""

    def __init__(self, space, is_test, op=OP_ID, scal=None, other=False, nlset=None):
        self.space, self.is_test, self.op, self.other = space, is_test, op, other
        self.scal = scal or Scalar(1.0)
        self.nlset = nlset   # level set of the normal in grad(u)*n (OP_DN)

    def Other(self):
        if self.op != OP_ID:
            raise XgpuError(".Other() is supported on u / v only")
        return Proxy(self.space, self.is_test, OP_ID, self.scal, other=True)

    def _with(self, op=None, scal=None, nlset=None):
        return Proxy(self.space, self.is_test, self.op if op is None else op, scal or self.scal, self.other,
                     nlset if nlset is not None else self.nlset)

    def __neg__(self):
        return self._with(scal=self.scal * -1.0)

    def __add__(self, o):
        if isinstance(o, (Proxy, FacetExpr)) and not isinstance(o, Form):
            return FacetExpr.of(self) + o
        return NotImplemented

    def __sub__(self, o):
        if isinstance(o, FacetExpr) or (isinstance(o, Proxy) and OP_FDN in (self.op, o.op)):
            return FacetExpr.of(self) - o
        if isinstance(o, Proxy) and o.other and not self.other and self.op == OP_ID and o.op == OP_ID \
                and o.is_test == self.is_test:
            if (self.scal.c, self.scal.hpow) != (o.scal.c, o.scal.hpow):   # 2*u - u.Other() is not a scaled jump
                raise XgpuError("unsupported difference of proxies: both sides of u - u.Other() must carry the same factor")
            return Proxy(self.space, self.is_test, OP_JUMP, self.scal)
        raise XgpuError("unsupported difference of proxies (menu: u - u.Other())")

    def __mul__(self, o):
        if isinstance(o, (numbers.Number, Scalar, MeshSize)):
            s = o if isinstance(o, Scalar) else (Scalar(float(o)) if isinstance(o, numbers.Number) else Scalar(1.0, o.power))
            return self._with(scal=self.scal * s)
        if isinstance(o, Normal):
            if self.op != OP_GRAD or self.other:
                raise XgpuError("only grad(u)*n is supported with the level-set normal")
            return self._with(op=OP_DN, nlset=o.lset, scal=self.scal * o.sign)
        if isinstance(o, FacetNormal):
            if self.op != OP_GRAD:
                raise XgpuError("only grad(u)*nF is supported with the facet normal")
            return self._with(op=OP_FDN)
        if isinstance(o, FacetExpr):
            return o * self
        if isinstance(o, Proxy) and (self.other or o.other or OP_FDN in (self.op, o.op)
                                     or (OP_JUMP in (self.op, o.op) and self.op != o.op and OP_GRAD not in (self.op, o.op))):
            return FacetExpr.of(self) * o   # traces on a facet: (u - u.Other()) * v, grad(u)*nF * (v - v.Other()), ...
        if isinstance(o, Proxy):
            return Form.from_pair(self, o)
        if isinstance(o, Form):
            return o.__rmul__(self)
        if isinstance(o, CoordMonomial):
            return LForm.from_proxy(self) * o
        if isinstance(o, CutDifferentialSymbol):
            return Integral(LForm.from_proxy(self), o)
        return NotImplemented

    __rmul__ = __mul__


# ------------------------------------------------------------------ facet integrands (dCut(..., skeleton / element_boundary))
OP_FDN = 4   # nF . grad(u): derivative along the facet normal of the first element (specialcf.normal / dn(u, 1))


class FacetNormal:
    """specialcf.normal(dim) inside skeleton / element-boundary integrals: outward unit normal of the first element"""


class FacetExpr:
    """linear combination of traces of ONE trial or test function on a facet: [(kind, side, c, hpow)], kind 0 = value,
    1 = nF . grad, side 0 = this element, 1 = .Other() — e.g. the DG jump u - u.Other() and the mean flux
    0.5 (grad(u) + grad(u.Other())) * nF of demos/fictdom_dg.py:108-125, the derivative jump dn(u,1) - dn(u.Other(),1)"""

    def __init__(self, space, is_test, items):
        self.space, self.is_test, self.items = space, is_test, items

    @staticmethod
    def of(x):
        if isinstance(x, FacetExpr):
            return x
        if isinstance(x, Proxy):
            s, side = x.scal, (1 if x.other else 0)
            if x.op == OP_ID:
                return FacetExpr(x.space, x.is_test, [(0, side, s.c, s.hpow)])
            if x.op == OP_FDN:
                return FacetExpr(x.space, x.is_test, [(1, side, s.c, s.hpow)])
            if x.op == OP_GRAD:   # a gradient waiting for the facet normal: (grad(u) + grad(u.Other())) * nF
                return FacetExpr(x.space, x.is_test, [(-1, side, s.c, s.hpow)])
            if x.op == OP_JUMP:
                return FacetExpr(x.space, x.is_test, [(0, 0, s.c, s.hpow), (0, 1, -s.c, s.hpow)])
        raise XgpuError("facet integrands: menu is u, u.Other(), grad(u)*nF, grad(u.Other())*nF and their linear combinations")

    def __add__(self, o):
        o = FacetExpr.of(o)
        if o.is_test != self.is_test:
            raise XgpuError("cannot add a trial and a test function")
        return FacetExpr(self.space, self.is_test, self.items + o.items)

    __radd__ = __add__

    def __neg__(self):
        return self * -1.0

    def __sub__(self, o):
        return self + (-FacetExpr.of(o))

    def __rsub__(self, o):
        return FacetExpr.of(o) + (-self)

    def __mul__(self, o):
        if isinstance(o, (numbers.Number, Scalar, MeshSize)):
            s = o if isinstance(o, Scalar) else (Scalar(float(o)) if isinstance(o, numbers.Number) else Scalar(1.0, o.power))
            return FacetExpr(self.space, self.is_test, [(k, sd, c * s.c, hp + s.hpow) for k, sd, c, hp in self.items])
        if isinstance(o, FacetNormal):
            if any(k != -1 for k, _, _, _ in self.items):
                raise XgpuError("only gradients can be multiplied with the facet normal")
            return FacetExpr(self.space, self.is_test, [(1, sd, c, hp) for _, sd, c, hp in self.items])
        if isinstance(o, (Proxy, FacetExpr)):
            o = FacetExpr.of(o)
            if any(k == -1 for k, _, _, _ in self.items + o.items):
                raise XgpuError("facet integrands: a gradient must be multiplied with the facet normal nF")
            if o.is_test == self.is_test:
                raise XgpuError("a bilinear form needs one trial and one test function per term")
            u, v = (o, self) if self.is_test else (self, o)
            return FacetForm([(tk, ts, ek, es, c1 * c2, h1 + h2) for tk, ts, c1, h1 in u.items for ek, es, c2, h2 in v.items], u.space)
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, MeshSize):
            return self * Scalar(1.0, -o.power)
        return self * (1.0 / float(o))


class FacetForm:
    """bilinear facet integrand: [(trial kind, trial side, test kind, test side, coef, hpow)] (xgpu_fterm)"""

    def __init__(self, terms, space, cf=None):
        self.terms, self.space, self.nlset, self.cf = terms, space, None, cf

    def _with_cf(self, cf):
        """cf(x, y[, z]) * integrand: evaluated at the physical points of the cut facet rules and folded into their weights
        (a coefficient in the integrand of T_CalcFacetMatrix, xfem/symboliccutbfi.cpp:1040-1100)"""
        return FacetForm(self.terms, self.space, cf if self.cf is None else self.cf * cf)

    @staticmethod
    def of(f):
        if isinstance(f, FacetForm):
            return f
        if isinstance(f, Form):   # u*v and (u-u.Other())*(v-v.Other()) written with plain proxies
            out = []
            for a, b, c, hp in f.terms:
                if a == OP_ID and b == OP_ID:
                    out.append((0, 0, 0, 0, c, hp))
                elif a == OP_JUMP and b == OP_JUMP:
                    out += [(0, 0, 0, 0, c, hp), (0, 1, 0, 0, -c, hp), (0, 0, 0, 1, -c, hp), (0, 1, 0, 1, c, hp)]
                else:
                    raise XgpuError("facet integrands: grad(u)*grad(v) / level-set normals are not traces on a facet (menu: values, nF . grad)")
            return FacetForm(out, f.space, f.cf)
        raise XgpuError("not a facet integrand")

    def __add__(self, o):
        o = FacetForm.of(o)
        return FacetForm(self.terms + o.terms, self.space, _merge_cf(self.cf, o.cf))

    __radd__ = __add__

    def __neg__(self):
        return self * -1.0

    def __sub__(self, o):
        return self + (-FacetForm.of(o))

    def __mul__(self, o):
        if isinstance(o, (numbers.Number, Scalar, MeshSize)):
            s = o if isinstance(o, Scalar) else (Scalar(float(o)) if isinstance(o, numbers.Number) else Scalar(1.0, o.power))
            return FacetForm([(a, b, c_, d, c * s.c, hp + s.hpow) for a, b, c_, d, c, hp in self.terms], self.space, self.cf)
        if isinstance(o, CutDifferentialSymbol):
            if not (o.skeleton or o.element_boundary):
                raise XgpuError("facet integrands need dCut(..., skeleton=True) or dCut(..., element_boundary=True)")
            return Integral(self, o)
        return NotImplemented

    __rmul__ = __mul__


def dn(u, order=1, comp=-1, hdiv=False):
    """xfem/ghostpenalty.cpp (DiffOpDuDnk): normal derivative across a facet.  Accelerated path: the first derivative,
    dn(u, 1) = grad(u) * nF; higher orders raise."""
    if order != 1 or comp != -1 or hdiv:
        raise XgpuError("dn: only the first normal derivative of a scalar H1 function is on the accelerated path")
    return grad(u) * FacetNormal()


class CoefficientFunction:
    """a coefficient evaluated at every quadrature point — NGSolve's CoefficientFunction on the accelerated path
    (xfem/symboliccutbfi.cpp:174-193, xfem/symboliccutlfi.cpp:96-118): fn(x, y[, z], *gf_values) -> tensor, evaluated with
    torch on the device at the PHYSICAL points of the cut rules (deformed mesh included, xgpu_rule_phys) and folded into
    the rule weights, so that one coefficient multiplies a whole integrand:  cf * (grad(u)*grad(v) + u*v) * dx,  cf * v * dx.
    gfs: H1 GridFunctions as (space, coefficient vector) pairs; their values at the points follow the coordinates."""

    def __init__(self, fn, gfs=()):
        if isinstance(fn, numbers.Number):
            c = float(fn)
            fn = lambda *a: torch.full_like(a[0], c)   # noqa: E731
        self.factors = [(fn, tuple(gfs))]   # a product of functions, each with its own GridFunction arguments
        self.scale = 1.0

    def _copy(self, factors=None, scale=None):
        r = CoefficientFunction.__new__(CoefficientFunction)
        r.factors = list(self.factors if factors is None else factors)
        r.scale = self.scale if scale is None else scale
        return r

    def __mul__(self, o):
        if isinstance(o, numbers.Number):
            return self._copy(scale=self.scale * float(o))
        if isinstance(o, CoefficientFunction):
            return self._copy(self.factors + o.factors, self.scale * o.scale)
        if isinstance(o, Proxy):
            if not o.is_test:
                raise XgpuError("a coefficient multiplies a whole integrand: cf * (u*v), cf * v")
            o = LForm.from_proxy(o)
        if isinstance(o, (Form, LForm)):
            return o._with_cf(self)
        if hasattr(o, "_st_with_cf"):   # space-time proxies / forms (spacetime.py): fn(x, y[, z], tref)
            return o._st_with_cf(self)
        if isinstance(o, FacetForm):
            return o._with_cf(self)
        if isinstance(o, FacetExpr):
            raise XgpuError("a coefficient multiplies a whole integrand: cf * (jump(u) * jump(v))")
        return NotImplemented

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0

    def evaluate_st(self, dim, fe_mesh, elems, off, pts, tv, ctx):
        """space-time rules: fn(x, y[, z], tref) at the points (x_q, tref_q) — the coordinates are physical, the time is the
        REFERENCE time of the slab as in ngsxfem (ReferenceTimeVariable, python/__init__.py:86-99; t = told + delta_t * tref
        is formed by the caller's function).  GridFunction arguments are spatial objects and are not taken here."""
        xyz = ops.rule_phys(dim, elems, fe_mesh.elements, fe_mesh.points, off, pts, ctx=ctx)
        args = [xyz[:, c] for c in range(dim)] + [tv]
        out = torch.full((xyz.shape[0],), self.scale, dtype=torch.float64, device=xyz.device)
        for fn, gfs in self.factors:
            if gfs:
                raise XgpuError("coefficient functions on space-time integrands: functions of (x, y[, z], tref) only")
            out = out * torch.as_tensor(fn(*args), dtype=torch.float64, device=xyz.device)
        return out.contiguous()

    def evaluate_facets(self, fe_mesh, facet_vtx, off, pts):
        """cut rules on a table of facets ((dim-1)-simplices in R^dim): fn(x, y[, z]) at x = sum_k lambda_k P[v_k], the
        barycentric weights of a rule point being (xi_0, .., xi_{dim-2}, 1 - sum xi) (reference vertex k <-> e_k, last <-> 0)"""
        d1 = facet_vtx.shape[1]                                         # vertices per facet = dim
        cnt = off[1:] - off[:-1]
        f_of_pt = torch.repeat_interleave(torch.arange(facet_vtx.shape[0], device=pts.device), cnt)
        lam = torch.cat([pts[:, :d1 - 1], 1.0 - pts[:, :d1 - 1].sum(dim=1, keepdim=True)], dim=1)
        X = fe_mesh.points[facet_vtx.long()][f_of_pt]                    # (npts, dim, dim)
        xyz = (lam.unsqueeze(2) * X).sum(dim=1)
        out = torch.full((xyz.shape[0],), self.scale, dtype=torch.float64, device=xyz.device)
        for fn, gfs in self.factors:
            if gfs:
                raise XgpuError("coefficient functions on facet integrands: functions of the coordinates only")
            out = out * torch.as_tensor(fn(*[xyz[:, c] for c in range(xyz.shape[1])]), dtype=torch.float64, device=xyz.device)
        return out.contiguous()

    def evaluate(self, dim, fe_mesh, elems, off, pts, ctx):
        """values at the points of the per-element rules (off, pts) on the listed elements"""
        xyz = ops.rule_phys(dim, elems, fe_mesh.elements, fe_mesh.points, off, pts, ctx=ctx)
        coords = [xyz[:, c] for c in range(dim)]
        out = torch.full((xyz.shape[0],), self.scale, dtype=torch.float64, device=xyz.device)
        for fn, gfs in self.factors:
            args = list(coords)
            for space, vec in gfs:
                args.append(ops.h1_eval_rule(dim, space.order, elems, fe_mesh.elements, space.el_dofs, vec, off, pts, ctx=ctx))
            out = out * torch.as_tensor(fn(*args), dtype=torch.float64, device=xyz.device)
        return out.contiguous()


class ElementCF(CoefficientFunction):
    """a piecewise constant coefficient: one value per mesh element (the reference's BitArrayCF / IndicatorCF(mesh, ba) and
    L2(order=0) GridFunctions such as CutRatioGF, python/__init__.py:441-462).  At the rule points of an element it is the
    element's value, so it folds into the weights like any coefficient: cf * (u*v) * dCut(...), cf * v * dCut(...),
    Integrate(..., cf=cf).  Products: with numbers and with other ElementCFs (pointwise coefficients and element-wise ones do
    not mix in one product: write two integrals or fold the element values into the callable's GridFunction arguments)."""

    def __init__(self, values, scale=1.0):
        self.values = torch.as_tensor(values, dtype=torch.float64).contiguous()
        if self.values.dim() != 1:
            raise XgpuError("ElementCF: one value per element")
        self.factors = []
        self.scale = float(scale)

    def _copy(self, factors=None, scale=None):
        return ElementCF(self.values, self.scale if scale is None else scale)

    def __mul__(self, o):
        if isinstance(o, numbers.Number):
            return ElementCF(self.values, self.scale * float(o))
        if isinstance(o, ElementCF):
            if o.values.shape != self.values.shape:
                raise XgpuError("ElementCF * ElementCF: different numbers of elements")
            return ElementCF(self.values * o.values.to(self.values.device), self.scale * o.scale)
        if isinstance(o, CoefficientFunction):
            raise XgpuError("a pointwise CoefficientFunction and an element-wise one do not mix in one product")
        return CoefficientFunction.__mul__(self, o)       # integrands: cf * form

    __rmul__ = __mul__

    def _per_point(self, fe_mesh, elems, off):
        if self.values.numel() != fe_mesh.ne:
            raise XgpuError("ElementCF: the number of values is not the number of mesh elements")
        v = self.values.to(off.device)
        if elems is not None:
            v = v[elems.long()]
        return (self.scale * torch.repeat_interleave(v, off[1:] - off[:-1])).contiguous()

    def evaluate(self, dim, fe_mesh, elems, off, pts, ctx):
        return self._per_point(fe_mesh, elems, off)

    def evaluate_st(self, dim, fe_mesh, elems, off, pts, tv, ctx):
        return self._per_point(fe_mesh, elems, off)

    def evaluate_facets(self, fe_mesh, facet_vtx, off, pts):
        raise XgpuError("element-wise coefficients on facet integrands are not on the accelerated path")


def _merge_cf(a, b):
    if a is b:
        return a
    raise XgpuError("integrands with different coefficient functions: add them as separate integrals (a += cf1*...*dx; a += cf2*...*dx)")


class LForm:
    """linear form integrand: sum of (op_test, coef, hpow, expo) = coef h^hpow x^expo op_test(v)
    (menu of xfem/symboliccutlfi.cpp's CoefficientFunction on the accelerated path)"""

    def __init__(self, terms, space, nlset=None, cf=None):
        self.terms, self.space, self.nlset, self.cf = terms, space, nlset, cf

    def _with_cf(self, cf):
        return LForm(self.terms, self.space, self.nlset, cf if self.cf is None else self.cf * cf)

    @staticmethod
    def from_proxy(v):
        if not v.is_test:
            raise XgpuError("a linear form needs a test function")
        if v.op not in (OP_ID, OP_DN) or v.other:
            raise XgpuError("linear forms: menu is f*v and f*(grad(v)*n)")
        return LForm([(v.op, v.scal.c, v.scal.hpow, (0, 0, 0))], v.space, v.nlset)

    def __add__(self, o):
        if isinstance(o, Proxy):
            o = LForm.from_proxy(o)
        if isinstance(o, LForm):
            return LForm(self.terms + o.terms, self.space, Form._merge_nlset(self.nlset, o.nlset), _merge_cf(self.cf, o.cf))
        return NotImplemented

    def __mul__(self, o):
        if isinstance(o, numbers.Number):
            o = Scalar(float(o))
        if isinstance(o, MeshSize):
            o = Scalar(1.0, o.power)
        if isinstance(o, Scalar):
            return LForm([(a, c * o.c, hp + o.hpow, ex) for a, c, hp, ex in self.terms], self.space, self.nlset, self.cf)
        if isinstance(o, CoordMonomial):
            return LForm([(a, c * o.coef, hp, tuple(e1 + e2 for e1, e2 in zip(ex, o.expo))) for a, c, hp, ex in self.terms],
                         self.space, self.nlset, self.cf)
        if isinstance(o, CutDifferentialSymbol):
            return Integral(self, o)
        return NotImplemented

    __rmul__ = __mul__


def TrialFunction(space):
    return Proxy(space, False)


def TestFunction(space):
    return Proxy(space, True)


def grad(u):
    if hasattr(u, "_st_grad"):  # proxies of a SpaceTimeFESpace (spacetime.py): the space gradient
        return u._st_grad()
    if isinstance(u, Proxy):
        if u.op != OP_ID:
            raise XgpuError("grad of this proxy is not supported")
        return u._with(op=OP_GRAD)   # grad(u.Other()) only makes sense times the facet normal (checked where it is used)
    if isinstance(u, GridFunction):
        return GradLset(u)
    raise XgpuError("grad: unsupported argument")


class Form:
    """sum of monomials (op_trial, op_test, coef, hpow)"""

    def __init__(self, terms, space, nlset=None, cf=None):
        self.terms, self.space, self.nlset, self.cf = terms, space, nlset, cf

    def _with_cf(self, cf):
        return Form(self.terms, self.space, self.nlset, cf if self.cf is None else self.cf * cf)

    @staticmethod
    def _merge_nlset(a, b):
        if a is None or b is None or a is b:
            return a if a is not None else b
        raise XgpuError("one integrand may use the normal of ONE level set only (split it into several integrals)")

    @staticmethod
    def from_pair(a, b):
        if a.is_test == b.is_test:
            raise XgpuError("a bilinear form needs one trial and one test function per term")
        u, v = (b, a) if a.is_test else (a, b)
        if (u.other or v.other) and not (u.op == OP_ID and v.op == OP_ID):
            raise XgpuError(".Other() appears in jumps (u - u.Other()) and facet traces only")
        if (u.op == OP_GRAD) != (v.op == OP_GRAD):
            raise XgpuError("grad(u) must be paired with grad(v) (dot product) or the normal n")
        if (u.op == OP_JUMP) != (v.op == OP_JUMP):
            raise XgpuError("jump terms must pair (u-u.Other()) with (v-v.Other())")
        s = u.scal * v.scal
        return Form([(u.op, v.op, s.c, s.hpow)], u.space, Form._merge_nlset(u.nlset, v.nlset))

    def __add__(self, o):
        if isinstance(o, FacetForm):
            return FacetForm.of(self) + o
        if isinstance(o, Form):
            return Form(self.terms + o.terms, self.space, Form._merge_nlset(self.nlset, o.nlset), _merge_cf(self.cf, o.cf))
        return NotImplemented

    def __sub__(self, o):
        return self + (-1.0) * o

    def __neg__(self):
        return (-1.0) * self

    def __mul__(self, o):
        if isinstance(o, (numbers.Number, Scalar, MeshSize)):
            s = o if isinstance(o, Scalar) else (Scalar(float(o)) if isinstance(o, numbers.Number) else Scalar(1.0, o.power))
            return Form([(a, b, c * s.c, hp + s.hpow) for a, b, c, hp in self.terms], self.space, self.nlset, self.cf)
        if isinstance(o, CutDifferentialSymbol) and (o.skeleton or o.element_boundary):
            return Integral(FacetForm.of(self), o)
        if isinstance(o, (CutDifferentialSymbol, FacetPatchDifferentialSymbol)):
            return Integral(self, o)
        return NotImplemented

    __rmul__ = __mul__


# ------------------------------------------------------------------ differential symbols
class CutDifferentialSymbol:
    def __init__(self, levelset, domain_type, order, definedonelements, deformation, time_order=-1, tref=None, vb=0,
                 skeleton=False, element_boundary=False):
        self.levelset, self.domain_type, self.order = levelset, domain_type, order
        self.definedonelements, self.deformation = definedonelements, deformation
        self.time_order, self.tref, self.vb = time_order, tref, vb
        self.skeleton, self.element_boundary = bool(skeleton), bool(element_boundary)


def dCut(levelset, domain_type, order=None, subdivlvl=None, time_order=-1, levelset_domain=None,
         definedonelements=None, deformation=None, definedon=None, vb=0, element_boundary=False,
         element_vb=0, skeleton=False, tref=None, **kwargs):
    """python/__init__.py:781-864"""
    if subdivlvl not in (None, 0):
        raise XgpuError("subdivlvl > 0 selects ngsxfem's legacy subdivision path: not provided")
    if element_vb != 0 or vb not in (0, 1):
        raise XgpuError("element_vb / vb: not on the accelerated path")
    if element_boundary or skeleton:
        # cut integrals over interior facets / element boundaries (xfem/symboliccutbfi.cpp:886-1290, :488-883)
        if element_boundary and skeleton:
            raise XgpuError("dCut: skeleton and element_boundary exclude each other (python/__init__.py:826-830)")
        if vb != 0 or deformation is not None or isinstance(levelset, (list, tuple)) or int(domain_type) == IF \
                or time_order >= 0 or tref is not None:
            raise XgpuError("dCut(skeleton / element_boundary): one level set, NEG / POS, affine mesh, no time slab "
                            "(the codim-2 IF case and deformed facets are not on the accelerated path)")
        if definedonelements is None:
            raise XgpuError("dCut(skeleton / element_boundary) needs definedonelements (facets / elements)")
    if vb == 1 and (deformation is not None or isinstance(levelset, (list, tuple))):
        raise XgpuError("dCut(vb=BND): single level set, no deformation (tests/pytests/test_bndcut.py)")
    if kwargs:
        raise XgpuError(f"unsupported dCut arguments: {sorted(kwargs)}")
    if time_order >= 0 or tref is not None:
        # space-time rules (cutint/xintegration.cpp:41-90): assembled by spacetime.py on a SpaceTimeFESpace
        if vb != 0 or deformation is not None or isinstance(levelset, (list, tuple)):
            raise XgpuError("space-time dCut: one level set, volume elements, no deformation")
    return CutDifferentialSymbol(levelset, domain_type, -1 if order is None else int(order), definedonelements, deformation,
                                 time_order=int(time_order), tref=tref, vb=vb, skeleton=skeleton, element_boundary=element_boundary)


class FacetPatchDifferentialSymbol:
    def __init__(self, definedonelements, order, deformation=None, time_order=-1, tref=None, downscale=None):
        self.definedonelements, self.order, self.deformation = definedonelements, order, deformation
        self.time_order, self.tref, self.downscale = time_order, tref, downscale


def dFacetPatch(definedonelements=None, deformation=None, time_order=-1, tref=None, downscale=None,
                bonus_intorder=0, **kwargs):
    """python/__init__.py:749-778; bonus_intorder is accepted and ignored as in the reference.  downscale: the patch rule
    of each element is scaled towards the common facet (SetIRScaling, symboliccutbfi.cpp:1849-1885, :2019-2033)"""
    if downscale is not None and not (0.0 < float(downscale) <= 1.0):
        raise XgpuError("SymbolicFacetPatchBFI: Only 0 < downscale <= 1 valid")
    if kwargs:
        raise XgpuError(f"unsupported dFacetPatch arguments: {sorted(kwargs)}")
    if definedonelements is None:
        raise XgpuError("dFacetPatch needs definedonelements (a BitArray over the facets)")
    if (time_order >= 0 or tref is not None) and deformation is not None:
        raise XgpuError("space-time dFacetPatch does not support a mesh deformation")
    return FacetPatchDifferentialSymbol(definedonelements, -1, deformation, time_order=int(time_order), tref=tref,
                                        downscale=None if downscale is None else float(downscale))


class Integral:
    def __init__(self, form, dx):
        self.form, self.dx = form, dx

    def __add__(self, o):
        return SumOfIntegrals([self]) + o


class SumOfIntegrals:
    def __init__(self, ints):
        self.ints = ints

    def __add__(self, o):
        if isinstance(o, Integral):
            return SumOfIntegrals(self.ints + [o])
        return SumOfIntegrals(self.ints + o.ints)


# ------------------------------------------------------------------ sparse matrix result
class SparseMatrixCSR:
    def __init__(self, ndof, rowptr, col, vals):
        self.ndof, self.rowptr, self.col, self.vals = ndof, rowptr, col, vals

    @property
    def nze(self):
        return self.col.numel()

    def CSR(self):
        return self.vals, self.col, self.rowptr

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals.cpu().numpy(), self.col.cpu().numpy(), self.rowptr.cpu().numpy()),
                             shape=(self.ndof, self.ndof))

    def __mul__(self, x):
        return ops.csr_spmv(self.rowptr, self.col, self.vals, x)


# ------------------------------------------------------------------ bilinear forms
class BilinearForm:
    """BilinearForm(space) / RestrictedBilinearForm(space, element_restriction, facet_restriction)

    Assemble() = graph (GetGraph) + batched element / facet-patch matrices + deterministic
    scatter, everything on the GPU.  Sparsity: union of the dof cliques of the active elements
    (element_restriction, else the elements carrying dofs) and, if the space has dgjumps, of the
    dof unions across the active facets (facet_restriction, else all interior facets between
    active elements) — utils/restrictedblf.cpp:8-163, SURVEY.md §9.14."""

    def __init__(self, space, symmetric=False, check_unused=True, element_restriction=None, facet_restriction=None,
                 name="bfa", **flags):
        if symmetric:
            raise XgpuError("symmetric storage is not provided (the cut BFI never symmetrises, symboliccutbfi.cpp:56)")
        self.space = space
        self.element_restriction, self.facet_restriction = element_restriction, facet_restriction
        self.integrals = []
        self.mat = None
        self.timings = {}

    def __iadd__(self, o):
        if isinstance(o, Integral):
            self.integrals.append(o)
        elif isinstance(o, SumOfIntegrals):
            self.integrals.extend(o.ints)
        else:
            raise XgpuError("BilinearForm += expects form * dCut(...) or form * dFacetPatch(...)")
        return self

    def _assemble_facet_integrals(self, skel, tables, mats, ctx):
        """dCut(lset, NEG|POS, skeleton=True, definedonelements=facets): one (2 nd) x (2 nd) matrix per interior facet over
        the dofs of both neighbours (T_CalcFacetMatrix, xfem/symboliccutbfi.cpp:886-1290); element_boundary=True: one nd x nd
        matrix per element, summed over its facets (T_CalcElementMatrixEBAdd, :488-883).  The cut rule of a facet is
        xgpu_cutquad on the facet-vertex table, order 2 * fe order (:978, :1029) unless given.  returns the BitArray of the
        facets that carry a skeleton integral (their couplings are in the graph already)."""
        sp, mesh = self.space, self.space.mesh
        dim, p, dev = mesh.dim, sp.order, mesh.device
        if dim < 2:
            raise XgpuError("facet integrals: 2D / 3D meshes")
        ents, el_f, fe = mesh.facets()
        covered = None
        for integ in skel:
            dx = integ.dx
            form = FacetForm.of(integ.form)
            phi = _lset_vec(dx.levelset, mesh)
            order = 2 * p if dx.order < 0 else dx.order
            if dx.skeleton:
                l_f = dx.definedonelements.indices()
                l_f = l_f[fe[l_f.long(), 1] >= 0].contiguous()   # interior facets only (LocalFacetNr2 == -1 throws, :913)
                if l_f.numel() == 0:
                    continue
                fv = ents[l_f.long()].to(torch.int32).contiguous()
                fel = fe[l_f.long()].contiguous()
                off, q, w = ops.cutquad(dim - 1, None, fv, mesh.points, phi, int(dx.domain_type), order, ctx=ctx)
                ctx.raise_on_geometry_error()
                if form.cf is not None and w.numel():
                    w = w * form.cf.evaluate_facets(mesh, fv, off, q)
                mm = ops.facet_elmat(dim, p, fv, fel, mesh.elements, mesh.points, (off, q, w), form.terms, 2, ctx=ctx)
                t_f = torch.cat([self._rows(fel[:, 0].contiguous()), self._rows(fel[:, 1].contiguous())], dim=1).contiguous()
                tables.append(t_f); mats.append(mm)
                fb = BitArray(mesh.nfacet, mesh.device)
                fb.bits.copy_(BitArray.from_numpy(_mask_from_list(l_f, mesh.nfacet), mesh.device).bits)
                covered = fb if covered is None else (covered | fb)
            else:
                if any(t[1] != 0 or t[3] != 0 for t in form.terms):
                    raise XgpuError("element-boundary integrals see the element's own functions only (no .Other())")
                l_e = dx.definedonelements.indices()
                if l_e.numel() == 0:
                    continue
                nl = dim + 1
                fids = el_f[l_e.long()].reshape(-1).long()                       # (n * (dim+1)) facet ids, element-major
                fv = ents[fids].to(torch.int32).contiguous()
                fel = torch.stack([l_e.repeat_interleave(nl), torch.full((l_e.numel() * nl,), -1, dtype=torch.int32, device=dev)], dim=1).contiguous()
                off, q, w = ops.cutquad(dim - 1, None, fv, mesh.points, phi, int(dx.domain_type), order, ctx=ctx)
                ctx.raise_on_geometry_error()
                if form.cf is not None and w.numel():
                    w = w * form.cf.evaluate_facets(mesh, fv, off, q)
                mm = ops.facet_elmat(dim, p, fv, fel, mesh.elements, mesh.points, (off, q, w), form.terms, 1, ctx=ctx)
                nd = sp.nd
                tables.append(self._rows(l_e.contiguous())); mats.append(mm.reshape(l_e.numel(), nl, nd, nd).sum(dim=1).contiguous())
        return covered

    # ---- helpers
    def _active_elements(self):
        if self.element_restriction is not None:
            return self.element_restriction
        act = getattr(self.space, "active_elements", None)
        if act is not None:
            return BitArray(self.space.mesh.ne, words=act)
        return BitArray(self.space.mesh.ne, self.space.mesh.device).Set()

    def _rows(self, elem_list):
        sp = self.space
        if hasattr(sp, "rows"):
            return sp.rows(elem_list)
        return ops.gather_rows(sp.el_dofs, elem_list, None)

    def _intorder(self, integ):
        """symboliccutbfi.cpp:100-118: ord_trial + ord_test - (min diff orders), overridden by the domain's order"""
        order = integ.dx.order
        if order < 0:
            terms = integ.form.terms
            dtr = min(0 if t[0] == OP_ID else 1 for t in terms)
            dte = min(0 if t[1] == OP_ID else 1 for t in terms)
            order = 2 * self.space.order - (dtr + dte)
        return order

    def _assemble_general(self, vol, active, tables, mats, ctx):
        """several level sets / multi-level-set domains (cutint/mlsetintegration.cpp:6-186): ONE batch over
        the active elements; every integrator builds its per-element rules on (definedonelements & active)
        — uncut relevant elements get the standard rule, irrelevant ones no points — and accumulates."""
        sp = self.space
        mesh = sp.mesh
        dim, p, ne, nd, dev = mesh.dim, sp.order, mesh.ne, sp.nd, mesh.device
        l_A = active.indices()
        if not l_A.numel():
            return
        slot_A = ops.bitset_rank(active.bits, ne, ctx=ctx)[0]
        m_A = torch.zeros((l_A.numel(), nd, nd), dtype=torch.float64, device=dev)
        bnd_batches = []
        for integ in vol:
            dx = integ.dx
            terms = list(integ.form.terms)
            if any(t[0] == OP_JUMP for t in terms):
                raise XgpuError("jump terms belong on dFacetPatch")
            order = self._intorder(integ)
            don = dx.definedonelements
            if dx.vb == 1:
                # boundary elements (dCut(..., vb=BND)): their cut rules are carried into the adjacent volume element;
                # one matrix per BOUNDARY element in its own batch (a volume element may own several boundary facets)
                phi = _lset_vec(dx.levelset, mesh)
                vol_els, off, q, w = _bnd_rule_in_volume(mesh, phi, int(dx.domain_type), order, ctx, facet_mask=don)
                keep = torch.nonzero(((active.bits[(vol_els >> 5).long()] >> (vol_els & 31)) & 1).bool()).reshape(-1)
                if keep.numel() != vol_els.numel():
                    raise XgpuError("dCut(vb=BND): a boundary element of an inactive volume element (restrict definedonelements)")
                if w.numel():
                    phi_n = _lset_vec(integ.form.nlset, mesh) if integ.form.nlset is not None else phi
                    m_B = ops.elmat_rule(dim, p, vol_els, mesh.elements, mesh.points, phi_n, off, q, w, terms, ctx=ctx)
                    bnd_batches.append((self._rows(vol_els), m_B))
                continue
            if don is None:
                l, slots = l_A, None
            else:
                sel = ops.bitset_op(ops.BIT_AND, active.bits, don.bits, ne, ctx=ctx)
                l = ops.bitset_to_list(sel, ne, ctx=ctx)
                slots = slot_A[l.long()].contiguous()
            if not l.numel():
                continue
            ls = dx.levelset
            with ctx.deformation(_eff_def(dx.deformation, mesh)):
                has_dn = any(t[0] == OP_DN or t[1] == OP_DN for t in terms)
                if isinstance(ls, (list, tuple)):
                    if has_dn and integ.form.nlset is None:
                        raise XgpuError("normal-derivative terms on a multi-level-set domain: build n from ONE of the level sets")
                    phis = torch.stack([_lset_vec(g, mesh) for g in ls]).contiguous()
                    dts = _normalize_dtlist(dx.domain_type, len(ls))
                    off, q, w = ops.cutquad_ml(dim, l, mesh.elements, mesh.points, phis, dts, order, ctx=ctx)
                    phi = phis[0].contiguous()
                else:
                    phi = _lset_vec(ls, mesh)
                    off, q, w = ops.cutquad(dim, l, mesh.elements, mesh.points, phi, int(dx.domain_type), order, ctx=ctx)
                # n = Normalize(grad(lset)) of the level set the integrand names (square.GetOuterNormals style forms)
                phi_n = _lset_vec(integ.form.nlset, mesh) if (has_dn and integ.form.nlset is not None) else phi
                ctx.raise_on_geometry_error()
                if w.numel() and integ.form.cf is not None:   # tabulated coefficient folded into the weights
                    w = w * integ.form.cf.evaluate(dim, mesh, l, off, q, ctx)
                if w.numel():
                    ops.elmat_rule(dim, p, l, mesh.elements, mesh.points, phi_n, off, q, w, terms, out=m_A, accumulate=True,
                                   slots=slots, ctx=ctx)
        tables.append(self._rows(l_A)); mats.append(m_A)
        for t_B, m_B in bnd_batches:
            tables.append(t_B); mats.append(m_B)

    def _assemble_single_lset(self, vol, active, tables, mats, ctx):
        sp = self.space
        mesh = sp.mesh
        dim, p, ne, nd, dev = mesh.dim, sp.order, mesh.ne, sp.nd, mesh.device
        phi0 = None
        for integ in vol:
            phi = _lset_vec(integ.dx.levelset, mesh)
            if phi0 is None:
                phi0 = phi
            elif phi.data_ptr() != phi0.data_ptr() and not torch.equal(phi, phi0):
                raise XgpuError("all cut integrators of one BilinearForm must use the same level set")
        if phi0 is not None:
            bn, bp, bi, _ = ops.classify(dim, mesh.elements, phi0, want_ratio=False, ctx=ctx)
            ctx.raise_on_geometry_error()
            b_unc = ops.bitset_op(ops.BIT_OR, bn, bp, ne, ctx=ctx)
            U = BitArray(ne, words=ops.bitset_op(ops.BIT_AND, b_unc, active.bits, ne, ctx=ctx))
            Cc = BitArray(ne, words=ops.bitset_op(ops.BIT_AND, bi, active.bits, ne, ctx=ctx))
        else:
            U, Cc = active, BitArray(ne, dev)
        l_U, l_C = U.indices(), Cc.indices()
        slot_U = ops.bitset_rank(U.bits, ne, ctx=ctx)[0] if l_U.numel() else None
        slot_C = ops.bitset_rank(Cc.bits, ne, ctx=ctx)[0] if l_C.numel() else None
        m_U = torch.empty((l_U.numel(), nd, nd), dtype=torch.float64, device=dev)
        m_C = torch.empty((l_C.numel(), nd, nd), dtype=torch.float64, device=dev)
        init_U = init_C = False
        for integ in vol:
            dx = integ.dx
            terms = list(integ.form.terms)
            if any(t[0] == OP_JUMP for t in terms):
                raise XgpuError("jump terms belong on dFacetPatch")
            dt = int(dx.domain_type)
            order = dx.order
            if order < 0:  # symboliccutbfi.cpp:100-118
                dtr = min(0 if t[0] == OP_ID else 1 for t in terms)
                dte = min(0 if t[1] == OP_ID else 1 for t in terms)
                order = 2 * p - (dtr + dte)
            don = dx.definedonelements
            if dt != IF and l_U.numel():
                sel = ops.bitset_op(ops.BIT_AND, bn if dt == NEG else bp, U.bits, ne, ctx=ctx)
                if don is not None:
                    sel = ops.bitset_op(ops.BIT_AND, sel, don.bits, ne, ctx=ctx)
                l = ops.bitset_to_list(sel, ne, ctx=ctx)
                if l.numel():
                    full = l.numel() == l_U.numel()
                    if not init_U and not full:
                        m_U.zero_()
                    with ctx.deformation(_eff_def(dx.deformation, mesh)):
                        if integ.form.cf is None:
                            ops.elmat_std(dim, p, l, mesh.elements, mesh.points, phi0, order, terms, out=m_U,
                                          accumulate=init_U or not full, slots=None if full else slot_U[l.long()].contiguous(), ctx=ctx)
                        else:   # a coefficient varies from point to point: per-element copies of the standard rule, weights scaled
                            off, q, w = ops.cutquad(dim, l, mesh.elements, mesh.points, phi0, dt, order, ctx=ctx)
                            w = w * integ.form.cf.evaluate(dim, mesh, l, off, q, ctx)
                            ops.elmat_rule(dim, p, l, mesh.elements, mesh.points, phi0, off, q, w, terms, out=m_U,
                                           accumulate=init_U or not full, slots=None if full else slot_U[l.long()].contiguous(), ctx=ctx)
                    init_U = True
            if l_C.numel():
                sel = Cc.bits if don is None else ops.bitset_op(ops.BIT_AND, Cc.bits, don.bits, ne, ctx=ctx)
                l = ops.bitset_to_list(sel, ne, ctx=ctx)
                if l.numel():
                    full = l.numel() == l_C.numel()
                    if not init_C and not full:
                        m_C.zero_()
                    with ctx.deformation(_eff_def(dx.deformation, mesh)):
                        off, q, w = ops.cutquad(dim, l, mesh.elements, mesh.points, phi0, dt, order, ctx=ctx)
                        ctx.raise_on_geometry_error()
                        if integ.form.cf is not None:
                            w = w * integ.form.cf.evaluate(dim, mesh, l, off, q, ctx)
                        ops.elmat_rule(dim, p, l, mesh.elements, mesh.points, phi0, off, q, w, terms, out=m_C,
                                       accumulate=init_C or not full, slots=None if full else slot_C[l.long()].contiguous(), ctx=ctx)
                    init_C = True
        if not init_U:
            m_U.zero_()
        if not init_C:
            m_C.zero_()
        if l_U.numel():
            tables.append(self._rows(l_U)); mats.append(m_U)
        if l_C.numel():
            tables.append(self._rows(l_C)); mats.append(m_C)


    def Assemble(self):
        sp = self.space
        unsup = [i for i in self.integrals if getattr(i.form, "cf", None) is not None and
                 (isinstance(i.dx, FacetPatchDifferentialSymbol) or i.dx.vb == 1)]
        if unsup:
            raise XgpuError("coefficient functions: volume cut integrals (dCut / dxtref / dmesh, vb=VOL) only")
        if hasattr(sp, "timeFE"):  # SpaceTimeFESpace
            from . import spacetime
            return spacetime.assemble_bilinear(self)
        if any(i.dx.time_order >= 0 or i.dx.tref is not None for i in self.integrals):
            raise XgpuError("space-time differential symbols (time_order / tref) need a SpaceTimeFESpace")
        mesh = sp.mesh
        dim, p = mesh.dim, sp.order
        ctx = default_ctx(mesh.device.index)
        dev = mesh.device
        ne = mesh.ne
        nd = sp.nd
        active = self._active_elements()

        # ---- volume (cut) integrators: group per level set; per element ONE matrix summing all integrators
        vol = [i for i in self.integrals if isinstance(i.dx, CutDifferentialSymbol) and not (i.dx.skeleton or i.dx.element_boundary)]
        skel = [i for i in self.integrals if isinstance(i.dx, CutDifferentialSymbol) and (i.dx.skeleton or i.dx.element_boundary)]
        fac = [i for i in self.integrals if isinstance(i.dx, FacetPatchDifferentialSymbol)]
        tables, mats = [], []
        # two batches: U = uncut active elements, C = cut active elements (w.r.t. the level set of the
        # integrators); every integrator accumulates into its elements' slot, so each element has ONE
        # matrix as in NGSolve's element loop (CalcElementMatrixAdd over all integrators, then one
        # AddElementMatrix).  A batch whose first integrator covers it completely is written without
        # a prior memset (write-only traffic).
        lsets = []
        for integ in vol:
            ls = integ.dx.levelset
            key = tuple(id(l) for l in ls) if isinstance(ls, (list, tuple)) else id(ls)
            if key not in lsets:
                lsets.append(key)
        general = len(lsets) > 1 or any(isinstance(i.dx.levelset, (list, tuple)) for i in vol) \
            or any(i.dx.vb == 1 for i in vol) \
            or any(i.form.nlset is not None and i.form.nlset is not i.dx.levelset for i in vol)
        if general:
            self._assemble_general(vol, active, tables, mats, ctx)
        else:
            self._assemble_single_lset(vol, active, tables, mats, ctx)

        # ---- facet patch integrators
        _, _, fe = mesh.facets() if (fac or skel or sp.dgjumps or self.facet_restriction is not None) else (None, None, None)
        skel_cover = self._assemble_facet_integrals(skel, tables, mats, ctx) if skel else None
        graph_facets = None
        if self.facet_restriction is not None:
            graph_facets = self.facet_restriction
        elif sp.dgjumps and mesh.dim > 1:
            # plain BilinearForm on a dgjumps space: couple across every facet between active elements
            graph_facets = BitArray(mesh.nfacet, words=ops.facets_with_neighbor_types(fe, active.bits, active.bits, False, False, True, ctx=ctx))
        facet_mats = {}
        for integ in fac:
            terms = integ.form.terms
            if not all(t[0] == OP_JUMP and t[1] == OP_JUMP for t in terms):
                raise XgpuError("dFacetPatch: only c*h^k*(u-u.Other())*(v-v.Other()) is supported")
            l_f = integ.dx.definedonelements.indices()
            # interior facets only: NGSolve's facet loop hands boundary facets to BND integrators, a facet-patch
            # integrator never sees them (T_CalcFacetMatrix throws for LocalFacetNr2 == -1, symboliccutbfi.cpp:1992)
            l_f = l_f[fe[l_f.long(), 1] >= 0].contiguous()
            key = id(integ.dx.definedonelements)
            for (_, _, c, hp) in terms:
                with ctx.deformation(_eff_def(integ.dx.deformation, mesh)):
                    mm = ops.facetpatch(dim, p, l_f, fe, mesh.elements, mesh.points, integ.dx.order, c, hp,
                                        downscale=integ.dx.downscale, ctx=ctx)
                if key in facet_mats:
                    facet_mats[key] = (l_f, facet_mats[key][1] + mm)
                else:
                    facet_mats[key] = (l_f, mm)
        covered = skel_cover
        for key, (l_f, mm) in facet_mats.items():
            fel = fe[l_f.long()]
            t_f = torch.cat([self._rows(fel[:, 0].contiguous()), self._rows(fel[:, 1].contiguous())], dim=1).contiguous()
            tables.append(t_f); mats.append(mm)
            fb = BitArray(mesh.nfacet, mesh.device)
            fb.bits.copy_(BitArray.from_numpy(_mask_from_list(l_f, mesh.nfacet), mesh.device).bits)
            covered = fb if covered is None else (covered | fb)
        if graph_facets is not None:
            # facets that only contribute sparsity (explicit zeros kept)
            rest = graph_facets if covered is None else BitArray(mesh.nfacet, words=ops.bitset_op(ops.BIT_ANDNOT, graph_facets.bits, covered.bits, mesh.nfacet, ctx=ctx))
            l_g = rest.indices()
            if l_g.numel():
                fel = fe[l_g.long()]
                inner = fel[:, 1] >= 0
                fel = fel[inner]
                if fel.shape[0]:
                    t_g = torch.cat([self._rows(fel[:, 0].contiguous()), self._rows(fel[:, 1].contiguous())], dim=1).contiguous()
                    tables.append(t_g)
                    mats.append(None)  # sparsity only

        if not tables:  # nothing active: the empty matrix
            self.mat = SparseMatrixCSR(sp.ndof, torch.zeros(sp.ndof + 1, dtype=torch.int64, device=dev),
                                       torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.float64, device=dev))
            return self
        plan = ops.csr_plan(sp.ndof, tables, ctx=ctx)
        rowptr, col, vals = ops.csr_assemble(plan, list(zip(tables, mats)))
        self.mat = SparseMatrixCSR(sp.ndof, rowptr, col, vals)
        return self


def _bnd_cutquad(mesh, bv, phi, dt, order, ctx):
    """cut rules of the boundary elements bv ((dim-1)-simplices) in their own reference coordinates.  dt = IF on the boundary
    segments of a 2D mesh is the codim-2 case of cutint/straightcutrule.cpp:646-656: the rule is the cut POINT with the
    weight w / |segment| (point measure: measure * weight = w).  The segment goes to the 1D interface rule with its arclength
    as coordinate, whose factor ||F^{-T} n_ref|| (:558-578) is exactly 1 / |segment|."""
    dim = mesh.dim
    if int(dt) != IF:
        return ops.cutquad(dim - 1, None, bv, mesh.points, phi, int(dt), order, ctx=ctx)
    if dim != 2:
        raise XgpuError("dCut(IF, vb=BND): boundary segments of 2D meshes only (codim-2 point measure, straightcutrule.cpp:646-656)")
    nb, dev = bv.shape[0], bv.device
    P = mesh.points
    length = (P[bv[:, 0].long()] - P[bv[:, 1].long()]).norm(dim=1)
    seg = torch.arange(2 * nb, dtype=torch.int32, device=dev).reshape(nb, 2)
    xs = torch.stack([length, torch.zeros_like(length)], dim=1).reshape(-1, 1).contiguous()   # vertex 0 <-> xi = 1
    return ops.cutquad(1, None, seg, xs, phi[bv.long()].reshape(-1).contiguous(), IF, order, ctx=ctx)


def _bnd_rule_in_volume(mesh, phi, dt, order, ctx, facet_mask=None):
    """dCut(lset, dt, vb=BND): the cut rule of every boundary facet (a (dim-1)-simplex, cutint/straightcutrule.cpp on
    the BND element) carried into the reference coordinates of the adjacent volume element, so that the volume
    kernels evaluate the traces: returns (volume element per boundary element, offsets, points, weights) with
    weights w * |facet| / |det F_vol| (the kernels multiply by the volume measure).  facet_mask: BitArray over the
    boundary ELEMENTS (definedonelements)."""
    dim = mesh.dim
    bv, fids = mesh.boundary_elements()
    _, _, fe = mesh.facets()
    vol = fe[fids.long(), 0].contiguous()                                  # adjacent volume element
    sel = None
    if facet_mask is not None:
        sel = facet_mask.indices()
        bv, vol = bv[sel.long()].contiguous(), vol[sel.long()].contiguous()
    off, q, w = _bnd_cutquad(mesh, bv, phi, dt, order, ctx)
    ctx.raise_on_geometry_error()
    nb = bv.shape[0]
    b_of_pt = torch.repeat_interleave(torch.arange(nb, device=bv.device), off[1:] - off[:-1])
    ev = mesh.elements[vol.long()].long()                                   # (nb, dim+1) volume element vertices
    # local index of every boundary vertex inside its volume element
    loc = (ev.unsqueeze(1) == bv.long().unsqueeze(2)).to(torch.int64).argmax(dim=2)   # (nb, dim)
    # barycentric weights of a facet point: reference simplex convention v_k <-> e_k, last vertex <-> origin
    lam = torch.cat([q[:, :dim - 1], 1.0 - q[:, :dim - 1].sum(dim=1, keepdim=True)], dim=1)          # (npts, dim)
    E = torch.zeros((dim + 1, 3), dtype=torch.float64, device=bv.device)     # reference coords of the volume vertices
    E[torch.arange(dim), torch.arange(dim)] = 1.0
    xh = (lam.unsqueeze(2) * E[loc[b_of_pt]]).sum(dim=1)                     # (npts, 3)
    # measures
    P = mesh.points
    pv = P[bv.long()]                                                        # (nb, dim, dim)
    if dim == 2:
        fm = (pv[:, 0] - pv[:, 1]).norm(dim=1)
    else:
        fm = torch.linalg.cross(pv[:, 0] - pv[:, 2], pv[:, 1] - pv[:, 2]).norm(dim=1)
    X = P[ev]                                                                # (nb, dim+1, dim)
    J = (X[:, :dim] - X[:, dim:dim + 1]).transpose(1, 2)
    vm = torch.linalg.det(J).abs()
    w2 = w * (fm / vm)[b_of_pt]
    return vol, off, xh.contiguous(), w2.contiguous()


class LinearForm:
    """LinearForm(space); f += coef * x**a * y**b * v * dCut(lset, dt, definedonelements=...); f.Assemble() -> f.vec
    (SymbolicCutLinearFormIntegrator, xfem/symboliccutlfi.cpp:51-140; demos/fictdom.py:132-137).  Element vectors
    on the GPU, summed per dof in a fixed (element, local index) order."""

    def __init__(self, space, name="lf", **flags):
        self.space, self.integrals, self.vec = space, [], None

    def __iadd__(self, o):
        ints = [o] if isinstance(o, Integral) else (o.ints if isinstance(o, SumOfIntegrals) else None)
        st = hasattr(self.space, "timeFE")
        if ints is None or not all((hasattr(i.form, "fixt") if st else isinstance(i.form, LForm))
                                   and isinstance(i.dx, CutDifferentialSymbol) for i in ints):
            raise XgpuError("LinearForm += expects f * v * dCut(...)")
        self.integrals.extend(ints)
        return self

    def Assemble(self):
        sp = self.space
        if any(getattr(i.form, "cf", None) is not None and i.dx.vb == 1 for i in self.integrals):
            raise XgpuError("coefficient functions: volume cut integrals (dCut / dxtref / dmesh, vb=VOL) only")
        if hasattr(sp, "timeFE"):  # SpaceTimeFESpace
            from . import spacetime
            return spacetime.assemble_linear(self)
        if any(i.dx.time_order >= 0 or i.dx.tref is not None for i in self.integrals):
            raise XgpuError("space-time differential symbols (time_order / tref) need a SpaceTimeFESpace")
        mesh = sp.mesh
        dim, p, ne, nd, dev = mesh.dim, sp.order, mesh.ne, sp.nd, mesh.device
        ctx = default_ctx(dev.index)
        act = getattr(sp, "active_elements", None)
        active = BitArray(ne, words=act) if act is not None else BitArray(ne, dev).Set()
        l_A = active.indices()
        self.vec = torch.zeros(sp.ndof, dtype=torch.float64, device=dev)
        if not l_A.numel() or not self.integrals:
            return self
        slot_A = ops.bitset_rank(active.bits, ne, ctx=ctx)[0]
        v_A = torch.zeros((l_A.numel(), nd), dtype=torch.float64, device=dev)
        rows_of = (lambda l: sp.rows(l)) if hasattr(sp, "rows") else (lambda l: ops.gather_rows(sp.el_dofs, l, None))
        extra = []
        for integ in self.integrals:
            dx = integ.dx
            order = dx.order if dx.order >= 0 else 2 * p  # symboliccutlfi.cpp:86-88
            don = dx.definedonelements
            if dx.vb == 1:  # boundary elements: own batch, see BilinearForm._assemble_general
                phi = _lset_vec(dx.levelset, mesh)
                vol_els, off, q, w = _bnd_rule_in_volume(mesh, phi, int(dx.domain_type), order, ctx, facet_mask=don)
                if w.numel():
                    phi_n = _lset_vec(integ.form.nlset, mesh) if integ.form.nlset is not None else phi
                    v_B = ops.elvec_rule(dim, p, vol_els, mesh.elements, mesh.points, phi_n, off, q, w, integ.form.terms, ctx=ctx)
                    extra.append((rows_of(vol_els), v_B))
                continue
            if don is None:
                l, slots = l_A, None
            else:
                l = ops.bitset_to_list(ops.bitset_op(ops.BIT_AND, active.bits, don.bits, ne, ctx=ctx), ne, ctx=ctx)
                slots = slot_A[l.long()].contiguous()
            if not l.numel():
                continue
            ls = dx.levelset
            with ctx.deformation(_eff_def(dx.deformation, mesh)):
                has_dn = any(t[0] == OP_DN for t in integ.form.terms)
                if isinstance(ls, (list, tuple)):
                    if has_dn and integ.form.nlset is None:
                        raise XgpuError("normal-derivative terms on a multi-level-set domain: build n from ONE of the level sets")
                    phis = torch.stack([_lset_vec(g, mesh) for g in ls]).contiguous()
                    dts = _normalize_dtlist(dx.domain_type, len(ls))
                    off, q, w = ops.cutquad_ml(dim, l, mesh.elements, mesh.points, phis, dts, order, ctx=ctx)
                    phi = phis[0].contiguous()
                else:
                    phi = _lset_vec(ls, mesh)
                    off, q, w = ops.cutquad(dim, l, mesh.elements, mesh.points, phi, int(dx.domain_type), order, ctx=ctx)
                ctx.raise_on_geometry_error()
                if has_dn and integ.form.nlset is not None:
                    phi = _lset_vec(integ.form.nlset, mesh)
                if w.numel() and integ.form.cf is not None:   # f(x_q) folded into the weights (xfem/symboliccutlfi.cpp:96-118)
                    w = w * integ.form.cf.evaluate(dim, mesh, l, off, q, ctx)
                if w.numel():
                    ops.elvec_rule(dim, p, l, mesh.elements, mesh.points, phi, off, q, w, integ.form.terms, out=v_A,
                                   accumulate=True, slots=slots, ctx=ctx)
        table = rows_of(l_A)
        batches = [(table, v_A)] + extra
        plan = ops.adjacency_plan(sp.ndof, [b[0] for b in batches], ctx=ctx)
        ops.vec_scatter(plan, batches, out=self.vec)
        return self


def _mask_from_list(lst, n):
    import numpy as np
    m = np.zeros(n, dtype=bool)
    m[lst.cpu().numpy()] = True
    return m


def RestrictedBilinearForm(space, name="blf", element_restriction=None, facet_restriction=None, check_unused=True, **flags):
    """python/__init__.py:973-1037 -> utils/python_utils.cpp:62-130"""
    return BilinearForm(space, element_restriction=element_restriction, facet_restriction=facet_restriction,
                        check_unused=check_unused, name=name, **flags)


# ------------------------------------------------------------------ Integrate
class CoordMonomial:
    """x**a * y**b * z**c  (the only coefficient functions Integrate supports besides constants)"""

    def __init__(self, expo, coef=1.0):
        self.expo, self.coef = tuple(expo), float(coef)

    def __mul__(self, o):
        if isinstance(o, numbers.Number):
            return CoordMonomial(self.expo, self.coef * o)
        if isinstance(o, (Proxy, LForm)):
            return o.__mul__(self)
        if not isinstance(o, CoordMonomial):
            return NotImplemented
        return CoordMonomial(tuple(a + b for a, b in zip(self.expo, o.expo)), self.coef * o.coef)

    __rmul__ = __mul__

    def __pow__(self, k):
        return CoordMonomial(tuple(a * k for a in self.expo), self.coef ** k)


x = CoordMonomial((1, 0, 0))
y = CoordMonomial((0, 1, 0))
z = CoordMonomial((0, 0, 1))


def _eff_def(deformation, mesh):
    """the `deformation=` argument, else the one set on the mesh (mesh.SetDeformation, python/lsetcurv.py:358-363)"""
    return deformation if deformation is not None else getattr(mesh, "deformation", None)


def _integrate_cf(ld, mesh, cf, order, deformation):
    """sum_q w_q |det F(xhat_q)| cf(x_q): the weights times the point measures (xgpu_rule_phys), summed in a fixed order"""
    if ld.get("subdivlvl", 0) != 0:
        raise XgpuError("subdivlvl != 0: legacy path not provided")
    if ld.get("vb", 0) == 1:
        raise XgpuError("Integrate over BND: constants and monomials only")
    if "order" in ld and ld["order"] is not None and ld["order"] >= 0:
        order = ld["order"]
    dim = mesh.dim
    ctx = default_ctx(mesh.device.index)
    ls, dt = ld["levelset"], ld["domain_type"]
    with ctx.deformation(_eff_def(deformation, mesh)):
        if isinstance(ls, (list, tuple)):
            phis = torch.stack([_lset_vec(l, mesh) for l in ls]).contiguous()
            off, p, w = ops.cutquad_ml(dim, None, mesh.elements, mesh.points, phis, _normalize_dtlist(dt, len(ls)), order, ctx=ctx)
        else:
            off, p, w = ops.cutquad(dim, None, mesh.elements, mesh.points, _lset_vec(ls, mesh), int(dt), order, ctx=ctx)
        ctx.raise_on_geometry_error()
        if not w.numel():
            return 0.0
        _, meas = ops.rule_phys(dim, None, mesh.elements, mesh.points, off, p, want_meas=True, ctx=ctx)
        vals = cf.evaluate(dim, mesh, None, off, p, ctx) * w * meas
    return ops.sum_f64(vals, ctx=ctx)


def Integrate(levelset_domain=None, mesh=None, cf=1.0, order=5, time_order=-1, deformation=None, heapsize=1000000):
    """python/__init__.py:361-439 -> cutint/python_cutint.cpp:21-163: sum_q w_q |det F| cf(x_q)."""
    ld = levelset_domain
    if "time_order" in ld and ld["time_order"] is not None and ld["time_order"] >= 0:
        time_order = ld["time_order"]
    if time_order >= 0 or ld.get("tref", None) is not None or hasattr(cf, "expo") and len(cf.expo) == 4:
        from . import spacetime
        return spacetime.integrate(ld, mesh, cf, order, time_order, deformation)
    if isinstance(cf, numbers.Number):
        cf = CoordMonomial((0, 0, 0), cf)
    if callable(cf) and not isinstance(cf, (CoordMonomial, CoefficientFunction)):
        cf = CoefficientFunction(cf)
    if isinstance(cf, CoefficientFunction):
        return _integrate_cf(ld, mesh, cf, order, deformation)
    if not isinstance(cf, CoordMonomial):
        raise XgpuError("Integrate: cf is a constant, a coordinate monomial, a CoefficientFunction or a callable of the coordinates")
    ld = levelset_domain
    if ld.get("subdivlvl", 0) != 0:
        raise XgpuError("subdivlvl != 0: legacy path not provided")
    if "order" in ld and ld["order"] is not None and ld["order"] >= 0:
        order = ld["order"]
    dim = mesh.dim
    ctx = default_ctx(mesh.device.index)
    ls, dt = ld["levelset"], ld["domain_type"]
    if ld.get("vb", 0) == 1:  # BND: dCut(lset, dt, vb=BND) of tests/pytests/test_bndcut.py — boundary facets as elements
        if isinstance(ls, (list, tuple)) or deformation is not None:
            raise XgpuError("Integrate over BND: single level set, no deformation")
        if dim < 2:
            raise XgpuError("Integrate over BND needs dim >= 2")
        bv, _ = mesh.boundary_elements()
        don = ld.get("definedonelements", None)
        if don is not None:                       # BitArray over the boundary elements (test_bndcut.py:203-217)
            bv = bv[don.indices().long()].contiguous()
        off, p, w = _bnd_cutquad(mesh, bv, _lset_vec(ls, mesh), dt, order, ctx)
        ctx.raise_on_geometry_error()
        return cf.coef * ops.integrate_monomial_bnd(dim - 1, dim, None, bv, mesh.points, off, p, w, cf.expo[:dim], ctx=ctx)
    with ctx.deformation(_eff_def(deformation, mesh)):
        if isinstance(ls, (list, tuple)):
            phis = torch.stack([_lset_vec(l, mesh) for l in ls]).contiguous()
            dts = _normalize_dtlist(dt, len(ls))
            off, p, w = ops.cutquad_ml(dim, None, mesh.elements, mesh.points, phis, dts, order, ctx=ctx)
        else:
            off, p, w = ops.cutquad(dim, None, mesh.elements, mesh.points, _lset_vec(ls, mesh), int(dt), order, ctx=ctx)
        ctx.raise_on_geometry_error()
        return cf.coef * ops.integrate_monomial(dim, None, mesh.elements, mesh.points, off, p, w, cf.expo[:dim], ctx=ctx)
