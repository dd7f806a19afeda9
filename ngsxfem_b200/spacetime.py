"""Space-time CutFEM surface (BASELINE config 5): tensor-product space-time FE spaces on time slabs, space-time cut
integrals and their assembly on the GPU.

Mirrors  spacetime/SpaceTimeFESpace.cpp / SpaceTimeFE.cpp (time-major tensor basis, nodal time FE), python/__init__.py:
129-153 (`dtref`, `dt`), :539-561 (`fix_tref`), :867-925 (`dxtref`), :927-972 (`dmesh(tref=...)`), `dCut(...,
time_order=k)` (:781-864) and the dispatcher cutint/xintegration.cpp:41-90 (space-time rule for time_order >= 0, a
spatial rule stamped with a fixed reference time for `tref`).  Element matrices: `xgpu_elmat_st`, vectors:
`xgpu_elvec_st`; sparse assembly through the same CSR plan as the spatial forms.

Menu (anything else raises XgpuError): sums of  c * h^k * A(u) * B(v)  with A, B in {u, dt(u)} or grad(u)*grad(v);
`fix_tref(u, t0) * fix_tref(v, t0)` on a spatial differential symbol; linear forms  c * h^k * x^a y^b z^c tref^d * B(v).
"""
import numbers

import torch

from . import ops
from ._lib import XgpuError, default_ctx
from .cutinfo import BitArray, GridFunction, POS
from .forms import (CoordMonomial, CutDifferentialSymbol, FacetPatchDifferentialSymbol, Integral, MeshSize, Scalar,
                    SparseMatrixCSR)

OP_ID, OP_GRAD, OP_DT = 0, 1, ops.OP_DT
OP_JUMP = 9   # u - u.Other() (facet patches only; not a kernel operator code)


class ScalarTimeFE:
    """NodalTimeFE of order k on [0,1] (spacetime/SpaceTimeFE.cpp:258-282): nodes 0,1 / 0,1/2,1 / Gauss-Lobatto"""

    def __init__(self, order=1):
        if not 0 <= int(order) <= 3:
            raise XgpuError("ScalarTimeFE: orders 0..3")
        self.order = int(order)

    def nodes(self):
        k = self.order
        if k == 0:
            return [0.0]
        if k == 1:
            return [0.0, 1.0]
        if k == 2:
            return [0.0, 0.5, 1.0]
        s = 1.0 / 5.0 ** 0.5
        return [0.0, 0.5 * (1.0 - s), 0.5 * (1.0 + s), 1.0]

    def shapes(self, t):
        nd = self.nodes()
        out = []
        for i, ni in enumerate(nd):
            s = 1.0
            for j, nj in enumerate(nd):
                if j != i:
                    s *= (t - nj) / (ni - nj)
            out.append(s)
        return out


class SpaceTimeFESpace:
    """SpaceTimeFESpace(spacefes, timefe): dof (j, i) = j * spacefes.ndof + i (time-major, SpaceTimeFE.hpp:185-218)"""

    def __init__(self, spacefes, timefe, dirichlet=None, dgjumps=False):
        self.spaceFES, self.timeFE = spacefes, timefe
        self.mesh, self.order, self.time_order = spacefes.mesh, spacefes.order, timefe.order
        self.nt = timefe.order + 1
        self.nds, self.ndof_space = spacefes.nd, spacefes.ndof
        self.nd, self.ndof = spacefes.nd * self.nt, spacefes.ndof * self.nt
        self.dgjumps = dgjumps
        self.active_elements = getattr(spacefes, "active_elements", None)

    def TimeFE_nodes(self):
        return self.timeFE.nodes()

    def rows(self, elem_list):
        sp = self.spaceFES
        r = sp.rows(elem_list) if hasattr(sp, "rows") else ops.gather_rows(sp.el_dofs, elem_list, None)
        shifted = [torch.where(r >= 0, r + j * self.ndof_space, r) for j in range(self.nt)]
        return torch.cat(shifted, dim=1).to(torch.int32).contiguous()

    def TnT(self):
        return STProxy(self, False), STProxy(self, True)

    def TrialFunction(self):
        return STProxy(self, False)

    def TestFunction(self):
        return STProxy(self, True)


class SpaceTimeGridFunction:
    """nodal-in-time P1 level set: vec is (k_t+1) x nv, time-major (the dof layout of a SpaceTimeFESpace over H1(1))"""

    def __init__(self, mesh, timefe, values=None):
        self.mesh, self.timeFE = mesh, timefe
        nt = timefe.order + 1
        self.vec = torch.zeros((nt, mesh.nv), dtype=torch.float64, device=mesh.device) if values is None \
            else torch.as_tensor(values, dtype=torch.float64).to(mesh.device).reshape(nt, mesh.nv).contiguous()


def SpaceTimeInterpolateToP1(fn, mesh, timefe, eps_perturbation=1e-14):
    """python/__init__.py:521-537 (SpaceTimeSet) with utils/p1interpol.cpp per time node: fn(x, y[, z], tref)"""
    p = mesh.points
    rows = []
    for t in timefe.nodes():
        v = torch.as_tensor(fn(*[p[:, i] for i in range(p.shape[1])], t), dtype=torch.float64, device=p.device)
        rows.append(torch.where(v.abs() < eps_perturbation, torch.full_like(v, eps_perturbation), v))
    return SpaceTimeGridFunction(mesh, timefe, torch.stack(rows))


# ------------------------------------------------------------------ symbolic layer
class STMonomial:
    """coef * x^a y^b z^c tref^d"""

    def __init__(self, expo, coef=1.0):
        self.expo, self.coef = tuple(expo), float(coef)

    def __mul__(self, o):
        if isinstance(o, numbers.Number):
            return STMonomial(self.expo, self.coef * o)
        if isinstance(o, CoordMonomial):
            o = STMonomial(tuple(o.expo) + (0,), o.coef)
        if isinstance(o, STMonomial):
            return STMonomial(tuple(a + b for a, b in zip(self.expo, o.expo)), self.coef * o.coef)
        if isinstance(o, (STProxy, STLForm)):
            return o.__mul__(self)
        return NotImplemented

    __rmul__ = __mul__

    def __pow__(self, k):
        return STMonomial(tuple(a * k for a in self.expo), self.coef ** k)


def ReferenceTimeVariable():
    return STMonomial((0, 0, 0, 1))


tref = ReferenceTimeVariable()


def _scal(o):
    if isinstance(o, Scalar):
        return o
    if isinstance(o, numbers.Number):
        return Scalar(float(o))
    return Scalar(1.0, o.power)


class STProxy:
    def __init__(self, space, is_test, op=OP_ID, scal=None, fixt=None, other=False):
        self.space, self.is_test, self.op, self.fixt, self.other = space, is_test, op, fixt, other
        self.scal = scal or Scalar(1.0)

    def _with(self, op=None, scal=None, fixt=None):
        return STProxy(self.space, self.is_test, self.op if op is None else op, scal or self.scal,
                       self.fixt if fixt is None else fixt, self.other)

    def Other(self):
        if self.op != OP_ID:
            raise XgpuError(".Other() is supported on u / v only")
        return STProxy(self.space, self.is_test, OP_ID, self.scal, self.fixt, other=True)

    def __sub__(self, o):
        if isinstance(o, STProxy) and o.other and not self.other and self.op == OP_ID and o.op == OP_ID \
                and o.is_test == self.is_test and o.fixt == self.fixt:
            if (self.scal.c, self.scal.hpow) != (o.scal.c, o.scal.hpow):
                raise XgpuError("unsupported difference of proxies: both sides of u - u.Other() must carry the same factor")
            return STProxy(self.space, self.is_test, OP_JUMP, self.scal, self.fixt)
        raise XgpuError("unsupported difference of proxies (menu: u - u.Other())")

    def _st_grad(self):
        if self.op != OP_ID:
            raise XgpuError("grad of this space-time proxy is not supported")
        return self._with(op=OP_GRAD)

    def __neg__(self):
        return self._with(scal=self.scal * -1.0)

    def __mul__(self, o):
        if isinstance(o, (numbers.Number, Scalar, MeshSize)):
            return self._with(scal=self.scal * _scal(o))
        if isinstance(o, STProxy):
            return STForm.from_pair(self, o)
        if isinstance(o, (STMonomial, CoordMonomial)):
            return STLForm.from_proxy(self) * o
        if isinstance(o, CutDifferentialSymbol):
            return Integral(STLForm.from_proxy(self), o)
        return NotImplemented

    __rmul__ = __mul__

    def _st_with_cf(self, cf):
        if not self.is_test:
            raise XgpuError("a coefficient multiplies a whole integrand: cf * (u*v), cf * v")
        return STLForm.from_proxy(self)._st_with_cf(cf)


def dtref(u):
    """python/__init__.py:129-136: derivative w.r.t. the REFERENCE time interval"""
    if not isinstance(u, STProxy) or u.op != OP_ID:
        raise XgpuError("cannot form dt")
    return u._with(op=OP_DT)


dt = dtref


def fix_tref(u, time):
    """python/__init__.py:539-561 for proxies: the time variable of u is fixed to the reference time `time`"""
    if not isinstance(u, STProxy):
        raise XgpuError("fix_tref: proxies of a SpaceTimeFESpace only")
    return u._with(fixt=float(time))


def _merge_st_cf(a, b):
    if a is b:
        return a
    raise XgpuError("integrands with different coefficient functions: add them as separate integrals")


class STForm:
    """cf: a CoefficientFunction fn(x, y[, z], tref) multiplying the whole integrand (folded into the rule weights,
    xfem/symboliccutbfi.cpp:174-193 with the space-time rule's reference times)"""

    def __init__(self, terms, space, fixt=None, cf=None):
        self.terms, self.space, self.fixt, self.cf = terms, space, fixt, cf

    def _st_with_cf(self, cf):
        return STForm(self.terms, self.space, self.fixt, cf if self.cf is None else self.cf * cf)

    @staticmethod
    def from_pair(a, b):
        if a.is_test == b.is_test:
            raise XgpuError("a bilinear form needs one trial and one test function per term")
        u, v = (b, a) if a.is_test else (a, b)
        if (u.op == OP_GRAD) != (v.op == OP_GRAD):
            raise XgpuError("grad(u) must be paired with grad(v)")
        if (u.op == OP_JUMP) != (v.op == OP_JUMP) or u.other or v.other:
            raise XgpuError("jump terms must pair (u-u.Other()) with (v-v.Other())")
        if u.fixt != v.fixt:
            raise XgpuError("fix_tref: trial and test function must be fixed to the same reference time")
        s = u.scal * v.scal
        return STForm([(u.op, v.op, s.c, s.hpow)], u.space, u.fixt)

    def __add__(self, o):
        if isinstance(o, STForm):
            if o.fixt != self.fixt:
                raise XgpuError("one integrand: one fixed reference time (split it into several integrals)")
            return STForm(self.terms + o.terms, self.space, self.fixt, _merge_st_cf(self.cf, o.cf))
        return NotImplemented

    def __neg__(self):
        return self * -1.0

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, o):
        if isinstance(o, (numbers.Number, Scalar, MeshSize)):
            s = _scal(o)
            return STForm([(a, b, c * s.c, hp + s.hpow) for a, b, c, hp in self.terms], self.space, self.fixt, self.cf)
        if isinstance(o, (CutDifferentialSymbol, FacetPatchDifferentialSymbol)):
            if self.cf is not None and isinstance(o, FacetPatchDifferentialSymbol):
                raise XgpuError("coefficient functions on facet-patch integrands are not on the accelerated path")
            return Integral(self, o)
        return NotImplemented

    __rmul__ = __mul__


class STLForm:
    """terms: (op_test, coef, hpow, (e0, e1, e2, e_tref))"""

    def __init__(self, terms, space, fixt=None, cf=None):
        self.terms, self.space, self.fixt, self.cf = terms, space, fixt, cf

    def _st_with_cf(self, cf):
        return STLForm(self.terms, self.space, self.fixt, cf if self.cf is None else self.cf * cf)

    @staticmethod
    def from_proxy(v):
        if not v.is_test:
            raise XgpuError("a linear form needs a test function")
        if v.op not in (OP_ID, OP_DT):
            raise XgpuError("space-time linear forms: menu is f*v and f*dt(v)")
        return STLForm([(v.op, v.scal.c, v.scal.hpow, (0, 0, 0, 0))], v.space, v.fixt)

    def __add__(self, o):
        if isinstance(o, STProxy):
            o = STLForm.from_proxy(o)
        if isinstance(o, STLForm):
            if o.fixt != self.fixt:
                raise XgpuError("one integrand: one fixed reference time")
            return STLForm(self.terms + o.terms, self.space, self.fixt, _merge_st_cf(self.cf, o.cf))
        return NotImplemented

    def __mul__(self, o):
        if isinstance(o, (numbers.Number, Scalar, MeshSize)):
            s = _scal(o)
            return STLForm([(a, c * s.c, hp + s.hpow, ex) for a, c, hp, ex in self.terms], self.space, self.fixt, self.cf)
        if isinstance(o, CoordMonomial):
            o = STMonomial(tuple(o.expo) + (0,), o.coef)
        if isinstance(o, STMonomial):
            return STLForm([(a, c * o.coef, hp, tuple(e1 + e2 for e1, e2 in zip(ex, o.expo))) for a, c, hp, ex in self.terms],
                           self.space, self.fixt, self.cf)
        if isinstance(o, CutDifferentialSymbol):
            return Integral(self, o)
        return NotImplemented

    __rmul__ = __mul__


# ------------------------------------------------------------------ differential symbols
def dxtref(mesh, order=None, time_order=-1, definedonelements=None, **kwargs):
    """python/__init__.py:867-925: all elements extruded by the reference interval — a POS domain of the constant
    level set +1 with a space-time rule"""
    if kwargs:
        raise XgpuError(f"unsupported dxtref arguments: {sorted(kwargs)}")
    if type(time_order) != int or (order is not None and type(order) != int):
        raise XgpuError("dxtref: order / time_order must be integers")
    one = SpaceTimeGridFunction(mesh, ScalarTimeFE(1), torch.ones((2, mesh.nv), dtype=torch.float64))
    return CutDifferentialSymbol(one, POS, -1 if order is None else order, definedonelements, None, time_order=time_order)


def dmesh(mesh=None, tref=None, order=None, definedonelements=None, **kwargs):
    """python/__init__.py:927-972: all elements; with tref the spatial rule is stamped with that reference time"""
    if kwargs:
        raise XgpuError(f"unsupported dmesh arguments: {sorted(kwargs)}")
    if mesh is None:
        raise XgpuError("dx(..,tref..) needs mesh")
    one = GridFunction(mesh, torch.ones(mesh.nv, dtype=torch.float64))
    return CutDifferentialSymbol(one, POS, -1 if order is None else order, definedonelements, None, tref=tref)


# ------------------------------------------------------------------ rules
def st_rule(mesh, elems, dx, space_order, fixt, ctx):
    """cutint/xintegration.cpp:41-90: (offsets, points, reference times, weights) of one integrator on `elems`"""
    dim = mesh.dim
    ls, dt_ = dx.levelset, int(dx.domain_type)
    if isinstance(ls, (list, tuple)):
        raise XgpuError("space-time rules: one level set (multi-level-set domains are not space-time)")
    if dx.deformation is not None:
        raise XgpuError("space-time rules do not support a mesh deformation")
    if dx.time_order >= 0:
        if fixt is not None or dx.tref is not None:
            raise XgpuError("fix_tref / tref belong on a spatial differential symbol (no time_order)")
        if isinstance(ls, SpaceTimeGridFunction):
            phi_st, kt = ls.vec, ls.timeFE.order
        else:  # a spatial level set is constant in time (xintegration.cpp:60-70)
            v = ls.vec if isinstance(ls, GridFunction) else ls
            phi_st, kt = v.reshape(1, -1).contiguous(), 0
        if phi_st.shape[1] != mesh.nv:
            raise XgpuError("level set must have one value per mesh vertex and time node (P1 in space)")
        off, q, tv, w = ops.cutquad_st(dim, elems, mesh.elements, mesh.points, phi_st.to(mesh.device).contiguous(), kt, dt_,
                                       dx.time_order, space_order, ctx=ctx)
        ctx.raise_on_geometry_error()
        return off, q, tv, w
    t0 = fixt if fixt is not None else dx.tref
    if t0 is None:
        raise XgpuError("SpaceTimeFE :: CalcShape called with a mere space IR")   # spacetime/SpaceTimeFE.cpp:41-42
    if dx.tref is not None and fixt is not None and float(dx.tref) != float(fixt):
        raise XgpuError("fix_tref and the differential symbol's tref disagree")
    if isinstance(ls, SpaceTimeGridFunction):  # the level set at the fixed reference time
        sh = torch.tensor(ls.timeFE.shapes(float(t0)), dtype=torch.float64, device=ls.vec.device)
        phi = (sh[:, None] * ls.vec).sum(dim=0)
        phi = torch.where(phi.abs() < 1e-14, torch.full_like(phi, 1e-14), phi).contiguous()   # utils/p1interpol.cpp:105-107
    else:
        phi = (ls.vec if isinstance(ls, GridFunction) else ls).to(mesh.device).contiguous()
    off, q, w = ops.cutquad(dim, elems, mesh.elements, mesh.points, phi, dt_, space_order, ctx=ctx)
    ctx.raise_on_geometry_error()
    return off, q, torch.full_like(w, float(t0)), w


def _element_sets(space, integ, active, slot_A, l_A, ctx):
    ne = space.mesh.ne
    don = integ.dx.definedonelements
    if don is None:
        return l_A, None
    l = ops.bitset_to_list(ops.bitset_op(ops.BIT_AND, active.bits, don.bits, ne, ctx=ctx), ne, ctx=ctx)
    return l, slot_A[l.long()].contiguous()


def _active(space):
    ne, dev = space.mesh.ne, space.mesh.device
    act = space.active_elements
    return BitArray(ne, words=act) if act is not None else BitArray(ne, dev).Set()


def assemble_bilinear(blf):
    """BilinearForm(SpaceTimeFESpace).Assemble(): one matrix per active element, every integrator accumulates"""
    sp = blf.space
    mesh = sp.mesh
    dim, p, kt, nd, dev = mesh.dim, sp.order, sp.time_order, sp.nd, mesh.device
    ctx = default_ctx(dev.index)
    active = blf.element_restriction if blf.element_restriction is not None else _active(sp)
    l_A = active.indices()
    if not l_A.numel():
        blf.mat = SparseMatrixCSR(sp.ndof, torch.zeros(sp.ndof + 1, dtype=torch.int64, device=dev),
                                  torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.float64, device=dev))
        return blf
    slot_A = ops.bitset_rank(active.bits, mesh.ne, ctx=ctx)[0]
    m_A = torch.zeros((l_A.numel(), nd, nd), dtype=torch.float64, device=dev)
    facet_batches = []
    for integ in blf.integrals:
        if not isinstance(integ.form, STForm):
            raise XgpuError("space-time BilinearForm += expects a space-time form * dCut / dxtref / dmesh / dFacetPatch")
        terms = list(integ.form.terms)
        if isinstance(integ.dx, FacetPatchDifferentialSymbol):
            # space-time ghost penalty (xfem/symboliccutbfi.cpp:1982-2377 with time_order / tref): jump . jump only
            if not all(t[0] == OP_JUMP and t[1] == OP_JUMP for t in terms):
                raise XgpuError("dFacetPatch: only c*h^k*(u-u.Other())*(v-v.Other()) is supported")
            t0 = integ.form.fixt if integ.form.fixt is not None else integ.dx.tref
            if integ.dx.time_order < 0 and t0 is None:
                raise XgpuError("SpaceTimeFE :: CalcShape called with a mere space IR")
            if integ.dx.time_order >= 0 and t0 is not None:
                raise XgpuError("fix_tref / tref and time_order exclude each other")
            _, _, fe = mesh.facets()
            l_f = integ.dx.definedonelements.indices()
            l_f = l_f[fe[l_f.long(), 1] >= 0].contiguous()   # interior facets only (symboliccutbfi.cpp:1992)
            if not l_f.numel():
                continue
            mm = None
            for (_, _, c, hp) in terms:
                m1 = ops.facetpatch_st(dim, p, kt, l_f, fe, mesh.elements, mesh.points, integ.dx.order, integ.dx.time_order, t0,
                                       c, hp, downscale=integ.dx.downscale, ctx=ctx)
                mm = m1 if mm is None else mm + m1
            fel = fe[l_f.long()]
            t_f = torch.cat([sp.rows(fel[:, 0].contiguous()), sp.rows(fel[:, 1].contiguous())], dim=1).contiguous()
            facet_batches.append((t_f, mm))
            continue
        if not isinstance(integ.dx, CutDifferentialSymbol) or any(t[0] == OP_JUMP for t in terms):
            raise XgpuError("jump terms belong on dFacetPatch")
        order = integ.dx.order
        if order < 0:  # symboliccutbfi.cpp:100-118 (dt has DIFFORDER 0, spacetime/diffopDt.hpp:19)
            dtr = min(1 if t[0] == OP_GRAD else 0 for t in terms)
            dte = min(1 if t[1] == OP_GRAD else 0 for t in terms)
            order = 2 * p - (dtr + dte)
        l, slots = _element_sets(sp, integ, active, slot_A, l_A, ctx)
        if not l.numel():
            continue
        off, q, tv, w = st_rule(mesh, l, integ.dx, order, integ.form.fixt, ctx)
        if w.numel() and integ.form.cf is not None:
            w = w * integ.form.cf.evaluate_st(dim, mesh, l, off, q, tv, ctx)
        if w.numel():
            ops.elmat_st(dim, p, kt, l, mesh.elements, mesh.points, off, q, tv, w, terms, out=m_A, accumulate=True,
                         slots=slots, ctx=ctx)
    table = sp.rows(l_A)
    batches = [(table, m_A)] + facet_batches
    plan = ops.csr_plan(sp.ndof, [b[0] for b in batches], ctx=ctx)
    rowptr, col, vals = ops.csr_assemble(plan, batches)
    blf.mat = SparseMatrixCSR(sp.ndof, rowptr, col, vals)
    return blf


def assemble_linear(lf):
    sp = lf.space
    mesh = sp.mesh
    dim, p, kt, nd, dev = mesh.dim, sp.order, sp.time_order, sp.nd, mesh.device
    ctx = default_ctx(dev.index)
    active = _active(sp)
    l_A = active.indices()
    lf.vec = torch.zeros(sp.ndof, dtype=torch.float64, device=dev)
    if not l_A.numel() or not lf.integrals:
        return lf
    slot_A = ops.bitset_rank(active.bits, mesh.ne, ctx=ctx)[0]
    v_A = torch.zeros((l_A.numel(), nd), dtype=torch.float64, device=dev)
    for integ in lf.integrals:
        order = integ.dx.order if integ.dx.order >= 0 else 2 * p  # symboliccutlfi.cpp:86-88
        l, slots = _element_sets(sp, integ, active, slot_A, l_A, ctx)
        if not l.numel():
            continue
        off, q, tv, w = st_rule(mesh, l, integ.dx, order, integ.form.fixt, ctx)
        if w.numel() and integ.form.cf is not None:   # f(x_q, tref_q) folded into the weights (xfem/symboliccutlfi.cpp:96-118)
            w = w * integ.form.cf.evaluate_st(dim, mesh, l, off, q, tv, ctx)
        if w.numel():
            ops.elvec_st(dim, p, kt, l, mesh.elements, mesh.points, off, q, tv, w, integ.form.terms, out=v_A, accumulate=True,
                         slots=slots, ctx=ctx)
    table = sp.rows(l_A)
    plan = ops.adjacency_plan(sp.ndof, [table], ctx=ctx)
    ops.vec_scatter(plan, [(table, v_A)], out=lf.vec)
    return lf


def integrate(ld, mesh, cf, order, time_order, deformation):
    """Integrate(levelset_domain, mesh, cf, order, time_order) on space-time rules (python/__init__.py:361-439,
    cutint/python_cutint.cpp:21-163): cf = constant or coef * x^a y^b z^c tref^d"""
    if deformation is not None:
        raise XgpuError("space-time Integrate does not support a mesh deformation")
    if ld.get("subdivlvl", 0) != 0:
        raise XgpuError("subdivlvl != 0: legacy path not provided")
    if isinstance(cf, numbers.Number):
        cf = STMonomial((0, 0, 0, 0), cf)
    elif isinstance(cf, CoordMonomial):
        cf = STMonomial(tuple(cf.expo) + (0,), cf.coef)
    if not isinstance(cf, STMonomial):
        raise XgpuError("Integrate: cf must be a constant or a monomial in x, y, z, tref")
    if "order" in ld and ld["order"] is not None and ld["order"] >= 0:
        order = ld["order"]
    ctx = default_ctx(mesh.device.index)
    dx = CutDifferentialSymbol(ld["levelset"], ld["domain_type"], order, None, None, time_order=time_order, tref=ld.get("tref", None))
    off, q, tv, w = st_rule(mesh, None, dx, order, None, ctx)
    return cf.coef * ops.integrate_monomial_st(mesh.dim, None, mesh.elements, mesh.points, off, q, tv, w, cf.expo, ctx=ctx)
