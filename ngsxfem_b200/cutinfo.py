"""CutInfo / MultiLevelsetCutInfo and the marker helpers — Python mirror of the ngsxfem surface
(xfem/python_xfem.cpp:47-199, 346-369, 510-720) on top of the CUDA kernels.

BitArrays live on the device as 32-bit word tensors (byte layout identical to NGSolve's BitArray).
`CutInfo.Update` refills the SAME BitArray objects in place: user code captures them once and
relies on later updates (xfem/cutinfo.cpp:72-78,127-132).
"""
import numpy as np
import torch

from . import ops
from ._lib import XgpuError, default_ctx

# DOMAIN_TYPE / COMBINED_DOMAIN_TYPE (utils/ngsxstd.hpp:16-30; python names python_ngsxfem.cpp:22-32)
NEG, POS, IF = 0, 1, 2
NO, CDOM_NEG, CDOM_POS, UNCUT, CDOM_IF, HASNEG, HASPOS, ANY = range(8)
VOL, BND = 0, 1


def _to_cdt(dt):
    """map a DOMAIN_TYPE or COMBINED_DOMAIN_TYPE to the combined bitmask (python/__init__.py:30-42)"""
    if isinstance(dt, CombinedDomainType):
        return dt.mask
    return {NEG: CDOM_NEG, POS: CDOM_POS, IF: CDOM_IF}[int(dt)]


class CombinedDomainType:
    def __init__(self, mask, name):
        self.mask, self.name = mask, name

    def __repr__(self):
        return self.name


# exported combined types (distinct objects so that HASNEG is not confused with a DOMAIN_TYPE int)
C_NO = CombinedDomainType(0, "NO")
C_NEG = CombinedDomainType(1, "CDOM_NEG")
C_POS = CombinedDomainType(2, "CDOM_POS")
C_UNCUT = CombinedDomainType(3, "UNCUT")
C_IF = CombinedDomainType(4, "CDOM_IF")
C_HASNEG = CombinedDomainType(5, "HASNEG")
C_HASPOS = CombinedDomainType(6, "HASPOS")
C_ANY = CombinedDomainType(7, "ANY")


def HAS(dt):
    """python/__init__.py:30-42"""
    if dt == NEG:
        return C_HASNEG
    if dt == POS:
        return C_HASPOS
    if dt == IF:
        return C_IF
    raise XgpuError("invalid domain type")


class BitArray:
    """device resident BitArray (NGSolve BitArray semantics: fixed size, |, &, ~, in-place ops)"""

    def __init__(self, n, device="cuda", words=None):
        self.n = int(n)
        self.bits = words if words is not None else ops.new_bits(self.n, device)

    def __len__(self):
        return self.n

    @property
    def device(self):
        return self.bits.device

    def NumSet(self):
        return ops.bitset_count(self.bits, self.n)

    def Clear(self):
        ops.bitset_op(ops.BIT_CLEAR, self.bits, self.bits, self.n, out=self.bits)
        return self

    def Set(self):
        ops.bitset_op(ops.BIT_SET, self.bits, self.bits, self.n, out=self.bits)
        return self

    def _bin(self, other, op, out=None):
        assert other.n == self.n
        r = ops.bitset_op(op, self.bits, other.bits, self.n, out=out)
        return BitArray(self.n, words=r) if out is None else self

    def __or__(self, o):
        return self._bin(o, ops.BIT_OR)

    def __and__(self, o):
        return self._bin(o, ops.BIT_AND)

    def __invert__(self):
        return BitArray(self.n, words=ops.bitset_op(ops.BIT_NOT, self.bits, self.bits, self.n))

    def __ior__(self, o):
        return self._bin(o, ops.BIT_OR, out=self.bits)

    def __iand__(self, o):
        return self._bin(o, ops.BIT_AND, out=self.bits)

    def indices(self):
        """ascending list of set bits (device int32 tensor)"""
        return ops.bitset_to_list(self.bits, self.n)

    def to_numpy(self):
        b = self.bits.cpu().numpy().view(np.uint8)
        return np.unpackbits(b, bitorder="little")[: self.n].astype(bool)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.to_numpy()[i]
        w = int(self.bits[i >> 5].item())
        return bool((w >> (i & 31)) & 1)

    def __eq__(self, o):
        return isinstance(o, BitArray) and o.n == self.n and torch.equal(self.bits, o.bits)

    __hash__ = None

    @staticmethod
    def from_numpy(mask, device="cuda"):
        mask = np.asarray(mask, dtype=bool)
        n = mask.size
        packed = np.packbits(mask, bitorder="little")
        buf = np.zeros(ops.nwords(n) * 4, np.uint8)
        buf[: packed.size] = packed
        return BitArray(n, words=torch.from_numpy(buf.view(np.int32).copy()).to(device))


class GridFunction:
    """minimal stand-in for a P1 H1 GridFunction: nodal values on the device (`vec`)"""

    def __init__(self, mesh, values=None):
        self.mesh = mesh
        self.vec = torch.zeros(mesh.nv, dtype=torch.float64, device=mesh.device) if values is None \
            else torch.as_tensor(values, dtype=torch.float64).to(mesh.device).contiguous()


def InterpolateToP1(fn, gf_or_mesh, eps_perturbation=1e-14):
    """utils/p1interpol.cpp:27-113: nodal interpolation, |v| < eps -> +eps (:105-107)."""
    gf = gf_or_mesh if isinstance(gf_or_mesh, GridFunction) else GridFunction(gf_or_mesh)
    p = gf.mesh.points
    v = fn(*[p[:, i] for i in range(p.shape[1])])
    v = torch.as_tensor(v, dtype=torch.float64, device=p.device)
    gf.vec = torch.where(v.abs() < eps_perturbation, torch.full_like(v, eps_perturbation), v).contiguous()
    return gf


def _lset_vec(levelset, mesh):
    if isinstance(levelset, GridFunction):
        v = levelset.vec
    elif isinstance(levelset, torch.Tensor):
        v = levelset
    else:
        raise XgpuError("the accelerated path needs the level set as a P1 GridFunction (nodal values); "
                        "CoefficientFunction level sets / subdivlvl>0 use ngsxfem's legacy path "
                        "(cutint/lsetintdomain.cpp:155-172) which is not provided")
    if v.numel() != mesh.nv:
        raise XgpuError("level set must have one value per mesh vertex (P1)")
    return v.to(mesh.device, torch.float64).contiguous()


class CutInfo:
    """xfem.CutInfo(mesh, levelset=None, subdivlvl=0, time_order=-1, heapsize=1000000)
    (xfem/python_xfem.cpp:54-86 -> CutInformation, xfem/cutinfo.cpp:14-261)."""

    def __init__(self, mesh, levelset=None, subdivlvl=0, time_order=-1, heapsize=1000000):
        self.mesh = mesh
        ne = mesh.ne
        dev = mesh.device
        self._ba = {m: BitArray(ne, dev) for m in range(8)}      # VOL BitArrays by combined mask
        self._ba[7].Set()
        self.cut_ratio = torch.zeros(ne, dtype=torch.float64, device=dev)
        # BND elements (cutinfo.cpp:82-125 loops over VOL and BND): boundary facets as (dim-1)-simplices
        self._bnd = None
        if mesh.dim >= 2:
            bv, _ = mesh.boundary_elements()
            nb = bv.shape[0]
            bb = {m: BitArray(nb, dev) for m in range(8)}
            bb[7].Set()
            self._bnd = (bv, bb, torch.zeros(nb, dtype=torch.float64, device=dev))
        self._updated = False
        if levelset is not None:
            self.Update(levelset, subdivlvl, time_order, heapsize)

    def Mesh(self):
        return self.mesh

    def Update(self, levelset, subdivlvl=0, time_order=-1, heapsize=1000000):
        if subdivlvl != 0:
            raise XgpuError("subdivlvl != 0 selects ngsxfem's legacy subdivision path: not provided")
        m = self.mesh
        ctx = default_ctx(m.device.index)
        b = self._ba
        if hasattr(levelset, "timeFE") or time_order >= 0:
            return self._update_spacetime(levelset, max(time_order, 0), ctx)
        phi = _lset_vec(levelset, m)
        ops.classify(m.dim, m.elements, phi, ctx=ctx,
                     out=(b[CDOM_NEG].bits, b[CDOM_POS].bits, b[CDOM_IF].bits, self.cut_ratio))
        # cutinfo.cpp:127-132
        ops.bitset_op(ops.BIT_OR, b[CDOM_NEG].bits, b[CDOM_POS].bits, m.ne, out=b[UNCUT].bits, ctx=ctx)
        ops.bitset_op(ops.BIT_OR, b[CDOM_NEG].bits, b[CDOM_IF].bits, m.ne, out=b[HASNEG].bits, ctx=ctx)
        ops.bitset_op(ops.BIT_OR, b[CDOM_POS].bits, b[CDOM_IF].bits, m.ne, out=b[HASPOS].bits, ctx=ctx)
        b[NO].Clear(); b[ANY].Set()
        if self._bnd is not None and self._bnd[0].shape[0] > 0:
            bv, bb, br = self._bnd
            nb = bv.shape[0]
            ops.classify(m.dim - 1, bv, phi, ctx=ctx, out=(bb[CDOM_NEG].bits, bb[CDOM_POS].bits, bb[CDOM_IF].bits, br))
            ops.bitset_op(ops.BIT_OR, bb[CDOM_NEG].bits, bb[CDOM_POS].bits, nb, out=bb[UNCUT].bits, ctx=ctx)
            ops.bitset_op(ops.BIT_OR, bb[CDOM_NEG].bits, bb[CDOM_IF].bits, nb, out=bb[HASNEG].bits, ctx=ctx)
            ops.bitset_op(ops.BIT_OR, bb[CDOM_POS].bits, bb[CDOM_IF].bits, nb, out=bb[HASPOS].bits, ctx=ctx)
        ctx.raise_on_geometry_error()
        self._updated = True

    def _update_spacetime(self, levelset, time_order, ctx):
        """cutinfo.cpp:90-124 with a space-time level set: the part volumes are the weight sums of the order-0
        space-time NEG / POS rules over the slab (time_order = 0 suffices to detect the cut topology, :66-67); a
        merely spatial level set is constant in time (cutint/xintegration.cpp:60-70).  BND elements: the reference
        expects no NEG part there (:101-102), they are marked POS."""
        m = self.mesh
        b = self._ba
        if hasattr(levelset, "timeFE"):
            phi_st, kt = levelset.vec, levelset.timeFE.order
        else:
            phi_st, kt = _lset_vec(levelset, m).reshape(1, -1).contiguous(), 0
        ops.classify_st(m.dim, m.elements, phi_st, kt, time_order,
                        out=(b[CDOM_NEG].bits, b[CDOM_POS].bits, b[CDOM_IF].bits, self.cut_ratio), ctx=ctx)
        ctx.raise_on_geometry_error()
        ops.bitset_op(ops.BIT_OR, b[CDOM_NEG].bits, b[CDOM_POS].bits, m.ne, out=b[UNCUT].bits, ctx=ctx)
        ops.bitset_op(ops.BIT_OR, b[CDOM_NEG].bits, b[CDOM_IF].bits, m.ne, out=b[HASNEG].bits, ctx=ctx)
        ops.bitset_op(ops.BIT_OR, b[CDOM_POS].bits, b[CDOM_IF].bits, m.ne, out=b[HASPOS].bits, ctx=ctx)
        b[NO].Clear(); b[ANY].Set()
        if self._bnd is not None and self._bnd[0].shape[0] > 0:
            _, bb, br = self._bnd
            for k in (CDOM_NEG, CDOM_IF, HASNEG, NO):
                bb[k].Clear()
            for k in (CDOM_POS, UNCUT, HASPOS, ANY):
                bb[k].Set()
            br.zero_()
        self._updated = True

    def _bnd_or_raise(self):
        if self._bnd is None:
            raise XgpuError("BND elements of a 1D mesh are points: not provided")
        return self._bnd

    def GetElementsOfType(self, dt, vb=VOL):
        if vb == BND:
            return self._bnd_or_raise()[1][_to_cdt(dt)]
        return self._ba[_to_cdt(dt)]

    def GetCutRatios(self, vb=VOL):
        if vb == BND:
            return self._bnd_or_raise()[2]
        return self.cut_ratio

    def GetElementsWithThresholdContribution(self, domain_type=NEG, threshold=1.0, vb=VOL):
        """xfem/cutinfo.cpp:263-284"""
        if vb == BND:   # cut_ratio_of_element[vb], ne = GetNE(vb)
            cr = self._bnd_or_raise()[2]
            return BitArray(int(cr.numel()), words=ops.threshold_elements(cr, int(domain_type), threshold))
        return BitArray(self.mesh.ne, words=ops.threshold_elements(self.cut_ratio, int(domain_type), threshold))


def _normalize_dtlist(dts, M):
    if hasattr(dts, "as_list"):
        dts = dts.as_list
    if isinstance(dts, tuple):
        dts = [dts]
    if not isinstance(dts, list) or not all(isinstance(t, tuple) for t in dts):
        raise XgpuError("domain type must be a tuple, a list of tuples or a DomainTypeArray")
    for t in dts:
        if len(t) != M:
            raise XgpuError("length of a domain-type tuple does not match the number of level sets")
    return [tuple(int(d) for d in t) for t in dts]


class MultiLevelsetCutInfo:
    """xfem.MultiLevelsetCutInfo (xfem/python_xfem.cpp:510-720, xfem/cutinfo.cpp:286-471):
    results are cached per requested tuple list (order-insensitive) and updated in place."""

    def __init__(self, mesh, levelsets):
        self.mesh = mesh
        self.lsets = [GridFunction(mesh, _lset_vec(l, mesh).clone()) for l in levelsets]
        self._of_type, self._with_contrib = [], []

    def _phis(self):
        return torch.stack([l.vec for l in self.lsets]).contiguous()

    def Mesh(self):
        return self.mesh

    @staticmethod
    def _key(dts):
        return tuple(sorted(dts))

    def GetElementsOfType(self, domain_type, vb=VOL, heapsize=1000000):
        dts = _normalize_dtlist(domain_type, len(self.lsets))
        key = self._key(dts)
        for k, d, ba in self._of_type:
            if k == key:
                return ba
        ba = BitArray(self.mesh.ne, self.mesh.device)
        ops.classify_ml(self.mesh.dim, self.mesh.elements, self._phis(), dts, out=ba.bits)
        self._of_type.append((key, dts, ba))
        return ba

    def _contrib(self, dts, ba):
        m = self.mesh
        # xfem/cutinfo.cpp:306-334: bit set iff the weights of the (order -1) multi-level-set rule sum to > 0
        off, pts, wts = ops.cutquad_ml(m.dim, None, m.elements, m.points, self._phis(), dts, -1)
        default_ctx(m.device.index).raise_on_geometry_error()
        # weight sums in rule order (`for (auto w : wei_arr) sum += w`, :322-325) and the > 0 decision on the device
        ctx = default_ctx(m.device.index)
        seg = ops.rule_element_sums(off, wts, ctx=ctx)
        none = torch.zeros_like(seg)
        scratch = (ops.new_bits(m.ne, m.device), ops.new_bits(m.ne, m.device))
        ops.classify_volumes(seg, none, (ba.bits, scratch[0], scratch[1], None), ctx=ctx)   # "NEG" bit = (sum > 0)

    def GetElementsWithContribution(self, domain_type, vb=VOL, heapsize=1000000):
        dts = _normalize_dtlist(domain_type, len(self.lsets))
        key = self._key(dts)
        for k, d, ba in self._with_contrib:
            if k == key:
                return ba
        ba = BitArray(self.mesh.ne, self.mesh.device)
        self._contrib(dts, ba)
        self._with_contrib.append((key, dts, ba))
        return ba

    def Update(self, levelsets, heapsize=1000000):
        for own, new in zip(self.lsets, levelsets):
            own.vec.copy_(_lset_vec(new, self.mesh))
        for _, dts, ba in self._with_contrib:
            self._contrib(dts, ba)
        for _, dts, ba in self._of_type:
            ops.classify_ml(self.mesh.dim, self.mesh.elements, self._phis(), dts, out=ba.bits)


def GetFacetsWithNeighborTypes(mesh, a, b=None, bnd_val_a=True, bnd_val_b=True, use_and=True, heapsize=1000000):
    """xfem/python_xfem.cpp:346-369 -> xfem/cutinfo.cpp:476-538.  Periodic meshes (mesh.SetPeriodicFacets /
    MakeStructured*Mesh(periodic_x=...)): the facet table carries the identified neighbour (:509-518)"""
    if b is None:
        b = a
    fe = mesh.facet_els_periodic()
    return BitArray(mesh.nfacet, words=ops.facets_with_neighbor_types(fe, a.bits, b.bits, bnd_val_a, bnd_val_b, use_and))


def GetElementsWithNeighborFacets(mesh, a, heapsize=1000000):
    """xfem/cutinfo.cpp:540-566"""
    _, _, fe = mesh.facets()
    return BitArray(mesh.ne, words=ops.elements_with_neighbor_facets(fe, a.bits, mesh.ne))


def GetElementsWithSharedVertex(mesh, a, heapsize=1000000):
    """xfem/cutinfo.cpp:568-597: the elements that share a vertex with an element of a (a itself included)"""
    vmarks = ops.dofs_of_elements(mesh.elements, a.bits, mesh.nv)
    return BitArray(mesh.ne, words=ops.elements_with_marked(mesh.elements, vmarks))


def GetDofsOfElements(space, a, heapsize=1000000):
    """xfem/cutinfo.cpp:599-622"""
    return BitArray(space.ndof, words=ops.dofs_of_elements(space.el_dofs, a.bits, space.ndof))
