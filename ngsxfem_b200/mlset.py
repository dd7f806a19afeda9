"""xfem.mlset on the accelerated path: the host-side description of regions bounded by several level sets
(python/mlset.py of the reference: `DomainTypeArray` :10-445, `TensorUnion` :448-483, `TensorIntersection` :486-519).

A region is a set of tuples over {NEG, POS, IF}, one entry per level set; a tuple with k entries IF is a piece of
co-dimension k.  The class is what `dCut(levelsets, dta)`, `Integrate({"levelset": ..., "domain_type": dta})` and
`MultiLevelsetCutInfo.GetElementsOfType / GetElementsWithContribution` take (they read `.as_list`); the algebra below (union,
intersection, complement, boundary, ANY expansion, compression to the pieces of positive measure) is plain set arithmetic on
the host — it is an input description, not part of the hot path — and follows the reference's semantics,
pinned by its own known answers (tests/pytests/test_mlset_module_functionality.py, test_mlset_integration.py:170-253).

Compression needs the measure of a piece: `Integrate(..., order=0)` on the device (cutint/python_cutint.cpp:21-163 through
xgpu_cutquad_ml_*).  The integrator can be replaced (`integrator=`: a function (level sets, tuple) -> measure) — the tests on a
host without a GPU pass their own.
"""
import itertools

from ._lib import XgpuError
from .cutinfo import NEG, POS, IF, C_ANY as ANY

_SIGNED = (POS, NEG)


def _is_any(d):
    return d is ANY


def _entry(d):
    """NEG / POS / IF as plain ints, ANY as the wildcard object; anything else is refused (python/mlset.py:85-88)"""
    if _is_any(d):
        return ANY
    if isinstance(d, bool) or not isinstance(d, int) or d not in (NEG, POS, IF):
        raise XgpuError("DomainTypeArray: initialised with something other than NEG, POS, IF, ANY")
    return int(d)


def _default_integrator(lsets, dtt):
    from .forms import Integrate
    mesh = lsets[0].mesh
    return Integrate({"levelset": list(lsets), "domain_type": dtt}, mesh=mesh, cf=1.0, order=0)


class DomainTypeArray:
    """DomainTypeArray(dtlist, lsets=None, persistent_compress=False, do_safety_checks=True)  (python/mlset.py:10-132)

    dtlist: a tuple or a list of tuples over NEG / POS / IF / ANY, all of one length and one co-dimension; every ANY stands
    for POS and NEG.  lsets: P1 level-set GridFunctions; when given, the pieces without measure are dropped at once, and kept
    dropped through later operations if persistent_compress is set.
    Attributes: as_list (expanded tuples, no duplicates), codim, lsets, persistent_compress."""

    def __init__(self, dtlist, lsets=None, persistent_compress=False, do_safety_checks=True, integrator=None):
        if type(dtlist) is tuple:
            dtlist = [dtlist]
        if type(dtlist) is not list or not dtlist or type(dtlist[0]) is not tuple:
            raise TypeError("Invalid input: dtlist must be a tuple or a list of tuples")
        width = len(dtlist[0])
        self.codim = sum(1 for d in dtlist[0] if (not _is_any(d)) and d == IF)
        if do_safety_checks:
            for t in dtlist:
                if type(t) is not tuple:
                    raise TypeError("Invalid input: dtlist must be a tuple or a list of tuples")
                ent = [_entry(d) for d in t]
                if sum(1 for d in ent if (not _is_any(d)) and d == IF) != self.codim:
                    raise XgpuError("DomainTypeArray initialised with tuples of different co-dimension!")
                if len(t) != width:
                    raise XgpuError("DomainTypeArray initialised with tuples of different length!")
            if lsets:
                if len(lsets) != width:
                    raise XgpuError("The number of level sets does not match the length of the arrays domain tuples!")
                if not all(hasattr(l, "vec") and hasattr(l, "mesh") for l in lsets):
                    raise XgpuError("The level set functions need to be P1 GridFunctions!")
        seen, out = set(), []
        for t in dtlist:
            where = [i for i, d in enumerate(t) if _is_any(d)]
            base = [d if _is_any(d) else int(d) for d in t]
            for choice in itertools.product(_SIGNED, repeat=len(where)):   # no ANY: the tuple itself
                cand = list(base)
                for i, d in zip(where, choice):
                    cand[i] = d
                cand = tuple(cand)
                if cand not in seen:
                    seen.add(cand)
                    out.append(cand)
        self.as_list = out
        self._integrator = integrator
        self.lsets = lsets
        self.persistent_compress = bool(persistent_compress)
        if self.lsets:
            self.Compress(self.lsets, self.persistent_compress)
        if not self.persistent_compress:
            self.lsets = None

    # ---- container protocol
    def __len__(self):
        return len(self.as_list)

    def __iter__(self):
        return iter(self.as_list)

    def __contains__(self, dtt):
        return tuple(dtt) in self.as_list

    def __eq__(self, o):
        if not isinstance(o, DomainTypeArray):
            return NotImplemented
        return self.codim == o.codim and len(self) == len(o) and set(self.as_list) == set(o.as_list)

    def __ne__(self, o):
        r = self.__eq__(o)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __str__(self):
        name = {NEG: "NEG", POS: "POS", IF: "IF"}
        doms = ", ".join("(" + ", ".join(name[d] for d in t) + ")" for t in self.as_list)
        return f"DomainTypeArray:\n  Co-dim : {self.codim}\n  Domains: [{doms}]"

    __repr__ = __str__

    # ---- algebra
    def _combined_state(self, o):
        """level sets / persistence of the result of a binary operation (python/mlset.py:160-176): kept only if BOTH
        operands compress persistently with respect to the same level sets"""
        if self.persistent_compress and o.persistent_compress:
            if len(self.lsets) != len(o.lsets) or any(a is not b for a, b in zip(self.lsets, o.lsets)):
                raise XgpuError("Operator only possible for domains with the same level sets")
            return self.lsets, True
        return None, False

    def _make(self, tuples, lsets, persistent):
        if not tuples:   # the reference's constructor cannot represent the empty region either (dtlist[0])
            raise XgpuError("DomainTypeArray: the operation leaves no region")
        return DomainTypeArray(list(tuples), lsets, persistent, integrator=self._integrator)   # safety checks on, as the reference

    def __or__(self, o):
        if self.codim != o.codim:
            raise XgpuError("Co-dims don't match: Union not possible!")
        lsets, pers = self._combined_state(o)
        return self._make(self.as_list + o.as_list, lsets, pers)

    def __and__(self, o):
        lsets, pers = self._combined_state(o)
        common = [t for t in self.as_list if t in o.as_list]
        if not common:
            # no tuple in common: pieces of higher co-dimension where the two regions touch — entries that differ become IF
            # unless they are POS against NEG (python/mlset.py:201-216)
            for a, b in itertools.product(self.as_list, o.as_list):
                piece = []
                for da, db in zip(a, b):
                    if da == db:
                        piece.append(da)
                    elif da in _SIGNED and db in _SIGNED:
                        piece = None
                        break
                    else:
                        piece.append(IF)
                if piece is not None:
                    common.append(tuple(piece))
        return self._make(common, lsets, pers)

    def __invert__(self):
        n = len(self.as_list[0])
        alphabet = _SIGNED if self.codim == 0 else (POS, NEG, IF)
        have = set(self.as_list)
        rest = [t for t in itertools.product(alphabet, repeat=n) if t.count(IF) == self.codim and t not in have]
        return self._make(rest, self.lsets, self.persistent_compress)

    def __ior__(self, o):
        return self.__or__(o)

    def __iand__(self, o):
        return self.__and__(o)

    # ---- measure-dependent operations
    def _measure(self, lsets, dtt):
        fn = self._integrator or _default_integrator
        return float(fn(lsets, dtt))

    def Compress(self, lsets, persistent=False):
        """drop the pieces whose measure with respect to `lsets` is not positive (|.| <= 1e-12, python/mlset.py:262-272);
        persistent: later Boundary() / operators compress too"""
        if len(lsets) != len(self.as_list[0]):
            raise XgpuError("The number of level sets does not match the length of the arrays domain tuples!")
        if not all(hasattr(l, "vec") and hasattr(l, "mesh") for l in lsets):
            raise XgpuError("The level set functions need to be P1 GridFunctions!")
        self.as_list = [t for t in self.as_list if abs(self._measure(lsets, t)) > 1e-12]
        if persistent:
            self.persistent_compress = True
            self.lsets = lsets
        return None

    def Boundary(self):
        """the boundary of the region: every signed entry of every tuple turned into IF, unless the region holds the tuple with
        POS AND the one with NEG in that place (an interior interface); compressed if the instance compresses persistently
        (python/mlset.py:280-323)"""
        if self.codim >= 3:
            raise XgpuError("Boundary does not make sense for codim >=3")
        have = set(self.as_list)
        out = []
        for t in self.as_list:
            for i, d in enumerate(t):
                if d == IF:
                    continue
                if t[:i] + (POS,) + t[i + 1:] in have and t[:i] + (NEG,) + t[i + 1:] in have:
                    continue
                piece = t[:i] + (IF,) + t[i + 1:]
                if self.lsets and abs(self._measure(self.lsets, piece)) < 1e-12:
                    continue
                out.append(piece)
        if self.persistent_compress:
            return self._make(out, self.lsets, True)
        return self._make(out, None, False)

    # ---- indicator functions (coefficient functions of the coordinates)
    def _lset_values(self, lsets):
        """lsets: callables of the coordinates or H1 GridFunctions given as (space, coefficient vector) pairs;
        returns (gfs for CoefficientFunction, function(all arguments) -> list of level-set values)"""
        gfs, kinds = [], []
        for l in lsets:
            if callable(l):
                kinds.append(("fn", l))
            elif isinstance(l, tuple) and len(l) == 2:
                kinds.append(("gf", len(gfs)))
                gfs.append(l)
            else:
                raise XgpuError("Indicator: level sets are callables of the coordinates or (H1 space, coefficient vector) pairs")
        ng = len(gfs)

        def values(*args):
            coords, gv = args[:len(args) - ng], args[len(args) - ng:]
            return [k[1](*coords) if k[0] == "fn" else gv[k[1]] for k in kinds]

        return gfs, values

    def Indicator(self, lsets):
        """1 in the region, 0 elsewhere (codim 0; python/mlset.py:325-362): a CoefficientFunction for Integrate / forms"""
        if self.codim != 0:
            return self.IndicatorSmoothed(lsets)
        from .forms import CoefficientFunction
        import torch
        if len(lsets) != len(self.as_list[0]):
            raise XgpuError("The number of level sets does not match the length of the arrays domain tuples!")
        gfs, values = self._lset_values(lsets)
        tuples = list(self.as_list)

        def fn(*args):
            v = values(*args)
            total = torch.zeros_like(torch.as_tensor(v[0], dtype=torch.float64))
            for t in tuples:
                ind = torch.ones_like(total)
                for i, d in enumerate(t):
                    if d == POS:
                        ind = ind * (v[i] > 0)
                    elif d == NEG:
                        ind = ind * (-v[i] > 0)
                total = total + ind
            return (total > 0).to(torch.float64)

        return CoefficientFunction(fn, gfs)

    def IndicatorSmoothed(self, lsets, eps=0.03):
        """1 within eps of the region of co-dimension > 0 (level sets taken as signed distances; python/mlset.py:364-412)"""
        if self.codim == 0:
            return self.Indicator(lsets)
        from .forms import CoefficientFunction
        import torch
        if len(lsets) != len(self.as_list[0]):
            raise XgpuError("The number of level sets does not match the length of the arrays domain tuples!")
        gfs, values = self._lset_values(lsets)
        tuples = list(self.as_list)

        def fn(*args):
            v = values(*args)
            total = torch.zeros_like(torch.as_tensor(v[0], dtype=torch.float64))
            for t in tuples:
                ind = torch.ones_like(total)
                for i, d in enumerate(t):
                    if d == IF:
                        ind = ind * (~(torch.abs(v[i]) - eps > 0))
                    elif d == POS:
                        ind = ind * (v[i] - eps > 0)
                    else:
                        ind = ind * (-v[i] + eps > 0)
                total = total + ind
            return (total > 0).to(torch.float64)

        return CoefficientFunction(fn, gfs)

    def GetOuterNormals(self, lsetsp1):
        """{boundary tuple: outward unit normal} over the pieces of Boundary() (codim 0; python/mlset.py:414-445): on the piece
        where entry i becomes IF the normal is +Normalize(grad(lsetsp1[i])) for a NEG entry, the negative for a POS entry.
        The values are the forms layer's level-set normals: grad(u) * n[tuple]."""
        if self.codim > 0:
            raise XgpuError("GetOuterNormals is only available for codim = 0 !")
        from .forms import Normal
        bnd = self.Boundary()
        out = {}
        for t in self.as_list:
            for i, d in enumerate(t):
                piece = t[:i] + (IF,) + t[i + 1:]
                if piece in bnd:
                    out[piece] = Normal(lsetsp1[i], sign=-1.0 if d == POS else 1.0)
        return out


def TensorUnion(*args):
    """the union of regions that depend on DIFFERENT sets of level sets: the tuples are placed side by side, the places of the
    other arrays filled with ANY (python/mlset.py:448-483)"""
    if not args:
        raise TypeError("TensorUnion needs at least one DomainTypeArray")
    for a in args:
        if type(a) is not DomainTypeArray:
            raise TypeError("TensorUnion only possible for DomainTypeArrays")
        if a.codim != args[0].codim:
            raise XgpuError("Cannot form TensorUnion for arrays of different codimension")
    widths = [len(a.as_list[0]) for a in args]
    total, out, at = sum(widths), [], 0
    for a, w in zip(args, widths):
        for t in a.as_list:
            out.append((ANY,) * at + t + (ANY,) * (total - at - w))
        at += w
    return DomainTypeArray(out)


def TensorIntersection(*args):
    """the intersection of regions that depend on different sets of level sets: all concatenations of one tuple of each
    array (python/mlset.py:486-519)"""
    if not args:
        raise TypeError("TensorIntersection needs at least one DomainTypeArray")
    for a in args:
        if type(a) is not DomainTypeArray:
            raise TypeError("TensorIntersection only possible for DomainTypeArrays")
    return DomainTypeArray([tuple(itertools.chain(*ts)) for ts in itertools.product(*[a.as_list for a in args])])
