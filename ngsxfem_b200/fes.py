"""H1 dof tables (input producer) and the restricted / compressed space of ngsxfem.

NGSolve's H1 FESpace numbering (third party) restated: vertex dofs, then (p-1) dofs per edge,
then (p-1)(p-2)/2 per face (3D) resp. cell (2D), then cell bubbles.  `Restrict` mirrors
utils/restrictedfespace.cpp:23-73 (active dofs = dofs of active elements, compressed ascending,
inactive elements have no dofs).
"""
import torch

from . import ops


class H1:
    def __init__(self, mesh, order=1, dgjumps=False, dirichlet=None):
        self.mesh, self.order, self.dgjumps = mesh, int(order), dgjumps
        p, dim = self.order, mesh.dim
        dev = mesh.device
        nv, ne = mesh.nv, mesh.ne
        cols = [mesh.elements.to(torch.int64)]
        off = nv
        if p >= 2:
            if dim == 1:
                ar = torch.arange(ne, device=dev, dtype=torch.int64)
                for k in range(p - 1):
                    cols.append((off + ar * (p - 1) + k).reshape(-1, 1))
                off += ne * (p - 1)
            else:
                _, el_edges = mesh.edges()
                ee = el_edges.to(torch.int64)
                for e in range(ee.shape[1]):
                    for k in range(p - 1):
                        cols.append((off + ee[:, e] * (p - 1) + k).reshape(-1, 1))
                off += mesh.nedge * (p - 1)
        if p >= 3 and dim >= 2:
            nfd = (p - 1) * (p - 2) // 2
            if dim == 3:
                _, el_faces = mesh.faces()
                ef = el_faces.to(torch.int64)
                for f in range(4):
                    for k in range(nfd):
                        cols.append((off + ef[:, f] * nfd + k).reshape(-1, 1))
                off += mesh.nface * nfd
                ncd = (p - 1) * (p - 2) * (p - 3) // 6
            else:
                ncd = nfd
            ar = torch.arange(ne, device=dev, dtype=torch.int64)
            for k in range(ncd):
                cols.append((off + ar * ncd + k).reshape(-1, 1))
            off += ne * ncd
        self.ndof = int(off)
        self.el_dofs = torch.cat(cols, dim=1).to(torch.int32).contiguous()
        self.nd = self.el_dofs.shape[1]
        assert self.nd == ops.h1_ndof(dim, p)
        self.active_elements = None

    def TnT(self):
        from .forms import TrialFunction, TestFunction
        return TrialFunction(self), TestFunction(self)

    def TrialFunction(self):
        return self.TnT()[0]

    def TestFunction(self):
        return self.TnT()[1]


class Deformation:
    """isoparametric mesh deformation: a GridFunction of H1(order)^dim (python/lsetcurv.py:281-356 produces it
    in the reference; here it is an INPUT).  vals: ndof x dim coefficients in the hierarchical basis of `space`."""

    def __init__(self, space, vals=None):
        self.space, self.order = space, space.order
        self.el_dofs = space.el_dofs
        dim = space.mesh.dim
        if vals is None:
            vals = torch.zeros((space.ndof, dim), dtype=torch.float64, device=space.mesh.device)
        self.vals = torch.as_tensor(vals, dtype=torch.float64).to(space.mesh.device).contiguous()
        assert self.vals.shape == (space.ndof, dim)

    @staticmethod
    def interpolate_p2(space, fn):
        """nodal interpolation of a displacement field into an order-2 space: vertex dofs d(v); the order-2
        edge function is L_2(lam_b-lam_a, lam_a+lam_b) = -2 lam_a lam_b (h1.cuh), so its coefficient is
        d(a) + d(b) - 2 d(midpoint).  Exact for quadratic fields."""
        assert space.order == 2
        mesh = space.mesh
        p = mesh.points
        dv = fn(p)
        edges, _ = mesh.edges()
        a, b = p[edges[:, 0].long()], p[edges[:, 1].long()]
        de = fn(a) + fn(b) - 2.0 * fn(0.5 * (a + b))
        return Deformation(space, torch.cat([dv, de], dim=0))

    def local(self):
        """per-element coefficients ne x nd x dim (gathered through el_dofs)"""
        return self.vals[self.el_dofs.long()]


class RestrictedH1:
    """Restrict(fes, active_elements): CompressedFESpace over the dofs of the active elements."""

    def __init__(self, base, active_bits):
        self.base, self.mesh, self.order, self.nd = base, base.mesh, base.order, base.nd
        self.dgjumps = base.dgjumps
        self.active_elements = active_bits
        self.active_dofs = ops.dofs_of_elements(base.el_dofs, active_bits, base.ndof)
        self.dofmap, self.ndof = ops.bitset_rank(self.active_dofs, base.ndof)
        self._el_dofs = None

    @property
    def el_dofs(self):
        """full table: rows of inactive elements are -1 (no dofs, restrictedfespace.cpp:62-73)"""
        if self._el_dofs is None:
            t = ops.gather_rows(self.base.el_dofs, None, self.dofmap)
            ne = self.mesh.ne
            idx = torch.arange(ne, device=t.device)
            act = ((self.active_elements[(idx >> 5)] >> (idx & 31)) & 1).bool()
            t[~act] = -1
            self._el_dofs = t
        return self._el_dofs

    def rows(self, elem_list):
        """dof rows of the listed elements in the compressed numbering"""
        return ops.gather_rows(self.base.el_dofs, elem_list, self.dofmap)

    TnT = H1.TnT
    TrialFunction = H1.TrialFunction
    TestFunction = H1.TestFunction


def Restrict(fes, active_elements):
    bits = active_elements.bits if hasattr(active_elements, "bits") else active_elements
    return RestrictedH1(fes, bits)
