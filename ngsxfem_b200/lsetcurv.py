"""Isoparametric mesh deformation from a level set — the mirror of python/lsetcurv.py:13-356 (LevelSetMeshAdaptation) and of
the pybind entry ProjectShift (lsetcurving/python_lsetcurving.cpp, lsetcurving/projshift.cpp:9-201) over the kernels of
csrc/lsetcurv.cu.  SURVEY.md §8 row f3: with it the cut integrals / forms with `deformation=` (fes.Deformation) run on a
deformation that is PRODUCED here instead of handed in.

What is mirrored: order, threshold, discontinuous_qn, lset_lower_bound / lset_upper_bound, eps_perturbation, levelset=,
CalcDeformation(levelset, ba=, blending=) with blending None / "none" / "quadratic" / "quartic" / callable, lset_p1, lset_ho,
qn, deform, the context manager (mesh.SetDeformation / UnsetDeformation), levelset_domain.
Not mirrored (raise): boundary_tangential, ProjectOnUpdate / ProjectGFs (shifted evaluation, lsetcurving/shiftedevaluate.cpp),
CalcMaxDistance (IntegrationPointExtrema), MarkForRefinement.

A level set is a callable of the coordinate tensors (x, y[, z]) -> tensor, evaluated on the device with torch (it is the
user's function, as the CoefficientFunction is in the reference); everything else is CUDA behind the C ABI."""
import math

import torch

from . import ops
from .cutinfo import GridFunction
from .fes import H1, Deformation
from ._lib import XgpuError


def _set_h1(space, fvals_points, order_rule, ctx=None):
    """GridFunction.Set: element-local L2 projection of the point values (ne x np [x ncomp]) + mean over the elements
    sharing a dof.  returns ndof [x ncomp] coefficients of the hierarchical basis"""
    mesh = space.mesh
    elres = ops.l2_project(mesh.dim, space.order, order_rule, None, mesh.elements, fvals_points, ctx=ctx)
    plan = space._all_elements_plan(ctx)
    ncomp = elres.shape[0]
    out = torch.zeros((space.ndof, ncomp), dtype=torch.float64, device=mesh.device)
    for c in range(ncomp):
        ops.vec_scatter_mean(plan, [(space.el_dofs, elres[c])], out, comp=c)
    return out[:, 0].contiguous() if fvals_points.dim() == 2 else out


def _all_elements_plan(space, ctx=None):
    """transposed table of the whole dof table (cached on the space: the mesh topology does not change)"""
    if getattr(space, "_plan_all", None) is None:
        space._plan_all = ops.adjacency_plan(space.ndof, [space.el_dofs], ctx=ctx)
    return space._plan_all


H1._all_elements_plan = _all_elements_plan


def ProjectShift(lset_ho, lset_p1, deform, qn, ba=None, blending=None, lower=0.0, upper=0.0, threshold=-1.0, ctx=None):
    """lsetcurving/projshift.cpp:9-201.  lset_ho: (space, coefficient vector); lset_p1: P1 GridFunction; deform: Deformation of
    the same space (overwritten); qn: None (grad lset_p1), "discontinuous" (grad lset_ho pointwise = an L2 qn field set from
    lset_ho.Deriv()) or an ndof x dim coefficient tensor of an H1(order)^dim field; ba: BitArray words of the elements to treat
    (None: ElementInRelevantBand(lower, upper)); blending: None or n_band x np values of alpha at the rule points —
    LevelSetMeshAdaptation passes a callable instead.  returns the number of point searches that did not converge."""
    space, vals_ho = lset_ho
    mesh, dim, p = space.mesh, space.mesh.dim, space.order
    ctx = ctx or ops.default_ctx(mesh.device.index)
    if deform.space is not space and (deform.space.order != p or deform.space.mesh is not mesh):
        raise XgpuError("ProjectShift: the deformation lives in H1(order)^dim of the level set's order")
    bits = ops.band_elements(dim, mesh.elements, lset_p1.vec, lower, upper, ctx=ctx) if ba is None else (ba.bits if hasattr(ba, "bits") else ba)
    lst = ops.bitset_to_list(bits, mesh.ne, ctx=ctx)
    deform.vals.zero_()
    if lst.numel() == 0:
        return 0
    if qn is None:
        mode, q = 0, None
    elif isinstance(qn, str):
        if qn != "discontinuous":
            raise XgpuError("ProjectShift: qn is None, 'discontinuous' or a coefficient tensor")
        mode, q = 1, None
    else:
        mode, q = 2, qn.contiguous()
    if callable(blending):
        xq = ops.rule_points_phys(dim, 2 * p, lst, mesh.elements, mesh.points, ctx=ctx)
        blending = blending(lst, xq).contiguous()
        if float(blending.max()) > 1 + 1e-6:
            raise XgpuError("alpha should not be larger than 1")   # shiftintegrators.cpp:93-94
    shift, nfail = ops.shift_points(dim, p, lst, mesh.elements, mesh.points, space.el_dofs, vals_ho, lset_p1.vec, mode, q,
                                    blending, threshold, ctx=ctx)
    elres = ops.l2_project(dim, p, 2 * p, lst, mesh.elements, shift, zero_vertex=True, ctx=ctx)
    table = space.el_dofs[lst.long()].contiguous()
    plan = ops.adjacency_plan(space.ndof, [table], ctx=ctx)
    for c in range(dim):
        ops.vec_scatter_mean(plan, [(table, elres[c])], deform.vals, comp=c)
    return nfail


class LevelSetMeshAdaptation:
    """python/lsetcurv.py:13-356"""

    def __init__(self, mesh, order=2, lset_lower_bound=0.0, lset_upper_bound=0.0, threshold=-1.0, discontinuous_qn=False,
                 eps_perturbation=1e-14, heapsize=1000000, levelset=None, boundary_tangential=False, boundary_normal=None):
        if boundary_tangential:
            raise XgpuError("LevelSetMeshAdaptation(boundary_tangential=...) is not provided (python/lsetcurv.py:160-240)")
        self.mesh = mesh
        self.order_deform = self.order_qn = self.order_lset = int(order)
        self.lset_lower_bound, self.lset_upper_bound = lset_lower_bound, lset_upper_bound
        self.threshold, self.eps_perturbation = threshold, eps_perturbation
        self.discontinuous_qn = bool(discontinuous_qn)
        self.v_ho = H1(mesh, order=self.order_lset)
        self.v_def = self.v_ho                      # H1(order)^dim shares the scalar dof table
        self.lset_ho = torch.zeros(self.v_ho.ndof, dtype=torch.float64, device=mesh.device)
        self.lset_p1 = GridFunction(mesh)
        self.qn = None
        self.deform = Deformation(self.v_def)
        self.deform_last = Deformation(self.v_def)
        self.nfail = 0
        if levelset is not None:
            self.CalcDeformation(levelset)

    def ProjectOnUpdate(self, *a, **k):
        raise XgpuError("ProjectOnUpdate / shifted_eval (lsetcurving/shiftedevaluate.cpp) is not provided")

    def CalcDeformation(self, levelset, ba=None, blending=None, dont_project_gfs=False):
        """python/lsetcurv.py:281-356"""
        mesh, dim, p = self.mesh, self.mesh.dim, self.order_lset
        ctx = ops.default_ctx(mesh.device.index)
        # lset_ho.Set(levelset)
        xq = ops.rule_points_phys(dim, 2 * p, None, mesh.elements, mesh.points, ctx=ctx)
        f = levelset(*[xq[..., c] for c in range(dim)])
        f = torch.as_tensor(f, dtype=torch.float64, device=mesh.device).expand(xq.shape[:2]).contiguous()
        self.lset_ho = _set_h1(self.v_ho, f, 2 * p, ctx)
        # qn.Set(lset_ho.Deriv()): an L2 field reproduces the gradient, an H1 field is its projection + mean
        if self.discontinuous_qn:
            self.qn = "discontinuous"
        else:
            g = ops.h1_grad_points(dim, p, 2 * p, None, mesh.elements, mesh.points, self.v_ho.el_dofs, self.lset_ho, ctx=ctx)
            self.qn = _set_h1(self.v_ho, g, 2 * p, ctx)
        # InterpolateToP1(lset_ho, lset_p1): the vertex dofs are the vertex values (utils/p1interpol.cpp:93-107)
        v = self.lset_ho[:mesh.nv]
        self.lset_p1.vec = torch.where(v.abs() < self.eps_perturbation, torch.full_like(v, self.eps_perturbation), v).contiguous()
        bl = None
        if blending is not None and blending != "none":
            if isinstance(blending, str):
                if blending not in ("quadratic", "quartic"):
                    raise XgpuError("blending: None, 'none', 'quadratic', 'quartic' or a callable of the coordinates")
                power = 2 if blending == "quadratic" else 4
                bl = lambda lst, xq_: self._lset_blending(lst, xq_, power)     # noqa: E731
            else:
                bl = lambda lst, xq_: torch.as_tensor(blending(*[xq_[..., c] for c in range(dim)]), dtype=torch.float64,  # noqa: E731
                                                      device=mesh.device).expand(xq_.shape[:2])
        self.deform_last.vals.copy_(self.deform.vals)
        self.nfail = ProjectShift((self.v_ho, self.lset_ho), self.lset_p1, self.deform, self.qn, ba, bl,
                                  lower=self.lset_lower_bound, upper=self.lset_upper_bound, threshold=self.threshold, ctx=ctx)
        return self.deform

    def _lset_blending(self, lst, xq, power):
        """lset_p1^k / (sqrt(dim) h)^k at the rule points, h = specialcf.mesh_size = |det F|^(1/dim) (python/lsetcurv.py:326-333)"""
        mesh, dim, p = self.mesh, self.mesh.dim, self.order_lset
        ev = mesh.elements[lst.long()].long()
        X = mesh.points[ev]                                     # n x (dim+1) x dim
        F = (X[:, :dim, :] - X[:, dim:dim + 1, :]).transpose(1, 2)
        h = torch.linalg.det(F).abs() ** (1.0 / dim)
        pr, _ = ops.get_rule(dim, 2 * p)
        pr = pr[:, :dim].to(mesh.device)
        lam = torch.cat([pr, 1.0 - pr.sum(dim=1, keepdim=True)], dim=1)   # np x (dim+1)
        l1 = self.lset_p1.vec[ev] @ lam.T                        # n x np
        return (l1 / (math.sqrt(dim) * h[:, None])) ** power

    def __enter__(self):
        self.mesh.SetDeformation(self.deform)
        return self.lset_p1

    def __exit__(self, type, value, tb):
        self.mesh.UnsetDeformation()

    def levelset_domain(self, domain_type=2):
        return {"levelset": self.lset_p1, "domain_type": domain_type}
