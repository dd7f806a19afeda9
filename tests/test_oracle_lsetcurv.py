"""Known-answer pin of the oracle's deformation producer (oracle/xo_shift.c: ProjectShift, lsetcurving/projshift.cpp:9-201 +
shiftintegrators.cpp:39-124 + calcpointshift.cpp:145-206) by the reference's own test tests/pytests/test_intcurved.py:17-91:
levelset sqrt(x^2+y^2) - 0.5 on [-1,1]^2 (triangles), LevelSetMeshAdaptation(order, threshold=0.2, discontinuous_qn=True);
the integrals of 1 over NEG / POS / IF on the DEFORMED mesh must hit pi/4, 4 - pi/4, pi with error < 1e-5 (order > 1) on
the finest mesh and an average convergence order > order + 0.75."""
from math import log, pi

import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from oracle import xoracle as xo

NEG, POS, IF = 0, 1, 2
REF = {NEG: 0.25 * pi, POS: 4.0 - 0.25 * pi, IF: pi}


def adaptation(n, order, threshold=0.2, qn_mode=1):
    """the oracle's LevelSetMeshAdaptation.CalcDeformation: returns mesh arrays, lset_p1, dof table, deformation"""
    mesh = xf.MakeStructured2DMesh(nx=n, ny=n, mapping=lambda x, y: (2 * x - 1, 2 * y - 1))
    V = xf.H1(mesh, order=order)
    pts, els, eld = mesh.points.numpy(), mesh.elements.numpy(), V.el_dofs.numpy()
    xq = xo.rule_points_phys(2, 2 * order, els, pts)
    f = np.sqrt(xq[..., 0] ** 2 + xq[..., 1] ** 2) - 0.5
    lset_ho = xo.set_from_points(2, order, 2 * order, els, eld, V.ndof, f)          # lset_ho.Set(levelset)
    lset_p1 = lset_ho[:mesh.nv].copy()                                              # InterpolateToP1: vertex dofs
    lset_p1[np.abs(lset_p1) < 1e-14] = 1e-14
    qn = xo.set_gradient(2, order, els, pts, eld, V.ndof, lset_ho) if qn_mode == 2 else None
    deform, nfail = xo.project_shift(2, order, els, pts, eld, V.ndof, lset_ho, lset_p1, qn_mode=qn_mode, qn=qn, threshold=threshold)
    assert nfail == 0
    return pts, els, eld, lset_p1, deform


@pytest.mark.parametrize("order", [1, 2, 3])
def test_intcurved_known_answers_and_orders(order):
    levels = 5 if order > 1 else 6
    err = {k: [] for k in REF}
    for lvl in range(levels):
        pts, els, eld, lset_p1, deform = adaptation(4 * 2 ** lvl, order)
        for k in REF:
            v = xo.integrate_one_deformed(2, els, pts, lset_p1, k, order, def_order=order, el_dofs_def=eld, def_vals=deform)
            err[k].append(abs(v - REF[k]))
    for k in REF:
        eoc = [log(a / b) / log(2) for a, b in zip(err[k][:-1], err[k][1:])]
        assert err[k][-1] < (1e-5 if order > 1 else 4e-4)     # the reference runs two more levels at order 1 (1e-4 there)
        s = 2 if order == 1 else 0
        assert sum(eoc[s:]) / len(eoc[s:]) > order + 0.75, (k, err[k], eoc)


def test_continuous_search_direction_gives_the_same_accuracy():
    """qn as an H1(order)^2 field set from grad lset_ho (the default of LevelSetMeshAdaptation, python/lsetcurv.py:309-310)"""
    pts, els, eld, lset_p1, deform = adaptation(32, 2, qn_mode=2)
    for k in REF:
        v = xo.integrate_one_deformed(2, els, pts, lset_p1, k, 2, def_order=2, el_dofs_def=eld, def_vals=deform)
        assert abs(v - REF[k]) < 2e-5
    # the uncurved P1 geometry is two orders worse
    v0 = xo.integrate_one_deformed(2, els, pts, lset_p1, IF, 2)
    assert abs(v0 - pi) > 1e-3


def test_deformation_vanishes_outside_the_band_and_at_vertices():
    pts, els, eld, lset_p1, deform = adaptation(16, 2)
    nv = pts.shape[0]
    assert np.all(deform[:nv] == 0.0)                       # vertex values to zero (projshift.cpp:108-109)
    cut = np.array([xo.element_in_band(lset_p1[e], 0.0, 0.0) for e in els])
    touched = np.zeros(deform.shape[0], dtype=bool)
    touched[eld[cut].reshape(-1)] = True
    assert np.all(deform[~touched] == 0.0) and np.abs(deform[touched]).max() > 0
    # threshold: |reference shift| <= 0.2 pointwise (shiftintegrators.cpp:108-111) — a tiny threshold shrinks the deformation
    _, _, _, _, d_small = adaptation(16, 2, threshold=1e-3)
    assert np.abs(d_small).max() < 0.1 * np.abs(deform).max()


def test_polynomial_level_set_is_reproduced_by_set():
    """GridFunction.Set restated: a quadratic is reproduced exactly by the order-2 space (projection + mean)"""
    mesh = xf.MakeStructured2DMesh(nx=5, ny=4)
    V = xf.H1(mesh, order=2)
    pts, els, eld = mesh.points.numpy(), mesh.elements.numpy(), V.el_dofs.numpy()
    xq = xo.rule_points_phys(2, 4, els, pts)
    f = 1.0 + xq[..., 0] - 2 * xq[..., 1] + 3 * xq[..., 0] * xq[..., 1] - xq[..., 1] ** 2
    c = xo.set_from_points(2, 2, 4, els, eld, V.ndof, f)
    fv = 1.0 + pts[:, 0] - 2 * pts[:, 1] + 3 * pts[:, 0] * pts[:, 1] - pts[:, 1] ** 2
    assert np.allclose(c[:mesh.nv], fv, atol=1e-12)
    g = xo.set_gradient(2, 2, els, pts, eld, V.ndof, c)     # gradient is linear: reproduced, vertex dofs = nodal values
    assert np.allclose(g[:mesh.nv, 0], 1 + 3 * pts[:, 1], atol=1e-11)
    assert np.allclose(g[:mesh.nv, 1], -2 + 3 * pts[:, 0] - 2 * pts[:, 1], atol=1e-11)
