"""Known-answer pin of the oracle's deformed-mesh arithmetic (x = x_affine + sum d_j phi_j, F, |det F|, interface
weights w ||F^{-T} n_ref||: NGSolve ElementTransformation + deformation as used by cutint/straightcutrule.cpp:558-578
and xfem/symboliccutbfi.cpp:135) on a configuration with closed-form answers:

  unit cube, Omega^- = {x < X}, X = 0.55 (planar P1 level set), deformation Psi(x) = x + (a y^2, b z^2, c x^2)
  (quadratic, exactly representable at order 2):  det DPsi = 1 + 8abc xyz, so
    |Psi(Omega^-)|        = X + 8abc X^2/8                        (rule order >= 3 is exact)
    int_{Psi(Omega^-)} x  = X^2/2 + a X/3 + 8abc (X^3/12 + a X^2/16)   (rule order >= 5 is exact)
    |Psi(Gamma)|          = int_0^1 int_0^1 sqrt(1 + 4a^2 y^2 + 16 a^2 b^2 y^2 z^2) dy dz  (smooth: converges fast)
"""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from oracle import xoracle as xo
from util import det_of

NEG, POS, IF = 0, 1, 2
A, B, Cc, X = 0.3, -0.25, 0.2, 0.55


def setup(n):
    mesh = xf.MakeStructured3DMesh(nx=n, device="cpu")
    V = xf.H1(mesh, order=2)
    defo = xf.Deformation.interpolate_p2(V, lambda p: torch.stack([A * p[:, 1] ** 2, B * p[:, 2] ** 2, Cc * p[:, 0] ** 2], dim=1))
    phi = (mesh.points[:, 0] - X).numpy()
    return mesh, V, defo, phi


def integrate(mesh, defo, phi, dt, order, f=None):
    pts, els = mesh.points.numpy(), mesh.elements.numpy()
    loc = defo.local().numpy()
    tot = 0.0
    for el in range(els.shape[0]):
        T = xo.make_trafo(pts[els[el]], els[el], def_order=2, deform=loc[el])
        p, w, _ = xo.straight_cut_rule(T, phi[els[el]], dt, order)
        for q in range(len(w)):
            d, x = det_of(T, xo, p[q])
            tot += w[q] * d * (1.0 if f is None else f(*x))
    return tot


def test_deformed_volume_and_first_moment_are_exact():
    mesh, V, defo, phi = setup(3)
    abc8 = 8 * A * B * Cc
    vol = integrate(mesh, defo, phi, NEG, 3)
    assert abs(vol - (X + abc8 * X ** 2 / 8)) < 1e-13
    volp = integrate(mesh, defo, phi, POS, 3)
    assert abs(vol + volp - (1 + abc8 / 8)) < 1e-13
    mx = integrate(mesh, defo, phi, NEG, 5, lambda x, y, z: x)
    assert abs(mx - (X ** 2 / 2 + A * X / 3 + abc8 * (X ** 3 / 12 + A * X ** 2 / 16))) < 1e-13


def test_deformed_interface_area_converges_to_the_closed_form():
    g, wg = np.polynomial.legendre.leggauss(40)
    y = 0.5 * (g + 1); w = 0.5 * wg
    Y, Z = np.meshgrid(y, y, indexing="ij")
    exact = float((np.sqrt(1 + 4 * A ** 2 * Y ** 2 + 16 * A ** 2 * B ** 2 * Y ** 2 * Z ** 2) * np.outer(w, w)).sum())
    mesh, V, defo, phi = setup(3)
    a4 = integrate(mesh, defo, phi, IF, 4)
    a8 = integrate(mesh, defo, phi, IF, 8)
    assert abs(a8 - exact) < 1e-9
    assert abs(a4 - exact) < 1e-5 and abs(a8 - exact) < abs(a4 - exact)


def test_zero_deformation_is_the_affine_rule():
    mesh, V, defo, phi = setup(2)
    pts, els = mesh.points.numpy(), mesh.elements.numpy()
    zero = np.zeros((V.nd, 3))
    for el in range(0, els.shape[0], 5):
        T0 = xo.make_trafo(pts[els[el]], els[el])
        T1 = xo.make_trafo(pts[els[el]], els[el], def_order=2, deform=zero)
        p0, w0, _ = xo.straight_cut_rule(T0, phi[els[el]], IF, 4)
        p1, w1, _ = xo.straight_cut_rule(T1, phi[els[el]], IF, 4)
        assert np.array_equal(p0, p1) and np.array_equal(w0, w1)
