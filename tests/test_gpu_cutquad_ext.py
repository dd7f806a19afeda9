"""GPU parity of the multi-level-set and space-time cut rules (xgpu_cutquad_ml_*, xgpu_cutquad_st_*) against the
oracle on the configurations of the reference's own tests (tests/pytests/test_mlset_integration.py,
test_spacetimecutrule.py): identical point counts per element, points and weights to 1e-13, plus the analytic
known answers through the GPU integrator.

Weights are compared with rtol 1e-13 plus an absolute floor of 1e-13 x the largest weight of the batch: level sets
through mesh vertices are perturbed by 1e-14 (InterpolateToP1), which creates sliver sub-simplices whose volume is pure
cancellation noise (the CUDA code contracts to FMA, the oracle is compiled with -ffp-contract=off)."""
import itertools
from math import sqrt

import numpy as np
import pytest
import torch

from ngsxfem_b200 import meshes, ops
from oracle import xoracle as xo
from util import interpolate_to_p1, mesh_np

pytestmark = pytest.mark.gpu
NEG, POS, IF = 0, 1, 2
DEV = "cuda"


def oracle_ml(pts, els, phis, dts, order):
    offs, P, W = [0], [], []
    for el in range(els.shape[0]):
        T = xo.make_trafo(pts[els[el]], els[el])
        ls = np.stack([p[els[el]] for p in phis], axis=1)
        p, w = xo.mlset_cut_rule(T, ls, dts, order)
        P.append(p); W.append(w); offs.append(offs[-1] + len(w))
    return np.array(offs), np.concatenate(P), np.concatenate(W)


def gpu_ml(m, phis, dts, order):
    md = m.to(DEV)
    gphis = torch.from_numpy(np.stack(phis)).to(DEV).contiguous()
    off, p, w = ops.cutquad_ml(m.dim, None, md.elements, md.points, gphis, dts, order)
    assert ops.default_ctx().status()[0] == 0
    return md, off, p, w


def check_ml(m, phis, dts, order):
    dts = [tuple(d) for d in (dts if isinstance(dts[0], (tuple, list)) else [dts])]
    pts, els = mesh_np(m)
    md, off, p, w = gpu_ml(m, phis, dts, order)
    oo, po, wo = oracle_ml(pts, els, phis, dts, order)
    assert np.array_equal(off.cpu().numpy(), oo)
    if len(wo):
        assert np.allclose(p.cpu().numpy(), po, rtol=1e-13, atol=1e-15)
        assert np.allclose(w.cpu().numpy(), wo, rtol=1e-13, atol=1e-13 * np.abs(wo).max())
    return md, off, p, w


def gpu_integral(m, md, off, p, w, expo):
    return ops.integrate_monomial(m.dim, None, md.elements, md.points, off, p, w, expo)


def test_ml_2d_triangle_matches_oracle_and_kats():
    # tests/pytests/test_mlset_integration.py:37-138
    m = meshes.MakeStructured2DMesh(nx=7, ny=9, mapping=lambda x, y: (2 * x - 1, 2 * y - 0.5))
    pts, _ = mesh_np(m)
    fns = [lambda x, y: y - 1, lambda x, y: 2 * x - y, lambda x, y: -2 * x - y]
    phis = [interpolate_to_p1(f, pts) for f in fns]
    tri = (NEG, NEG, NEG)
    for order in (0, 2, 5):
        md, off, p, w = check_ml(m, phis, tri, order)
        assert abs(gpu_integral(m, md, off, p, w, (0, 0)) - 0.5) < 1e-12
    inv = [t for t in itertools.product((NEG, POS), repeat=3) if t != tri]
    md, off, p, w = check_ml(m, phis, inv, 1)
    assert abs(gpu_integral(m, md, off, p, w, (0, 0)) - 3.5) < 1e-12
    for dom, ref in (((IF, NEG, NEG), 1.0), ((NEG, IF, NEG), sqrt(5 / 4)), ((NEG, NEG, IF), sqrt(5 / 4))):
        md, off, p, w = check_ml(m, phis, dom, 3)
        assert abs(gpu_integral(m, md, off, p, w, (0, 0)) - ref) < 1e-12
    # codim 2: point evaluations of x + y
    for dom, ref in (((IF, IF, NEG), 1.5), ((IF, NEG, IF), 0.5), ((NEG, IF, IF), 0.0)):
        md, off, p, w = check_ml(m, phis, dom, 2)
        val = gpu_integral(m, md, off, p, w, (1, 0)) + gpu_integral(m, md, off, p, w, (0, 1))
        assert abs(val - ref) < 1e-12


def test_ml_3d_cube_matches_oracle_and_kats():
    # tests/pytests/test_mlset_integration.py:304-415
    m = meshes.MakeStructured3DMesh(nx=3, mapping=lambda x, y, z: (1.6 * x - 0.8, 1.6 * y - 0.8, 1.6 * z - 0.8))
    pts, _ = mesh_np(m)
    fns = [lambda x, y, z: x - 0.5, lambda x, y, z: x + 0.5, lambda x, y, z: y - 0.5,
           lambda x, y, z: y + 0.5, lambda x, y, z: z - 0.5, lambda x, y, z: z + 0.5]
    phis = [interpolate_to_p1(f, pts) for f in fns]
    cube = (NEG, POS, NEG, POS, NEG, POS)
    md, off, p, w = check_ml(m, phis, cube, 2)
    assert abs(gpu_integral(m, md, off, p, w, (0, 0, 0)) - 1) < 1e-12
    side = (IF, POS, NEG, POS, NEG, POS)
    md, off, p, w = check_ml(m, phis, side, 2)
    assert abs(gpu_integral(m, md, off, p, w, (0, 0, 0)) - 1) < 1e-12
    edge = (IF, POS, IF, POS, NEG, POS)
    md, off, p, w = check_ml(m, phis, edge, 2)
    assert abs(gpu_integral(m, md, off, p, w, (0, 0, 0)) - 1) < 1e-12
    corner = (IF, POS, IF, POS, IF, POS)
    md, off, p, w = check_ml(m, phis, corner, 1)
    val = sum(gpu_integral(m, md, off, p, w, e) for e in ((1, 0, 0), (0, 1, 0), (0, 0, 1)))
    assert abs(val - 1.5) < 1e-12
    # several tuples concatenated
    sides = []
    for i in range(6):
        d = list(cube); d[i] = IF
        sides.append(tuple(d))
    md, off, p, w = check_ml(m, phis, sides, 1)
    assert abs(gpu_integral(m, md, off, p, w, (0, 0, 0)) - 6) < 1e-12


def test_ml_sphere_and_planes_on_element_list():
    # config-4 style: cube cut by 3 planes and a sphere, domain (NEG,NEG,NEG,NEG), element list input
    m = meshes.MakeStructured3DMesh(nx=5)
    pts, els = mesh_np(m)
    fns = [lambda x, y, z: x + 0.3 * y - 0.71, lambda x, y, z: 0.2 * x - y + 0.1 * z + 0.23,
           lambda x, y, z: z - 0.77 + 0.1 * x, lambda x, y, z: np.sqrt((x - .5) ** 2 + (y - .5) ** 2 + (z - .5) ** 2) - 0.41]
    phis = [interpolate_to_p1(f, pts) for f in fns]
    dts = [(NEG, NEG, NEG, NEG)]
    md = m.to(DEV)
    gphis = torch.from_numpy(np.stack(phis)).to(DEV).contiguous()
    lst = torch.arange(0, m.ne, 3, dtype=torch.int32, device=DEV)
    off, p, w = ops.cutquad_ml(3, lst, md.elements, md.points, gphis, dts, 4)
    oo, po, wo = oracle_ml(pts, els[::3], phis, dts, 4)
    assert np.array_equal(off.cpu().numpy(), oo)
    assert np.allclose(p.cpu().numpy(), po, rtol=1e-13, atol=1e-15)
    assert np.allclose(w.cpu().numpy(), wo, rtol=1e-13, atol=1e-13 * np.abs(wo).max())
    # boundary of the domain piece on the sphere, and an edge (codim 2)
    for dom in ((NEG, NEG, NEG, IF), (IF, NEG, NEG, IF), (IF, IF, NEG, NEG)):
        check_ml(m, phis, dom, 2)


def oracle_st(pts, els, phi_st, tord, dt, ot, os_):
    offs, P, TV, W = [0], [], [], []
    for el in range(els.shape[0]):
        T = xo.make_trafo(pts[els[el]], els[el])
        p, tv, w = xo.spacetime_cut_rule(T, phi_st[:, els[el]], tord, dt, ot, os_)
        P.append(p); TV.append(tv); W.append(w); offs.append(offs[-1] + len(w))
    return np.array(offs), np.concatenate(P), np.concatenate(TV), np.concatenate(W)


def time_nodes(order):
    if order >= 4:   # Gauss-Lobatto-Legendre points (spacetime/SpaceTimeFE.cpp:258-282): 0, 1 and the roots of P_k'
        c = np.zeros(order + 1); c[order] = 1
        r = np.sort(np.polynomial.legendre.legroots(np.polynomial.legendre.legder(c)))
        return [0.0] + list(0.5 * (1 + r)) + [1.0]
    return {0: [0.0], 1: [0.0, 1.0], 2: [0.0, 0.5, 1.0],
            3: [0.0, 0.5 * (1 - 1 / sqrt(5)), 0.5 * (1 + 1 / sqrt(5)), 1.0]}[order]


@pytest.mark.parametrize("k", [4, 5, 6, 8])
def test_spacetime_high_order_time_nodes_kat(k):
    # level set x - t^k through its values on the reference's time nodes: NEG = int_0^1 t^k dt = 1/(k+1); a kernel that
    # interpolated on other nodes (e.g. Chebyshev points) would integrate a different function
    m = meshes.MakeStructured1DMesh(3)
    pts, els = mesh_np(m)
    phi_st = np.stack([pts[:, 0] - t ** k for t in time_nodes(k)])
    md = m.to(DEV)
    g = torch.from_numpy(phi_st).to(DEV).contiguous()
    for dt, ref in ((NEG, 1 / (k + 1)), (POS, k / (k + 1))):
        off, p, tv, w = ops.cutquad_st(1, None, md.elements, md.points, g, k, dt, k, 1)
        assert ops.default_ctx().status()[0] == 0
        val = ops.integrate_monomial(1, None, md.elements, md.points, off, p, w, (0,))
        assert abs(val - ref) < 1e-12
        oo, po, tvo, wo = oracle_st(pts, els, phi_st, k, dt, k, 1)
        assert np.array_equal(off.cpu().numpy(), oo)
        assert np.allclose(w.cpu().numpy(), wo, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("dim,n", [(1, 9), (2, 7), (3, 4)])
@pytest.mark.parametrize("tord", [1, 2, 3])
@pytest.mark.parametrize("dt", [NEG, POS, IF])
def test_spacetime_rule_matches_oracle(dim, n, tord, dt):
    m = meshes.MakeStructured1DMesh(n) if dim == 1 else (meshes.MakeStructured2DMesh(nx=n, ny=n) if dim == 2
                                                         else meshes.MakeStructured3DMesh(nx=n))
    pts, els = mesh_np(m)
    # moving sphere / point: centre 0.5 + 0.1 sin(2 pi t) in the first coordinate, radius 0.3
    def lset(t):
        c = np.full(dim, 0.5); c[0] += 0.1 * np.sin(2 * np.pi * t)
        return interpolate_to_p1(lambda *x: np.sqrt(sum((xi - ci) ** 2 for xi, ci in zip(x, c))) - 0.3, pts)
    phi_st = np.stack([lset(t) for t in time_nodes(tord)])
    md = m.to(DEV)
    g = torch.from_numpy(phi_st).to(DEV).contiguous()
    off, p, tv, w = ops.cutquad_st(dim, None, md.elements, md.points, g, tord, dt, 3, 2)
    assert ops.default_ctx().status()[0] == 0
    oo, po, tvo, wo = oracle_st(pts, els, phi_st, tord, dt, 3, 2)
    assert np.array_equal(off.cpu().numpy(), oo)
    assert len(wo) > 0
    assert np.allclose(p.cpu().numpy(), po, rtol=1e-13, atol=1e-15)
    assert np.allclose(tv.cpu().numpy(), tvo, rtol=1e-13, atol=1e-16)
    assert np.allclose(w.cpu().numpy(), wo, rtol=1e-13, atol=1e-13 * np.abs(wo).max())


@pytest.mark.parametrize("dt,ref", [(POS, 1 / 8), (NEG, 7 / 8), (IF, 1 / 2)])
def test_spacetime_kat_1d(dt, ref):
    # tests/pytests/test_spacetimecutrule.py:41-69: lset 1 - 2x - 2t on the unit interval x [0,1]
    m = meshes.MakeStructured1DMesh(4)
    pts, els = mesh_np(m)
    phi_st = np.stack([1 - 2 * pts[:, 0] - 2 * t for t in (0.0, 1.0)])
    md = m.to(DEV)
    off, p, tv, w = ops.cutquad_st(1, None, md.elements, md.points, torch.from_numpy(phi_st).to(DEV).contiguous(), 1, dt, 2, 2)
    val = ops.integrate_monomial(1, None, md.elements, md.points, off, p, w, (0,))
    assert abs(val - ref) < 5e-15 * 9
