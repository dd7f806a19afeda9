"""Pin the oracle against the analytic known-answer values of the reference's own tests
(SURVEY.md §4 / §8c).  Each test cites the reference test it restates; meshes are our own
structured ones (the values are mesh independent)."""
from math import pi, sqrt

import numpy as np
import pytest

from ngsxfem_b200 import meshes
from oracle import xoracle as xo
from util import interpolate_to_p1, mesh_np, integrate_rule_fn, unpack_bits

NEG, POS, IF = xo.NEG, xo.POS, xo.IF


def integ(m, phi, dt, order, expo=None):
    pts, els = mesh_np(m)
    if expo is None:
        expo = [0] * m.dim
    return xo.integrate_monomial(m.dim, els, pts, phi, dt, order, expo)


# tests/pytests/test_localgeom.py:24-45
def test_cut_triangle_orders_0_to_15():
    import torch
    P = np.array([[0, 0], [1, 0], [0, 1], [0.5, 0], [0.5, 0.5], [0, 0.5]], dtype=float)
    E = np.array([[0, 3, 5], [3, 1, 4], [5, 4, 2], [3, 4, 5]], dtype=np.int32)
    phi = interpolate_to_p1(lambda x, y: x + y - 0.25, P)
    for order in range(16):
        mn = xo.integrate_monomial(2, E, P, phi, NEG, order, [0, 0])
        mp = xo.integrate_monomial(2, E, P, phi, POS, order, [0, 0])
        tol = 5e-16 * (order + 1) ** 2
        assert abs(mn - 1 / 32) < tol
        assert abs(mp - 0.5 + 1 / 32) < tol
        assert abs(mn + mp - 0.5) < tol


# tests/pytests/test_straightcutrule.py:11-31
@pytest.mark.parametrize("domain", [NEG, POS, IF])
@pytest.mark.parametrize("alpha", [2, 4, 8])
def test_polynomial_segm(domain, alpha):
    m = meshes.MakeStructured1DMesh(1)
    x_ast = 0.78522
    pts, els = mesh_np(m)
    phi = x_ast - pts[:, 0]
    ref = {POS: x_ast ** (alpha + 1) / (alpha + 1), NEG: (1 - x_ast ** (alpha + 1)) / (alpha + 1), IF: x_ast ** alpha}
    val = integ(m, phi, domain, alpha, [alpha])
    assert abs(val - ref[domain]) < 5e-15 * (alpha + 1) ** 2


# tests/pytests/test_straightcutrule.py:106-126
@pytest.mark.parametrize("order", [2, 4, 8])
@pytest.mark.parametrize("domain", [NEG, POS, IF])
@pytest.mark.parametrize("N", [1, 10, 30])
def test_straight_cut_square(order, domain, N):
    m = meshes.MakeStructured2DMesh(nx=N, ny=N)
    phi = interpolate_to_p1(lambda x, y: 1 - 2 * x - 2 * y, mesh_np(m)[0])
    ref = {NEG: 7 / 8, POS: 1 / 8, IF: 1 / sqrt(2)}
    assert abs(integ(m, phi, domain, order) - ref[domain]) < 5e-15 * (order + 1) ** 2


# tests/pytests/test_straightcutrule.py:128-154
@pytest.mark.parametrize("domain", [IF, NEG, POS])
@pytest.mark.parametrize("dim", [0, 1])
@pytest.mark.parametrize("eps", [1e-1, 1e-2, 5e-3, 1e-3, 0])
def test_orth_cut_square_eps(domain, dim, eps):
    order = 2
    m = meshes.MakeStructured2DMesh(nx=1, ny=1)
    if dim == 0:
        f = lambda x, y: 1 - 2 * x + eps * (y - 0.5)
    else:
        f = lambda x, y: 1 - 2 * y + eps * (x - 0.5)
    phi = interpolate_to_p1(f, mesh_np(m)[0])
    ref = {NEG: 0.5, POS: 0.5, IF: sqrt(1 + eps * eps / 4)}
    assert abs(integ(m, phi, domain, order) - ref[domain]) < 5e-15 * (order + 1) ** 2


# tests/pytests/test_straightcutrule.py:157-185
@pytest.mark.parametrize("order", [2, 4, 8])
@pytest.mark.parametrize("domain", [NEG, POS])
@pytest.mark.parametrize("alpha", [0, 1, 2])
@pytest.mark.parametrize("dim", [0, 1])
def test_straight_cut_square_polynomial(order, domain, alpha, dim):
    m = meshes.MakeStructured2DMesh(nx=1, ny=1)
    phi = interpolate_to_p1(lambda x, y: 1 - 2 * x - 2 * y, mesh_np(m)[0])
    val_pos = 2 ** (-alpha - 2) / (alpha * alpha + 3 * alpha + 2)
    ref = {POS: val_pos, NEG: 1. / (alpha + 1) - val_pos}
    expo = [0, 0]; expo[dim] = alpha
    assert abs(integ(m, phi, domain, order, expo) - ref[domain]) < 5e-15 * (order + 1) ** 2


# tests/pytests/test_straightcutrule.py:187-208
@pytest.mark.parametrize("order", [2, 4, 8])
@pytest.mark.parametrize("domain", [NEG, POS, IF])
@pytest.mark.parametrize("N", [1, 3])
def test_straight_cut_cube(order, domain, N):
    m = meshes.MakeStructured3DMesh(nx=N)
    phi = interpolate_to_p1(lambda x, y, z: 1 - 2 * x - 2 * y - 2 * z, mesh_np(m)[0])
    ref = {POS: 1. / 48, NEG: 47. / 48, IF: sqrt(3) / 8}
    assert abs(integ(m, phi, domain, order) - ref[domain]) < 5e-15 * (order + 1) ** 2


# tests/pytests/test_straightcutrule.py:210-235
@pytest.mark.parametrize("domain", [NEG, POS])
@pytest.mark.parametrize("alpha", [0, 1, 2])
@pytest.mark.parametrize("dim", [0, 1, 2])
def test_straight_cut_cube_polynomial(domain, alpha, dim):
    order = 4
    m = meshes.MakeStructured3DMesh(nx=1)
    phi = interpolate_to_p1(lambda x, y, z: 1 - 2 * x - 2 * y - 2 * z, mesh_np(m)[0])
    val_pos = 2 ** (-alpha - 3) / (alpha ** 3 + 6 * alpha * alpha + 11 * alpha + 6)
    ref = {POS: val_pos, NEG: 1. / (alpha + 1) - val_pos}
    expo = [0, 0, 0]; expo[dim] = alpha
    assert abs(integ(m, phi, domain, order, expo) - ref[domain]) < 5e-15 * (order + 1) ** 2


# tests/pytests/test_straightcutrule.py:237-280
@pytest.mark.parametrize("domain", [NEG, POS, IF])
@pytest.mark.parametrize("dim", [0, 1])
def test_orth_cut_square(domain, dim):
    m = meshes.MakeStructured2DMesh(nx=1, ny=1)
    phi = interpolate_to_p1(lambda x, y: 1 - 3 * (x if dim == 0 else y), mesh_np(m)[0])
    ref = {NEG: 2. / 3, POS: 1. / 3, IF: 1.}
    assert abs(integ(m, phi, domain, 2) - ref[domain]) < 5e-15 * 9


@pytest.mark.parametrize("domain", [NEG, POS, IF])
@pytest.mark.parametrize("dim", [0, 1, 2])
def test_orth_cut_cube(domain, dim):
    m = meshes.MakeStructured3DMesh(nx=1)
    phi = interpolate_to_p1(lambda x, y, z: 1 - 2 * (x, y, z)[dim], mesh_np(m)[0])
    ref = {POS: 0.5, NEG: 0.5, IF: 1.}
    assert abs(integ(m, phi, domain, 2) - ref[domain]) < 5e-15 * 9


# tests/pytests/test_straightcutrule.py:33-104 (circle r=0.6, EOC of the P1 geometry > 1.75)
@pytest.mark.parametrize("domain", [NEG, POS, IF])
def test_circle_convergence(domain):
    r = 0.6
    ref = {POS: 1 - pi * r * r / 4, NEG: pi * r * r / 4, IF: r * pi / 2}
    errs = []
    for i in range(1, 8):
        m = meshes.MakeStructured2DMesh(nx=2 ** i, ny=2 ** i)
        phi = interpolate_to_p1(lambda x, y: np.sqrt(x * x + y * y) - r, mesh_np(m)[0])
        errs.append(abs(integ(m, phi, domain, 2) - ref[domain]))
    eoc = [np.log(errs[i - 1] / errs[i]) / np.log(2) for i in range(1, len(errs))]
    assert np.mean(eoc[1:]) > 1.75


# tests/pytests/test_backuprule.py / test_straightcutrule.py sphere convergence
@pytest.mark.parametrize("domain", [NEG, IF])
def test_sphere_convergence(domain):
    r = 0.7234436998
    ref = {NEG: 4 / 3 * pi * r ** 3 / 8, IF: 4 * pi * r * r / 8}
    errs = []
    for i in range(1, 6):
        m = meshes.MakeStructured3DMesh(nx=2 ** i)
        phi = interpolate_to_p1(lambda x, y, z: np.sqrt(x * x + y * y + z * z) - r, mesh_np(m)[0])
        errs.append(abs(integ(m, phi, domain, 2) - ref[domain]))
    eoc = [np.log(errs[i - 1] / errs[i]) / np.log(2) for i in range(1, len(errs))]
    assert np.mean(eoc[1:]) > 1.75


# ---------------------------------------------------------------- multi level set
def expand_any(dtl):
    out = []
    for t in dtl:
        locs = [i for i, d in enumerate(t) if d == "ANY"]
        if not locs:
            out.append(tuple(t)); continue
        import itertools
        for comb in itertools.product((POS, NEG), repeat=len(locs)):
            l = list(t)
            for k, c in zip(locs, comb):
                l[k] = c
            out.append(tuple(l))
    return sorted(set(out))


def ml_integrate(m, phis, dts, order=0, f=None):
    pts, els = mesh_np(m)
    dts = expand_any(dts if isinstance(dts, list) else [dts])

    def rule(T, el):
        ls = np.stack([p[els[el]] for p in phis], axis=1)
        return xo.mlset_cut_rule(T, ls, dts, order)
    return integrate_rule_fn(xo, pts, els, rule, f)


# tests/pytests/test_mlset_integration.py:37-138
def test_mlset_2d_triangle():
    m = meshes.MakeStructured2DMesh(nx=7, ny=9, mapping=lambda x, y: (2 * x - 1, 2 * y - 0.5))
    pts, els = mesh_np(m)
    fns = [lambda x, y: y - 1, lambda x, y: 2 * x - y, lambda x, y: -2 * x - y]
    phis = [interpolate_to_p1(f, pts) for f in fns]
    tol = 1e-12
    assert abs(xo.integrate_monomial(2, els, pts, phis[0], NEG, 0, [0, 0]) - 3) < tol
    tri = (NEG, NEG, NEG)
    assert abs(ml_integrate(m, phis, tri) - 0.5) < tol
    import itertools
    inv = [t for t in itertools.product((NEG, POS), repeat=3) if t != tri]
    assert abs(ml_integrate(m, phis, inv) - 3.5) < tol
    assert abs(ml_integrate(m, phis, (IF, NEG, NEG)) - 1) < tol
    assert abs(ml_integrate(m, phis, (NEG, IF, NEG)) - sqrt(5 / 4)) < tol
    assert abs(ml_integrate(m, phis, (NEG, NEG, IF)) - sqrt(5 / 4)) < tol
    bnd = [(IF, NEG, NEG), (NEG, IF, NEG), (NEG, NEG, IF)]
    assert abs(ml_integrate(m, phis, bnd) - 1 - sqrt(5)) < tol
    xy = lambda x, y: x + y
    assert abs(ml_integrate(m, phis, (IF, IF, NEG), f=xy) - 1.5) < tol
    assert abs(ml_integrate(m, phis, (IF, NEG, IF), f=xy) - 0.5) < tol
    assert abs(ml_integrate(m, phis, (NEG, IF, IF), f=xy)) < tol
    assert abs(ml_integrate(m, phis, [(IF, IF, NEG), (IF, NEG, IF), (NEG, IF, IF)], f=xy) - 2.0) < tol
    # markers (:60-82): ElementsOfType vs ElementsWithContribution
    b_neg = unpack_bits(xo.ml_elements_of_type(2, els, np.stack(phis), [tri]), m.ne)
    b_has = unpack_bits(xo.ml_elements_with_contribution(2, els, pts, np.stack(phis), [tri]), m.ne)
    b_if = unpack_bits(xo.ml_elements_with_contribution(2, els, pts, np.stack(phis), bnd), m.ne)
    b_not = unpack_bits(xo.ml_elements_with_contribution(2, els, pts, np.stack(phis), inv), m.ne)
    assert np.array_equal(~b_neg, b_not)
    assert np.array_equal(b_if, b_has & ~b_neg)


# tests/pytests/test_mlset_integration.py:304-415
def test_mlset_3d_cube():
    m = meshes.MakeStructured3DMesh(nx=3, mapping=lambda x, y, z: (1.6 * x - 0.8, 1.6 * y - 0.8, 1.6 * z - 0.8))
    pts, els = mesh_np(m)
    fns = [lambda x, y, z: x - 0.5, lambda x, y, z: x + 0.5, lambda x, y, z: y - 0.5,
           lambda x, y, z: y + 0.5, lambda x, y, z: z - 0.5, lambda x, y, z: z + 0.5]
    phis = [interpolate_to_p1(f, pts) for f in fns]
    tol = 1e-12
    cube = (NEG, POS, NEG, POS, NEG, POS)
    assert abs(ml_integrate(m, phis, cube) - 1) < tol
    import itertools
    inv = [t for t in itertools.product((NEG, POS), repeat=6) if t != cube]
    assert abs(ml_integrate(m, phis, inv) - 3.096) < tol
    sides = []
    for i in range(6):
        d = list(cube); d[i] = IF
        sides.append(tuple(d))
        assert abs(ml_integrate(m, phis, sides[-1]) - 1) < tol
    assert abs(ml_integrate(m, phis, sides) - 6) < tol
    edges = [(IF, POS, IF, POS, NEG, POS), (IF, POS, NEG, IF, NEG, POS),
             (IF, POS, NEG, POS, IF, POS), (IF, POS, NEG, POS, NEG, IF),
             (NEG, IF, IF, POS, NEG, POS), (NEG, IF, NEG, IF, NEG, POS),
             (NEG, IF, NEG, POS, IF, POS), (NEG, IF, NEG, POS, NEG, IF),
             (NEG, POS, IF, POS, IF, POS), (NEG, POS, NEG, IF, IF, POS),
             (NEG, POS, IF, POS, NEG, IF), (NEG, POS, NEG, IF, NEG, IF)]
    for e in edges:
        assert abs(ml_integrate(m, phis, e) - 1) < tol
    assert abs(ml_integrate(m, phis, edges) - 12) < tol
    point_domains = {"ppp": (IF, POS, IF, POS, IF, POS), "ppn": (IF, POS, IF, POS, NEG, IF),
                     "pnp": (IF, POS, NEG, IF, IF, POS), "npp": (NEG, IF, IF, POS, IF, POS),
                     "pnn": (IF, POS, NEG, IF, NEG, IF), "npn": (NEG, IF, IF, POS, NEG, IF),
                     "nnp": (NEG, IF, NEG, IF, IF, POS), "nnn": (NEG, IF, NEG, IF, NEG, IF)}
    vals = {"ppp": 1.5, "ppn": 0.5, "pnp": 0.5, "npp": 0.5, "pnn": -0.5, "npn": -0.5, "nnp": -0.5, "nnn": -1.5}
    for key, dom in point_domains.items():
        assert abs(ml_integrate(m, phis, dom, f=lambda x, y, z: x + y + z) - vals[key]) < tol


# tests/pytests/test_mlset_integration.py:418-485
def test_mlset_3d_codim2_cross():
    m = meshes.MakeStructured3DMesh(nx=3, mapping=lambda x, y, z: (1.6 * x - 0.8, 1.6 * y - 0.8, 1.6 * z - 0.8))
    pts, els = mesh_np(m)
    mk = lambda fns: [interpolate_to_p1(f, pts) for f in fns]
    phis = mk([lambda x, y, z: x - 0.5, lambda x, y, z: x + 0.5, lambda x, y, z: x - y, lambda x, y, z: z - 0])
    assert abs(ml_integrate(m, phis, (NEG, POS, IF, IF)) - sqrt(2)) < 1e-12
    phis = mk([lambda x, y, z: x - 0.5, lambda x, y, z: -x - 0.5, lambda x, y, z: z - y,
               lambda x, y, z: x - z, lambda x, y, z: x + z])
    cross = [(NEG, NEG, IF, IF, "ANY"), (NEG, NEG, IF, "ANY", IF)]
    assert abs(ml_integrate(m, phis, cross) - 2 * sqrt(3)) < 1e-12
    phis = mk([lambda x, y, z: x - 0.5, lambda x, y, z: -x - 0.5, lambda x, y, z: -z - 0.5,
               lambda x, y, z: z + y, lambda x, y, z: x + y, lambda x, y, z: z + y + 0.1])
    lines = [(NEG, NEG, NEG, IF, IF, "ANY"), (NEG, NEG, NEG, "ANY", IF, IF)]
    assert abs(ml_integrate(m, phis, lines) - 1.9 * sqrt(3)) < 1e-12


# ---------------------------------------------------------------- space-time
def st_integrate(m, lset_nodes, time_fe_order, dt, order_time, order_space, f=None):
    """sum w * |det F| * f(x, t); lset_nodes: list (per time node) of nodal arrays"""
    pts, els = mesh_np(m)
    total = 0.0
    from util import det_of
    for el in range(els.shape[0]):
        T = xo.make_trafo(pts[els[el]], els[el])
        lst = np.stack([l[els[el]] for l in lset_nodes], axis=0)
        p, tv, w = xo.spacetime_cut_rule(T, lst, time_fe_order, dt, order_time, order_space)
        for q in range(len(w)):
            d, x = det_of(T, xo, p[q])
            total += w[q] * d * (1.0 if f is None else f(*x, tv[q]))
    return total


# tests/pytests/test_spacetimecutrule.py:19-69  (levelset 1-2x-2t on a 1D mesh: POS 1/8, NEG 7/8, IF 1/2)
@pytest.mark.parametrize("domain,ref", [(POS, 1 / 8), (NEG, 7 / 8), (IF, 1 / 2)])
@pytest.mark.parametrize("N", [1, 4])
def test_spacetime_1d(domain, ref, N):
    m = meshes.MakeStructured1DMesh(N)
    pts, _ = mesh_np(m)
    lsn = [interpolate_to_p1(lambda x: 1 - 2 * x - 2 * t, pts) for t in (0.0, 1.0)]
    val = st_integrate(m, lsn, 1, domain, 2, 2)
    assert abs(val - ref) < 5e-15 * 9


# moving circle (tests/pytests/test_spacetimecutrule.py:583-596 style): int_0^1 area(t) dt, P2 in time
def test_spacetime_moving_circle_area():
    r0 = 0.3
    errs = []
    for n in (8, 16, 32):
        m = meshes.MakeStructured2DMesh(nx=n, ny=n)
        pts, _ = mesh_np(m)
        lsn = [interpolate_to_p1(lambda x, y: np.sqrt((x - 0.4 - 0.2 * t) ** 2 + (y - 0.5) ** 2) - r0, pts)
               for t in (0.0, 0.5, 1.0)]
        val = st_integrate(m, lsn, 2, NEG, 3, 2)
        errs.append(abs(val - pi * r0 * r0))
    eoc = [np.log(errs[i - 1] / errs[i]) / np.log(2) for i in range(1, len(errs))]
    assert errs[-1] < 2e-3 and np.mean(eoc) > 1.5


# NodalTimeFE nodes of order >= 4 are Gauss-Lobatto-Legendre points (spacetime/SpaceTimeFE.cpp:258-282, :229-256), not
# Chebyshev points.  Level set x - t^k given by its values on THOSE nodes is exactly x - t^k, so NEG = int t^k = 1/(k+1);
# with other nodes the interpolant (and the measure) would differ.
def lobatto_nodes(k):
    if k == 4:
        return [0.0, 0.5 * (1 - np.sqrt(3 / 7)), 0.5, 0.5 * (1 + np.sqrt(3 / 7)), 1.0]
    # interior nodes: roots of P_k' mapped to [0, 1]
    c = np.zeros(k + 1); c[k] = 1
    r = np.sort(np.polynomial.legendre.legroots(np.polynomial.legendre.legder(c)))
    return [0.0] + list(0.5 * (1 + r)) + [1.0]


@pytest.mark.parametrize("k", [4, 5, 6, 8])
def test_spacetime_high_order_time_nodes(k):
    m = meshes.MakeStructured1DMesh(3)
    pts, _ = mesh_np(m)
    nodes = lobatto_nodes(k)
    lsn = [interpolate_to_p1(lambda x: x - t ** k, pts) for t in nodes]
    neg = st_integrate(m, lsn, k, NEG, k, 1)
    pos = st_integrate(m, lsn, k, POS, k, 1)
    assert abs(neg - 1 / (k + 1)) < 1e-12 and abs(pos - k / (k + 1)) < 1e-12


def test_graph_timing_variant_equals_the_two_call_graph():
    """xo_graph_once (one sort per row, thread-local column buffers; what bench.py's CPU arm times) == xo_graph, for every
    thread count"""
    rng = np.random.default_rng(0)
    nt0 = xo.num_threads()
    try:
        for ndof, nr, w in [(500, 700, 10), (50, 0, 4), (64, 40, 3), (3000, 9000, 10)]:
            t = rng.integers(0, ndof, size=(nr, w)).astype(np.int32)
            if nr:
                t[rng.random(t.shape) < 0.05] = -1
            a = xo.graph(ndof, t)
            for nt in (1, 3, 8):
                xo.set_num_threads(nt)
                b = xo.graph(ndof, t, once=True)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    finally:
        xo.set_num_threads(nt0)
