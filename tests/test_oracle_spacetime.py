"""Pins of the oracle's space-time FE restatement (oracle/xo_fe.c: st_basis, xo_elmat_st_from_rule, xo_elvec_st_from_rule)
— spacetime/SpaceTimeFE.cpp:57-135, :258-282 — against closed forms: the Lagrange property on NodalTimeFE's nodes, the
time derivative of polynomials, and the reference's own space-time known answers (tests/pytests/
test_spacetimecutrule.py:14-69: int tref = 1/2, int tref^3 = 1/4, int tref^2 (x^2+y^2) = 2/9 on the unit square;
level set 1-2x-2t: NEG 7/8, POS 1/8, IF 1/2) computed through element MATRICES and VECTORS."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import xoracle as xo

NEG, POS, IF = 0, 1, 2
ID, GRAD, DT = 0, 1, 3
NODES = {0: [0.0], 1: [0.0, 1.0], 2: [0.0, 0.5, 1.0],
         3: [0.0, 0.5 * (1 - 1 / np.sqrt(5)), 0.5 * (1 + 1 / np.sqrt(5)), 1.0]}


def time_shapes(k, t):
    sh = np.zeros(k + 1); dsh = np.zeros(k + 1)
    fn = xo.lib().xo_time_shapes; fn.restype = None
    fn(k, C.c_double(t), sh.ctypes.data_as(C.POINTER(C.c_double)))
    fn = xo.lib().xo_time_dshapes; fn.restype = None
    fn(k, C.c_double(t), dsh.ctypes.data_as(C.POINTER(C.c_double)))
    return sh, dsh


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_time_basis_is_nodal_and_differentiates_polynomials(k):
    for i, ti in enumerate(NODES[k]):
        sh, _ = time_shapes(k, ti)
        assert np.allclose(sh, np.eye(k + 1)[i], atol=1e-15)
    for t in (0.0, 0.3, 0.77, 1.0):
        sh, dsh = time_shapes(k, t)
        assert abs(sh.sum() - 1) < 1e-14 and abs(dsh.sum()) < 1e-13
        for m in range(1, k + 1):   # d/dt t^m = m t^(m-1), t^m is in the space
            c = np.array(NODES[k]) ** m
            assert abs(c @ sh - t ** m) < 1e-14
            assert abs(c @ dsh - m * t ** (m - 1)) < 1e-13


def unit_square():
    P = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=float)
    E = np.array([[0, 1, 3], [0, 3, 2]], dtype=np.int32)
    return P, E


def st_sum(P, E, lset_st, kl, dom, ot, os_, p, kt, terms=None, lterms=None):
    """sum over the two triangles of 1^T M 1 (terms) or 1^T f (lterms), 1 = constant one (vertex dofs, every time node)"""
    nds = xo.h1_ndof(2, p)
    one = np.zeros(nds * (kt + 1))
    for j in range(kt + 1):
        one[j * nds: j * nds + 3] = 1.0
    tot = 0.0
    for el in range(E.shape[0]):
        T = xo.make_trafo(P[E[el]], E[el])
        q, tv, w = xo.spacetime_cut_rule(T, lset_st[:, E[el]], kl, dom, ot, os_)
        if not len(w):
            continue
        if terms is not None:
            tot += one @ xo.elmat_st_from_rule(T, p, kt, q, tv, w, terms) @ one
        else:
            tot += one @ xo.elvec_st_from_rule(T, p, kt, q, tv, w, lterms)
    return tot


@pytest.mark.parametrize("p,kt", [(1, 1), (2, 2), (3, 3)])
def test_reference_known_answers_through_matrices_and_vectors(p, kt):
    P, E = unit_square()
    neg = -np.ones((2, 4))
    # test_spacetime_integrate_no_cut: (tref, 1/2), (tref^3, 1/4), (x, 1/2), (tref^2 (x^2+y^2), 2/9)
    assert abs(st_sum(P, E, neg, 1, NEG, 1, 0, p, kt, lterms=[(ID, 1.0, 0, (0, 0, 0, 1))]) - 0.5) < 5e-15
    assert abs(st_sum(P, E, neg, 1, NEG, 3, 0, p, kt, lterms=[(ID, 1.0, 0, (0, 0, 0, 3))]) - 0.25) < 5e-15
    assert abs(st_sum(P, E, neg, 1, NEG, 0, 1, p, kt, lterms=[(ID, 1.0, 0, (1, 0, 0, 0))]) - 0.5) < 5e-15
    got = st_sum(P, E, neg, 1, NEG, 2, 2, p, kt, lterms=[(ID, 1.0, 0, (2, 0, 0, 2)), (ID, 1.0, 0, (0, 2, 0, 2))])
    assert abs(got - 2 / 9) < 5e-15
    # test_spacetime_integrateX_via_straight_cutted_quad2Dplus1D: lset = 1 - 2x - 2t
    ls = np.stack([1 - 2 * P[:, 0] - 2 * t for t in (0.0, 1.0)])
    ls = np.where(np.abs(ls) < 1e-14, 1e-14, ls)
    for dom, ref in ((POS, 1 / 8), (NEG, 7 / 8), (IF, 0.5)):
        assert abs(st_sum(P, E, ls, 1, dom, 2 * kt, 2 * p, p, kt, terms=[(ID, ID, 1.0, 0)]) - ref) < 5e-15
    # dt(u) v with u = 1: zero row sums; grad.grad annihilates constants
    assert abs(st_sum(P, E, ls, 1, NEG, 2 * kt, 2 * p, p, kt, terms=[(DT, ID, 1.0, 0)])) < 1e-14
    assert abs(st_sum(P, E, ls, 1, NEG, 2 * kt, 2 * p, p, kt, terms=[(GRAD, GRAD, 1.0, 0)])) < 1e-14


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_uncut_p1_element_matrices_are_the_closed_form_tensor_products(dim):
    """layout and operators of the oracle's space-time basis against closed forms on one uncut affine simplex:
    mass = M_time (x) M_space, dt-form = D_time (x) M_space, stiffness = M_time (x) K_space with
    M_space = |T| (1 + delta_ik) / ((D+1)(D+2)),  K_space = |T| grad(lam_i).grad(lam_k),
    P1 time: M_time = [[1/3, 1/6], [1/6, 1/3]], D_time[j][l] = int L_j L_l' = [[-1/2, 1/2], [-1/2, 1/2]]  (rows = test);
    P2 time (nodes 0, 1/2, 1): M_time = [[4, 2, -1], [2, 16, 2], [-1, 2, 4]] / 30  (time-major dofs: j * nd_s + i)"""
    rng = np.random.default_rng(5)
    V = np.eye(dim + 1, dim) + 0.2 * rng.standard_normal((dim + 1, dim))      # a generic simplex
    vnum = np.arange(dim + 1, dtype=np.int32)
    T = xo.make_trafo(V, vnum)
    E = (V[:dim] - V[dim]).T                                                    # F: columns v_k - v_D (v_D <-> origin)
    vol = abs(np.linalg.det(E)) / math.factorial(dim)
    Ms = vol * (np.ones((dim + 1, dim + 1)) + np.eye(dim + 1)) / ((dim + 1) * (dim + 2))
    G = np.vstack([np.linalg.inv(E), -np.linalg.inv(E).sum(axis=0)])            # rows: physical gradients of lam_k
    Ks = vol * G @ G.T
    neg = -np.ones((2, dim + 1))
    for kt, Mt, Dt in ((1, np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]]), np.array([[-.5, .5], [-.5, .5]])),
                       (2, np.array([[4, 2, -1], [2, 16, 2], [-1, 2, 4]]) / 30.0, None)):
        q, tv, w = xo.spacetime_cut_rule(T, neg, 1, NEG, 2 * kt, 2)
        m = xo.elmat_st_from_rule(T, 1, kt, q, tv, w, [(ID, ID, 1.0, 0)])
        assert np.allclose(m, np.kron(Mt, Ms), rtol=1e-13, atol=1e-16)
        k = xo.elmat_st_from_rule(T, 1, kt, q, tv, w, [(GRAD, GRAD, 1.0, 0)])
        assert np.allclose(k, np.kron(Mt, Ks), rtol=1e-12, atol=1e-15)
        if Dt is not None:
            d = xo.elmat_st_from_rule(T, 1, kt, q, tv, w, [(DT, ID, 1.0, 0)])   # (op_trial, op_test): dt(u) v
            assert np.allclose(d, np.kron(Dt, Ms), rtol=1e-13, atol=1e-16)
            dT = xo.elmat_st_from_rule(T, 1, kt, q, tv, w, [(ID, DT, 1.0, 0)])  # u dt(v): the transpose
            assert np.allclose(dT, d.T, rtol=1e-13, atol=1e-16)


def test_oracle_reproduces_the_committed_golden_vectors():
    """tests/golden/spacetime_elmat.json (tools/gen_golden_spacetime.py): rules, element matrices and vectors of three cut
    space-time elements — freezes the restatement against drift"""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import gen_golden_spacetime as gg
    gold = json.load(open(os.path.join(root, "tests", "golden", "spacetime_elmat.json")))
    assert [tuple(t) for t in gold["terms"]] == gg.TERMS
    for c in gold["cases"]:
        q, tv, w, m, v = gg.compute(c)
        assert len(w) == len(c["wts"]) > 0
        assert np.allclose(q, np.array(c["pts"]), rtol=0, atol=1e-15)
        assert np.allclose(tv, np.array(c["tvals"]), rtol=0, atol=1e-15)
        assert np.allclose(w, np.array(c["wts"]), rtol=1e-14, atol=1e-18)
        assert np.allclose(m, np.array(c["elmat"]), rtol=1e-14, atol=1e-16)
        assert np.allclose(v, np.array(c["elvec"]), rtol=1e-14, atol=1e-16)


@pytest.mark.parametrize("kt", [1, 2])
def test_moving_plane_in_3d_known_answers(kt):
    """3D analogue of the reference's moving-plane test: lset = x + t - 1 on the unit cube (6 Kuhn tets): the NEG domain
    {x < 1 - t} has space-time volume 1/2, POS 1/2, the interface (a unit square for every t) space-time measure 1;
    first moments int_NEG x = 1/6, int_NEG tref = 1/6 — through the mass matrix / element vectors of a P1 x P_kt space"""
    import itertools
    P = np.array(list(itertools.product([0.0, 1.0], repeat=3)))[:, ::-1].copy()      # vertex id = i + 2 j + 4 k
    vid = lambda c: int(c[0] + 2 * c[1] + 4 * c[2])
    els = []
    for perm in itertools.permutations(range(3)):
        c = [0, 0, 0]
        path = [vid(c)]
        for ax in perm:
            c[ax] = 1
            path.append(vid(c))
        els.append(path)
    E = np.array(els, dtype=np.int32)
    nodes = NODES[kt]
    ls = np.stack([P[:, 0] + t - 1 for t in nodes])
    ls = np.where(np.abs(ls) < 1e-14, 1e-14, ls)
    one = np.zeros(4 * (kt + 1)); one[:] = 1.0                                    # P1: all shape functions sum to 1
    res = {}
    for dom in (NEG, POS, IF):
        tot = mx = mt = 0.0
        for el in range(6):
            T = xo.make_trafo(P[E[el]], E[el])
            q, tv, w = xo.spacetime_cut_rule(T, ls[:, E[el]], kt, dom, 2 * kt, 2)
            if not len(w):
                continue
            tot += one @ xo.elmat_st_from_rule(T, 1, kt, q, tv, w, [(ID, ID, 1.0, 0)]) @ one
            mx += one @ xo.elvec_st_from_rule(T, 1, kt, q, tv, w, [(ID, 1.0, 0, (1, 0, 0, 0))])
            mt += one @ xo.elvec_st_from_rule(T, 1, kt, q, tv, w, [(ID, 1.0, 0, (0, 0, 0, 1))])
        res[dom] = (tot, mx, mt)
    assert abs(res[NEG][0] - 0.5) < 1e-13 and abs(res[POS][0] - 0.5) < 1e-13 and abs(res[IF][0] - 1.0) < 1e-13
    assert abs(res[NEG][1] - 1 / 6) < 1e-13 and abs(res[NEG][2] - 1 / 6) < 1e-13
    assert abs(res[IF][1] - 0.5) < 1e-13                                             # int_0^1 (1 - t) dt on the interface


def test_quadratic_in_time_level_sets_known_answers():
    """P2-in-time level sets on the unit square (two triangles), space-time rule with time_order 2:
    (a) lset = x - (1/4 + t^2/2): no vertex changes sign, NEG area(t) = 1/4 + t^2/2, space-time volume 5/12;
    (b) lset = x - 2 t^2: the vertices with x = 1 change sign at t = 1/sqrt(2) (quadratic root finding,
        cutint/spacetimecutrule.cpp:20-48), NEG volume = int_0^{1/sqrt 2} 2 t^2 dt + 1 - 1/sqrt 2 = 1 - sqrt(2)/3"""
    P, E = unit_square()
    for fn, ref in ((lambda x, t: x - (0.25 + 0.5 * t * t), 5 / 12), (lambda x, t: x - 2 * t * t, 1 - np.sqrt(2) / 3)):
        ls = np.stack([fn(P[:, 0], t) for t in NODES[2]])
        ls = np.where(np.abs(ls) < 1e-14, 1e-14, ls)
        neg = st_sum(P, E, ls, 2, NEG, 2, 2, 1, 2, terms=[(ID, ID, 1.0, 0)])
        pos = st_sum(P, E, ls, 2, POS, 2, 2, 1, 2, terms=[(ID, ID, 1.0, 0)])
        assert abs(neg - ref) < 1e-13 and abs(neg + pos - 1.0) < 1e-13
