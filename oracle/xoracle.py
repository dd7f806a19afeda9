"""ctypes binding of the CPU restatement ORACLE (oracle/libxoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package ngsxfem_b200 never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NEG, POS, IF = 0, 1, 2
OP_ID, OP_GRAD, OP_DN = 0, 1, 2


class Term(C.Structure):
    _fields_ = [("op_trial", C.c_int), ("op_test", C.c_int), ("coef", C.c_double), ("hpow", C.c_int)]


class Trafo(C.Structure):
    _fields_ = [("dim", C.c_int), ("vtx", (C.c_double * 3) * 4), ("vnum", C.c_int * 4),
                ("def_order", C.c_int), ("deform", C.POINTER(C.c_double))]


def build(force=False):
    so = os.path.join(_HERE, "libxoracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h", ".inc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libxoracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.xo_integrate_monomial.restype = C.c_double
        L.xo_simplex_volume.restype = C.c_double
        L.xo_graph.restype = C.c_int64
        L.xo_batch_cut_rules.restype = C.c_int64
    return _LIB


def _p(a, t=C.c_double):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def make_trafo(vtx, vnum=None, def_order=0, deform=None):
    vtx = np.asarray(vtx, dtype=np.float64)
    dim = vtx.shape[1]
    T = Trafo()
    T.dim = dim
    for k in range(dim + 1):
        for i in range(dim):
            T.vtx[k][i] = vtx[k, i]
        T.vnum[k] = int(vnum[k]) if vnum is not None else k
    T.def_order = def_order
    if deform is not None:
        d = _f64(deform)
        T._keep = d
        T.deform = _p(d)
    return T


def ref_trafo(dim):
    vt = np.zeros((dim + 1, dim))
    for i in range(dim):
        vt[i, i] = 1.0
    return make_trafo(vt)


def select_rule(dim, order):
    L = lib()
    n = L.xo_rule_size(dim, order)
    pts = np.zeros((n, 3)); wts = np.zeros(n)
    L.xo_select_rule(dim, order, _p(pts), _p(wts))
    return pts, wts


def gauss_jacobi01(n, alpha):
    x = np.zeros(n); w = np.zeros(n)
    lib().xo_gauss_jacobi01(n, alpha, _p(x), _p(w))
    return x, w


def check_if_straight_cut(v, eps=0.0):
    v = _f64(v)
    return lib().xo_check_if_straight_cut(_p(v), C.c_int(len(v)), C.c_double(eps))


def straight_cut_rule(T, lsetvals, dt, order, cap=20000):
    v = _f64(lsetvals)
    pts = np.zeros((cap, 3)); wts = np.zeros(cap)
    std = C.c_int(0)
    n = lib().xo_straight_cut_rule(C.byref(T), _p(v), dt, order, _p(pts), _p(wts), cap, C.byref(std))
    if n < 0:
        raise RuntimeError(f"oracle error {n}")
    return pts[:n].copy(), wts[:n].copy(), bool(std.value)


def mlset_cut_rule(T, lsets, dts, order, cap=200000):
    lsets = _f64(lsets)  # (dim+1) x M
    M = lsets.shape[1]
    dts = np.ascontiguousarray(np.atleast_2d(dts), dtype=np.int32)
    pts = np.zeros((cap, 3)); wts = np.zeros(cap)
    n = lib().xo_mlset_cut_rule(C.byref(T), _p(lsets), M, _p(dts, C.c_int), dts.shape[0], order,
                                _p(pts), _p(wts), cap)
    if n < 0:
        raise RuntimeError(f"oracle error {n}")
    return pts[:n].copy(), wts[:n].copy()


def spacetime_cut_rule(T, lset_st, time_fe_order, dt, order_time, order_space, cap=200000):
    l = _f64(lset_st)
    pts = np.zeros((cap, 3)); tv = np.zeros(cap); wts = np.zeros(cap)
    n = lib().xo_spacetime_cut_rule(C.byref(T), _p(l), time_fe_order, dt, order_time, order_space,
                                    _p(pts), _p(tv), _p(wts), cap)
    if n < 0:
        raise RuntimeError(f"oracle error {n}")
    return pts[:n].copy(), tv[:n].copy(), wts[:n].copy()


def cutinfo_update(dim, el_vtx, phi):
    el_vtx = _i32(el_vtx); phi = _f64(phi)
    ne = el_vtx.shape[0]
    nb = (ne + 7) // 8
    bn = np.zeros(nb, np.uint8); bp = np.zeros(nb, np.uint8); bi = np.zeros(nb, np.uint8)
    cr = np.zeros(ne)
    err_el = C.c_int64(-1)
    err = lib().xo_cutinfo_update(dim, C.c_int64(ne), _p(el_vtx, C.c_int32), _p(phi), _p(bn, C.c_uint8),
                                  _p(bp, C.c_uint8), _p(bi, C.c_uint8), _p(cr), C.byref(err_el))
    return bn, bp, bi, cr, err, err_el.value


def ml_elements_of_type(dim, el_vtx, phis, dts):
    el_vtx = _i32(el_vtx); phis = _f64(phis)
    dts = np.ascontiguousarray(np.atleast_2d(dts), dtype=np.int32)
    ne = el_vtx.shape[0]
    bits = np.zeros((ne + 7) // 8, np.uint8)
    lib().xo_ml_elements_of_type(dim, C.c_int64(ne), _p(el_vtx, C.c_int32), _p(phis), C.c_int64(phis.shape[1]),
                                 phis.shape[0], _p(dts, C.c_int), dts.shape[0], _p(bits, C.c_uint8))
    return bits


def ml_elements_with_contribution(dim, el_vtx, vtx_xyz, phis, dts):
    el_vtx = _i32(el_vtx); phis = _f64(phis); vtx_xyz = _f64(vtx_xyz)
    dts = np.ascontiguousarray(np.atleast_2d(dts), dtype=np.int32)
    ne = el_vtx.shape[0]
    bits = np.zeros((ne + 7) // 8, np.uint8)
    err = lib().xo_ml_elements_with_contribution(dim, C.c_int64(ne), _p(el_vtx, C.c_int32), _p(vtx_xyz), _p(phis),
                                                 C.c_int64(phis.shape[1]), phis.shape[0], _p(dts, C.c_int),
                                                 dts.shape[0], _p(bits, C.c_uint8))
    if err:
        raise RuntimeError(f"oracle error {err}")
    return bits


def facets_with_neighbor_types(facet_els, a, b, bnd_val_a=True, bnd_val_b=True, use_and=True, periodic_facet=None):
    fe = _i32(facet_els)
    nf = fe.shape[0]
    out = np.zeros((nf + 7) // 8, np.uint8)
    if periodic_facet is not None:
        pf = _i32(periodic_facet)
        lib().xo_facets_with_neighbor_types_periodic(C.c_int64(nf), _p(fe, C.c_int32), _p(pf, C.c_int32), _p(a, C.c_uint8),
                                                     _p(b, C.c_uint8), int(bnd_val_a), int(bnd_val_b), int(use_and),
                                                     _p(out, C.c_uint8))
        return out
    lib().xo_facets_with_neighbor_types(C.c_int64(nf), _p(fe, C.c_int32), _p(a, C.c_uint8), _p(b, C.c_uint8),
                                        int(bnd_val_a), int(bnd_val_b), int(use_and), _p(out, C.c_uint8))
    return out


def elements_with_neighbor_facets(facet_els, facets, ne):
    fe = _i32(facet_els)
    out = np.zeros((ne + 7) // 8, np.uint8)
    lib().xo_elements_with_neighbor_facets(C.c_int64(fe.shape[0]), _p(fe, C.c_int32), _p(facets, C.c_uint8),
                                           C.c_int64(ne), _p(out, C.c_uint8))
    return out


def dofs_of_elements(el_dofs, elems, ndof):
    ed = _i32(el_dofs)
    out = np.zeros((ndof + 7) // 8, np.uint8)
    lib().xo_dofs_of_elements(C.c_int64(ed.shape[0]), ed.shape[1], _p(ed, C.c_int32), _p(elems, C.c_uint8),
                              C.c_int64(ndof), _p(out, C.c_uint8))
    return out


def h1_ndof(dim, order):
    return lib().xo_h1_ndof(dim, order)


def h1_shapes(dim, order, vnum, xhat):
    nd = h1_ndof(dim, order)
    sh = np.zeros(nd); dsh = np.zeros((nd, dim))
    vn = np.ascontiguousarray(vnum, dtype=np.int32)
    x = np.zeros(3); x[:len(xhat)] = xhat
    lib().xo_h1_shapes(dim, order, _p(vn, C.c_int), _p(x), _p(sh), _p(dsh))
    return sh, dsh


def trafo_eval(T, xhat):
    """(F dim x dim, x) of the element transformation at xhat"""
    F = np.zeros(9); x = np.zeros(3)
    xh = np.zeros(3); xh[:len(xhat)] = xhat
    lib().xo_trafo_eval(C.byref(T), _p(xh), _p(F), _p(x))
    d = T.dim
    return F[:d * d].reshape(d, d).copy(), x[:d].copy()


def terms_array(terms):
    arr = (Term * len(terms))()
    for i, (ou, ot, c, hp) in enumerate(terms):
        arr[i].op_trial = ou; arr[i].op_test = ot; arr[i].coef = c; arr[i].hpow = hp
    return arr


def cut_elmat(T, fe_order, lsetvals, dt, order, terms):
    nd = h1_ndof(T.dim, fe_order)
    m = np.zeros((nd, nd))
    v = _f64(lsetvals)
    ta = terms_array(terms)
    n = lib().xo_cut_elmat(C.byref(T), fe_order, _p(v), dt, order, ta, len(terms), _p(m))
    if n < 0:
        raise RuntimeError(f"oracle error {n}")
    return m, n


def elmat_from_rule(T, fe_order, lsetvals, pts, wts, terms, out=None):
    """xo_elmat_from_rule: ADDS sum_q coef h^k (w_q |det F|) op_test op_trial into out (nd x nd)"""
    nd = h1_ndof(T.dim, fe_order)
    m = np.zeros((nd, nd)) if out is None else out
    v = None if lsetvals is None else _f64(lsetvals)
    pts = _f64(pts); wts = _f64(wts)
    ta = terms_array(terms)
    lib().xo_elmat_from_rule(C.byref(T), fe_order, _p(v), _p(pts), _p(wts), len(wts), ta, len(terms), _p(m))
    return m


class FTerm(C.Structure):
    _fields_ = [("trial_kind", C.c_int), ("trial_side", C.c_int), ("test_kind", C.c_int), ("test_side", C.c_int),
                ("coef", C.c_double), ("hpow", C.c_int)]


def cut_facet_elmat(T1, T2, fe_order, facet_vnum, lset_facet, dt, order, terms):
    """xo_cut_facet_elmat (symboliccutbfi.cpp:886-1290 / :488-883): terms = [(trial_kind, trial_side, test_kind, test_side,
    coef, hpow)], kind 0 = value / 1 = nF.grad, side 0 = this element / 1 = .Other(); T2 None = one-sided (element boundary).
    returns (matrix (ns*nd)^2 rows = test, number of points)"""
    nd = h1_ndof(T1.dim, fe_order)
    N = nd * (2 if T2 is not None else 1)
    m = np.zeros((N, N))
    ta = (FTerm * max(1, len(terms)))()
    for i, t in enumerate(terms):
        ta[i].trial_kind, ta[i].trial_side, ta[i].test_kind, ta[i].test_side = int(t[0]), int(t[1]), int(t[2]), int(t[3])
        ta[i].coef, ta[i].hpow = float(t[4]), int(t[5])
    fv = np.ascontiguousarray(facet_vnum, dtype=np.int32)
    ls = _f64(lset_facet)
    n = lib().xo_cut_facet_elmat(C.byref(T1), C.byref(T2) if T2 is not None else None, fe_order,
                                 fv.ctypes.data_as(C.POINTER(C.c_int)), _p(ls), dt, order, ta, len(terms), _p(m))
    if n < 0:
        raise RuntimeError(f"oracle error {n}")
    return m, n


class LTerm(C.Structure):
    _fields_ = [("op_test", C.c_int), ("coef", C.c_double), ("hpow", C.c_int), ("expo", C.c_int * 3)]


def lterms_array(terms):
    """terms: (op_test, coef, hpow, (e0, e1, e2))"""
    arr = (LTerm * len(terms))()
    for i, (ot, c, hp, ex) in enumerate(terms):
        arr[i].op_test = ot; arr[i].coef = c; arr[i].hpow = hp
        for k in range(3):
            arr[i].expo[k] = int(ex[k]) if k < len(ex) else 0
    return arr


def batch_cut_elvecs(dim, fe_order, elems, el_vtx, vtx_xyz, phi, dt, order, terms, def_order=0, deform=None):
    el_vtx = _i32(el_vtx); vtx_xyz = _f64(vtx_xyz); phi = _f64(phi)
    el = None if elems is None else _i32(elems)
    n = el_vtx.shape[0] if el is None else len(el)
    nd = h1_ndof(dim, fe_order)
    vecs = np.zeros((n, nd))
    ta = lterms_array(terms)
    d = None if deform is None else _f64(deform)
    err = lib().xo_batch_cut_elvecs(dim, fe_order, C.c_int64(n), _p(el, C.c_int32), _p(el_vtx, C.c_int32), _p(vtx_xyz),
                                    _p(phi), dt, order, ta, len(terms), def_order, _p(d), _p(vecs))
    if err:
        raise RuntimeError(f"oracle error {err}")
    return vecs


def facet_patch_elmat(T1, T2, fe_order, order, coef, hpow):
    nd = h1_ndof(T1.dim, fe_order)
    m = np.zeros((2 * nd, 2 * nd))
    nfail = lib().xo_facet_patch_elmat(C.byref(T1), C.byref(T2), fe_order, order, C.c_double(coef), hpow, _p(m))
    return m, nfail


def num_threads():
    """threads the OpenMP loops of libxoracle.so use"""
    return int(lib().xo_num_threads())


def set_num_threads(n):
    lib().xo_set_num_threads(int(n))


def graph(ndof, table, once=False):
    """once=True: the one-call / one-sort-per-row variant the CPU baseline times (same result)"""
    t = _i32(table)
    nrows, width = t.shape
    rowptr = np.zeros(ndof + 1, np.int64)
    if once:
        L = lib()
        L.xo_graph_once.restype = C.c_int64
        L.xo_free.restype = None
        cp = C.POINTER(C.c_int32)()
        nnz = L.xo_graph_once(C.c_int64(ndof), C.c_int64(nrows), width, _p(t, C.c_int32), _p(rowptr, C.c_int64), C.byref(cp))
        col = np.ctypeslib.as_array(cp, shape=(max(nnz, 1),))[:nnz].copy()
        L.xo_free(cp)
        return rowptr, col
    nnz = lib().xo_graph(C.c_int64(ndof), C.c_int64(nrows), width, _p(t, C.c_int32), _p(rowptr, C.c_int64), None)
    col = np.zeros(nnz, np.int32)
    lib().xo_graph(C.c_int64(ndof), C.c_int64(nrows), width, _p(t, C.c_int32), _p(rowptr, C.c_int64), _p(col, C.c_int32))
    return rowptr, col


def scatter(ndof, table, mats, rowptr, col, vals=None, threads=False):
    """threads=True: timing-only variant (atomic adds over the host threads, order not fixed)"""
    t = _i32(table); m = _f64(mats)
    if vals is None:
        vals = np.zeros(len(col))
    fn = lib().xo_scatter_threads if threads else lib().xo_scatter
    fn.restype = None
    fn(C.c_int64(ndof), C.c_int64(t.shape[0]), t.shape[1], _p(t, C.c_int32), _p(m),
                     _p(rowptr, C.c_int64), _p(col, C.c_int32), _p(vals))
    return vals


def batch_cut_rules(dim, elems, el_vtx, vtx_xyz, phi, dt, order):
    el_vtx = _i32(el_vtx); vtx_xyz = _f64(vtx_xyz); phi = _f64(phi)
    el = None if elems is None else _i32(elems)
    n = el_vtx.shape[0] if el is None else len(el)
    off = np.zeros(n + 1, np.int64)
    tot = lib().xo_batch_cut_rules(dim, C.c_int64(n), _p(el, C.c_int32), _p(el_vtx, C.c_int32), _p(vtx_xyz), _p(phi),
                                   dt, order, _p(off, C.c_int64), None, None)
    if tot < 0:
        raise RuntimeError(f"oracle error {tot}")
    pts = np.zeros((tot, 3)); wts = np.zeros(tot)
    lib().xo_batch_cut_rules(dim, C.c_int64(n), _p(el, C.c_int32), _p(el_vtx, C.c_int32), _p(vtx_xyz), _p(phi),
                             dt, order, _p(off, C.c_int64), _p(pts), _p(wts))
    return off, pts, wts


def batch_cut_elmats(dim, fe_order, elems, el_vtx, vtx_xyz, phi, dt, order, terms, def_order=0, deform=None):
    el_vtx = _i32(el_vtx); vtx_xyz = _f64(vtx_xyz); phi = _f64(phi)
    el = None if elems is None else _i32(elems)
    n = el_vtx.shape[0] if el is None else len(el)
    nd = h1_ndof(dim, fe_order)
    mats = np.zeros((n, nd, nd))
    ta = terms_array(terms)
    d = None if deform is None else _f64(deform)
    err = lib().xo_batch_cut_elmats(dim, fe_order, C.c_int64(n), _p(el, C.c_int32), _p(el_vtx, C.c_int32), _p(vtx_xyz),
                                    _p(phi), dt, order, ta, len(terms), def_order, _p(d), _p(mats))
    if err:
        raise RuntimeError(f"oracle error {err}")
    return mats


def batch_facet_patch(dim, fe_order, facets, facet_els, el_vtx, vtx_xyz, order, coef, hpow, def_order=0, deform=None, downscale=1.0):
    el_vtx = _i32(el_vtx); vtx_xyz = _f64(vtx_xyz); fe = _i32(facet_els)
    fl = None if facets is None else _i32(facets)
    n = fe.shape[0] if fl is None else len(fl)
    nd = h1_ndof(dim, fe_order)
    mats = np.zeros((n, 2 * nd, 2 * nd))
    d = None if deform is None else _f64(deform)
    nfail = lib().xo_batch_facet_patch_scaled(dim, fe_order, C.c_int64(n), _p(fl, C.c_int32), _p(fe, C.c_int32),
                                              _p(el_vtx, C.c_int32), _p(vtx_xyz), order, C.c_double(coef), hpow,
                                              def_order, _p(d), C.c_double(downscale), _p(mats))
    if nfail < 0:
        raise ValueError("Only 0 < downscale <= 1 valid")
    return mats, nfail


def integrate_monomial(dim, el_vtx, vtx_xyz, phi, dt, order, expo):
    el_vtx = _i32(el_vtx); vtx_xyz = _f64(vtx_xyz); phi = _f64(phi)
    ex = np.ascontiguousarray(expo, dtype=np.int32)
    err = C.c_int(0)
    v = lib().xo_integrate_monomial(dim, C.c_int64(el_vtx.shape[0]), _p(el_vtx, C.c_int32), _p(vtx_xyz), _p(phi),
                                    dt, order, _p(ex, C.c_int), C.byref(err))
    if err.value:
        raise RuntimeError(f"oracle error {err.value}")
    return v


# ------------------------------------------------------------------ space-time FE on space-time rules
OP_DT = 3


class StLTerm(C.Structure):
    _fields_ = [("op_test", C.c_int), ("coef", C.c_double), ("hpow", C.c_int), ("expo", C.c_int * 4)]


def elmat_st_from_rule(T, fe_order, time_fe_order, pts, tvals, wts, terms, out=None):
    """xo_elmat_st_from_rule: ADDS into out (nd x nd), nd = nd_s*(time_fe_order+1), time-major"""
    nd = h1_ndof(T.dim, fe_order) * (time_fe_order + 1)
    m = np.zeros((nd, nd)) if out is None else out
    pts = _f64(pts); tv = _f64(tvals); wts = _f64(wts)
    ta = terms_array(terms)
    fn = lib().xo_elmat_st_from_rule
    fn.restype = None
    fn(C.byref(T), fe_order, time_fe_order, _p(pts), _p(tv), _p(wts), len(wts), ta, len(terms), _p(m))
    return m


def elvec_st_from_rule(T, fe_order, time_fe_order, pts, tvals, wts, terms, out=None):
    """terms: (op_test, coef, hpow, (e0, e1, e2, e_t)); ADDS into out (nd)"""
    nd = h1_ndof(T.dim, fe_order) * (time_fe_order + 1)
    v = np.zeros(nd) if out is None else out
    pts = _f64(pts); tv = _f64(tvals); wts = _f64(wts)
    arr = (StLTerm * len(terms))()
    for i, (ot, c, hp, ex) in enumerate(terms):
        arr[i].op_test = ot; arr[i].coef = c; arr[i].hpow = hp
        for k in range(4):
            arr[i].expo[k] = int(ex[k]) if k < len(ex) else 0
    fn = lib().xo_elvec_st_from_rule
    fn.restype = None
    fn(C.byref(T), fe_order, time_fe_order, _p(pts), _p(tv), _p(wts), len(wts), arr, len(terms), _p(v))
    return v


# ------------------------------------------------------------------ isoparametric deformation producer (xo_shift.c)
def rule_points_phys(dim, order_rule, el_vtx, vtx_xyz):
    el_vtx = _i32(el_vtx); vtx_xyz = _f64(vtx_xyz)
    ne = el_vtx.shape[0]
    np_ = lib().xo_rule_size(dim, order_rule)
    out = np.zeros((ne, np_, dim))
    lib().xo_rule_points_phys(dim, order_rule, C.c_int64(ne), _p(el_vtx, C.c_int32), _p(vtx_xyz), _p(out))
    return out


def set_from_points(dim, fe_order, order_rule, el_vtx, el_dofs, ndof, fvals):
    """GridFunction.Set restated: element-wise L2 projection of the point values (ne x np [x ncomp]) + mean per dof"""
    el_vtx = _i32(el_vtx); el_dofs = _i32(el_dofs); fv = _f64(fvals)
    ne = el_vtx.shape[0]
    ncomp = 1 if fv.ndim == 2 else fv.shape[2]
    out = np.zeros((ndof, ncomp))
    lib().xo_set_from_points(dim, fe_order, order_rule, C.c_int64(ne), _p(el_vtx, C.c_int32), _p(el_dofs, C.c_int32),
                             C.c_int64(ndof), ncomp, _p(fv), _p(out))
    return out[:, 0] if fv.ndim == 2 else out


def set_gradient(dim, fe_order, el_vtx, vtx_xyz, el_dofs, ndof, lset_ho):
    el_vtx = _i32(el_vtx); el_dofs = _i32(el_dofs); vtx_xyz = _f64(vtx_xyz); lh = _f64(lset_ho)
    out = np.zeros((ndof, dim))
    lib().xo_set_gradient(dim, fe_order, C.c_int64(el_vtx.shape[0]), _p(el_vtx, C.c_int32), _p(vtx_xyz), _p(el_dofs, C.c_int32),
                          C.c_int64(ndof), _p(lh), _p(out))
    return out


def project_shift(dim, fe_order, el_vtx, vtx_xyz, el_dofs, ndof, lset_ho, lset_p1, qn_mode=0, qn=None, ba=None, blending=None,
                  lower=0.0, upper=0.0, threshold=-1.0):
    """ProjectShift (lsetcurving/projshift.cpp:9-201): returns (deform ndof x dim, #searches that did not converge)"""
    el_vtx = _i32(el_vtx); el_dofs = _i32(el_dofs); vtx_xyz = _f64(vtx_xyz)
    lh = _f64(lset_ho); l1 = _f64(lset_p1)
    q = None if qn is None else _f64(qn)
    b = None if ba is None else np.ascontiguousarray(ba, dtype=np.uint8)
    bl = None if blending is None else _f64(blending)
    out = np.zeros((ndof, dim))
    L = lib()
    L.xo_project_shift.restype = C.c_int64
    nf = L.xo_project_shift(dim, fe_order, C.c_int64(el_vtx.shape[0]), _p(el_vtx, C.c_int32), _p(vtx_xyz), _p(el_dofs, C.c_int32),
                            C.c_int64(ndof), _p(lh), _p(l1), qn_mode, _p(q), _p(b, C.c_uint8), _p(bl), C.c_double(lower),
                            C.c_double(upper), C.c_double(threshold), _p(out))
    return out, int(nf)


def integrate_one_deformed(dim, el_vtx, vtx_xyz, phi, dt, order, def_order=0, el_dofs_def=None, def_vals=None):
    el_vtx = _i32(el_vtx); vtx_xyz = _f64(vtx_xyz); phi = _f64(phi)
    ed = None if el_dofs_def is None else _i32(el_dofs_def)
    dv = None if def_vals is None else _f64(def_vals)
    err = C.c_int(0)
    L = lib()
    L.xo_integrate_one_deformed.restype = C.c_double
    v = L.xo_integrate_one_deformed(dim, C.c_int64(el_vtx.shape[0]), _p(el_vtx, C.c_int32), _p(vtx_xyz), _p(phi), dt, order,
                                    def_order, _p(ed, C.c_int32), _p(dv), C.byref(err))
    if err.value:
        raise RuntimeError(f"oracle error {err.value}")
    return v


def element_in_band(vals, lower, upper):
    v = _f64(vals)
    return bool(lib().xo_element_in_band(_p(v), len(v), C.c_double(lower), C.c_double(upper)))
