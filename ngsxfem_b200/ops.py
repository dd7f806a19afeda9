"""Thin tensor-level wrappers of the C ABI (one function per xgpu_* entry point).

Tensors live on the CUDA device of the context; outputs are allocated with torch (device memory
plumbing only) and filled by the hand-written kernels in libxgpu.so.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import Batch, LTerm, Term, default_ctx, ptr

NEG, POS, IF = 0, 1, 2
OP_ID, OP_GRAD, OP_DN = 0, 1, 2
XGPU_ERR_UNSUPPORTED, XGPU_ERR_CAPACITY = -3, -14
BIT_OR, BIT_AND, BIT_ANDNOT, BIT_NOT, BIT_COPY, BIT_CLEAR, BIT_SET = range(7)


def nwords(nbits):
    return (int(nbits) + 31) // 32


def _i32(t):
    assert t.dtype == torch.int32 and t.is_cuda and t.is_contiguous(), "int32 contiguous CUDA tensor expected"
    return t


def _f64(t):
    assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous(), "float64 contiguous CUDA tensor expected"
    return t


def new_bits(nbits, device):
    return torch.zeros(nwords(nbits), dtype=torch.int32, device=device)


def terms_array(terms):
    arr = (Term * max(1, len(terms)))()
    for i, (ou, ot, c, hp) in enumerate(terms):
        arr[i].op_trial = int(ou); arr[i].op_test = int(ot); arr[i].coef = float(c); arr[i].hpow = int(hp)
    return arr


# ---------------------------------------------------------------- rules / basis (host tables)
def get_rule(dim, order, ctx=None):
    ctx = ctx or default_ctx()
    n = ctx.lib.xgpu_rule_size(ctx.h, dim, order)
    if n < 0:
        ctx.check(n)
    pts = torch.zeros((n, 3), dtype=torch.float64)
    wts = torch.zeros(n, dtype=torch.float64)
    rc = ctx.lib.xgpu_get_rule(ctx.h, dim, order, C.cast(pts.data_ptr(), _lib.c_f64p), C.cast(wts.data_ptr(), _lib.c_f64p))
    if rc < 0:
        ctx.check(rc)
    return pts, wts


def set_rule(dim, order, pts, wts, ctx=None):
    ctx = ctx or default_ctx()
    pts = pts.to(torch.float64).cpu().contiguous(); wts = wts.to(torch.float64).cpu().contiguous()
    ctx.check(ctx.lib.xgpu_set_rule(ctx.h, dim, order, wts.numel(), C.cast(pts.data_ptr(), _lib.c_f64p),
                                    C.cast(wts.data_ptr(), _lib.c_f64p)))


def h1_ndof(dim, order):
    return _lib.load().xgpu_h1_ndof(dim, order)


def h1_shapes_host(dim, order, vnum, xhat):
    lib = _lib.load()
    nd = lib.xgpu_h1_ndof(dim, order)
    sh = torch.zeros(nd, dtype=torch.float64); dsh = torch.zeros((nd, dim), dtype=torch.float64)
    vn = torch.as_tensor(vnum, dtype=torch.int32).contiguous()
    xh = torch.zeros(3, dtype=torch.float64); xh[:len(xhat)] = torch.as_tensor(xhat, dtype=torch.float64)
    rc = lib.xgpu_h1_shapes_host(dim, order, C.cast(vn.data_ptr(), _lib.c_i32p), C.cast(xh.data_ptr(), _lib.c_f64p),
                                 C.cast(sh.data_ptr(), _lib.c_f64p), C.cast(dsh.data_ptr(), _lib.c_f64p))
    if rc < 0:
        raise _lib.XgpuError(f"h1_shapes_host failed ({rc})")
    return sh, dsh


# ---------------------------------------------------------------- (a) CutInfo
def classify(dim, el_vtx, phi, want_ratio=True, ctx=None, out=None):
    ctx = ctx or default_ctx()
    ne = el_vtx.shape[0]
    dev = el_vtx.device
    if out is None:
        bn, bp, bi = new_bits(ne, dev), new_bits(ne, dev), new_bits(ne, dev)
        cr = torch.empty(ne, dtype=torch.float64, device=dev) if want_ratio else None
    else:
        bn, bp, bi, cr = out
    ctx.check(ctx.lib.xgpu_classify(ctx.h, dim, ne, ptr(_i32(el_vtx)), ptr(_f64(phi)), ptr(bn), ptr(bp), ptr(bi), ptr(cr)))
    return bn, bp, bi, cr


def bitset_op(op, a, b, nbits, out=None, ctx=None):
    ctx = ctx or default_ctx()
    if out is None:
        out = torch.empty_like(a)
    ctx.check(ctx.lib.xgpu_bitset_op(ctx.h, op, ptr(a), ptr(b), ptr(out), nbits))
    return out


def bitset_count(bits, nbits, ctx=None):
    ctx = ctx or default_ctx()
    c = C.c_int64(0)
    ctx.check(ctx.lib.xgpu_bitset_count(ctx.h, ptr(bits), nbits, C.byref(c)))
    return c.value


def bitset_counts(bitsets, nbits, ctx=None):
    """numbers of set bits of several bitsets, one host wait for all (xgpu_bitset_count_many); nbits: one length for all or one
    per bitset — bits with index >= nbits are not counted (a prefix of a longer bitset)"""
    ctx = ctx or default_ctx()
    k = len(bitsets)
    nb = [int(nbits)] * k if not isinstance(nbits, (list, tuple)) else [int(v) for v in nbits]
    ptrs = (_lib.VP * k)(*[_lib.VP(ptr(b)) for b in bitsets])
    nba = (C.c_int64 * k)(*nb)
    cnt = (C.c_int64 * k)()
    ctx.check(ctx.lib.xgpu_bitset_count_many(ctx.h, k, ptrs, nba, cnt))
    return [int(c) for c in cnt]


def bitset_to_list(bits, nbits, ctx=None, count=None):
    ctx = ctx or default_ctx()
    if count is None:
        count = bitset_count(bits, nbits, ctx)
    lst = torch.empty(count, dtype=torch.int32, device=bits.device)
    ctx.check(ctx.lib.xgpu_bitset_to_list(ctx.h, ptr(bits), nbits, ptr(lst)))
    return lst


def threshold_elements(cut_ratio, dt, threshold, ctx=None):
    ctx = ctx or default_ctx()
    ne = cut_ratio.numel()
    out = new_bits(ne, cut_ratio.device)
    ctx.check(ctx.lib.xgpu_threshold_elements(ctx.h, ne, ptr(_f64(cut_ratio)), dt, float(threshold), ptr(out)))
    return out


def _dts_array(dts):
    flat = [int(d) for tup in dts for d in tup]
    return (C.c_int * max(1, len(flat)))(*flat)


def classify_ml(dim, el_vtx, phis, dts, ctx=None, out=None):
    ctx = ctx or default_ctx()
    ne = el_vtx.shape[0]
    M, nv = phis.shape
    if out is None:
        out = new_bits(ne, el_vtx.device)
    ctx.check(ctx.lib.xgpu_classify_ml(ctx.h, dim, ne, ptr(_i32(el_vtx)), ptr(_f64(phis)), nv, M, _dts_array(dts),
                                       len(dts), ptr(out)))
    return out


def facets_with_neighbor_types(facet_els, a, b, bnd_val_a=True, bnd_val_b=True, use_and=True, ctx=None):
    ctx = ctx or default_ctx()
    nf = facet_els.shape[0]
    out = new_bits(nf, facet_els.device)
    ctx.check(ctx.lib.xgpu_facets_with_neighbor_types(ctx.h, nf, ptr(_i32(facet_els)), ptr(a), ptr(b), int(bnd_val_a),
                                                      int(bnd_val_b), int(use_and), ptr(out)))
    return out


def elements_with_neighbor_facets(facet_els, facet_bits, ne, ctx=None):
    ctx = ctx or default_ctx()
    out = new_bits(ne, facet_els.device)
    ctx.check(ctx.lib.xgpu_elements_with_neighbor_facets(ctx.h, facet_els.shape[0], ptr(_i32(facet_els)), ptr(facet_bits),
                                                         ne, ptr(out)))
    return out


def dofs_of_elements(el_dofs, el_bits, ndof, ctx=None):
    ctx = ctx or default_ctx()
    out = new_bits(ndof, el_dofs.device)
    ctx.check(ctx.lib.xgpu_dofs_of_elements(ctx.h, el_dofs.shape[0], el_dofs.shape[1], ptr(_i32(el_dofs)), ptr(el_bits),
                                            ndof, ptr(out)))
    return out


def bitset_rank(bits, nbits, ctx=None):
    ctx = ctx or default_ctx()
    mp = torch.empty(nbits, dtype=torch.int32, device=bits.device)
    n = C.c_int64(0)
    ctx.check(ctx.lib.xgpu_bitset_rank(ctx.h, ptr(bits), nbits, ptr(mp), C.byref(n)))
    return mp, n.value


def gather_rows(table, lst=None, mp=None, ctx=None):
    ctx = ctx or default_ctx()
    n = table.shape[0] if lst is None else lst.numel()
    out = torch.empty((n, table.shape[1]), dtype=torch.int32, device=table.device)
    ctx.check(ctx.lib.xgpu_gather_rows(ctx.h, n, ptr(lst), table.shape[1], ptr(_i32(table)), ptr(mp), ptr(out)))
    return out


# ---------------------------------------------------------------- (b) cut quadrature
def cutquad_count(dim, elems, el_vtx, phi, dt, order, ctx=None):
    """first half of cutquad: (offsets int64 [n+1], total number of points) — synchronises the context's stream"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    off = torch.empty(n + 1, dtype=torch.int64, device=el_vtx.device)
    tot = C.c_int64(0)
    ctx.check(ctx.lib.xgpu_cutquad_count(ctx.h, dim, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(phi)), dt, order,
                                         ptr(off), C.byref(tot)))
    return off, tot.value


def cutquad_fill(dim, elems, el_vtx, vtx_xyz, phi, dt, order, off, total, ctx=None):
    """second half of cutquad: (pts [total,3], wts [total]); only enqueues"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    pts = torch.empty((total, 3), dtype=torch.float64, device=el_vtx.device)
    wts = torch.empty(total, dtype=torch.float64, device=el_vtx.device)
    ctx.check(ctx.lib.xgpu_cutquad_fill(ctx.h, dim, n, ptr(elems), ptr(el_vtx), ptr(_f64(vtx_xyz)), ptr(phi), dt, order,
                                        ptr(off), ptr(pts), ptr(wts)))
    return pts, wts


def cutquad(dim, elems, el_vtx, vtx_xyz, phi, dt, order, ctx=None):
    """returns (offsets int64 [n+1], pts [total,3], wts [total])"""
    off, total = cutquad_count(dim, elems, el_vtx, phi, dt, order, ctx=ctx)
    pts, wts = cutquad_fill(dim, elems, el_vtx, vtx_xyz, phi, dt, order, off, total, ctx=ctx)
    return off, pts, wts


def cutquad_ml(dim, elems, el_vtx, vtx_xyz, phis, dts, order, ctx=None):
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    dev = el_vtx.device
    M, nv = phis.shape
    off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    tot = C.c_int64(0)
    da = _dts_array(dts)
    ctx.check(ctx.lib.xgpu_cutquad_ml_count(ctx.h, dim, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(phis)), nv, M, da,
                                            len(dts), order, ptr(off), C.byref(tot)))
    pts = torch.empty((tot.value, 3), dtype=torch.float64, device=dev)
    wts = torch.empty(tot.value, dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.xgpu_cutquad_ml_fill(ctx.h, dim, n, ptr(elems), ptr(el_vtx), ptr(_f64(vtx_xyz)), ptr(phis), nv, M, da,
                                           len(dts), order, ptr(off), ptr(pts), ptr(wts)))
    return off, pts, wts


def cutquad_st(dim, elems, el_vtx, vtx_xyz, phi_st, time_fe_order, dt, order_time, order_space, ctx=None):
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    dev = el_vtx.device
    nv = phi_st.shape[1]
    off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    tot = C.c_int64(0)
    ctx.check(ctx.lib.xgpu_cutquad_st_count(ctx.h, dim, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(phi_st)), nv,
                                            time_fe_order, dt, order_time, order_space, ptr(off), C.byref(tot)))
    pts = torch.empty((tot.value, 3), dtype=torch.float64, device=dev)
    tv = torch.empty(tot.value, dtype=torch.float64, device=dev)
    wts = torch.empty(tot.value, dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.xgpu_cutquad_st_fill(ctx.h, dim, n, ptr(elems), ptr(el_vtx), ptr(_f64(vtx_xyz)), ptr(phi_st), nv,
                                           time_fe_order, dt, order_time, order_space, ptr(off), ptr(pts), ptr(tv), ptr(wts)))
    return off, pts, tv, wts


def integrate_monomial(dim, elems, el_vtx, vtx_xyz, off, pts, wts, expo, ctx=None):
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    ex = (C.c_int * 3)(*([int(e) for e in expo] + [0] * (3 - len(expo))))
    res = C.c_double(0.0)
    ctx.check(ctx.lib.xgpu_integrate_monomial(ctx.h, dim, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(off),
                                              ptr(pts), ptr(wts), ex, C.byref(res)))
    return res.value


def classify_st(dim, el_vtx, phi_st, time_fe_order, time_order, out=None, want_ratio=True, ctx=None):
    """CutInfo on a time slab in one pass: (bits_neg, bits_pos, bits_if, cut_ratio)"""
    ctx = ctx or default_ctx()
    ne, dev = el_vtx.shape[0], el_vtx.device
    if out is None:
        out = (new_bits(ne, dev), new_bits(ne, dev), new_bits(ne, dev),
               torch.empty(ne, dtype=torch.float64, device=dev) if want_ratio else None)
    ctx.check(ctx.lib.xgpu_classify_st(ctx.h, dim, ne, ptr(_i32(el_vtx)), ptr(_f64(phi_st)), phi_st.shape[1], int(time_fe_order),
                                       int(time_order), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3])))
    return out


def rule_element_sums(off, wts, ctx=None):
    """per-element sums of the rule weights, added in rule order"""
    ctx = ctx or default_ctx()
    n = off.numel() - 1
    out = torch.empty(n, dtype=torch.float64, device=off.device)
    ctx.check(ctx.lib.xgpu_rule_element_sums(ctx.h, n, ptr(off), ptr(wts), ptr(out)))
    return out


def classify_volumes(vol_neg, vol_pos, out, ctx=None):
    """out = (bits_neg, bits_pos, bits_if, cut_ratio or None), cutinfo.cpp:104-124"""
    ctx = ctx or default_ctx()
    ctx.check(ctx.lib.xgpu_classify_volumes(ctx.h, vol_neg.numel(), ptr(vol_neg), ptr(vol_pos), ptr(out[0]), ptr(out[1]),
                                            ptr(out[2]), ptr(out[3])))
    return out


def integrate_monomial_st(dim, elems, el_vtx, vtx_xyz, off, pts, tvals, wts, expo, ctx=None):
    """space-time rule batch: expo = (e0, e1, e2, e_tref)"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    ex = (C.c_int * 4)(*[int(e) for e in expo])
    res = C.c_double(0.0)
    ctx.check(ctx.lib.xgpu_integrate_monomial_st(ctx.h, dim, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(off),
                                                 ptr(pts), ptr(tvals), ptr(wts), ex, C.byref(res)))
    return res.value


def integrate_monomial_bnd(dim, sdim, elems, el_vtx, vtx_xyz, off, pts, wts, expo, ctx=None):
    """boundary elements: simplices of dimension dim embedded in R^sdim"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    ex = (C.c_int * 3)(*([int(e) for e in expo] + [0] * (3 - len(expo))))
    res = C.c_double(0.0)
    ctx.check(ctx.lib.xgpu_integrate_monomial_bnd(ctx.h, dim, sdim, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)),
                                                  ptr(off), ptr(pts), ptr(wts), ex, C.byref(res)))
    return res.value


# ---------------------------------------------------------------- (c,d) matrices
def elmat_rule(dim, fe_order, elems, el_vtx, vtx_xyz, phi, off, pts, wts, terms, out=None, accumulate=False, slots=None, ctx=None):
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    nd = h1_ndof(dim, fe_order)
    if out is None:
        out = torch.empty((n, nd, nd), dtype=torch.float64, device=el_vtx.device)
        accumulate = False
    ctx.check(ctx.lib.xgpu_elmat_rule(ctx.h, dim, fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(phi),
                                      ptr(off), ptr(pts), ptr(wts), terms_array(terms), len(terms), int(accumulate), ptr(slots), ptr(out)))
    return out


def elmat_rule_pair(dim, fe_order, elems, el_vtx, vtx_xyz, phi, rule_a, terms_a, rule_b, terms_b, out=None, accumulate=False,
                    slots=None, ctx=None):
    """two integrators in one pass: rule_x = (offsets, pts, wts)"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    nd = h1_ndof(dim, fe_order)
    if out is None:
        out = torch.empty((n, nd, nd), dtype=torch.float64, device=el_vtx.device)
    ctx.check(ctx.lib.xgpu_elmat_rule_pair(ctx.h, dim, fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(phi),
                                           ptr(rule_a[0]), ptr(rule_a[1]), ptr(rule_a[2]), terms_array(terms_a), len(terms_a),
                                           ptr(rule_b[0]), ptr(rule_b[1]), ptr(rule_b[2]), terms_array(terms_b), len(terms_b),
                                           int(accumulate), ptr(slots), ptr(out)))
    return out


def elmat_cut_pair(dim, fe_order, elems, el_vtx, vtx_xyz, phi, dt_a, order_a, terms_a, dt_b=None, order_b=0, terms_b=(),
                   out=None, accumulate=False, slots=None, ctx=None):
    """cut-element matrices with the cut rules made on the fly (xgpu_elmat_cut_pair): integrator A = (dt_a, order_a, terms_a),
    optional integrator B; orders 1..2, mass / stiffness terms, affine mesh — raises XgpuError (unsupported) otherwise"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    nd = h1_ndof(dim, fe_order)
    if out is None:
        out = torch.empty((n, nd, nd), dtype=torch.float64, device=el_vtx.device)
        accumulate = False
    tb = list(terms_b)
    ctx.check(ctx.lib.xgpu_elmat_cut_pair(ctx.h, dim, fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(phi),
                                          int(dt_a), int(order_a), terms_array(terms_a), len(terms_a),
                                          -1 if dt_b is None else int(dt_b), int(order_b), terms_array(tb) if tb else None, len(tb),
                                          int(accumulate), ptr(slots), ptr(out)))
    return out


def elmat_std(dim, fe_order, elems, el_vtx, vtx_xyz, phi, order, terms, out=None, accumulate=False, slots=None, ctx=None):
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    nd = h1_ndof(dim, fe_order)
    if out is None:
        out = torch.empty((n, nd, nd), dtype=torch.float64, device=el_vtx.device)
        accumulate = False
    ctx.check(ctx.lib.xgpu_elmat_std(ctx.h, dim, fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(phi),
                                     order, terms_array(terms), len(terms), int(accumulate), ptr(slots), ptr(out)))
    return out


OP_DT = 3  # include/xgpu.h XGPU_OP_DT: derivative w.r.t. the reference time (dt(u))


def elmat_st(dim, fe_order, time_fe_order, elems, el_vtx, vtx_xyz, off, pts, tvals, wts, terms, out=None, accumulate=False,
             slots=None, ctx=None):
    """space-time element matrices on a space-time rule (cutquad_st): n x nd x nd, nd = nd_s*(time_fe_order+1), time-major"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    nd = h1_ndof(dim, fe_order) * (time_fe_order + 1)
    if out is None:
        out = torch.empty((n, nd, nd), dtype=torch.float64, device=el_vtx.device)
        accumulate = False
    ctx.check(ctx.lib.xgpu_elmat_st(ctx.h, dim, fe_order, time_fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)),
                                    ptr(off), ptr(pts), ptr(tvals), ptr(wts), terms_array(terms), len(terms), int(accumulate),
                                    ptr(slots), ptr(out)))
    return out


def elvec_st(dim, fe_order, time_fe_order, elems, el_vtx, vtx_xyz, off, pts, tvals, wts, terms, out=None, accumulate=False,
             slots=None, ctx=None):
    """space-time element vectors; terms: (op_test, coef, hpow, (e0, e1, e2, e_tref))"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    nd = h1_ndof(dim, fe_order) * (time_fe_order + 1)
    if out is None:
        out = torch.empty((n, nd), dtype=torch.float64, device=el_vtx.device)
        accumulate = False
    arr = (_lib.StLTerm * max(1, len(terms)))()
    for i, (ot, c, hp, ex) in enumerate(terms):
        arr[i].op_test = int(ot); arr[i].coef = float(c); arr[i].hpow = int(hp)
        for k in range(4):
            arr[i].expo[k] = int(ex[k]) if k < len(ex) else 0
    ctx.check(ctx.lib.xgpu_elvec_st(ctx.h, dim, fe_order, time_fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)),
                                    ptr(off), ptr(pts), ptr(tvals), ptr(wts), arr, len(terms), int(accumulate),
                                    ptr(slots), ptr(out)))
    return out


def lterms_array(terms):
    """terms: (op_test, coef, hpow, (e0, e1, e2))"""
    arr = (LTerm * max(1, len(terms)))()
    for i, (ot, c, hp, ex) in enumerate(terms):
        arr[i].op_test = int(ot); arr[i].coef = float(c); arr[i].hpow = int(hp)
        for k in range(3):
            arr[i].expo[k] = int(ex[k]) if k < len(ex) else 0
    return arr


def elvec_rule(dim, fe_order, elems, el_vtx, vtx_xyz, phi, off, pts, wts, terms, out=None, accumulate=False, slots=None, ctx=None):
    """element vectors on per-element rules (xfem/symboliccutlfi.cpp:51-140)"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    nd = h1_ndof(dim, fe_order)
    if out is None:
        out = torch.empty((n, nd), dtype=torch.float64, device=el_vtx.device)
    ctx.check(ctx.lib.xgpu_elvec_rule(ctx.h, dim, fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(phi),
                                      ptr(off), ptr(pts), ptr(wts), lterms_array(terms), len(terms), int(accumulate),
                                      ptr(slots), ptr(out)))
    return out


def elvec_std(dim, fe_order, elems, el_vtx, vtx_xyz, phi, order, terms, out=None, accumulate=False, slots=None, ctx=None):
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    nd = h1_ndof(dim, fe_order)
    if out is None:
        out = torch.empty((n, nd), dtype=torch.float64, device=el_vtx.device)
    ctx.check(ctx.lib.xgpu_elvec_std(ctx.h, dim, fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(phi),
                                     order, lterms_array(terms), len(terms), int(accumulate), ptr(slots), ptr(out)))
    return out


class _Downscale:
    """dFacetPatch(downscale=s) = SetIRScaling on the integrator (xfem/python_xfem.cpp:966-967): state of the context around
    one facet-patch call; 0 < s <= 1, else XgpuError (symboliccutbfi.cpp:2031-2033)"""

    def __init__(self, ctx, s):
        self.ctx, self.s = ctx, 1.0 if s is None else float(s)

    def __enter__(self):
        if self.s != 1.0:
            self.ctx.check(self.ctx.lib.xgpu_facetpatch_set_downscale(self.ctx.h, self.s))

    def __exit__(self, *exc):
        if self.s != 1.0:
            self.ctx.lib.xgpu_facetpatch_set_downscale(self.ctx.h, 1.0)
        return False


def facetpatch(dim, fe_order, facets, facet_els, el_vtx, vtx_xyz, order, coef, hpow, out=None, downscale=None, ctx=None):
    ctx = ctx or default_ctx()
    n = facet_els.shape[0] if facets is None else facets.numel()
    nd = h1_ndof(dim, fe_order)
    if out is None:
        out = torch.empty((n, 2 * nd, 2 * nd), dtype=torch.float64, device=el_vtx.device)
    with _Downscale(ctx, downscale):
        ctx.check(ctx.lib.xgpu_facetpatch(ctx.h, dim, fe_order, n, ptr(facets), ptr(_i32(facet_els)), ptr(_i32(el_vtx)),
                                          ptr(_f64(vtx_xyz)), order, float(coef), int(hpow), ptr(out)))
    return out


def facetpatch_st(dim, fe_order, time_fe_order, facets, facet_els, el_vtx, vtx_xyz, order, time_order, tref, coef, hpow,
                  downscale=None, ctx=None):
    """space-time ghost-penalty patch matrices: n x (2 nd) x (2 nd), nd = nd_s*(time_fe_order+1)"""
    ctx = ctx or default_ctx()
    n = facet_els.shape[0] if facets is None else facets.numel()
    nd = h1_ndof(dim, fe_order) * (time_fe_order + 1)
    out = torch.empty((n, 2 * nd, 2 * nd), dtype=torch.float64, device=el_vtx.device)
    with _Downscale(ctx, downscale):
        ctx.check(ctx.lib.xgpu_facetpatch_st(ctx.h, dim, fe_order, time_fe_order, n, ptr(facets), ptr(_i32(facet_els)),
                                             ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), int(order), int(time_order),
                                             float(0.0 if tref is None else tref), float(coef), int(hpow), ptr(out)))
    return out


# ---------------------------------------------------------------- (e) sparse assembly
class CsrPlan:
    def __init__(self, ctx, handle, nnz, ndof, nb):
        self.ctx, self.h, self.nnz, self.ndof, self.nb = ctx, handle, nnz, ndof, nb

    def phase_ms(self):
        """(transposed-table ms, pattern-pass ms) of xgpu_csr_plan, CUDA events on its stream"""
        ms = (C.c_double * 2)()
        self.ctx.check(self.ctx.lib.xgpu_plan_phase_ms(self.h, ms))
        return ms[0], ms[1]

    def __del__(self):
        try:
            if self.h:
                self.ctx.lib.xgpu_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass


class AffineCoefs:
    """a PROCEDURAL batch: per-element coefficients (n x 8) of uncut affine elements; the matrices
    sum_m coefs[e][m] * tref[m] are generated inside xgpu_csr_assemble_fused and never reach memory"""

    def __init__(self, coefs, tref_ptr, ncoef):
        self.coefs, self.tref_ptr, self.ncoef = coefs, tref_ptr, ncoef


def affine_coefs(dim, fe_order, elems, el_vtx, vtx_xyz, order, terms, ctx=None):
    """xgpu_affine_coefs + xgpu_reftensor: the procedural form of elmat_std's affine fast path"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    coefs = torch.empty((n, 8), dtype=torch.float64, device=el_vtx.device)
    ta = terms_array(terms)
    if order < 0:  # symboliccutbfi.cpp:100-118
        order = 2 * fe_order - min(0 if t[0] == OP_ID else 1 for t in terms) - min(0 if t[1] == OP_ID else 1 for t in terms)
    tref = _lib.VP()
    nc = C.c_int(0)
    ctx.check(ctx.lib.xgpu_reftensor(ctx.h, dim, fe_order, order, C.byref(tref), C.byref(nc)))
    ctx.check(ctx.lib.xgpu_affine_coefs(ctx.h, dim, n, ptr(_i32(elems)) if elems is not None else None, ptr(_i32(el_vtx)),
                                        ptr(_f64(vtx_xyz)), ta, len(terms), ptr(coefs)))
    return AffineCoefs(coefs, tref.value, nc.value)


def _batches(batches):
    arr = (Batch * len(batches))()
    for i, (table, mats) in enumerate(batches):
        arr[i].table = ptr(_i32(table))
        if isinstance(mats, AffineCoefs):
            arr[i].mats = None
            arr[i].coefs = ptr(mats.coefs)
        else:
            arr[i].mats = ptr(mats) if mats is not None else None
            arr[i].coefs = None
        arr[i].nrows = table.shape[0]
        arr[i].width = table.shape[1]
    return arr


def csr_plan(ndof, tables, ctx=None):
    ctx = ctx or default_ctx()
    arr = _batches([(t, None) for t in tables])
    h = _lib.VP()
    nnz = C.c_int64(0)
    ctx.check(ctx.lib.xgpu_csr_plan(ctx.h, ndof, arr, len(tables), C.byref(h), C.byref(nnz)))
    return CsrPlan(ctx, h, nnz.value, ndof, len(tables))


def adjacency_plan(ndof, tables, ctx=None):
    """transposed table only (LinearForm assembly)"""
    ctx = ctx or default_ctx()
    arr = _batches([(t, None) for t in tables])
    h = _lib.VP()
    ctx.check(ctx.lib.xgpu_adjacency_plan(ctx.h, ndof, arr, len(tables), C.byref(h)))
    return CsrPlan(ctx, h, 0, ndof, len(tables))


def vec_scatter(plan, batches, out=None):
    """batches: (table, vecs[nrows x width]); out[d] = fixed-order sum of the entries of dof d"""
    if out is None:
        out = torch.empty(plan.ndof, dtype=torch.float64, device=batches[0][0].device)
    arr = _batches(batches)
    plan.ctx.check(plan.ctx.lib.xgpu_vec_scatter(plan.ctx.h, plan.h, arr, len(batches), ptr(out)))
    return out


def vec_scatter_mean(plan, batches, out, comp=0):
    """like vec_scatter with the MEAN over the contributing entries; out: ndof x stride, component `comp` is written"""
    arr = _batches(batches)
    stride = 1 if out.dim() == 1 else out.shape[1]
    plan.ctx.check(plan.ctx.lib.xgpu_vec_scatter_mean(plan.ctx.h, plan.h, arr, len(batches), ptr(out), stride, comp))
    return out


# ---------------------------------------------------------------- tabulated coefficient functions (csrc/deform.cu)
def rule_phys(dim, elems, el_vtx, vtx_xyz, off, pts, want_meas=False, ctx=None):
    """physical coordinates (total x dim) [and measures |det F|] of the points of per-element rules; honours the deformation"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    tot = pts.shape[0]
    xyz = torch.empty((tot, dim), dtype=torch.float64, device=el_vtx.device)
    meas = torch.empty(tot, dtype=torch.float64, device=el_vtx.device) if want_meas else None
    if tot:
        ctx.check(ctx.lib.xgpu_rule_phys(ctx.h, dim, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(off), ptr(_f64(pts)),
                                         ptr(xyz), ptr(meas)))
    return (xyz, meas) if want_meas else xyz


def elements_with_marked(table, marks, ctx=None):
    """BitArray words of the rows of `table` (ne x nd) that hold at least one marked entry"""
    ctx = ctx or default_ctx()
    ne, nd = table.shape
    out = new_bits(ne, table.device)
    ctx.check(ctx.lib.xgpu_elements_with_marked(ctx.h, ne, nd, ptr(_i32(table)), ptr(marks), ptr(out)))
    return out


def sum_f64(vals, ctx=None):
    """deterministic sum of a device vector (xgpu_reduce_sum)"""
    ctx = ctx or default_ctx()
    r = C.c_double(0.0)
    ctx.check(ctx.lib.xgpu_reduce_sum(ctx.h, vals.numel(), ptr(_f64(vals)), C.byref(r)))
    return float(r.value)


def h1_eval_rule(dim, fe_order, elems, el_vtx, el_dofs, vals, off, pts, ctx=None):
    """values of an H1(fe_order) function (coefficient vector `vals`) at the points of per-element rules"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    out = torch.empty(pts.shape[0], dtype=torch.float64, device=el_vtx.device)
    if pts.shape[0]:
        ctx.check(ctx.lib.xgpu_h1_eval_rule(ctx.h, dim, fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_i32(el_dofs)), ptr(_f64(vals)),
                                            ptr(off), ptr(_f64(pts)), ptr(out)))
    return out


# ---------------------------------------------------------------- isoparametric deformation producer (csrc/lsetcurv.cu)
def band_elements(dim, el_vtx, lset_p1, lower=0.0, upper=0.0, ctx=None):
    """ElementInRelevantBand for every element -> BitArray words (lsetcurving/calcpointshift.cpp:120-141)"""
    ctx = ctx or default_ctx()
    ne = el_vtx.shape[0]
    bits = new_bits(ne, el_vtx.device)
    ctx.check(ctx.lib.xgpu_band_elements(ctx.h, dim, ne, ptr(_i32(el_vtx)), ptr(_f64(lset_p1)), float(lower), float(upper), ptr(bits)))
    return bits


def rule_points_phys(dim, order_rule, elems, el_vtx, vtx_xyz, ctx=None):
    """physical coordinates of the rule points on the listed elements: n x np x dim"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    np_ = ctx.lib.xgpu_rule_size(ctx.h, dim, order_rule)
    out = torch.empty((n, np_, dim), dtype=torch.float64, device=el_vtx.device)
    ctx.check(ctx.lib.xgpu_rule_points_phys(ctx.h, dim, order_rule, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(out)))
    return out


def h1_grad_points(dim, fe_order, order_rule, elems, el_vtx, vtx_xyz, el_dofs, vals, ctx=None):
    """physical gradient of an H1(fe_order) function at the rule points: n x np x dim"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    np_ = ctx.lib.xgpu_rule_size(ctx.h, dim, order_rule)
    out = torch.empty((n, np_, dim), dtype=torch.float64, device=el_vtx.device)
    ctx.check(ctx.lib.xgpu_h1_grad_points(ctx.h, dim, fe_order, order_rule, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)),
                                          ptr(_i32(el_dofs)), ptr(_f64(vals)), ptr(out)))
    return out


def l2_project(dim, fe_order, order_rule, elems, el_vtx, fvals, zero_vertex=False, ctx=None):
    """element-local L2 projection of point values (n x np x ncomp, or n x np) onto H1(fe_order): ncomp x n x nd"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    fv = _f64(fvals)
    ncomp = 1 if fv.dim() == 2 else fv.shape[2]
    nd = h1_ndof(dim, fe_order)
    out = torch.empty((ncomp, n, nd), dtype=torch.float64, device=el_vtx.device)
    ctx.check(ctx.lib.xgpu_l2_project(ctx.h, dim, fe_order, order_rule, n, ptr(elems), ptr(_i32(el_vtx)), ncomp, ptr(fv),
                                      int(bool(zero_vertex)), ptr(out)))
    return out


def shift_points(dim, fe_order, elems, el_vtx, vtx_xyz, el_dofs, lset_ho, lset_p1, qn_mode=0, qn=None, blending=None,
                 threshold=-1.0, ctx=None):
    """pointwise shift of ShiftIntegrator (xgpu_shift_points): (n x np x dim, #searches that did not converge)"""
    ctx = ctx or default_ctx()
    n = el_vtx.shape[0] if elems is None else elems.numel()
    np_ = ctx.lib.xgpu_rule_size(ctx.h, dim, 2 * fe_order)
    out = torch.empty((n, np_, dim), dtype=torch.float64, device=el_vtx.device)
    nf = C.c_int64(0)
    ctx.check(ctx.lib.xgpu_shift_points(ctx.h, dim, fe_order, n, ptr(elems), ptr(_i32(el_vtx)), ptr(_f64(vtx_xyz)), ptr(_i32(el_dofs)),
                                        ptr(_f64(lset_ho)), ptr(_f64(lset_p1)), int(qn_mode), ptr(_f64(qn)) if qn is not None else None,
                                        ptr(_f64(blending)) if blending is not None else None, float(threshold), ptr(out), C.byref(nf)))
    return out, int(nf.value)


def csr_pattern(plan, device):
    rowptr = torch.empty(plan.ndof + 1, dtype=torch.int64, device=device)
    col = torch.empty(plan.nnz, dtype=torch.int32, device=device)
    plan.ctx.check(plan.ctx.lib.xgpu_csr_pattern(plan.ctx.h, plan.h, ptr(rowptr), ptr(col)))
    return rowptr, col


def csr_scatter(plan, batches, rowptr, col, vals=None):
    """vals is OVERWRITTEN (every entry of the pattern is written once; explicit zeros kept);
    a batch (table, None) contributes sparsity only"""
    if vals is None:
        vals = torch.empty(plan.nnz, dtype=torch.float64, device=rowptr.device)
    arr = _batches(batches)
    plan.ctx.check(plan.ctx.lib.xgpu_csr_scatter(plan.ctx.h, plan.h, arr, len(batches), ptr(rowptr), ptr(col), ptr(vals)))
    return vals


def csr_assemble(plan, batches):
    """pattern + values in one pass (xgpu_csr_assemble): returns rowptr, col, vals"""
    dev = batches[0][0].device
    rowptr = torch.empty(plan.ndof + 1, dtype=torch.int64, device=dev)
    col = torch.empty(plan.nnz, dtype=torch.int32, device=dev)
    vals = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
    arr = _batches(batches)
    plan.ctx.check(plan.ctx.lib.xgpu_csr_assemble(plan.ctx.h, plan.h, arr, len(batches), ptr(rowptr), ptr(col), ptr(vals)))
    return rowptr, col, vals


def csr_plan_rows(ndof, tables, ctx=None):
    """transposed table + count pass (xgpu_csr_plan_rows): the plan of the one-row-pass assembly; raises XgpuError with
    code XGPU_ERR_UNSUPPORTED for shapes the one-pass kernels do not take (duplicate dofs in a table row, rows wider than 64)"""
    ctx = ctx or default_ctx()
    arr = _batches([(t, None) for t in tables])
    h = _lib.VP()
    nnz = C.c_int64(0)
    ctx.check(ctx.lib.xgpu_csr_plan_rows(ctx.h, ndof, arr, len(tables), C.byref(h), C.byref(nnz)))
    return CsrPlan(ctx, h, nnz.value, ndof, len(tables))


def csr_count_rows(plan):
    """count pass on an existing transposed table (xgpu_csr_count_rows); sets plan.nnz"""
    nnz = C.c_int64(0)
    plan.ctx.check(plan.ctx.lib.xgpu_csr_count_rows(plan.ctx.h, plan.h, C.byref(nnz)))
    plan.nnz = nnz.value
    return plan


def restrict_plan(el_dofs, lists, ndof_full, ctx=None):
    """Restrict + compressed dof rows + transposed table (xgpu_restrict_plan).  lists: element lists (int32 tensors).
    returns (dofmap, ndof, tables, plan); plan is an adjacency plan of the tables (csr_count_rows / vec_scatter)"""
    ctx = ctx or default_ctx()
    dev = el_dofs.device
    nd = el_dofs.shape[1]
    n = len(lists)
    tables = [torch.empty((l.numel(), nd), dtype=torch.int32, device=dev) for l in lists]
    dofmap = torch.empty(max(1, ndof_full), dtype=torch.int32, device=dev)
    lp = (_lib.VP * n)(*[ptr(_i32(l)) if l.numel() else None for l in lists])
    tp = (_lib.VP * n)(*[ptr(t) if t.numel() else None for t in tables])
    cnts = (C.c_int64 * n)(*[l.numel() for l in lists])
    ndof = C.c_int64(0)
    h = _lib.VP()
    ctx.check(ctx.lib.xgpu_restrict_plan(ctx.h, nd, ptr(_i32(el_dofs)), n, lp, cnts, ndof_full, ptr(dofmap), C.byref(ndof), tp,
                                         C.byref(h)))
    return dofmap[:ndof_full], ndof.value, tables, CsrPlan(ctx, h, 0, ndof.value, n)


def csr_assemble_fused(plan, batches, rows=None, out=None):
    """ONE visit per row (xgpu_csr_assemble_fused): pattern + values, `AffineCoefs` in place of the matrices = procedural
    batch.  plan: csr_plan_rows.  returns rowptr, col, vals.
    rows=(lo, hi): only these rows are finished (xgpu_csr_assemble_fused_rows) into out=(rowptr, col, vals) of a previous
    call (None: fresh arrays; the row offsets are copied by the call that allocates)"""
    ctx = plan.ctx
    dev = batches[0][0].device
    arr = _batches(batches)
    proc = [m for _, m in batches if isinstance(m, AffineCoefs)]
    tref, nc = (proc[0].tref_ptr, proc[0].ncoef) if proc else (None, 0)
    fresh = out is None
    if fresh:
        out = (torch.empty(plan.ndof + 1, dtype=torch.int64, device=dev), torch.empty(plan.nnz, dtype=torch.int32, device=dev),
               torch.empty(plan.nnz, dtype=torch.float64, device=dev))
    rowptr, col, vals = out
    if rows is None:
        ctx.check(ctx.lib.xgpu_csr_assemble_fused(ctx.h, plan.h, arr, len(batches), tref, nc, ptr(rowptr), ptr(col), ptr(vals)))
    else:
        ctx.check(ctx.lib.xgpu_csr_assemble_fused_rows(ctx.h, plan.h, arr, len(batches), tref, nc, ptr(rowptr), ptr(col), ptr(vals),
                                                       int(rows[0]), int(rows[1]), int(fresh)))
    return rowptr, col, vals


def rows_offsets(rowptr, rows, ctx=None):
    """offsets (n+1, int64, device) of the packed values of the listed CSR rows (rows < 0: length 0); no host sync"""
    ctx = ctx or default_ctx()
    off = torch.empty(rows.numel() + 1, dtype=torch.int64, device=rows.device)
    ctx.check(ctx.lib.xgpu_rows_offsets(ctx.h, rows.numel(), ptr(rowptr), ptr(_i32(rows)), ptr(off), None))
    return off


def rows_pack(rowptr, rows, off, vals, total, ctx=None):
    """total: size of the message buffer (the exact packed size, or a static capacity >= it: the tail stays unwritten)"""
    ctx = ctx or default_ctx()
    out = torch.empty(total, dtype=torch.float64, device=vals.device)
    ctx.check(ctx.lib.xgpu_rows_pack(ctx.h, rows.numel(), ptr(rowptr), ptr(_i32(rows)), ptr(off), ptr(vals), ptr(out)))
    return out


def rows_add(rowptr, rows, off, buf, vals, ctx=None):
    ctx = ctx or default_ctx()
    ctx.check(ctx.lib.xgpu_rows_add(ctx.h, rows.numel(), ptr(rowptr), ptr(_i32(rows)), ptr(off), ptr(buf), ptr(vals)))
    return vals


def facet_elmat(dim, fe_order, facet_vtx, facet_els, el_vtx, vtx_xyz, rule, terms, nsides, ctx=None):
    """xgpu_facet_elmat: matrices of cut integrals over facets (rule = (offsets, pts, wts) of xgpu_cutquad on the facet-vertex
    table); terms = [(trial kind, trial side, test kind, test side, coef, hpow)]; returns n x (nsides nd)^2"""
    ctx = ctx or default_ctx()
    n = facet_vtx.shape[0]
    nd = h1_ndof(dim, fe_order)
    N = nsides * nd
    mats = torch.empty((n, N, N), dtype=torch.float64, device=facet_vtx.device)
    ta = (_lib.FTerm * max(1, len(terms)))()
    for i, t in enumerate(terms):
        ta[i].trial_kind, ta[i].trial_side, ta[i].test_kind, ta[i].test_side = int(t[0]), int(t[1]), int(t[2]), int(t[3])
        ta[i].coef, ta[i].hpow = float(t[4]), int(t[5])
    off, q, w = rule
    ctx.check(ctx.lib.xgpu_facet_elmat(ctx.h, dim, fe_order, n, ptr(_i32(facet_vtx)), ptr(_i32(facet_els)), ptr(_i32(el_vtx)),
                                       ptr(_f64(vtx_xyz)), ptr(off), ptr(_f64(q)), ptr(_f64(w)), ta, len(terms), nsides, ptr(mats)))
    return mats


def iface_marks(has_bits, pair_elem, pair_gpos, nflags, ctx=None):
    """activity flags (uint8, nflags) of static interface candidates: flags[pair_gpos[i]] = 1 if element pair_elem[i] is active"""
    ctx = ctx or default_ctx()
    flags = torch.empty(max(1, nflags), dtype=torch.uint8, device=has_bits.device)
    ctx.check(ctx.lib.xgpu_iface_marks(ctx.h, ptr(has_bits), pair_elem.numel(), ptr(pair_elem), ptr(pair_gpos), nflags, ptr(flags)))
    return flags[:nflags]


def iface_rows(cand, flags, dofmap, ctx=None):
    """rows[i] = flags[i] ? dofmap[cand[i]] : -1"""
    ctx = ctx or default_ctx()
    rows = torch.empty(cand.numel(), dtype=torch.int32, device=cand.device)
    ctx.check(ctx.lib.xgpu_iface_rows(ctx.h, cand.numel(), ptr(cand), ptr(flags), ptr(dofmap), ptr(rows)))
    return rows


def iface_index(cand, flags, dofmap, n_lo, ctx=None):
    """idx[i] = flags[i] ? dofmap[cand[i]] - n_lo : -1 (int64: the owner-local index message)"""
    ctx = ctx or default_ctx()
    idx = torch.empty(cand.numel(), dtype=torch.int64, device=cand.device)
    ctx.check(ctx.lib.xgpu_iface_index(ctx.h, cand.numel(), ptr(cand), ptr(flags), ptr(dofmap), int(n_lo), ptr(idx)))
    return idx


def l2g_own(ndof, n_lo, n_own, counts, rank, ctx=None):
    ctx = ctx or default_ctx()
    l2g = torch.empty(ndof, dtype=torch.int64, device=counts.device)
    ctx.check(ctx.lib.xgpu_l2g_own(ctx.h, ndof, int(n_lo), int(n_own), ptr(counts), int(rank), ptr(l2g)))
    return l2g


def l2g_ghosts(l2g, dofmap_range, rbuf, counts, q, ctx=None):
    ctx = ctx or default_ctx()
    ctx.check(ctx.lib.xgpu_l2g_ghosts(ctx.h, dofmap_range.numel(), ptr(dofmap_range), ptr(rbuf), ptr(counts), int(q), ptr(l2g)))
    return l2g


def csr_spmv(rowptr, col, vals, x, ctx=None):
    ctx = ctx or default_ctx()
    y = torch.empty(rowptr.numel() - 1, dtype=torch.float64, device=x.device)
    ctx.check(ctx.lib.xgpu_csr_spmv(ctx.h, y.numel(), ptr(rowptr), ptr(col), ptr(vals), ptr(_f64(x)), ptr(y)))
    return y
