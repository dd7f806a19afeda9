"""GPU parity tests: every CUDA kernel of libxgpu.so (through the C ABI) against the CPU oracle on
the same seeded inputs.  Bars: bit-exact for bitsets / lists / offsets / CSR pattern; 1e-13
relative for quadrature points and weights (north_star b); 1e-12 relative (Frobenius) for
element matrices and assembled values (north_star e)."""
import numpy as np
import pytest
import torch

from ngsxfem_b200 import meshes, ops, fes
from ngsxfem_b200._lib import XgpuError
from oracle import xoracle as xo
from util import interpolate_to_p1, mesh_np, unpack_bits

pytestmark = pytest.mark.gpu

NEG, POS, IF = 0, 1, 2
DEV = "cuda"


def bits_np(t):
    return t.cpu().numpy().view(np.uint8)


def to_dev(m):
    return m.to(DEV)


def level_set(m, kind, seed=0):
    pts, _ = mesh_np(m)
    rng = np.random.default_rng(seed)
    c = 0.5 + 0.05 * rng.standard_normal(m.dim)
    if kind == "sphere":
        r = 0.31 + 0.02 * rng.random()
        f = lambda *x: np.sqrt(sum((xi - ci) ** 2 for xi, ci in zip(x, c))) - r
    else:
        nrm = rng.standard_normal(m.dim)
        f = lambda *x: sum(n * (xi - ci) for n, xi, ci in zip(nrm, x, c))
    return interpolate_to_p1(f, pts)


def make_mesh(dim, n):
    if dim == 1:
        return meshes.MakeStructured1DMesh(n)
    if dim == 2:
        return meshes.MakeStructured2DMesh(nx=n, ny=n + 1)
    return meshes.MakeStructured3DMesh(nx=n, ny=n + 1, nz=n - 1)


# ------------------------------------------------------------------ tables
def test_rules_and_shapes_match_oracle():
    for dim in (1, 2, 3):
        for order in range(0, 13):
            p, w = ops.get_rule(dim, order)
            po, wo = xo.select_rule(dim, order)
            assert p.shape[0] == po.shape[0]
            assert np.allclose(p.numpy(), po, rtol=0, atol=2e-16 * (order + 1))
            assert np.allclose(w.numpy(), wo, rtol=4e-16 * (order + 1), atol=0)
    rng = np.random.default_rng(1)
    for dim in (1, 2, 3):
        for order in (1, 2, 3, 4):
            for _ in range(5):
                vnum = rng.permutation(50)[:dim + 1]
                x = rng.random(dim) / dim
                sh, dsh = ops.h1_shapes_host(dim, order, vnum, x)
                so, dso = xo.h1_shapes(dim, order, vnum, x)
                assert np.allclose(sh.numpy(), so, rtol=1e-14, atol=1e-15)
                assert np.allclose(dsh.numpy(), dso, rtol=1e-14, atol=1e-14)


# ------------------------------------------------------------------ (a) classify
@pytest.mark.parametrize("dim,n", [(1, 37), (2, 23), (3, 9)])
@pytest.mark.parametrize("kind", ["sphere", "plane"])
def test_classify_matches_oracle(dim, n, kind):
    m = make_mesh(dim, n)
    phi = level_set(m, kind, seed=dim)
    pts, els = mesh_np(m)
    bn, bp, bi, cr, err, _ = xo.cutinfo_update(dim, els, phi)
    assert err == 0
    md = to_dev(m)
    gn, gp, gi, gcr = ops.classify(dim, md.elements, torch.from_numpy(phi).to(DEV))
    nb = len(bn)
    assert np.array_equal(bits_np(gn)[:nb], bn)
    assert np.array_equal(bits_np(gp)[:nb], bp)
    assert np.array_equal(bits_np(gi)[:nb], bi)
    assert unpack_bits(bi, m.ne).sum() > 0
    assert np.allclose(gcr.cpu().numpy(), cr, rtol=1e-13, atol=1e-15)
    assert ops.default_ctx().status()[0] == 0


def test_classify_zero_vertex_values():
    # exact zeros: uncut elements with a zero vertex stay NEG/POS; a cut element with a zero vertex makes
    # the reference throw (straightcutrule.cpp:45-49) -> error status with the first failing element
    els = np.array([[0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 5]], dtype=np.int32)
    phi = np.array([0.0, 1.0, 2.0, -1.0, 0.0, 0.0])
    bn, bp, bi, cr, err, err_el = xo.cutinfo_update(2, els, phi)
    ctx = ops.default_ctx()
    gn, gp, gi, gcr = ops.classify(2, torch.from_numpy(els).to(DEV), torch.from_numpy(phi).to(DEV))
    code, el = ctx.status()
    ctx.lib.xgpu_clear_status(ctx.h)
    assert err != 0 and code == -11 and el == err_el
    assert np.array_equal(bits_np(gn)[:1], bn) and np.array_equal(bits_np(gp)[:1], bp) and np.array_equal(bits_np(gi)[:1], bi)


def test_classify_empty_and_ragged_sizes():
    for ne in (0, 1, 31, 32, 33, 1025):
        rng = np.random.default_rng(ne)
        nv = 50
        els = np.stack([rng.permutation(nv)[:4] for _ in range(ne)]).astype(np.int32) if ne else np.zeros((0, 4), np.int32)
        phi = rng.standard_normal(nv)
        gn, gp, gi, gcr = ops.classify(3, torch.from_numpy(els).to(DEV), torch.from_numpy(phi).to(DEV))
        if ne == 0:
            continue
        bn, bp, bi, cr, err, _ = xo.cutinfo_update(3, els, phi)
        nb = len(bn)
        assert np.array_equal(bits_np(gn)[:nb], bn) and np.array_equal(bits_np(gi)[:nb], bi)
        assert np.allclose(gcr.cpu().numpy(), cr, rtol=1e-13, atol=1e-15)
        # tail bits beyond ne are zero
        tot = sum(int(np.unpackbits(bits_np(b)).sum()) for b in (gn, gp, gi))
        assert tot == ne


def test_bitset_algebra_and_compaction():
    rng = np.random.default_rng(5)
    for n in (1, 31, 64, 1000, 100003):
        a = rng.random(n) < 0.3
        b = rng.random(n) < 0.5
        pa = np.packbits(a, bitorder="little"); pb = np.packbits(b, bitorder="little")
        nw = ops.nwords(n)
        ta = torch.zeros(nw * 4, dtype=torch.uint8); ta[:len(pa)] = torch.from_numpy(pa)
        tb = torch.zeros(nw * 4, dtype=torch.uint8); tb[:len(pb)] = torch.from_numpy(pb)
        ga = ta.view(torch.int32).to(DEV); gb = tb.view(torch.int32).to(DEV)
        for op, ref in ((ops.BIT_OR, a | b), (ops.BIT_AND, a & b), (ops.BIT_ANDNOT, a & ~b), (ops.BIT_NOT, ~a)):
            r = ops.bitset_op(op, ga, gb, n)
            assert np.array_equal(unpack_bits(bits_np(r), n), ref)
            assert np.unpackbits(bits_np(r)).sum() == ref.sum()  # tail stays clear
        assert ops.bitset_count(ga, n) == a.sum()
        lst = ops.bitset_to_list(ga, n)
        assert np.array_equal(lst.cpu().numpy(), np.nonzero(a)[0])
        mp, na = ops.bitset_rank(ga, n)
        ref = np.where(a, np.cumsum(a) - 1, -1)
        assert na == a.sum() and np.array_equal(mp.cpu().numpy(), ref)


def test_marker_helpers_match_oracle():
    m = meshes.MakeStructured3DMesh(nx=6)
    phi = level_set(m, "sphere", 3)
    pts, els = mesh_np(m)
    md = to_dev(m)
    _, _, fe = md.facets()
    gn, gp, gi, _ = ops.classify(3, md.elements, torch.from_numpy(phi).to(DEV))
    ghas = ops.bitset_op(ops.BIT_OR, gn, gi, m.ne)
    nbe = (m.ne + 7) // 8
    hasneg, ifb = bits_np(ghas)[:nbe].copy(), bits_np(gi)[:nbe].copy()
    fe_np = fe.cpu().numpy()
    for use_and in (True, False):
        for bv in (True, False):
            gf = ops.facets_with_neighbor_types(fe, ghas, gi, bv, bv, use_and)
            of = xo.facets_with_neighbor_types(fe_np, hasneg, ifb, bv, bv, use_and)
            assert np.array_equal(bits_np(gf)[:len(of)], of)
    gf = ops.facets_with_neighbor_types(fe, ghas, gi)
    of = xo.facets_with_neighbor_types(fe_np, hasneg, ifb)
    ge = ops.elements_with_neighbor_facets(fe, gf, m.ne)
    oe = xo.elements_with_neighbor_facets(fe_np, of, m.ne)
    assert np.array_equal(bits_np(ge)[:len(oe)], oe)
    V = fes.H1(md, order=2)
    gd = ops.dofs_of_elements(V.el_dofs, ghas, V.ndof)
    od = xo.dofs_of_elements(V.el_dofs.cpu().numpy(), hasneg, V.ndof)
    assert np.array_equal(bits_np(gd)[:len(od)], od)
    # GetElementsWithSharedVertex (xfem/cutinfo.cpp:568-597) against its definition in numpy
    import ngsxfem_b200 as xf
    shared = xf.GetElementsWithSharedVertex(md, xf.BitArray(m.ne, words=gi))
    in_if = unpack_bits(bits_np(gi), m.ne).astype(bool)
    vmark = np.zeros(m.nv, dtype=bool); vmark[els[in_if].reshape(-1)] = True
    assert np.array_equal(unpack_bits(bits_np(shared.bits), m.ne).astype(bool), vmark[els].any(axis=1))
    assert shared.NumSet() > int(in_if.sum())
    thr = ops.threshold_elements(ops.classify(3, md.elements, torch.from_numpy(phi).to(DEV))[3], NEG, 0.3)
    cr = xo.cutinfo_update(3, els, phi)[3]
    assert np.array_equal(unpack_bits(bits_np(thr), m.ne), cr >= 0.3)


def test_classify_ml_matches_oracle():
    m = meshes.MakeStructured3DMesh(nx=7)
    pts, els = mesh_np(m)
    phis = np.stack([level_set(m, "plane", s) for s in range(3)] + [level_set(m, "sphere", 9)])
    md = to_dev(m)
    for dts in ([(NEG, NEG, NEG, NEG)], [(IF, NEG, POS, NEG), (NEG, IF, NEG, NEG)], [(POS, POS, IF, IF)]):
        g = ops.classify_ml(3, md.elements, torch.from_numpy(phis).to(DEV), dts)
        o = xo.ml_elements_of_type(3, els, phis, dts)
        assert np.array_equal(bits_np(g)[:len(o)], o)


# ------------------------------------------------------------------ (b) cut quadrature
@pytest.mark.parametrize("dim,n", [(1, 11), (2, 13), (3, 6)])
@pytest.mark.parametrize("dt", [NEG, POS, IF])
@pytest.mark.parametrize("order", [0, 1, 2, 4, 5, 8])
def test_cutquad_matches_oracle(dim, n, dt, order):
    m = make_mesh(dim, n)
    phi = level_set(m, "sphere", 10 + dim)
    pts, els = mesh_np(m)
    md = to_dev(m)
    off, p, w = ops.cutquad(dim, None, md.elements, md.points, torch.from_numpy(phi).to(DEV), dt, order)
    oo, po, wo = xo.batch_cut_rules(dim, None, els, pts, phi, dt, order)
    assert np.array_equal(off.cpu().numpy(), oo)
    assert len(wo) > 0
    assert np.allclose(p.cpu().numpy(), po, rtol=1e-13, atol=1e-15)
    assert np.allclose(w.cpu().numpy(), wo, rtol=1e-13, atol=1e-300)
    assert ops.default_ctx().status()[0] == 0


def test_cutquad_on_element_list_and_integrals():
    # the IF list of a sphere cut, integrals through the GPU path against closed forms (KATs of
    # tests/pytests/test_straightcutrule.py:187-208 on the unit cube)
    m = meshes.MakeStructured3DMesh(nx=3)
    pts, els = mesh_np(m)
    phi = interpolate_to_p1(lambda x, y, z: 1 - 2 * x - 2 * y - 2 * z, pts)
    md = to_dev(m)
    gphi = torch.from_numpy(phi).to(DEV)
    ref = {POS: 1. / 48, NEG: 47. / 48, IF: np.sqrt(3) / 8}
    for order in (2, 4, 8):
        for dt in (NEG, POS, IF):
            off, p, w = ops.cutquad(3, None, md.elements, md.points, gphi, dt, order)
            val = ops.integrate_monomial(3, None, md.elements, md.points, off, p, w, [0, 0, 0])
            assert abs(val - ref[dt]) < 5e-15 * (order + 1) ** 2
    gn, gp, gi, _ = ops.classify(3, md.elements, gphi)
    lst = ops.bitset_to_list(gi, m.ne)
    off, p, w = ops.cutquad(3, lst, md.elements, md.points, gphi, NEG, 3)
    oo, po, wo = xo.batch_cut_rules(3, lst.cpu().numpy(), els, pts, phi, NEG, 3)
    assert np.array_equal(off.cpu().numpy(), oo) and np.allclose(w.cpu().numpy(), wo, rtol=1e-13)


# ------------------------------------------------------------------ (c) element matrices
def rel_err(a, b):
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    den = np.maximum(np.linalg.norm(b, axis=1), 1e-300)
    return (np.linalg.norm(a - b, axis=1) / den).max()


TERMS_VOL = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 1.0, 0)]
TERMS_NITSCHE = [(ops.OP_DN, ops.OP_ID, -1.0, 0), (ops.OP_ID, ops.OP_DN, -1.0, 0), (ops.OP_ID, ops.OP_ID, 40.0, -1)]


@pytest.mark.parametrize("dim,n", [(2, 9), (3, 5)])
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_elmat_cut_matches_oracle(dim, n, p):
    m = make_mesh(dim, n)
    phi = level_set(m, "sphere", 20 + dim)
    pts, els = mesh_np(m)
    md = to_dev(m)
    gphi = torch.from_numpy(phi).to(DEV)
    _, _, gi, _ = ops.classify(dim, md.elements, gphi)
    lst = ops.bitset_to_list(gi, m.ne)
    assert lst.numel() > 0
    lnp = lst.cpu().numpy()
    for dt, terms in ((NEG, TERMS_VOL), (POS, TERMS_VOL), (IF, TERMS_NITSCHE)):
        order = 2 * p
        off, q, w = ops.cutquad(dim, lst, md.elements, md.points, gphi, dt, order)
        g = ops.elmat_rule(dim, p, lst, md.elements, md.points, gphi, off, q, w, terms)
        o = xo.batch_cut_elmats(dim, p, lnp, els, pts, phi, dt, order, terms)
        assert rel_err(g.cpu().numpy(), o) < 1e-12
    # accumulate: NEG part + IF part in one matrix (CalcElementMatrixAdd semantics)
    off, q, w = ops.cutquad(dim, lst, md.elements, md.points, gphi, NEG, 2 * p)
    g = ops.elmat_rule(dim, p, lst, md.elements, md.points, gphi, off, q, w, TERMS_VOL)
    off2, q2, w2 = ops.cutquad(dim, lst, md.elements, md.points, gphi, IF, 2 * p)
    ops.elmat_rule(dim, p, lst, md.elements, md.points, gphi, off2, q2, w2, TERMS_NITSCHE, out=g, accumulate=True)
    o = xo.batch_cut_elmats(dim, p, lnp, els, pts, phi, NEG, 2 * p, TERMS_VOL) + \
        xo.batch_cut_elmats(dim, p, lnp, els, pts, phi, IF, 2 * p, TERMS_NITSCHE)
    assert rel_err(g.cpu().numpy(), o) < 1e-12


@pytest.mark.parametrize("dim,n", [(1, 9), (2, 7), (3, 4)])
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_elmat_uncut_matches_oracle(dim, n, p):
    m = make_mesh(dim, n)
    pts, els = mesh_np(m)
    if dim > 1:  # distort so that elements are not congruent
        rng = np.random.default_rng(7)
        pts = pts + 0.15 / n * rng.random(pts.shape)
        m = meshes.Mesh(torch.from_numpy(pts), m.elements)
    phi = -np.ones(m.nv)  # all NEG
    md = to_dev(m)
    gphi = torch.from_numpy(phi).to(DEV)
    for terms, order in ((TERMS_VOL, -1), (TERMS_VOL[:1], -1), ([(0, 0, 2.5, -2)], 2 * p), (TERMS_NITSCHE, 2 * p)):
        g = ops.elmat_std(dim, p, None, md.elements, md.points, gphi, order, terms)
        o = xo.batch_cut_elmats(dim, p, None, els, pts, phi, NEG, order, terms)
        assert rel_err(g.cpu().numpy(), o) < 1e-12, (terms, order)
    # basis independent invariants: sum of mass matrix entries = element volume; stiffness kills constants
    g = ops.elmat_std(dim, p, None, md.elements, md.points, gphi, -1, [(0, 0, 1.0, 0)]).cpu().numpy()
    from math import factorial
    vols = np.abs(np.linalg.det(np.stack([pts[els[:, k]] - pts[els[:, dim]] for k in range(dim)], axis=1))) / factorial(dim)
    nv = dim + 1  # vertex functions form a partition of unity
    assert np.allclose(g[:, :nv, :nv].sum(axis=(1, 2)), vols, rtol=1e-12)


# ------------------------------------------------------------------ (d) facet patch
@pytest.mark.parametrize("dim,n", [(2, 6), (3, 3)])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_facetpatch_matches_oracle(dim, n, p):
    m = make_mesh(dim, n)
    pts, els = mesh_np(m)
    rng = np.random.default_rng(3)
    pts = pts + 0.1 / n * rng.random(pts.shape)
    m = meshes.Mesh(torch.from_numpy(pts), m.elements)
    md = to_dev(m)
    _, _, fe = md.facets()
    fe_np = fe.cpu().numpy()
    inner = np.nonzero(fe_np[:, 1] >= 0)[0].astype(np.int32)
    sel = inner[:: max(1, len(inner) // 40)]
    g = ops.facetpatch(dim, p, torch.from_numpy(sel).to(DEV), fe, md.elements, md.points, -1, 0.1, -2)
    o, nfail = xo.batch_facet_patch(dim, p, sel, fe_np, els, pts, -1, 0.1, -2)
    assert nfail == 0
    assert rel_err(g.cpu().numpy(), o) < 1e-12
    # jump of a global polynomial of degree <= p vanishes: J^T M J = 0 cannot be tested basis-free
    # without interpolation; instead check symmetry and positive semi-definiteness
    gm = g.cpu().numpy()
    assert np.allclose(gm, gm.transpose(0, 2, 1), rtol=1e-12, atol=1e-14 * np.abs(gm).max())


# ------------------------------------------------------------------ (e) CSR
@pytest.mark.parametrize("dim,n,p", [(2, 8, 1), (2, 6, 2), (3, 4, 2), (3, 3, 3)])
def test_csr_pattern_and_values_match_oracle(dim, n, p):
    m = make_mesh(dim, n)
    phi = level_set(m, "sphere", 30 + dim)
    pts, els = mesh_np(m)
    md = to_dev(m)
    gphi = torch.from_numpy(phi).to(DEV)
    gn, gp, gi, _ = ops.classify(dim, md.elements, gphi)
    ghas = ops.bitset_op(ops.BIT_OR, gn, gi, m.ne)
    V = fes.Restrict(fes.H1(md, order=p), ghas)
    l_neg = ops.bitset_to_list(gn, m.ne)
    l_if = ops.bitset_to_list(gi, m.ne)
    t_neg, t_if = V.rows(l_neg), V.rows(l_if)
    m_neg = ops.elmat_std(dim, p, l_neg, md.elements, md.points, gphi, -1, TERMS_VOL)
    off, q, w = ops.cutquad(dim, l_if, md.elements, md.points, gphi, NEG, 2 * p)
    m_if = ops.elmat_rule(dim, p, l_if, md.elements, md.points, gphi, off, q, w, TERMS_VOL)
    # ghost penalty facets between HASNEG and IF
    _, _, fe = md.facets()
    gf = ops.facets_with_neighbor_types(fe, ghas, gi, False, False, True)
    l_f = ops.bitset_to_list(gf, m.nfacet)
    m_f = ops.facetpatch(dim, p, l_f, fe, md.elements, md.points, -1, 0.1, -2)
    fe_l = fe[l_f.long()]
    t_f = torch.cat([V.rows(fe_l[:, 0].contiguous()), V.rows(fe_l[:, 1].contiguous())], dim=1).contiguous()
    tables = [t_neg, t_if, t_f]
    mats = [m_neg, m_if, m_f]
    plan = ops.csr_plan(V.ndof, tables)
    rowptr, col = ops.csr_pattern(plan, DEV)
    vals = ops.csr_scatter(plan, list(zip(tables, mats)), rowptr, col)
    vals2 = ops.csr_scatter(plan, list(zip(tables, mats)), rowptr, col)
    assert torch.equal(vals, vals2)  # bit-reproducible
    # oracle: pad the narrower tables to the facet width with -1
    W = t_f.shape[1]
    def pad(t, mm):
        tn = np.full((t.shape[0], W), -1, np.int32); tn[:, :t.shape[1]] = t.cpu().numpy()
        mn = np.zeros((t.shape[0], W, W)); mn[:, :t.shape[1], :t.shape[1]] = mm.cpu().numpy()
        return tn, mn
    tabs, mts = zip(*[pad(t, mm) for t, mm in zip(tables, mats)])
    tab = np.concatenate(tabs); mt = np.concatenate(mts)
    orow, ocol = xo.graph(V.ndof, tab)
    assert np.array_equal(rowptr.cpu().numpy(), orow)
    assert np.array_equal(col.cpu().numpy(), ocol)
    oval = xo.scatter(V.ndof, tab, mt, orow, ocol)
    gv = vals.cpu().numpy()
    assert np.linalg.norm(gv - oval) <= 1e-12 * np.linalg.norm(oval)
    # SpMV against scipy
    import scipy.sparse as sp
    A = sp.csr_matrix((oval, ocol, orow), shape=(V.ndof, V.ndof))
    x = np.random.default_rng(0).standard_normal(V.ndof)
    y = ops.csr_spmv(rowptr, col, vals, torch.from_numpy(x).to(DEV)).cpu().numpy()
    assert np.allclose(y, A @ x, rtol=1e-11, atol=1e-12 * np.abs(A @ x).max())


@pytest.mark.parametrize("dim,n", [(2, 9), (3, 5)])
@pytest.mark.parametrize("p", [2, 3, 4])
def test_elmat_tensor_tiles_match_cuda_core_path_and_oracle(dim, n, p):
    """north_star (c): the FP64 tensor-core (DMMA) contraction against the CUDA-core kernel and the oracle,
    cut elements (per-element rules: NEG part + IF mass accumulated) and uncut elements (shared rule)"""
    from ngsxfem_b200._lib import default_ctx
    ctx = default_ctx()
    m = make_mesh(dim, n)
    phi = level_set(m, "sphere", 30 + dim)
    pts, els = mesh_np(m)
    md = to_dev(m)
    gphi = torch.from_numpy(phi).to(DEV)
    _, _, gi, _ = ops.classify(dim, md.elements, gphi)
    lst = ops.bitset_to_list(gi, m.ne)
    lnp = lst.cpu().numpy()
    off, q, w = ops.cutquad(dim, lst, md.elements, md.points, gphi, NEG, 2 * p)
    off2, q2, w2 = ops.cutquad(dim, lst, md.elements, md.points, gphi, IF, 2 * p)
    res = {}
    try:
        for mode in (-1, 1):
            ctx.set_option("elmat_dmma", mode)
            g = ops.elmat_rule(dim, p, lst, md.elements, md.points, gphi, off, q, w, TERMS_VOL)
            ops.elmat_rule(dim, p, lst, md.elements, md.points, gphi, off2, q2, w2, [(0, 0, 3.0, -1)], out=g, accumulate=True)
            u = ops.elmat_std(dim, 3 if p == 2 else p, None, md.elements, md.points, gphi, 2 * p, TERMS_VOL)
            res[mode] = (g.cpu().numpy(), u.cpu().numpy())
    finally:
        ctx.set_option("elmat_dmma", 0)
    o = xo.batch_cut_elmats(dim, p, lnp, els, pts, phi, NEG, 2 * p, TERMS_VOL) + \
        xo.batch_cut_elmats(dim, p, lnp, els, pts, phi, IF, 2 * p, [(0, 0, 3.0, -1)])
    assert rel_err(res[1][0], o) < 1e-12
    assert rel_err(res[1][0], res[-1][0]) < 1e-12
    assert rel_err(res[1][1], res[-1][1]) < 1e-12
    # the tensor path computes the tiles ti <= tj and mirrors the off-diagonal ones
    assert rel_err(res[1][0], res[1][0].transpose(0, 2, 1)) < 1e-14


@pytest.mark.parametrize("dim,n,p", [(3, 5, 2), (3, 4, 1), (2, 9, 2), (3, 3, 3), (3, 3, 4)])
def test_elmat_rule_pair_equals_two_accumulating_passes(dim, n, p):
    """xgpu_elmat_rule_pair (NEG volume terms + IF interface term in one pass) against the oracle and against the two
    separate accumulating calls (same sums, different association: 1e-14)"""
    m = (meshes.MakeStructured3DMesh(nx=n) if dim == 3 else meshes.MakeStructured2DMesh(nx=n, ny=n)).to(DEV)
    c = (.47, .52, .5)
    phi = torch.sqrt(sum((m.points[:, i] - c[i]) ** 2 for i in range(dim))) - 0.34
    phi = torch.where(phi.abs() < 1e-14, torch.full_like(phi, 1e-14), phi).contiguous()
    bn, bp, bi, _ = ops.classify(dim, m.elements, phi)
    l_if = ops.bitset_to_list(bi, m.ne)
    TV = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 0.5, 0)]
    TI = [(ops.OP_ID, ops.OP_ID, 3.0, -1)]
    ra = ops.cutquad(dim, l_if, m.elements, m.points, phi, ops.NEG, 2 * p)
    rb = ops.cutquad(dim, l_if, m.elements, m.points, phi, ops.IF, 2 * p)
    ctx = ops.default_ctx()
    for mode in (0, -1):
        ctx.set_option("elmat_dmma", mode)
        try:
            pair = ops.elmat_rule_pair(dim, p, l_if, m.elements, m.points, phi, ra, TV, rb, TI)
            two = ops.elmat_rule(dim, p, l_if, m.elements, m.points, phi, ra[0], ra[1], ra[2], TV)
            ops.elmat_rule(dim, p, l_if, m.elements, m.points, phi, rb[0], rb[1], rb[2], TI, out=two, accumulate=True)
        finally:
            ctx.set_option("elmat_dmma", 0)
        assert torch.allclose(pair, two, rtol=1e-13, atol=1e-14 * float(two.abs().max()))
        assert torch.allclose(pair, pair.transpose(1, 2), rtol=1e-13, atol=1e-14 * float(two.abs().max()))
    pts, els = m.points.cpu().numpy(), m.elements.cpu().numpy()
    lo = l_if.cpu().numpy()
    mo = xo.batch_cut_elmats(dim, p, lo, els, pts, phi.cpu().numpy(), xo.NEG, 2 * p, TV)
    mo += xo.batch_cut_elmats(dim, p, lo, els, pts, phi.cpu().numpy(), xo.IF, 2 * p, TI)
    assert np.linalg.norm(pair.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)


@pytest.mark.parametrize("dim,n,p", [(3, 2, 4), (2, 4, 4), (3, 2, 3)])
def test_facetpatch_tensor_tiles_match_cuda_core_kernel_and_oracle(dim, n, p):
    """order >= 3: the facet-patch contraction on FP64 tensor tiles (4 patch points per DMMA k-step) against the
    CUDA-core rank-1 kernel (elmat_dmma = -1) and the oracle"""
    m = make_mesh(dim, n)
    pts, els = mesh_np(m)
    rng = np.random.default_rng(5)
    pts = pts + 0.1 / n * rng.random(pts.shape)
    m = meshes.Mesh(torch.from_numpy(pts), m.elements)
    md = to_dev(m)
    _, _, fe = md.facets()
    fe_np = fe.cpu().numpy()
    sel = np.nonzero(fe_np[:, 1] >= 0)[0].astype(np.int32)[:: 3]
    ctx = ops.default_ctx()
    out = {}
    for mode in (0, -1):
        ctx.set_option("elmat_dmma", mode)
        try:
            out[mode] = ops.facetpatch(dim, p, torch.from_numpy(sel).to(DEV), fe, md.elements, md.points, -1, 0.37, -2).cpu().numpy()
        finally:
            ctx.set_option("elmat_dmma", 0)
    o, nfail = xo.batch_facet_patch(dim, p, sel, fe_np, els, pts, -1, 0.37, -2)
    assert nfail == 0
    assert rel_err(out[0], o) < 1e-12 and rel_err(out[-1], o) < 1e-12
    assert rel_err(out[0], out[-1]) < 1e-13
    assert np.allclose(out[0], out[0].transpose(0, 2, 1), rtol=1e-13, atol=1e-15 * np.abs(o).max())


@pytest.mark.parametrize("dim,n,p", [(3, 6, 2), (3, 5, 1), (2, 11, 2), (2, 9, 1), (1, 9, 2)])
def test_elmat_lane_per_point_kernel_equals_staged_kernel_and_oracle(dim, n, p):
    """k_elmat_lp (order <= 2, one lane per quadrature point, matrix in registers, elmat_lp.cu) against the staged half-warp
    kernel k_elmat_ms (option "elmat_lp" = 0) and the oracle: single rule, rule pair, accumulate, symmetric, reproducible"""
    if dim == 3:
        m = meshes.MakeStructured3DMesh(nx=n).to(DEV)
    elif dim == 2:
        m = meshes.MakeStructured2DMesh(nx=n, ny=n).to(DEV)
    else:
        m = make_mesh(1, n); m = to_dev(m)
    c = (.47, .52, .5)
    phi = torch.sqrt(sum((m.points[:, i] - c[i]) ** 2 for i in range(dim))) - 0.34
    phi = torch.where(phi.abs() < 1e-14, torch.full_like(phi, 1e-14), phi).contiguous()
    ne = m.elements.shape[0]
    bn, bp, bi, _ = ops.classify(dim, m.elements, phi)
    l_if = ops.bitset_to_list(bi, ne)
    assert l_if.numel() > 0
    TV = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 0.5, 0)]
    TI = [(ops.OP_ID, ops.OP_ID, 3.0, -1)]
    ra = ops.cutquad(dim, l_if, m.elements, m.points, phi, ops.NEG, 2 * p)
    rb = ops.cutquad(dim, l_if, m.elements, m.points, phi, ops.IF, 2 * p)
    ctx = ops.default_ctx()
    res = {}
    for mode in (1, 0):
        ctx.set_option("elmat_lp", mode)
        try:
            pair = ops.elmat_rule_pair(dim, p, l_if, m.elements, m.points, phi, ra, TV, rb, TI)
            pair2 = ops.elmat_rule_pair(dim, p, l_if, m.elements, m.points, phi, ra, TV, rb, TI)
            one = ops.elmat_rule(dim, p, l_if, m.elements, m.points, phi, ra[0], ra[1], ra[2], TV)
            acc = one.clone()
            ops.elmat_rule(dim, p, l_if, m.elements, m.points, phi, rb[0], rb[1], rb[2], TI, out=acc, accumulate=True)
        finally:
            ctx.set_option("elmat_lp", 1)
        assert torch.equal(pair, pair2)
        res[mode] = (pair, one, acc)
    scale = float(res[0][0].abs().max())
    for a, b in zip(res[1], res[0]):
        assert torch.allclose(a, b, rtol=1e-12, atol=1e-14 * scale)
    assert torch.allclose(res[1][0], res[1][2], rtol=1e-12, atol=1e-14 * scale)
    assert torch.equal(res[1][0], res[1][0].transpose(1, 2))
    pts, els = m.points.cpu().numpy(), m.elements.cpu().numpy()
    lo = l_if.cpu().numpy()
    mo = xo.batch_cut_elmats(dim, p, lo, els, pts, phi.cpu().numpy(), xo.NEG, 2 * p, TV)
    mo += xo.batch_cut_elmats(dim, p, lo, els, pts, phi.cpu().numpy(), xo.IF, 2 * p, TI)
    assert np.linalg.norm(res[1][0].cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)


@pytest.mark.parametrize("dim,n,p", [(3, 6, 2), (3, 5, 1), (2, 11, 2), (2, 9, 1), (1, 9, 2), (1, 7, 1)])
def test_elmat_cut_pair_rules_on_the_fly_equal_stored_rules_and_oracle(dim, n, p):
    """xgpu_elmat_cut_pair (the cut rule of an element is made inside the element-matrix kernel, as T_CalcElementMatrixAdd does,
    symboliccutbfi.cpp:105-121) against xgpu_cutquad_* + xgpu_elmat_rule_pair and against the oracle — on a list that also
    holds UNCUT elements (HASNEG: standard rule for the NEG integrator, nothing for the IF one)"""
    if dim == 3:
        m = meshes.MakeStructured3DMesh(nx=n).to(DEV)
    elif dim == 2:
        m = meshes.MakeStructured2DMesh(nx=n, ny=n).to(DEV)
    else:
        m = to_dev(make_mesh(1, n))
    c = (.47, .52, .5)
    phi = torch.sqrt(sum((m.points[:, i] - c[i]) ** 2 for i in range(dim))) - 0.34
    phi = torch.where(phi.abs() < 1e-14, torch.full_like(phi, 1e-14), phi).contiguous()
    ne = m.elements.shape[0]
    bn, bp, bi, _ = ops.classify(dim, m.elements, phi)
    lst = ops.bitset_to_list(ops.bitset_op(ops.BIT_OR, bn, bi, ne), ne)
    l_if = ops.bitset_to_list(bi, ne)
    assert lst.numel() > l_if.numel() > 0
    TV = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 0.5, 0)]
    TI = [(ops.OP_ID, ops.OP_ID, 3.0, -1)]
    order = 2 * p
    ra = ops.cutquad(dim, lst, m.elements, m.points, phi, ops.NEG, order)
    rb = ops.cutquad(dim, lst, m.elements, m.points, phi, ops.IF, order)
    stored = ops.elmat_rule_pair(dim, p, lst, m.elements, m.points, phi, ra, TV, rb, TI)
    fly = ops.elmat_cut_pair(dim, p, lst, m.elements, m.points, phi, ops.NEG, order, TV, ops.IF, order, TI)
    fly2 = ops.elmat_cut_pair(dim, p, lst, m.elements, m.points, phi, ops.NEG, order, TV, ops.IF, order, TI)
    assert torch.equal(fly, fly2)
    scale = float(stored.abs().max())
    assert torch.allclose(fly, stored, rtol=1e-12, atol=1e-14 * scale)
    assert torch.equal(fly, fly.transpose(1, 2))
    # one integrator, then the second one accumulated = the pair
    one = ops.elmat_cut_pair(dim, p, lst, m.elements, m.points, phi, ops.NEG, order, TV)
    ops.elmat_cut_pair(dim, p, lst, m.elements, m.points, phi, ops.IF, order, TI, out=one, accumulate=True)
    assert torch.allclose(one, fly, rtol=1e-12, atol=1e-14 * scale)
    # POS side of the cut elements
    rp = ops.cutquad(dim, l_if, m.elements, m.points, phi, ops.POS, order)
    sp = ops.elmat_rule(dim, p, l_if, m.elements, m.points, phi, rp[0], rp[1], rp[2], TV)
    fp = ops.elmat_cut_pair(dim, p, l_if, m.elements, m.points, phi, ops.POS, order, TV)
    assert torch.allclose(fp, sp, rtol=1e-12, atol=1e-14 * scale)
    pts, els = m.points.cpu().numpy(), m.elements.cpu().numpy()
    lo = lst.cpu().numpy()
    mo = xo.batch_cut_elmats(dim, p, lo, els, pts, phi.cpu().numpy(), xo.NEG, order, TV)
    mo += xo.batch_cut_elmats(dim, p, lo, els, pts, phi.cpu().numpy(), xo.IF, order, TI)
    assert np.linalg.norm(fly.cpu().numpy() - mo) <= 1e-12 * np.linalg.norm(mo)
    # outside the menu: the caller is told to take the two-step path
    with pytest.raises(Exception):
        ops.elmat_cut_pair(dim, 3, lst, m.elements, m.points, phi, ops.NEG, 6, TV)
    with pytest.raises(Exception):
        ops.elmat_cut_pair(dim, p, l_if, m.elements, m.points, phi, ops.IF, order, TERMS_NITSCHE)


# ------------------------------------------------------------------ (d) facet patch, dFacetPatch(downscale=)
@pytest.mark.parametrize("dim,n,p", [(1, 7, 2), (2, 5, 1), (2, 5, 2), (2, 4, 3), (2, 3, 4), (3, 3, 1), (3, 3, 2), (3, 2, 3), (3, 2, 4)])
@pytest.mark.parametrize("s", [0.5, 0.13])
def test_facetpatch_downscale_matches_oracle(dim, n, p, s):
    """xgpu_facetpatch_set_downscale (SetIRScaling; ScaleAndMapIntegrationRule, symboliccutbfi.cpp:1849-1885) on the CUDA-core
    (p <= 2) and the tensor-tile (p >= 3) kernels; the state is reset after the call"""
    m = make_mesh(dim, n)
    pts, els = mesh_np(m)
    rng = np.random.default_rng(5)
    pts = pts + 0.1 / n * rng.random(pts.shape)
    m = meshes.Mesh(torch.from_numpy(pts), m.elements)
    md = to_dev(m)
    _, _, fe = md.facets()
    fe_np = fe.cpu().numpy()
    inner = np.nonzero(fe_np[:, 1] >= 0)[0].astype(np.int32)
    sel = inner[:: max(1, len(inner) // 40)]
    gsel = torch.from_numpy(sel).to(DEV)
    g1 = ops.facetpatch(dim, p, gsel, fe, md.elements, md.points, -1, 0.37, -2)
    g = ops.facetpatch(dim, p, gsel, fe, md.elements, md.points, -1, 0.37, -2, downscale=s)
    o, nfail = xo.batch_facet_patch(dim, p, sel, fe_np, els, pts, -1, 0.37, -2, downscale=s)
    assert nfail == 0
    assert rel_err(g.cpu().numpy(), o) < 1e-12
    assert rel_err(g.cpu().numpy(), g1.cpu().numpy()) > 1e-3           # it is another matrix
    # u = 1 on element 1, 0 on element 2: coef h^-2 s^2 (|T1| + |T2|) — closed form on the device result
    gm = g.cpu().numpy()
    nd = gm.shape[1] // 2
    one = np.zeros(2 * nd); one[:dim + 1] = 1.0
    import math
    for k, f in enumerate(sel):
        X1, X2 = pts[els[fe_np[f, 0]]], pts[els[fe_np[f, 1]]]
        v1 = abs(np.linalg.det(X1[:dim] - X1[dim])) / math.factorial(dim)
        v2 = abs(np.linalg.det(X2[:dim] - X2[dim])) / math.factorial(dim)
        h1 = (v1 * math.factorial(dim)) ** (1.0 / dim)                    # specialcf.mesh_size of element 1: |det F|^(1/D)
        want = 0.37 * h1 ** -2 * s * s * (v1 + v2)
        assert abs(one @ gm[k] @ one - want) < 1e-11 * want
    # the state does not leak: the next plain call is the unscaled matrix, bit for bit
    g2 = ops.facetpatch(dim, p, gsel, fe, md.elements, md.points, -1, 0.37, -2)
    assert torch.equal(g1, g2)
    with pytest.raises(XgpuError):
        ops.facetpatch(dim, p, gsel, fe, md.elements, md.points, -1, 0.37, -2, downscale=1.5)
