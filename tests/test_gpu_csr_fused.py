"""xgpu_csr_assemble_fused (one pass over the rows: pattern + values + row offsets by decoupled look-back, procedural affine
batches) against the two-pass path (xgpu_csr_plan + xgpu_csr_assemble) and a dense numpy reference, through the C ABI.
Bit-exact pattern; values bit-identical to the two-pass path (same summation order)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from ngsxfem_b200._lib import XgpuError

pytestmark = pytest.mark.gpu
DEV = "cuda"


def random_tables(rng, ndof, shapes, holes=0.0):
    """table rows with distinct dofs per row (like element dof rows), optional -1 entries"""
    out = []
    for nrows, w in shapes:
        t = np.empty((nrows, w), dtype=np.int32)
        for r in range(nrows):
            # clustered like a mesh: dofs near a random centre
            c = rng.integers(0, ndof)
            pool = np.unique(np.clip(c + rng.integers(-3 * w, 3 * w + 1, size=6 * w), 0, ndof - 1))
            while len(pool) < w:
                pool = np.unique(np.concatenate([pool, rng.integers(0, ndof, size=w)]))
            t[r] = rng.permutation(pool)[:w]
        if holes > 0:
            t[rng.random(t.shape) < holes] = -1
        out.append(t)
    return out


def dense_reference(ndof, tables, mats):
    A = np.zeros((ndof, ndof))
    P = np.zeros((ndof, ndof), dtype=bool)
    for t, m in zip(tables, mats):
        for r in range(t.shape[0]):
            idx = t[r]
            ok = idx >= 0
            P[np.ix_(idx[ok], idx[ok])] = True
            if m is not None:
                A[np.ix_(idx[ok], idx[ok])] += m[r][np.ix_(ok, ok)]
    return A, P


@pytest.mark.parametrize("ndof,shapes,holes", [
    (300, [(400, 10)], 0.0),
    (500, [(300, 4), (100, 4)], 0.0),
    (200, [(150, 10), (60, 10), (20, 10)], 0.1),
    (1000, [(700, 20)], 0.0),
    (400, [(200, 35)], 0.05),      # two trips (rows wider than a warp)
    (64, [(40, 3)], 0.0),
    (50, [(30, 1)], 0.0),
    (700, [(500, 12), (100, 6)], 0.0),  # mixed widths
])
def test_fused_rows_match_two_pass_and_dense(ndof, shapes, holes):
    rng = np.random.default_rng(hash((ndof, len(shapes))) % 2 ** 31)
    tabs = random_tables(rng, ndof, shapes, holes)
    mats = [rng.standard_normal((t.shape[0], t.shape[1], t.shape[1])) for t in tabs]
    if len(tabs) >= 3:
        mats[2] = None  # sparsity-only batch
    gt = [torch.from_numpy(t).to(DEV) for t in tabs]
    gm = [None if m is None else torch.from_numpy(m).to(DEV) for m in mats]
    batches = list(zip(gt, gm))
    plan2 = ops.csr_plan(ndof, gt)
    rp2, c2, v2 = ops.csr_assemble(plan2, batches)
    plan = ops.csr_plan_rows(ndof, gt)
    assert plan.nnz == plan2.nnz
    rp, c, v = ops.csr_assemble_fused(plan, batches)
    assert torch.equal(rp, rp2) and torch.equal(c, c2)
    if max(w for _, w in shapes) <= 32:
        assert torch.equal(v, v2)   # same summation order
    else:  # rows wider than a warp: the two-pass kernel sums trip-major inside a batch of entries, the one-pass kernel in strict entry order
        assert torch.allclose(v, v2, rtol=1e-13, atol=1e-13)
    A, P = dense_reference(ndof, tabs, mats)
    rpn, cn, vn = rp.cpu().numpy(), c.cpu().numpy(), v.cpu().numpy()
    assert rpn[-1] == P.sum() == len(cn)
    for d in range(ndof):
        cols = cn[rpn[d]:rpn[d + 1]]
        assert np.array_equal(cols, np.flatnonzero(P[d]))
        assert np.allclose(vn[rpn[d]:rpn[d + 1]], A[d, cols], rtol=1e-13, atol=1e-13)
    # the same call again: bit reproducible (the summation order is fixed)
    rp3, c3, v3 = ops.csr_assemble_fused(plan, batches)
    assert torch.equal(rp3, rp) and torch.equal(c3, c) and torch.equal(v3, v)


def test_unsupported_shapes_are_declined_not_emulated():
    # a dof twice in one table row (facet-patch rows): the one-pass kernels decline, the two-pass path takes it
    t = torch.tensor([[0, 1, 2, 1], [2, 3, 4, 5]], dtype=torch.int32, device=DEV)
    with pytest.raises(XgpuError) as ei:
        ops.csr_plan_rows(6, [t])
    assert ei.value.code == ops.XGPU_ERR_UNSUPPORTED
    # a plan without row counts is rejected by the one-pass assembly
    t2 = torch.tensor([[0, 1, 2, 3]], dtype=torch.int32, device=DEV)
    with pytest.raises(XgpuError):
        ops.csr_assemble_fused(ops.adjacency_plan(6, [t2]), [(t2, torch.ones(1, 4, 4, dtype=torch.float64, device=DEV))])


def test_empty_rows_and_empty_matrix():
    t = torch.tensor([[5, 7, 9]], dtype=torch.int32, device=DEV)
    m = torch.arange(9, dtype=torch.float64, device=DEV).reshape(1, 3, 3)
    plan = ops.csr_plan_rows(12, [t])
    rp, c, v = ops.csr_assemble_fused(plan, [(t, m)])
    assert rp.tolist() == [0] * 6 + [3, 3, 6, 6, 9, 9, 9]
    assert c.tolist() == [5, 7, 9] * 3 and v.tolist() == list(range(9))
    t0 = torch.empty((0, 3), dtype=torch.int32, device=DEV)
    plan0 = ops.csr_plan_rows(4, [t0])
    rp0, c0, v0 = ops.csr_assemble_fused(plan0, [(t0, torch.empty((0, 3, 3), dtype=torch.float64, device=DEV))])
    assert rp0.tolist() == [0] * 5 and c0.numel() == 0


@pytest.mark.parametrize("dim,p", [(3, 2), (3, 1), (2, 2), (2, 1), (1, 2)])
def test_procedural_affine_batch_equals_matrices_in_memory(dim, p):
    """a batch given by xgpu_affine_coefs (64 B per element) assembles to the same bits as the same elements'
    matrices from xgpu_elmat_std's affine path"""
    n = {1: 40, 2: 12, 3: 5}[dim]
    mesh = {1: lambda: xf.MakeStructured1DMesh(nx=n, device=DEV),
            2: lambda: xf.MakeStructured2DMesh(nx=n, ny=n, device=DEV),
            3: lambda: xf.MakeStructured3DMesh(nx=n, device=DEV)}[dim]()
    V = xf.H1(mesh, order=p)
    terms = [(ops.OP_GRAD, ops.OP_GRAD, 0.7, 0), (ops.OP_ID, ops.OP_ID, 1.3, -1)]
    elems = torch.arange(0, mesh.ne, 2, dtype=torch.int32, device=DEV)
    table = ops.gather_rows(V.el_dofs, elems, None)
    mats = ops.elmat_std(dim, p, elems, mesh.elements, mesh.points, None, 2 * p, terms)
    plan = ops.csr_plan_rows(V.ndof, [table])
    rp, c, v = ops.csr_assemble_fused(plan, [(table, mats)])
    proc = ops.affine_coefs(dim, p, elems, mesh.elements, mesh.points, 2 * p, terms)
    ctx = ops.default_ctx()
    ctx.set_option("csr_closed_form", 0)   # rows from the reference tensors: the FMA chain of xgpu_elmat_std's affine path
    try:
        rp2, c2, v2 = ops.csr_assemble_fused(plan, [(table, proc)])
    finally:
        ctx.set_option("csr_closed_form", 1)
    assert torch.equal(rp, rp2) and torch.equal(c, c2) and torch.equal(v, v2)
    # default: P2-tet rows of the thread-per-row kernel in closed form (barycentric Gram matrix) — equal up to rounding
    rp3, c3, v3 = ops.csr_assemble_fused(plan, [(table, proc)])
    assert torch.equal(rp, rp3) and torch.equal(c, c3)
    assert (v3 - v).abs().max().item() <= 1e-14 * v.abs().max().item()


@pytest.mark.parametrize("n,p", [(6, 2), (5, 1), (4, 3)])
def test_restrict_plan_equals_the_separate_calls(n, p):
    """xgpu_restrict_plan (one atomic pass + one table pass) against xgpu_dofs_of_elements + xgpu_bitset_rank +
    xgpu_gather_rows + xgpu_csr_plan_rows: same compressed numbering, same dof rows, same CSR"""
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    V = xf.H1(mesh, order=p)
    rng = np.random.default_rng(n * 10 + p)
    sel = torch.from_numpy(rng.random(mesh.ne) < 0.4).to(DEV)
    l_all = torch.nonzero(sel).flatten().int()
    l_a, l_b = l_all[::2].contiguous(), l_all[1::2].contiguous()
    bits = ops.new_bits(mesh.ne, DEV)
    words = torch.zeros(ops.nwords(mesh.ne) * 32, dtype=torch.int64, device=DEV)
    words[l_all.long()] = 1
    w = (words.view(-1, 32) << torch.arange(32, device=DEV)).sum(1)
    bits.copy_(torch.where(w >= 2 ** 31, w - 2 ** 32, w).int())
    act = ops.dofs_of_elements(V.el_dofs, bits, V.ndof)
    dofmap, ndof = ops.bitset_rank(act, V.ndof)
    t_a, t_b = ops.gather_rows(V.el_dofs, l_a, dofmap), ops.gather_rows(V.el_dofs, l_b, dofmap)
    dofmap2, ndof2, (u_a, u_b), plan = ops.restrict_plan(V.el_dofs, [l_a, l_b], V.ndof)
    assert ndof2 == ndof and torch.equal(dofmap2, dofmap)
    assert torch.equal(u_a, t_a) and torch.equal(u_b, t_b)
    nd = V.el_dofs.shape[1]
    m_a = torch.from_numpy(rng.standard_normal((l_a.numel(), nd, nd))).to(DEV)
    m_b = torch.from_numpy(rng.standard_normal((l_b.numel(), nd, nd))).to(DEV)
    ref = ops.csr_assemble_fused(ops.csr_plan_rows(ndof, [t_a, t_b]), [(t_a, m_a), (t_b, m_b)])
    ops.csr_count_rows(plan)
    got = ops.csr_assemble_fused(plan, [(u_a, m_a), (u_b, m_b)])
    for x, y in zip(ref, got):
        assert torch.equal(x, y)
    # an empty list is a batch without rows
    e = torch.empty(0, dtype=torch.int32, device=DEV)
    _, nd3, (v_a, v_e), plan3 = ops.restrict_plan(V.el_dofs, [l_a, e], V.ndof)
    assert v_e.shape == (0, nd) and nd3 <= ndof


@pytest.mark.parametrize("n,p,proc", [(6, 2, True), (6, 2, False), (7, 1, False)])
def test_row_ranges_assemble_to_the_same_bits(n, p, proc):
    """xgpu_csr_assemble_fused_rows: three row ranges finished in any order into the same arrays = one call over all rows
    (the element-partitioned assembly finishes the rows it sends first); an empty range and a range past the end are no-ops"""
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    V = xf.H1(mesh, order=p)
    rng = np.random.default_rng(n + p)
    # a ball of elements (the shape of HASNEG): every row goes through the thread-per-row / warp-per-row kernels.  A sparse
    # random subset has short rows with more than 32 columns, which belong to the lane-group kernel (next test)
    cen = mesh.points[mesh.elements.long()].mean(dim=1)
    elems = torch.nonzero(((cen - 0.5) ** 2).sum(dim=1) < 0.42 ** 2).reshape(-1).to(torch.int32)
    dofmap, ndof, (table,), plan = ops.restrict_plan(V.el_dofs, [elems], V.ndof)
    ops.csr_count_rows(plan)
    if proc:
        m = ops.affine_coefs(3, p, elems, mesh.elements, mesh.points, 2 * p, [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 1.0, 0)])
    else:
        nd = V.el_dofs.shape[1]
        m = torch.from_numpy(rng.standard_normal((elems.numel(), nd, nd))).to(DEV)
    batches = [(table, m)]
    rp, c, v = ops.csr_assemble_fused(plan, batches)
    a, b = ndof // 3 + 5, 2 * ndof // 3 - 7          # not multiples of 32
    out = ops.csr_assemble_fused(plan, batches, rows=(b, ndof))
    assert torch.equal(out[0], rp)
    lo, hi = int(rp[b]), int(rp[ndof])
    assert torch.equal(out[1][lo:hi], c[lo:hi]) and torch.equal(out[2][lo:hi], v[lo:hi])
    out[2][:lo] = float("nan")                         # rows outside the range must not be read or written
    ops.csr_assemble_fused(plan, batches, rows=(0, a), out=out)
    assert torch.isnan(out[2][int(rp[a]):lo]).all()
    ops.csr_assemble_fused(plan, batches, rows=(a, a), out=out)
    ops.csr_assemble_fused(plan, batches, rows=(a, b), out=out)
    ops.csr_assemble_fused(plan, batches, rows=(ndof, ndof + 100), out=out)
    assert torch.equal(out[1], c) and torch.equal(out[2], v)
    with pytest.raises(XgpuError):
        ops.csr_assemble_fused(plan, batches, rows=(5, 2), out=out)


def test_row_ranges_decline_plans_with_lane_group_rows():
    """rows outside the thread- / warp-per-row kernels cannot be split by range: XGPU_ERR_UNSUPPORTED (-3), nothing written,
    and the caller assembles all rows at once (distributed.py falls back to the exchange after the row pass)"""
    mesh = xf.MakeStructured3DMesh(nx=6, device=DEV)
    V = xf.H1(mesh, order=2)
    rng = np.random.default_rng(8)
    elems = torch.from_numpy(np.flatnonzero(rng.random(mesh.ne) < 0.7).astype(np.int32)).to(DEV)
    dofmap, ndof, (table,), plan = ops.restrict_plan(V.el_dofs, [elems], V.ndof)
    ops.csr_count_rows(plan)
    nd = V.el_dofs.shape[1]
    m = torch.from_numpy(rng.standard_normal((elems.numel(), nd, nd))).to(DEV)
    full = ops.csr_assemble_fused(plan, [(table, m)])
    try:
        part = ops.csr_assemble_fused(plan, [(table, m)], rows=(0, ndof // 2))
    except XgpuError as e:
        assert e.code == -3 and "lane-group" in str(e)
    else:   # the policy took every row on the fast kernels: then the halves must agree
        ops.csr_assemble_fused(plan, [(table, m)], rows=(ndof // 2, ndof), out=part)
        assert torch.equal(part[2], full[2])


@pytest.mark.parametrize("n,p", [(5, 2), (6, 1), (4, 3)])
def test_half_width_sorting_network_gives_the_same_records(n, p):
    """warp-per-row pattern pass: rows with at most 128 candidates sort 4 keys per lane ("csr_warp_half", default) — same
    row lengths, columns and values, bit for bit, as the 8-keys-per-lane network"""
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    V = xf.H1(mesh, order=p)
    rng = np.random.default_rng(n * 7 + p)
    cen = mesh.points[mesh.elements.long()].mean(dim=1)
    elems = torch.nonzero(((cen - 0.5) ** 2).sum(dim=1) < 0.45 ** 2).reshape(-1).to(torch.int32)
    nd = V.el_dofs.shape[1]
    m = torch.from_numpy(rng.standard_normal((elems.numel(), nd, nd))).to(DEV)
    ctx = ops.default_ctx()
    out = []
    for half in (1, 0):
        ctx.set_option("csr_warp_half", half)
        try:
            dofmap, ndof, (table,), plan = ops.restrict_plan(V.el_dofs, [elems], V.ndof)
            ops.csr_count_rows(plan)
            out.append(ops.csr_assemble_fused(plan, [(table, m)]))
        finally:
            ctx.set_option("csr_warp_half", 1)
    for x, y in zip(*out):
        assert torch.equal(x, y)
    ref = ops.csr_assemble(ops.csr_plan(ndof, [table]), [(table, m)])     # the two-pass path
    assert torch.equal(out[0][0], ref[0]) and torch.equal(out[0][1], ref[1]) and torch.equal(out[0][2], ref[2])
