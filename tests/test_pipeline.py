"""End-to-end parity of the config-2 hot path (CutInfo -> Restrict -> cut quadrature -> element
matrices -> graph -> CSR) through the C ABI against the CPU oracle pipeline on the same inputs:
bit-exact BitArrays and sparsity pattern, cut ratios 1e-13, assembled values 1e-12 (Frobenius)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from ngsxfem_b200.pipeline import CutAssembly, HostBuffers, drain_host, sphere_levelset, step_host
from oracle.pipeline_ref import oracle_step

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bits_np(t):
    return t.cpu().numpy().view(np.uint8)


def check(r, o, ne):
    nb = (ne + 7) // 8
    for g, c in zip(r["bits"], o["bits"]):
        assert np.array_equal(bits_np(g)[:nb], c)
    assert r["ndof"] == o["ndof"] and r["nnz"] == o["nnz"] and r["units"] == o["units"]
    assert np.allclose(r["cut_ratio"].cpu().numpy(), o["cut_ratio"], rtol=1e-13, atol=1e-15)
    assert np.array_equal(r["rowptr"].cpu().numpy(), o["rowptr"])
    assert np.array_equal(r["col"].cpu().numpy(), o["col"])
    gv = r["vals"].cpu().numpy() if torch.is_tensor(r["vals"]) else r["vals"]
    assert np.linalg.norm(gv - o["vals"]) <= 1e-12 * np.linalg.norm(o["vals"])


@pytest.mark.parametrize("n,p,center,radius", [(6, 2, (.5, .5, .5), 0.4), (5, 1, (.45, .55, .5), 0.33),
                                               (4, 3, (.5, .5, .5), 0.4), (9, 2, (.52, .47, .5), 0.31)])
def test_sphere_cut_assembly_matches_oracle(n, p, center, radius):
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    V0 = xf.H1(mesh, order=p)
    phi = sphere_levelset(mesh.points, [center], radius)
    prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, p)
    r = prob.step(phi)
    o, _ = oracle_step(3, mesh.elements.cpu().numpy(), mesh.points.cpu().numpy(), V0.el_dofs.cpu().numpy(), V0.ndof, p,
                       phi.cpu().numpy())
    check(r, o, mesh.ne)
    # run to run bit reproducibility of the whole pass
    r2 = prob.step(phi)
    assert torch.equal(r["vals"], r2["vals"]) and torch.equal(r["col"], r2["col"])


def test_fused_assemble_equals_pattern_plus_scatter():
    """xgpu_csr_assemble (compaction fused into the scatter) against xgpu_csr_pattern + xgpu_csr_scatter: bit identical"""
    mesh = xf.MakeStructured3DMesh(nx=8, device=DEV)
    V0 = xf.H1(mesh, order=2)
    phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
    r = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2).step(phi, keep=True, mode="twopass")
    batches = [(r["t_neg"], r["m_neg"]), (r["t_if"], r["m_if"])]
    plan = ops.csr_plan(r["ndof"], [b[0] for b in batches])
    rowptr, col = ops.csr_pattern(plan, mesh.device)
    vals = ops.csr_scatter(plan, batches, rowptr, col)
    assert torch.equal(rowptr, r["rowptr"]) and torch.equal(col, r["col"]) and torch.equal(vals, r["vals"])
    # a second plan takes over the shared pattern scratch: the first one must redo its pass transparently
    plan2 = ops.csr_plan(r["ndof"], [batches[0][0]])
    rp2, c2, v2 = ops.csr_assemble(plan, batches)
    assert torch.equal(rp2, rowptr) and torch.equal(c2, col) and torch.equal(v2, vals)
    del plan2


@pytest.mark.parametrize("n,p,radius", [(8, 2, 0.4), (7, 1, 0.33), (5, 3, 0.4), (13, 2, 0.27)])
@pytest.mark.parametrize("concurrent", [False, True])
def test_one_row_pass_equals_the_two_pass_path_bit_for_bit(n, p, radius, concurrent):
    """mode="fused" (procedural affine batch + xgpu_csr_assemble_fused: pattern, values and row offsets in one pass over the
    rows) against mode="twopass" (all element matrices in memory, xgpu_csr_plan + xgpu_csr_assemble): identical bits"""
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    V0 = xf.H1(mesh, order=p)
    phi = sphere_levelset(mesh.points, [(.5, .48, .52)], radius)
    prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, p)
    for c in (prob.ctx, prob.ctx_hp):
        c.set_option("csr_closed_form", 0)   # rows of the procedural batch from the reference tensors: the two-pass FMA chain
    try:
        a = prob.step(phi, mode="fused", concurrent=concurrent)
    finally:
        for c in (prob.ctx, prob.ctx_hp):
            c.set_option("csr_closed_form", 1)
    b = prob.step(phi, mode="twopass")
    assert a["nnz"] == b["nnz"] and a["ndof"] == b["ndof"]
    assert torch.equal(a["rowptr"], b["rowptr"]) and torch.equal(a["col"], b["col"])
    assert torch.equal(a["vals"], b["vals"])
    # default configuration: P2-tet rows in closed form — equal up to rounding, bit-reproducible run to run
    a1 = prob.step(phi, mode="fused", concurrent=concurrent)
    assert torch.equal(a1["rowptr"], b["rowptr"]) and torch.equal(a1["col"], b["col"])
    assert (a1["vals"] - b["vals"]).abs().max().item() <= 1e-13 * b["vals"].abs().max().item()
    a2 = prob.step(phi, mode="fused", concurrent=concurrent)
    assert torch.equal(a1["vals"], a2["vals"]) and torch.equal(a1["col"], a2["col"])


def test_cut_rules_on_the_fly_equal_stored_rules_in_the_step():
    """step(cut_rules="fly") (xgpu_elmat_cut_pair) against step(cut_rules="stored") (xgpu_cutquad_* + xgpu_elmat_rule_pair)"""
    mesh = xf.MakeStructured3DMesh(nx=9, device=DEV)
    V0 = xf.H1(mesh, order=2)
    phi = sphere_levelset(mesh.points, [(.5, .48, .52)], 0.37)
    prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2)
    a = prob.step(phi)
    b = prob.step(phi, cut_rules="stored")
    assert a["npts"] == (None, None) and b["npts"] == prob.cut_point_counts(phi)
    assert torch.equal(a["rowptr"], b["rowptr"]) and torch.equal(a["col"], b["col"])
    assert (a["vals"] - b["vals"]).abs().max().item() <= 1e-13 * b["vals"].abs().max().item()


def test_host_buffer_path_matches_device_path():
    mesh = xf.MakeStructured3DMesh(nx=7, device=DEV)
    V0 = xf.H1(mesh, order=2)
    phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
    r = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2).step(phi)
    hb = HostBuffers(mesh.elements, mesh.points, V0.el_dofs, phi)
    rh, h2d, d2h = step_host(hb, 3, V0.ndof, 2, torch.device(DEV, 0))
    assert h2d > 0 and d2h > 0
    assert torch.equal(hb.out["vals"][:r["nnz"]], r["vals"].cpu())
    assert torch.equal(hb.out["col"][:r["nnz"]], r["col"].cpu())


def test_pipelined_host_steps_land_the_same_results():
    """software-pipelined e2e steps (bench.py e2e leg): step k+1 starts while step k's results are still leaving; after the
    drain both alternating landing sets hold the values of the plain step, bit for bit"""
    n = 12
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    V0 = xf.H1(mesh, order=2)
    dev = torch.device(DEV, 0)
    phis = [sphere_levelset(mesh.points, [(.5, .5, .5)], rr) for rr in (0.4, 0.33, 0.4)]
    refs = [CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2).step(ph) for ph in phis]
    hb = HostBuffers(mesh.elements, mesh.points, V0.el_dofs, phis[0])
    for k, ph in enumerate(phis):
        hb.phi = ph.cpu().pin_memory()
        step_host(hb, 3, V0.ndof, 2, dev, pipelined=True)
        if k >= 1:   # the landing set of the previous step is complete once its event has fired
            hb.done[1 - hb.parity].synchronize()
            r = refs[k - 1]
            assert torch.equal(hb.out[f"vals#{1 - hb.parity}"][:r["nnz"]], r["vals"].cpu())
            assert torch.equal(hb.out[f"col#{1 - hb.parity}"][:r["nnz"]], r["col"].cpu())
    drain_host(hb, dev)
    torch.cuda.current_stream(dev).synchronize()
    r = refs[-1]
    assert torch.equal(hb.out[f"vals#{hb.parity}"][:r["nnz"]], r["vals"].cpu())
    assert torch.equal(hb.out[f"rowptr#{hb.parity}"][:r["ndof"] + 1], r["rowptr"].cpu())
    assert torch.equal(hb.out[f"cut_ratio#{hb.parity}"][:mesh.ne], r["cut_ratio"].cpu())


def test_full_size_matches_oracle():
    """BASELINE config 2 at FULL size (12 582 912 tets) against the threaded CPU oracle on the same inputs: BitArrays, counts
    and the CSR pattern (134.7 M column indices) exact, cut ratios 1e-13, assembled values 1e-12 (Frobenius)"""
    n = 128
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    V0 = xf.H1(mesh, order=2)
    phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
    r = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2).step(phi)
    o, _ = oracle_step(3, mesh.elements.cpu().numpy(), mesh.points.cpu().numpy(), V0.el_dofs.cpu().numpy(), V0.ndof, 2,
                       phi.cpu().numpy(), threads=True)
    assert (r["n_neg"], r["n_if"], r["nnz"]) == (3263736, 226020, o["nnz"])
    check(r, o, mesh.ne)


def test_full_size_invariants():
    """config-2 size: size independent properties —
    sum of all entries of the mass part = |Omega^-_h| (partition of unity), element counts of the
    Kuhn mesh / sphere cut, row sums of the stiffness part vanish."""
    n = 128
    mesh = xf.MakeStructured3DMesh(nx=n, device=DEV)
    V0 = xf.H1(mesh, order=2)
    phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
    prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2)
    r = prob.step(phi, keep=True)
    assert (r["n_neg"], r["n_if"]) == (3263736, 226020)
    bn, bp, bi = r["bits"]
    assert ops.bitset_count(bp, mesh.ne) == 9093156
    # A = K + M + M_Gamma ; 1^T A 1 = |Omega^-_h| + |Gamma_h| since K 1 = 0 and sum of shapes = 1 is
    # NOT true for the hierarchical basis: use the vertex-hat coefficient vector of the constant 1
    one = torch.zeros(r["ndof"], dtype=torch.float64, device=DEV)
    tn = torch.cat([r["t_neg"][:, :4].reshape(-1), r["t_if"][:, :4].reshape(-1)]).long()
    one[tn] = 1.0  # vertex dofs = 1, edge dofs = 0 represents u = 1 in the hierarchical P2 basis
    y = ops.csr_spmv(r["rowptr"], r["col"], r["vals"], one)
    total = float((one * y).sum())
    vol = 4.0 / 3.0 * np.pi * 0.4 ** 3
    area = 4.0 * np.pi * 0.4 ** 2
    # P1 geometry error is O(h^2): |Omega_h| and |Gamma_h| are within 1e-3 relative at h = 1/128
    assert abs(total - (vol + area)) < 2e-3 * (vol + area)
    # exact identity on the discrete geometry: the same quantity from the cut rules directly
    off, q, w = ops.cutquad(3, r["l_if"], mesh.elements, mesh.points, phi, ops.NEG, 4)
    v_cut = ops.integrate_monomial(3, r["l_if"], mesh.elements, mesh.points, off, q, w, (0, 0, 0))
    off2, q2, w2 = ops.cutquad(3, r["l_if"], mesh.elements, mesh.points, phi, ops.IF, 4)
    a_cut = ops.integrate_monomial(3, r["l_if"], mesh.elements, mesh.points, off2, q2, w2, (0, 0, 0))
    v_unc = r["n_neg"] / (6.0 * n ** 3)
    assert abs(total - (v_cut + v_unc + a_cut)) < 1e-11 * total
