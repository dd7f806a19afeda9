"""The config-2 hot path as ONE host-side call sequence (BASELINE.json configs[1]):

    ci   = CutInfo(mesh, lsetp1)                                   xfem/cutinfo.cpp:61-133
    Vh   = Restrict(H1(mesh, order=p), ci.GetElementsOfType(HASNEG)) utils/restrictedfespace.cpp:23-73
    a    = RestrictedBilinearForm(Vh, element_restriction=hasneg)   utils/restrictedblf.cpp:142-163
    a   += (grad(u)*grad(v) + u*v) * dCut(lsetp1, NEG, definedonelements=hasneg)
    a   += u*v * dCut(lsetp1, IF, definedonelements=ci.GetElementsOfType(IF))
    a.Assemble()                                                   xfem/symboliccutbfi.cpp:76-486

`CutAssembly.step()` runs exactly that through the C ABI (ngsxfem_b200.ops) on device-resident
inputs; `step_host()` runs it from pinned HOST buffers with the host<->device copies inside
(the e2e leg of bench.py).  Used by bench.py, __graft_entry__.smoke() and the parity tests.
"""
import os

import torch

from . import ops
from ._lib import Ctx, default_ctx

TERMS_VOL = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0), (ops.OP_ID, ops.OP_ID, 1.0, 0)]
TERMS_IF = [(ops.OP_ID, ops.OP_ID, 1.0, 0)]


class PhaseTimer:
    """CUDA-event stopwatch per phase on the current stream (the stream the kernels run on)."""

    def __init__(self, enabled=True):
        self.enabled = enabled
        self.marks = []
        self.split = {}

    def split_phase(self, name, parts):
        """replace phase `name` by named parts (ms measured inside the library); the remainder stays under `name`"""
        if self.enabled:
            self.split[name] = parts

    def mark(self, name):
        if self.enabled:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def result(self):
        """{phase: ms}; call after a synchronize"""
        out = {}
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        for name, parts in self.split.items():
            if name in out:
                for k, v in parts.items():
                    out[k] = out.get(k, 0.0) + v
                    out[name] -= v
        return out


_SIDE = {}


def _side_context(dev):
    """per device: a LOW-priority (stream, xgpu context) pair for the concurrent cut-element branch (its kernels use
    short-lived blocks, so they fill whatever the main stream leaves free) and a HIGH-priority pair for the main chain"""
    if dev.index not in _SIDE:
        swap = os.environ.get("XGPU_SIDE_PRIO", "low") == "high"   # experiment switch (tools/perf_overlap.py)
        lo, hi = torch.cuda.Stream(dev, priority=-1 if swap else 0), torch.cuda.Stream(dev, priority=0 if swap else -1)
        with torch.cuda.stream(lo):
            c_lo = Ctx(dev.index)
            c_lo.set_option("short_blocks", int(os.environ.get("XGPU_SIDE_SHORT", "0" if swap else "1")))
            c_lo.set_option("cut_blocks_per_sm", int(os.environ.get("XGPU_CUT_BPSM", "2")))
        with torch.cuda.stream(hi):
            c_hi = Ctx(dev.index)
        _SIDE[dev.index] = (lo, c_lo, hi, c_hi)
    return _SIDE[dev.index]


class CutAssembly:
    """device-resident mesh / space tables + one assembly pass"""

    def __init__(self, dim, el_vtx, vtx_xyz, el_dofs, ndof_full, fe_order, ctx=None):
        self.dim, self.p = dim, fe_order
        self.el_vtx, self.vtx_xyz, self.el_dofs, self.ndof_full = el_vtx, vtx_xyz, el_dofs, int(ndof_full)
        self.ne = el_vtx.shape[0]
        self.dev = el_vtx.device
        self.ctx = ctx or default_ctx(self.dev.index)
        self.nd = ops.h1_ndof(dim, fe_order)
        # second context on a side stream: the element matrices (HBM-write / FP64 bound) run
        # concurrently with the sparsity graph (shared-memory / issue bound); they only meet in the scatter
        self.side_stream, self.ctx2, self.main_hp, self.ctx_hp = _side_context(self.dev)

    def step(self, phi, timer=None, keep=False, hooks=None, concurrent=False, mode="fused", cut_rules="fly"):
        """one pass of the hot path; returns a dict with the CSR matrix and the unit counts.
        mode="fused" (default): the uncut affine elements are a PROCEDURAL batch (7 coefficients per element,
        `xgpu_affine_coefs`, no 800 B matrices in HBM), a count pass gives the row offsets (`xgpu_csr_plan_rows`) and every
        row is finished in ONE visit (`xgpu_csr_assemble_fused`: sorting networks in registers / over the warp, no hash
        tables).  mode="twopass": the round-1 path — every element matrix written to memory, `xgpu_csr_plan` (pattern
        pass) + `xgpu_csr_assemble` (gather); kept as the cross-check.  Both give the same bits.
        hooks: optional {"classified"|"need_dofs"|"need_xyz"|"pattern"|"values": fn(partial results)} called at those
        points (the host-buffer path overlaps its H2D / D2H copies with the kernels through them).
        concurrent=True runs the cut-element branch (cut quadrature + cut element matrices: FP64 / latency bound, low
        occupancy) on a side stream next to Restrict + the transposed table (memory / atomics bound); they meet in the
        row pass.  Off when per-phase CUDA-event times are wanted.
        cut_rules="fly" (default, order <= 2): the cut rules are made inside the element-matrix kernel
        (`xgpu_elmat_cut_pair`); "stored": `xgpu_cutquad_count/_fill` write them out, `xgpu_elmat_rule_pair` reads them
        (the result's "npts" is (None, None) on the "fly" path: `cut_point_counts` gives the numbers)."""
        if concurrent and torch.cuda.current_stream(self.dev) != self.main_hp:
            # the main chain on a high-priority stream, the cut-element branch on a low-priority one
            caller = torch.cuda.current_stream(self.dev)
            self.main_hp.wait_stream(caller)
            with torch.cuda.stream(self.main_hp):
                out = self._step(phi, timer, keep, hooks, True, mode, self.ctx_hp, cut_rules)
            caller.wait_stream(self.main_hp)
            for v_ in out.values():
                for t_ in (v_ if isinstance(v_, tuple) else (v_,)):
                    if torch.is_tensor(t_):
                        t_.record_stream(caller)
            return out
        return self._step(phi, timer, keep, hooks, concurrent, mode, self.ctx, cut_rules)

    def cut_point_counts(self, phi, l_if=None):
        """(#NEG points, #IF points) of the cut rules on the cut elements — reporting only (the "fly" path never counts)"""
        if l_if is None:
            _, _, bi, _ = ops.classify(self.dim, self.el_vtx, phi, ctx=self.ctx)
            l_if = ops.bitset_to_list(bi, self.ne, ctx=self.ctx)
        _, n1 = ops.cutquad_count(self.dim, l_if, self.el_vtx, phi, ops.NEG, 2 * self.p, ctx=self.ctx)
        _, n2 = ops.cutquad_count(self.dim, l_if, self.el_vtx, phi, ops.IF, 2 * self.p, ctx=self.ctx)
        return n1, n2

    def _step(self, phi, timer, keep, hooks, concurrent, mode, ctx, cut_rules="fly"):
        t = timer or PhaseTimer(False)
        hooks = hooks or {}
        dim, p, ne = self.dim, self.p, self.ne
        els, xyz = self.el_vtx, self.vtx_xyz
        fused = mode == "fused"
        t.mark("start")
        # (a) CutInfo.Update
        bn, bp, bi, cr = ops.classify(dim, els, phi, ctx=ctx)
        t.mark("classify")
        if "classified" in hooks:
            hooks["classified"](dict(bits=(bn, bp, bi), cut_ratio=cr))
        has = ops.bitset_op(ops.BIT_OR, bn, bi, ne, ctx=ctx)
        n_neg, n_if = ops.bitset_counts([bn, bi], ne, ctx=ctx)      # one host wait for both list lengths
        l_neg = ops.bitset_to_list(bn, ne, ctx=ctx, count=n_neg)
        l_if = ops.bitset_to_list(bi, ne, ctx=ctx, count=n_if)
        t.mark("lists")
        order = 2 * p  # symboliccutbfi.cpp:100-118 (ID*ID present: min diff order 0)
        main = torch.cuda.current_stream(self.dev)
        side, ctx2 = (self.side_stream, self.ctx2) if concurrent else (main, ctx)
        m_if = torch.empty((l_if.numel(), self.nd, self.nd), dtype=torch.float64, device=self.dev)

        def cut_branch():
            # (b) cut quadrature: the two count passes first (each synchronises its stream only), then everything else of
            # this branch is enqueued without a host wait.  (c) cut elements: the NEG volume terms and the IF interface
            # term in one pass (one matrix per element, as in NGSolve's element loop)
            if "need_xyz" in hooks:
                hooks["need_xyz"](None)
            if concurrent:
                side.wait_stream(main)
            with torch.cuda.stream(side):
                if concurrent and t.enabled:
                    t.side = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    t.side[0].record(side)
                if cut_rules == "fly" and p <= 2:
                    # the reference's own structure: the cut rule of an element is made inside the element-matrix call and
                    # never stored (xgpu_elmat_cut_pair) — no count pass (host wait), no 32 B per point through HBM
                    if concurrent and t.enabled:
                        t.side[1].record(side)
                    ops.elmat_cut_pair(dim, p, l_if, els, xyz, phi, ops.NEG, order, TERMS_VOL, ops.IF, order, TERMS_IF,
                                       out=m_if, ctx=ctx2)
                    n1 = n2 = None
                    m_if.record_stream(side)
                else:
                    off, n1 = ops.cutquad_count(dim, l_if, els, phi, ops.NEG, order, ctx=ctx2)
                    off2, n2 = ops.cutquad_count(dim, l_if, els, phi, ops.IF, order, ctx=ctx2)
                    q, w = ops.cutquad_fill(dim, l_if, els, xyz, phi, ops.NEG, order, off, n1, ctx=ctx2)
                    q2, w2 = ops.cutquad_fill(dim, l_if, els, xyz, phi, ops.IF, order, off2, n2, ctx=ctx2)
                    t.mark("cutquad")
                    if concurrent and t.enabled:
                        t.side[1].record(side)
                    ops.elmat_rule_pair(dim, p, l_if, els, xyz, phi, (off, q, w), TERMS_VOL, (off2, q2, w2), TERMS_IF, out=m_if, ctx=ctx2)
                    for t_ in (off, off2, q, w, q2, w2, m_if):
                        t_.record_stream(side)
                if concurrent and t.enabled:
                    t.side[2].record(side)
            t.mark("elmat_cut")
            return n1, n2

        def restrict():
            # Restrict(H1, hasneg): active dofs, compressed numbering, dof rows of the two batches
            if "need_dofs" in hooks:
                hooks["need_dofs"](None)
            act = ops.dofs_of_elements(self.el_dofs, has, self.ndof_full, ctx=ctx)
            dofmap, ndof = ops.bitset_rank(act, self.ndof_full, ctx=ctx)
            self._dofmap = dofmap
            t_neg = ops.gather_rows(self.el_dofs, l_neg, dofmap, ctx=ctx)
            t_if = ops.gather_rows(self.el_dofs, l_if, dofmap, ctx=ctx)
            t.mark("restrict")
            return ndof, t_neg, t_if

        if fused and p <= 2:
            def restrict_and_table():
                # Restrict(H1, hasneg) + the dof rows of the two batches + their transposed table: one sequence
                if "need_dofs" in hooks:
                    hooks["need_dofs"](None)
                dm, nd_, tabs, pl = ops.restrict_plan(self.el_dofs, [l_neg, l_if], self.ndof_full, ctx=ctx)
                self._dofmap = dm
                t.mark("restrict_plan")
                return nd_, tabs[0], tabs[1], pl

            if concurrent:
                npts = cut_branch()
                ndof, t_neg, t_if, plan = restrict_and_table()
            else:
                ndof, t_neg, t_if, plan = restrict_and_table()
                npts = cut_branch()
            # (e) row lengths, the 7 coefficients of every uncut element, then the row pass
            ops.csr_count_rows(plan)
            t.mark("csr_count_pass")
            m_neg = ops.affine_coefs(dim, p, l_neg, els, xyz, order, TERMS_VOL, ctx=ctx)
            t.mark("affine_coefs")
            if concurrent:
                main.wait_stream(side)
            ctx2.raise_on_geometry_error()
            t.mark("join")
            rowptr, col, vals = ops.csr_assemble_fused(plan, [(t_neg, m_neg), (t_if, m_if)])
            t.mark("csr_rows_fused")
            if "pattern" in hooks:
                hooks["pattern"](dict(rowptr=rowptr, col=col))
        else:
            ndof, t_neg, t_if = restrict()
            tables = [t_neg, t_if]
            # host-buffer path: the sparsity graph needs only the dof rows, so it can be built BEFORE the element matrices and
            # its pattern leaves for the host while they are computed and summed
            early_graph = "pattern" in hooks
            plan = rowptr = col = None

            def graph():
                pl = ops.csr_plan(ndof, tables, ctx=ctx)
                t.mark("csr_plan")
                if t.enabled:
                    a_ms, p_ms = pl.phase_ms()
                    t.split_phase("csr_plan", {"csr_adjacency": a_ms, "csr_pattern_pass": p_ms})
                return pl

            if early_graph:
                plan = graph()
                rowptr, col = ops.csr_pattern(plan, self.dev)
                t.mark("csr_pattern")
                hooks["pattern"](dict(rowptr=rowptr, col=col))
            npts = cut_branch()
            m_neg = torch.empty((l_neg.numel(), self.nd, self.nd), dtype=torch.float64, device=self.dev)
            with torch.cuda.stream(side):
                ops.elmat_std(dim, p, l_neg, els, xyz, phi, order, TERMS_VOL, out=m_neg, ctx=ctx2)
                m_neg.record_stream(side)
            t.mark("elmat")
            if not early_graph:
                plan = graph()
            if concurrent:
                main.wait_stream(side)
            ctx2.raise_on_geometry_error()
            t.mark("elmat_join")
            if early_graph:
                vals = ops.csr_scatter(plan, [(t_neg, m_neg), (t_if, m_if)], rowptr, col)
            else:
                # pattern compaction fused into the scatter (xgpu_csr_assemble): rowptr, colidx and vals in one pass
                rowptr, col, vals = ops.csr_assemble(plan, [(t_neg, m_neg), (t_if, m_if)])
            t.mark("csr_scatter")
        if "values" in hooks:
            hooks["values"](dict(vals=vals))
        out = dict(ndof=ndof, nnz=plan.nnz, n_neg=l_neg.numel(), n_if=l_if.numel(),
                   units=l_neg.numel() + l_if.numel(), rowptr=rowptr, col=col, vals=vals,
                   bits=(bn, bp, bi), cut_ratio=cr, npts=npts, dofmap=self._dofmap)
        if keep:
            out.update(l_neg=l_neg, l_if=l_if, t_neg=t_neg, t_if=t_if, m_neg=m_neg, m_if=m_if)
        return out


class HostBuffers:
    """pinned host copies of the inputs and pinned landing buffers for the outputs (e2e leg)"""

    def __init__(self, el_vtx, vtx_xyz, el_dofs, phi):
        self.el_vtx = el_vtx.cpu().pin_memory()
        self.vtx_xyz = vtx_xyz.cpu().pin_memory()
        self.el_dofs = el_dofs.cpu().pin_memory()
        self.phi = phi.cpu().pin_memory()
        self.out = {}

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.el_vtx, self.vtx_xyz, self.el_dofs, self.phi))

    def land(self, name, t, streams=None):
        """device tensor -> pinned host buffer (grown on demand); returns bytes copied.
        streams: optional list of copy streams — a large tensor is split into one slice per stream so
        that several DMA engines share the PCIe link"""
        buf = self.out.get(name)
        if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
            buf = torch.empty(int(t.numel() * 1.05) + 16, dtype=t.dtype).pin_memory()
            self.out[name] = buf
        flat = t.reshape(-1)
        n = flat.numel()
        if streams and n * t.element_size() > (32 << 20):
            k = len(streams)
            step = (n + k - 1) // k
            for i, st in enumerate(streams):
                a, b = i * step, min(n, (i + 1) * step)
                if a < b:
                    with torch.cuda.stream(st):
                        buf[a:b].copy_(flat[a:b], non_blocking=True)
        else:
            buf[:n].copy_(flat, non_blocking=True)
        return n * t.element_size()


def step_host(hb, dim, ndof_full, fe_order, device, ctx=None, overlap=True, pipelined=False, trace=None):
    """the same pass starting from HOST buffers: H2D of every input, assembly, D2H of the CutInfo
    BitArrays / cut ratios and of the CSR matrix.  returns (result dict, h2d bytes, d2h bytes).

    overlap=True: the copies run on two side streams ordered against the kernels by events — the
    connectivity and level set go first (classification needs only those), the dof table follows while
    the classification runs, the CutInfo outputs / the sparsity pattern / the values leave as soon
    as their kernels have finished (PCIe is full duplex; the D2H of 1.7 GB is the long pole).

    pipelined=True (needs overlap): the call does not wait for its own D2H — the next step's H2D and kernels run while
    this step's results are still leaving (software pipelining over steps, two alternating sets of pinned landing
    buffers).  The caller ends a sequence with `drain_host(hb, device)`; results of step k are valid after the drain or
    once `hb.done[k % 2]` (a CUDA event) has completed."""
    main = torch.cuda.current_stream(device)
    if not overlap:
        els = hb.el_vtx.to(device, non_blocking=True)
        xyz = hb.vtx_xyz.to(device, non_blocking=True)
        eld = hb.el_dofs.to(device, non_blocking=True)
        phi = hb.phi.to(device, non_blocking=True)
        prob = CutAssembly(dim, els, xyz, eld, ndof_full, fe_order, ctx=ctx)
        r = prob.step(phi)
        d2h = sum(hb.land(f"bits{i}", b) for i, b in enumerate(r["bits"]))
        for k in ("cut_ratio", "rowptr", "col", "vals"):
            d2h += hb.land(k, r[k])
        main.synchronize()
        return r, hb.h2d_bytes(), d2h
    if not hasattr(hb, "s_in"):
        hb.s_in, hb.s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    s_in, s_out = hb.s_in, hb.s_out
    par = ""
    if pipelined:
        hb.parity = 1 - getattr(hb, "parity", 1)
        par = f"#{hb.parity}"
        if not hasattr(hb, "done"):
            hb.done = [None, None]
        if hb.done[hb.parity] is not None:
            hb.done[hb.parity].synchronize()   # this landing set was filled two steps ago: its consumer has had a full step
    s_in.wait_stream(main)
    s_out.wait_stream(main)
    def tr(name, stream):
        if trace is not None:   # timeline of the copies (tools/e2e_timeline.py)
            ev = torch.cuda.Event(enable_timing=True); ev.record(stream); trace.append((name, ev))

    tr("h2d_begin", s_in)
    with torch.cuda.stream(s_in):
        els = hb.el_vtx.to(device, non_blocking=True)
        phi = hb.phi.to(device, non_blocking=True)
        e_first = torch.cuda.Event(); e_first.record(s_in)
        eld = hb.el_dofs.to(device, non_blocking=True)       # Restrict needs the dof table next ...
        e_rest = torch.cuda.Event(); e_rest.record(s_in)
        xyz = hb.vtx_xyz.to(device, non_blocking=True)       # ... the coordinates only the quadrature / element matrices
        e_xyz = torch.cuda.Event(); e_xyz.record(s_in)
    tr("h2d_end", s_in)
    for t_ in (els, phi, xyz, eld):
        t_.record_stream(main)
    main.wait_event(e_first)
    d2h = [0]

    def send(items):
        ev = torch.cuda.Event(); ev.record(main)
        s_out.wait_event(ev)
        tr("d2h_begin:" + items[0][0], s_out)
        with torch.cuda.stream(s_out):
            for name, t_ in items:
                t_.record_stream(s_out)
                d2h[0] += hb.land(name + par, t_)  # one D2H stream: two did not add bandwidth (PCIe bound, profiles/r01_notes.md)

    hooks = {
        "classified": lambda p: send([(f"bits{i}", b) for i, b in enumerate(p["bits"])] + [("cut_ratio", p["cut_ratio"])]),
        "need_dofs": lambda p: main.wait_event(e_rest),
        "need_xyz": lambda p: main.wait_event(e_xyz),
        "pattern": lambda p: send([("rowptr", p["rowptr"]), ("col", p["col"])]),
        "values": lambda p: send([("vals", p["vals"])]),
    }
    prob = CutAssembly(dim, els, xyz, eld, ndof_full, fe_order, ctx=ctx)
    r = prob.step(phi, hooks=hooks)
    tr("d2h_end", s_out)
    tr("compute_end", main)
    if pipelined:
        hb.done[hb.parity] = torch.cuda.Event()
        hb.done[hb.parity].record(s_out)
        return r, hb.h2d_bytes(), d2h[0]
    s_out.synchronize()
    main.synchronize()
    return r, hb.h2d_bytes(), d2h[0]


def drain_host(hb, device):
    """end of a pipelined sequence of step_host calls: the launching stream waits for the outstanding copies"""
    main = torch.cuda.current_stream(device)
    if hasattr(hb, "s_out"):
        main.wait_stream(hb.s_out)
        main.wait_stream(hb.s_in)


def sphere_levelset(points, centers, radius, eps=1e-14):
    """nodal P1 values of min_k |x - c_k| - r with the InterpolateToP1 perturbation
    (utils/p1interpol.cpp:105-107: |v| < eps -> +eps)"""
    phi = None
    for c in centers:
        d = torch.sqrt(sum((points[:, i] - c[i]) ** 2 for i in range(points.shape[1]))) - radius
        phi = d if phi is None else torch.minimum(phi, d)
    return torch.where(phi.abs() < eps, torch.full_like(phi, eps), phi).contiguous()
