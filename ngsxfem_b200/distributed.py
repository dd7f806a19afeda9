"""Element-partitioned cut assembly over several GPUs (north_star: "elements are
partitioned by contiguous index ranges across the GPUs of one box, and the per-rank CSR pieces are
merged with NCCL over NVLink only at the final interface-row reduction").

One process per GPU (`torch.distributed`, backend nccl; gloo on CPU for the host-logic tests).
Each rank holds a LOCAL mesh = its own contiguous element range followed by one layer of HALO
elements of its neighbours (connectivity / level set only: halo element matrices are never
computed here).  Every local dof carries a global key `gid` (any total order that all ranks
agree on) and an `owner` rank.

Per assembly:
  1. classify all local elements (own + halo)                      [compute backend]
  2. `mark`: OR the activity of shared dofs with the neighbours — the reference's
     `AllReduceNodalData(…, MPI_LOR)` (xfem/cutinfo.cpp:194-199).  With a complete halo layer this is
     already implied locally (a dof shared with a neighbour is touched by that neighbour's halo
     elements), so no message is needed.
  3. number the active local dofs by (owner, gid): every rank sees shared dofs in the same relative
     order, hence the rows of an interface dof have IDENTICAL column order on both sides
  4. element matrices of the OWN active elements                    [compute backend]
  5. local CSR over own batches + the halo batch as sparsity only (mats = NULL), so the owner's
     pattern of an interface row is already final ("union of cliques" over both sides)
  6. interface-row reduction: for each neighbour q, the values of the rows {owner == q, touched by my own elements}
     are packed (xgpu_rows_pack) and sent, the owner adds what it receives (xgpu_rows_add) in a fixed order (lower rank
     first).  All candidate lists are static; inactive candidates have length 0 (masks, no compaction).
  7. global numbering: offset[owner] + index among the owner's active dofs; the owner-local indices of the interface
     dofs travel in the SAME batched NCCL isend/irecv round as the values (static-size int message per neighbour),
     the offsets come from one all-reduce of the owned counts.  These are the only data-path communications.
The result is a row-distributed CSR (rows = owned dofs) with global column ids.

The compute backend is duck-typed (`classify`, `elmats`, `csr`): `GpuBackend` below drives the
CUDA kernels through the C ABI; the CPU gloo tests inject their own checker backend (tests/dist_backend.py).
"""
import os

import torch
import torch.distributed as dist

from ._lib import XgpuError


# ------------------------------------------------------------------ helpers (torch plumbing, CPU or CUDA)
def _excl_cumsum(x):
    c = torch.cumsum(x, 0)
    return c - x


def row_positions(rowptr, rows):
    """flat positions in vals of the listed CSR rows (row after row)"""
    rows = rows.long()
    start = rowptr[rows]
    ln = rowptr[rows + 1] - start
    tot = int(ln.sum().item()) if rows.numel() else 0
    if tot == 0:
        return torch.zeros(0, dtype=torch.int64, device=rowptr.device), ln
    off = _excl_cumsum(ln)
    idx = torch.repeat_interleave(start - off, ln) + torch.arange(tot, device=rowptr.device)
    return idx, ln


class LocalPartition:
    """inputs of one rank (all tensors on the rank's device):
    el_vtx (ne_loc x D+1 int32, own elements first), ne_own, el_rank (ne_loc int32: owning rank of
    every local element), vtx_xyz, el_dofs (ne_loc x nd int32, local dof ids), ndof_loc,
    gid (ndof_loc int64 global key), owner (ndof_loc int32), halo_for {q: int32 list of OWN elements
    that rank q holds as halo}.

    The constructor renumbers the local dofs ONCE into (owner, gid) order, so that per assembly the
    compressed numbering of the active dofs is the plain ascending compression of the single-GPU
    path (Restrict), dofs of one owner form a contiguous range, and the small static interface
    tables below can be prepared up front."""

    def __init__(self, dim, el_vtx, ne_own, el_rank, vtx_xyz, el_dofs, ndof_loc, gid, owner, halo_for, fe_order, rank):
        self.dim, self.el_vtx, self.ne_own, self.el_rank, self.vtx_xyz = dim, el_vtx, int(ne_own), el_rank, vtx_xyz
        self.ndof_loc, self.fe_order, self.rank = int(ndof_loc), fe_order, rank
        self.neighbors = sorted(halo_for.keys())
        dev = el_vtx.device
        assert int(gid.max().item()) < (1 << 52)
        key = (owner.to(torch.int64) << 52) | gid
        order = torch.argsort(key)                       # old local dof ids in key order
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel(), device=dev)
        self.el_dofs = inv[el_dofs.long()].to(torch.int32).contiguous()   # key-order dof ids
        self.gid, self.owner = gid[order].contiguous(), owner[order].contiguous()
        # contiguous static dof range per owner
        ow = self.owner.to(torch.int64)
        self.range_of = {}
        for q in [rank] + self.neighbors:
            idx = torch.nonzero(ow == q).reshape(-1)
            self.range_of[q] = (int(idx[0].item()), int(idx[-1].item()) + 1) if idx.numel() else (0, 0)
        # static interface tables per neighbour q
        self.iface = {}
        for q in self.neighbors:
            E = halo_for[q].long()                                        # my elements that q holds as halo
            H = torch.nonzero(el_rank.to(torch.int64) == q).reshape(-1)     # my halo elements owned by q
            TE, TH = self.el_dofs[E].long(), self.el_dofs[H].long()
            ent = {"E": E, "H": H}
            # candidate rows I send to q: dofs owned by q among the dofs of my elements E
            ent["C_send"], pos = self._candidates(TE, ow, q)
            ent["pos_send"] = _compact_pos(pos)
            # candidate dofs owned by me that q can see / contribute to
            D, pos = self._candidates(torch.cat([TE, TH]), ow, rank)
            ent["D"], ent["posE"], ent["posH"] = D, _compact_pos(pos[:TE.shape[0]]), _compact_pos(pos[TE.shape[0]:])
            self.iface[q] = ent

    def kernel_tables(self):
        """static inputs of the interface kernels (xgpu_iface_marks / _rows / _index; built once, on first use):
        all candidate lists of all neighbours concatenated — rows part [C_send_q, D_q]_q, then the `seen` part [D_q]_q —
        the (element, flag position) pairs that mark them, and the static message capacities: cap_recv[q] = sum of the FULL
        row lengths (every element active) of my candidate rows D_q = an upper bound of what q can send me."""
        if getattr(self, "_kt", None) is not None:
            return self._kt
        dev = self.el_vtx.device
        i32 = lambda t: t.to(torch.int32).contiguous()
        cand_rows, pe, pg, views = [], [], [], {}
        base = 0
        for q in self.neighbors:
            it = self.iface[q]
            nC, nD = it["C_send"].numel(), it["D"].numel()
            r, pos = it["pos_send"]
            pe.append(it["E"][r]); pg.append(pos + base)
            r, pos = it["posH"]
            pe.append(it["H"][r]); pg.append(pos + base + nC)
            views[q] = (base, nC, nD)
            cand_rows += [it["C_send"], it["D"]]
            base += nC + nD
        nrows = base
        seen_base = {}
        for q in self.neighbors:
            it = self.iface[q]
            nD = it["D"].numel()
            seen_base[q] = base
            for key, els in (("posH", it["H"]), ("posE", it["E"])):
                r, pos = it[key]
                pe.append(els[r]); pg.append(pos + base)
            base += nD
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        cand_seen = [self.iface[q]["D"] for q in self.neighbors]
        cap_recv = {}
        for q in self.neighbors:
            it = self.iface[q]
            T = torch.cat([self.el_dofs[it["E"]], self.el_dofs[it["H"]]]).long()
            nd = T.shape[1]
            inc = torch.zeros(self.ndof_loc, dtype=torch.bool, device=dev)
            inc[it["D"]] = True
            rows = T.unsqueeze(2).expand(-1, nd, nd)
            cols = T.unsqueeze(1).expand(-1, nd, nd)
            m = (rows >= 0) & (cols >= 0)
            m &= inc[rows.clamp(min=0)]
            cap_recv[q] = int(torch.unique(rows[m] * self.ndof_loc + cols[m]).numel())
        self._kt = dict(cand_rows=i32(torch.cat(cand_rows)) if cand_rows else i32(z), nrows=nrows,
                        cand_seen=i32(torch.cat(cand_seen)) if cand_seen else i32(z), nflags=base,
                        pair_elem=i32(torch.cat(pe)) if pe else i32(z), pair_gpos=i32(torch.cat(pg)) if pg else i32(z),
                        views=views, seen_base=seen_base, cap_recv=cap_recv)
        return self._kt

    @staticmethod
    def _candidates(T, ow, q):
        """sorted unique dofs of table T owned by q, and for every entry of T its position in that list (-1)"""
        m = ow[T] == q
        C = torch.unique(T[m])                           # ascending = key order
        pos = torch.full(T.shape, -1, dtype=torch.int64, device=T.device)
        pos[m] = torch.searchsorted(C, T[m])
        return C, pos


def _compact_pos(pos):
    """static (row, candidate position) pairs of a position table (rows x width, -1 = not a candidate)"""
    r, c = torch.nonzero(pos >= 0, as_tuple=True)
    return r.contiguous(), pos[r, c].contiguous()


def _marked(rp, act, n):
    """flags over a candidate list of length n: positions hit by the ACTIVE rows; rp = _compact_pos(table).
    Mask arithmetic over the static pair list only: no data-dependent sizes, no host synchronisation."""
    rows, posv = rp
    f = torch.zeros(n + 1, dtype=torch.bool, device=posv.device)
    f[torch.where(act[rows], posv, torch.full_like(posv, n))] = True
    return f[:n]


class _StepResult(dict):
    """result dict of a step; entries given as callables are evaluated on first access (a device read-back that the
    step itself does not wait for)"""

    def __init__(self, d, **lazy):
        super().__init__(d)
        self._lazy = lazy

    def __getitem__(self, k):
        if k not in self.keys() and k in self._lazy:
            self[k] = self._lazy.pop(k)()
        return super().__getitem__(k)

    def __contains__(self, k):
        return super().__contains__(k) or k in self._lazy


class DistributedCutAssembly:
    def __init__(self, part, backend, rank, world, group=None):
        self.part, self.be, self.rank, self.world, self.group = part, backend, rank, world, group
        dev = part.el_vtx.device
        self._ne_own = torch.tensor([part.ne_own], dtype=torch.int64, device=dev)

    # ---- communication primitives
    def _exchange_all(self, msgs):
        """msgs: [(q, send_tensor, recv_tensor)].  NCCL: one batched group of isend/irecv (all neighbours in
        flight together, ordered on the current stream).  gloo (CPU tests): blocking pairs, lower rank sends first."""
        if not msgs:
            return
        if dist.get_backend(self.group) == "nccl":
            ops = []
            for q, s_t, r_t in msgs:
                ops.append(dist.P2POp(dist.isend, s_t, q, self.group))
                ops.append(dist.P2POp(dist.irecv, r_t, q, self.group))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            return
        for q, s_t, r_t in msgs:
            if self.rank < q:
                dist.send(s_t, q, group=self.group)
                dist.recv(r_t, q, group=self.group)
            else:
                dist.recv(r_t, q, group=self.group)
                dist.send(s_t, q, group=self.group)

    def _handshake(self):
        """once: message capacities (what every neighbour may send me at most) travel to the senders, and the candidate
        lists are validated — my list D towards q and q's ghost range of my dofs must be the same dofs in the same order
        (true for complete one-layer halos; a general partition that breaks it fails here instead of hanging in NCCL)"""
        p, dev = self.part, self.part.el_vtx.device
        kt = p.kernel_tables()
        msgs = []
        for q in p.neighbors:
            qa, qb = p.range_of[q]
            mine = torch.tensor([kt["cap_recv"][q], p.iface[q]["D"].numel(), qb - qa,
                                 int(p.gid[p.iface[q]["D"]].sum().item()) % (1 << 40), int(p.gid[qa:qb].sum().item()) % (1 << 40)],
                                dtype=torch.int64, device=dev)
            msgs.append((q, mine, torch.empty_like(mine)))
        self._exchange_all(msgs)
        self.cap_send = {}
        for q, mine, theirs in msgs:
            t, m = theirs.tolist(), mine.tolist()
            if t[1] != m[2] or t[2] != m[1] or t[3] != m[4] or t[4] != m[3]:
                raise RuntimeError(f"rank {self.rank}: the interface candidate lists towards rank {q} do not match its ghost range "
                                   "(the partition needs a complete one-layer halo)")
            self.cap_send[q] = t[0]

    def step(self, phi, timer=None):
        if getattr(self.be, "kernel_path", False) and (self.world == 1 or dist.get_backend(self.group) == "nccl"):
            return self._step_kernels(phi, timer)
        return self._step_generic(phi, timer)

    def _step_kernels(self, phi, timer=None):
        """the step on the CUDA backend: the single-GPU kernel sequence (Restrict + transposed table in one sequence, cut
        elements on a side stream, procedural affine batch, one-row-pass CSR) over own + halo batches, the interface
        bookkeeping in merge.cu kernels over the static candidate tables, messages of STATIC capacity (no host round trip
        for their sizes), numbering by kernels.  Same protocol and same results as _step_generic."""
        p, be, r = self.part, self.be, self.rank
        ops, ctx = be.ops, be.ctx
        dev = p.el_vtx.device
        mark = timer.mark if timer is not None else (lambda name: None)
        kt = p.kernel_tables()
        if self.world > 1 and not hasattr(self, "cap_send"):
            self._handshake()
        mark("start")
        ne = p.el_vtx.shape[0]
        if type(be) is GpuBackend:
            # classification + all four list lengths (own elements come first: prefix counts) with ONE host wait; the geometry
            # status is checked once, inside local_assembly
            dim = p.el_vtx.shape[1] - 1
            cls = ops.classify(dim, p.el_vtx, phi, ctx=ctx)
            bn, bp, bi, cr = cls
            has = ops.bitset_op(ops.BIT_OR, bn, bi, ne, ctx=ctx)
            n_neg, n_if, kn, ki = ops.bitset_counts([bn, bi, bn, bi], [ne, ne, p.ne_own, p.ne_own], ctx=ctx)
            l_neg = ops.bitset_to_list(bn, ne, ctx=ctx, count=n_neg)
            l_if = ops.bitset_to_list(bi, ne, ctx=ctx, count=n_if)
        else:
            l_neg, l_if, cls = be.classify(p.el_vtx, phi)
            bn, bp, bi, cr = cls
            has = ops.bitset_op(ops.BIT_OR, bn, bi, ne, ctx=ctx)
            ne_own = self._ne_own.to(l_neg.dtype)
            kn, ki = torch.stack([torch.searchsorted(l_neg, ne_own)[0], torch.searchsorted(l_if, ne_own)[0]]).tolist()
        l_has = ops.bitset_to_list(has, ne, ctx=ctx, count=l_neg.numel() + l_if.numel())
        l_neg_own, l_if_own, l_halo = l_neg[:kn], l_if[:ki], l_has[kn + ki:]
        mark("classify")
        a, b = p.range_of[r]
        if not hasattr(self, "_pin"):
            self._pin = torch.zeros(2, dtype=torch.int64).pin_memory()

        def owned_range(dofmap):
            # active dofs below the two ends of my owner range: two small reductions + an async copy into pinned memory,
            # enqueued right after Restrict — the count pass's own host synchronisation (nnz) completes it, so the
            # distributed logic adds no host round trip of its own here
            def below(i):
                if i <= 0:
                    return torch.zeros((), dtype=torch.int64, device=dev)
                return (dofmap[:i].max().to(torch.int64) + 1).clamp(min=0)
            self._pin.copy_(torch.stack([below(a), below(b)]), non_blocking=True)

        state = {}

        def bookkeeping(dofmap, rowptr, vals, n_lo):
            """interface flags, row lists, index messages of ALL neighbours (three launches), then the packed value messages"""
            flags = ops.iface_marks(has, kt["pair_elem"], kt["pair_gpos"], kt["nflags"], ctx=ctx)
            rows_all = ops.iface_rows(kt["cand_rows"], flags, dofmap, ctx=ctx)
            idx_all = ops.iface_index(kt["cand_seen"], flags[kt["nrows"]:], dofmap, n_lo, ctx=ctx)
            mark("merge_rows")
            msgs, lists, ib = [], [], 0
            for q in p.neighbors:
                base, nC, nD = kt["views"][q]
                srows, rrows = rows_all[base:base + nC], rows_all[base + nC:base + nC + nD]
                soff, roff = ops.rows_offsets(rowptr, srows, ctx=ctx), ops.rows_offsets(rowptr, rrows, ctx=ctx)
                lists.append((rrows, roff))
                msgs.append((q, ops.rows_pack(rowptr, srows, soff, vals, self.cap_send[q], ctx=ctx),
                             torch.empty(kt["cap_recv"][q], dtype=vals.dtype, device=dev)))
                qa, qb = p.range_of[q]
                msgs.append((q, idx_all[ib:ib + nD], torch.empty(qb - qa, dtype=torch.int64, device=dev)))
                ib += nD
            mark("merge_pack")
            return msgs, lists

        def split_values(plan, batches, dofmap, ndof):
            """the rows this rank SENDS (ghost dofs: the two ends of the (owner, key) numbering) are finished first and travel
            on a side stream while the rows it owns are summed; the received values are added afterwards"""
            n_lo, n_hi = self._pin.tolist()      # completed by the count pass's host wait
            main = torch.cuda.current_stream(dev)
            try:
                out = ops.csr_assemble_fused(plan, batches, rows=(0, n_lo))
                ops.csr_assemble_fused(plan, batches, rows=(n_hi, ndof), out=out)
            except XgpuError:                     # rows outside the thread- / warp-per-row kernels: one pass, exchange after it
                out = ops.csr_assemble_fused(plan, batches)
                state["late"] = True
                return out
            rowptr, col, vals = out
            mark("csr_rows_ghost")
            msgs, lists = bookkeeping(dofmap, rowptr, vals, n_lo)
            if not hasattr(self, "_xstream"):
                self._xstream = torch.cuda.Stream(dev)
            xs = self._xstream
            xs.wait_stream(main)
            with torch.cuda.stream(xs):
                for _, s_t, r_t in msgs:
                    s_t.record_stream(xs); r_t.record_stream(xs)
                self._exchange_all(msgs)
            ops.csr_assemble_fused(plan, batches, rows=(n_lo, n_hi), out=out)
            state.update(msgs=msgs, lists=lists, xs=xs)
            return out

        # opt-in (XGPU_OVERLAP=1): measured at N = 2 the device-resident step is the same within noise (strong 3.91 vs 3.93 ms,
        # the wait moves into the numbering phase) and the pipelined e2e leg loses 2x (side-stream blocks are not reusable by
        # the caching allocator until the exchange ends); profiles/r02_notes.md
        overlap = self.world > 1 and bool(p.neighbors) and type(be) is GpuBackend and os.environ.get("XGPU_OVERLAP", "0") == "1"
        res = be.local_assembly(p, phi, l_neg_own, l_if_own, l_halo, mark, after_restrict=owned_range if self.world > 1 else None,
                                values=split_values if overlap else None)
        dofmap, ndof, rowptr, col, vals = res
        if self.world > 1 and ndof > 0:
            n_lo, n_hi = self._pin.tolist()
        else:
            n_lo, n_hi = 0, ndof
        n_own = n_hi - n_lo
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        counts[r] = n_own
        red = None
        if self.world > 1:
            red = dist.all_reduce(counts, group=self.group, async_op=True)
        msgs, lists = [], []
        if p.neighbors:
            if "msgs" in state:                                            # the exchange has been running under the own rows
                msgs, lists = state["msgs"], state["lists"]
                torch.cuda.current_stream(dev).wait_stream(state["xs"])
                mark("merge_exchange")
            else:
                msgs, lists = bookkeeping(dofmap, rowptr, vals, n_lo)
                self._exchange_all(msgs)
                mark("merge_exchange")
            for k, q in enumerate(p.neighbors):                           # neighbours ascending: lower rank first
                rrows, roff = lists[k]
                ops.rows_add(rowptr, rrows, roff, msgs[2 * k][2], vals, ctx=ctx)
            mark("merge")
        if red is not None:
            red.wait()
        l2g = ops.l2g_own(ndof, n_lo, n_own, counts, r, ctx=ctx)
        for k, q in enumerate(p.neighbors):
            qa, qb = p.range_of[q]
            ops.l2g_ghosts(l2g, dofmap[qa:qb], msgs[2 * k + 1][2], counts, q, ctx=ctx)
        mark("numbering")
        return _StepResult(dict(ndof_local=ndof, n_lo=n_lo, n_own=n_own, rowptr=rowptr, col=col, vals=vals, l2g=l2g, dofmap=dofmap,
                                units=kn + ki, n_neg=kn, n_if=ki, nnz=int(col.numel()), cls=cls),
                           ndof_global=lambda: int(counts.sum().item()))

    def _step_generic(self, phi, timer=None):
        p, be, r = self.part, self.be, self.rank
        dev = p.el_vtx.device
        mark = timer.mark if timer is not None else (lambda name: None)
        mark("start")
        # 1. classification of own + halo elements; own elements come first, lists are ascending
        l_neg, l_if, cls = be.classify(p.el_vtx, phi)
        mark("classify")
        # 2./3. active dofs of own + halo elements; dof ids are already in (owner, gid) order
        dofmap, ndof = be.active_dofs(p, cls, l_neg, l_if)

        def below(i):  # number of active dofs with local id < i (dofmap is the ascending compression, -1 = inactive)
            if i <= 0:
                return torch.zeros((), dtype=torch.int64, device=dev)
            return (dofmap[:i].max().to(torch.int64) + 1).clamp(min=0)

        ne_own = self._ne_own.to(l_neg.dtype)
        sizes = [torch.searchsorted(l_neg, ne_own)[0].to(torch.int64), torch.searchsorted(l_if, ne_own)[0].to(torch.int64)]
        a, b = p.range_of[r]
        sizes += [below(a), below(b)]
        sizes = torch.stack(sizes).tolist()          # the ONE host synchronisation of this phase
        kn, ki = sizes[0], sizes[1]
        n_lo, n_own = sizes[2], sizes[3] - sizes[2]
        mark("restrict_sizes")
        l_neg_own, l_if_own = l_neg[:kn], l_if[:ki]
        l_halo = torch.cat([l_neg[kn:], l_if[ki:]]).sort().values
        t_neg = be.gather_rows(p.el_dofs, l_neg_own, dofmap)
        t_if = be.gather_rows(p.el_dofs, l_if_own, dofmap)
        t_halo = be.gather_rows(p.el_dofs, l_halo, dofmap)
        mark("restrict_rows")
        # activity of the static interface candidates (masks over the static lists: no compaction, no host sync)
        m_snd, m_rcv, m_seen = {}, {}, {}
        for q in p.neighbors:
            it = p.iface[q]
            actE, actH = be.element_activity(cls, it["E"]), be.element_activity(cls, it["H"])
            m_snd[q] = _marked(it["pos_send"], actE, it["C_send"].numel())
            m_rcv[q] = _marked(it["posH"], actH, it["D"].numel())
            m_seen[q] = m_rcv[q] | _marked(it["posE"], actE, it["D"].numel())
        mark("restrict")
        # 4. element matrices of the own active elements
        m_neg, m_if = be.elmats(p, phi, l_neg_own, l_if_own)
        mark("elmat")
        # 5. local CSR; the halo batch contributes sparsity only
        batches = [(t_neg, m_neg), (t_if, m_if), (t_halo, None)]
        batches = [bt for bt in batches if bt[0].shape[0] > 0]
        if batches:
            rowptr, col, vals = be.csr(ndof, batches)
        else:  # nothing active on this rank (not even in the halo): an empty local matrix
            rowptr = torch.zeros(ndof + 1, dtype=torch.int64, device=dev)
            col = torch.zeros(0, dtype=torch.int32, device=dev)
            vals = torch.zeros(0, dtype=torch.float64, device=dev)
        mark("csr")
        # 6. interface-row reduction: the values of the rows {owner == q, touched by my own elements} are packed and
        #    sent, the owner adds what it receives in a fixed order (lower rank first)
        lists = []
        for q in p.neighbors:
            it = p.iface[q]
            for cand, m in ((it["C_send"], m_snd[q]), (it["D"], m_rcv[q])):
                loc = dofmap[cand]
                rows = torch.where(m, loc, torch.full_like(loc, -1)).contiguous()   # -1: inactive candidate
                lists.append((rows, be.rows_offsets(rowptr, rows)))
        mark("merge_rows")
        tots = torch.stack([off[-1] for _, off in lists]).tolist() if lists else []   # the one host sync of this phase
        # 7. global numbering = offset[owner] + index among the owner's active dofs (ascending key).  The owner-local
        #    indices of the interface dofs travel in the SAME message round as the interface values: one STATIC-size
        #    int message per neighbour over all interface candidates (-1 = inactive) — my candidate list D towards q
        #    and q's ghost range of my dofs are the same dofs in the same key order (LocalPartition); the offsets come
        #    from one all-reduce of the owned counts, issued first so that it overlaps the packing.
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        counts[r] = n_own
        red = None
        if self.world > 1:
            red = dist.all_reduce(counts, group=self.group, async_op=True)
        msgs = []
        for k, q in enumerate(p.neighbors):
            (srows, soff), (rrows, roff) = lists[2 * k], lists[2 * k + 1]
            msgs.append((q, be.rows_pack(rowptr, srows, soff, vals, tots[2 * k]),
                         torch.empty(tots[2 * k + 1], dtype=vals.dtype, device=dev)))
            loc = dofmap[p.iface[q]["D"]].long()
            qa, qb = p.range_of[q]
            msgs.append((q, torch.where(m_seen[q], loc - n_lo, torch.full_like(loc, -1)).contiguous(),
                         torch.empty(qb - qa, dtype=torch.int64, device=dev)))
        mark("merge_pack")
        self._exchange_all(msgs)
        mark("merge_exchange")
        for k, q in enumerate(p.neighbors):                               # neighbours ascending: lower rank first
            rrows, roff = lists[2 * k + 1]
            be.rows_add(rowptr, rrows, roff, msgs[2 * k][2], vals)
        mark("merge")
        if red is not None:
            red.wait()
        offsets = _excl_cumsum(counts)
        l2g = torch.full((ndof + 1,), -1, dtype=torch.int64, device=dev)   # slot ndof: sink of the inactive candidates
        l2g[n_lo:n_lo + n_own] = offsets[r] + torch.arange(n_own, device=dev)
        for k, q in enumerate(p.neighbors):
            rbuf = msgs[2 * k + 1][2]
            qa, qb = p.range_of[q]
            loc = dofmap[qa:qb].long()                                     # my ghosts owned by q
            l2g.scatter_(0, torch.where(loc >= 0, loc, torch.full_like(loc, ndof)),
                         torch.where(rbuf >= 0, rbuf + offsets[q], rbuf))
        l2g = l2g[:ndof]
        ndof_global = counts.sum()
        mark("numbering")
        return dict(ndof_local=ndof, n_lo=n_lo, n_own=n_own, rowptr=rowptr, col=col, vals=vals, l2g=l2g, dofmap=dofmap,
                    ndof_global=int(ndof_global.item()), units=int(l_neg_own.numel() + l_if_own.numel()),
                    n_neg=int(l_neg_own.numel()), n_if=int(l_if_own.numel()), nnz=int(col.numel()), cls=cls)

    @staticmethod
    def owned_rows_global(res):
        """(global row ids, rowptr slice, global cols, vals) of the rows this rank owns"""
        a, b = res["n_lo"], res["n_lo"] + res["n_own"]
        rp = res["rowptr"]
        lo, hi = int(rp[a].item()), int(rp[b].item())
        cols = res["l2g"][res["col"][lo:hi].long()]
        if res["ndof_global"] < 2 ** 31:   # same index width as the single-GPU CSR
            cols = cols.to(torch.int32)
        return res["l2g"][a:b], rp[a:b + 1] - lo, cols, res["vals"][lo:hi]


# ------------------------------------------------------------------ consistency of the merged matrix (bench.py, tests)
def _key_weight(g, a):
    """deterministic weight in [0.5, 1.5) of a global dof key"""
    return 0.5 + ((g * a) % 1048573).to(torch.float64) / 1048573.0


def _block_checksums(rowptr, col, vals, keys_of_local):
    """linear functionals of a CSR block whose rows / columns carry global keys: sum of the entries and two
    position-dependent weighted sums  sum_ij A_ij w(key_i) w'(key_j)"""
    n = rowptr.numel() - 1
    ln = rowptr[1:] - rowptr[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=vals.device), ln)
    kr, kc = keys_of_local[rows], keys_of_local[col.long()]
    v = vals
    return torch.stack([v.sum(), (v * _key_weight(kr, 7919) * _key_weight(kc, 104729)).sum(),
                        (v * _key_weight(kr, 611953) * _key_weight(kc, 15485863)).sum(), v.abs().sum()])


def merge_check(asm, res, phi, group=None, rtol=1e-10):
    """Does the row-distributed matrix (after the NCCL interface-row merge) hold exactly the contributions of all ranks?
    Every rank assembles ITS OWN elements alone (single-GPU path, no halo, no merge); since every element belongs to one
    rank, the sum over the ranks of any linear functional of those local matrices equals the functional of the global
    matrix.  Three functionals with position-dependent weights over the GLOBAL dof keys (sum of the entries, two weighted
    sums) are all-reduced and compared with the same functionals of the owned rows of the merged result.  Returns a dict
    (ok, relative differences, totals) — bench.py prints it as "check" and fails on a mismatch."""
    from .pipeline import CutAssembly
    p = asm.part
    dev = p.el_vtx.device
    # merged result: owned rows, columns in local compressed numbering; keys of the local compressed dofs
    keys = p.gid[res["dofmap"] >= 0]
    a, b = res["n_lo"], res["n_lo"] + res["n_own"]
    rp = res["rowptr"]
    lo, hi = int(rp[a].item()), int(rp[b].item())
    rows_local = rp[a:b + 1] - lo
    ln = rows_local[1:] - rows_local[:-1]
    rr = torch.repeat_interleave(torch.arange(a, b, device=dev), ln)
    v, c = res["vals"][lo:hi], res["col"][lo:hi].long()
    got = torch.stack([v.sum(), (v * _key_weight(keys[rr], 7919) * _key_weight(keys[c], 104729)).sum(),
                       (v * _key_weight(keys[rr], 611953) * _key_weight(keys[c], 15485863)).sum(),
                       torch.tensor(float(hi - lo), dtype=torch.float64, device=dev)])
    # reference: own elements only, no halo, no merge
    ne = p.ne_own
    loc = CutAssembly(p.dim, p.el_vtx[:ne].contiguous(), p.vtx_xyz, p.el_dofs[:ne].contiguous(), p.ndof_loc, p.fe_order,
                      ctx=asm.be.ctx).step(phi)
    k2 = p.gid[loc["dofmap"] >= 0]
    ref = _block_checksums(loc["rowptr"], loc["col"], loc["vals"], k2)
    tot = torch.cat([got, ref])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, group=group)
    g, r = tot[:4].tolist(), tot[4:].tolist()
    scale = max(r[3], 1e-300)
    rel = [abs(g[i] - r[i]) / scale for i in range(3)]
    return {"ok": bool(max(rel) < rtol), "rel_diff": rel, "sum_entries": g[0], "sum_entries_ref": r[0], "nnz_owned_total": int(g[3]),
            "sum_abs_ref": r[3], "tolerance": rtol,
            "what": "sum over ranks of three linear functionals (weights from the global dof keys) of the owned rows of the merged "
                    "matrix vs the same functionals of every rank's own elements assembled alone (no halo, no merge)"}


# ------------------------------------------------------------------ CUDA compute backend
class GpuBackend:
    """drives libxgpu.so (C ABI) — there is no CPU fallback"""

    fused = True   # one-row-pass CSR (xgpu_csr_plan_rows + xgpu_csr_assemble_fused) where the shapes allow it
    kernel_path = True   # DistributedCutAssembly._step_kernels (subclasses with other element kernels: generic step)

    def __init__(self, ctx=None):
        from . import ops
        self.ops = ops
        self.ctx = ctx or ops.default_ctx()

    def local_assembly(self, part, phi, l_neg, l_if, l_halo, mark=lambda name: None, after_restrict=None, values=None):
        """the single-GPU sequence of pipeline.CutAssembly._step over the rank's own batches + the halo batch (sparsity
        only): Restrict + dof rows + transposed table in one sequence (xgpu_restrict_plan), the cut-element branch (cut
        quadrature + cut element matrices) on the side stream, count pass, 7 coefficients per uncut element, one row pass.
        returns dofmap, ndof, rowptr, col, vals"""
        from .pipeline import TERMS_IF, TERMS_VOL, _side_context
        ops, ctx, dim, p = self.ops, self.ctx, part.dim, part.fe_order
        els, xyz, dev = part.el_vtx, part.vtx_xyz, part.el_vtx.device
        order = 2 * p
        nd = ops.h1_ndof(dim, p)
        l_neg, l_if, l_halo = l_neg.contiguous(), l_if.contiguous(), l_halo.contiguous()
        if p > 2 or not self.fused:
            raise NotImplementedError("kernel path: orders 1, 2")
        main = torch.cuda.current_stream(dev)
        if os.environ.get("XGPU_CONCURRENT", "0") == "1":   # the cut-element branch on a low-priority side stream: not steadier,
            side, ctx2, _, _ = _side_context(dev)               # not faster than one stream (profiles/r02_notes.md)
        else:
            side, ctx2 = main, ctx
        m_if = torch.empty((l_if.numel(), nd, nd), dtype=torch.float64, device=dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if l_if.numel():
                # cut rules made inside the element-matrix kernel (no count pass / host wait, no stored rules)
                ops.elmat_cut_pair(dim, p, l_if, els, xyz, phi, ops.NEG, order, TERMS_VOL, ops.IF, order, TERMS_IF, out=m_if, ctx=ctx2)
            m_if.record_stream(side)
        mark("cut_branch_enqueued")
        lists = [l_neg, l_if, l_halo]
        keep = [i for i, l in enumerate(lists) if l.numel() > 0]
        if not keep:   # nothing active on this rank (not even in the halo): an empty local matrix
            main.wait_stream(side)
            dofmap = torch.full((part.ndof_loc,), -1, dtype=torch.int32, device=dev)
            return (dofmap, 0, torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int32, device=dev),
                    torch.zeros(0, dtype=torch.float64, device=dev))
        dofmap, ndof, tabs, plan = ops.restrict_plan(part.el_dofs, [lists[i] for i in keep], part.ndof_loc, ctx=ctx)
        if after_restrict is not None:
            after_restrict(dofmap)   # enqueued before the count pass (whose host synchronisation completes it)
        mark("restrict_plan")
        ops.csr_count_rows(plan)
        mark("csr_count_pass")
        m_neg = ops.affine_coefs(dim, p, l_neg, els, xyz, order, TERMS_VOL, ctx=ctx)
        main.wait_stream(side)
        ctx2.raise_on_geometry_error()
        mark("elmat")
        mats = [m_neg, m_if, None]
        batches = [(tabs[k], mats[i]) for k, i in enumerate(keep)]
        if values is None:
            rowptr, col, vals = ops.csr_assemble_fused(plan, batches)
        else:   # the caller splits the values pass (rows to send first, exchange under the rows it owns)
            rowptr, col, vals = values(plan, batches, dofmap, ndof)
        mark("csr_rows_fused")
        return dofmap, ndof, rowptr, col, vals

    def classify(self, el_vtx, phi):
        ops, ctx = self.ops, self.ctx
        dim, ne = el_vtx.shape[1] - 1, el_vtx.shape[0]
        bn, bp, bi, cr = ops.classify(dim, el_vtx, phi, ctx=ctx)
        ctx.raise_on_geometry_error()
        return ops.bitset_to_list(bn, ne, ctx=ctx), ops.bitset_to_list(bi, ne, ctx=ctx), (bn, bp, bi, cr)

    def element_activity(self, cls, elems):
        """bool per listed element: HASNEG bit (NEG | IF) looked up in the BitArray words"""
        bn, bp, bi, cr = cls
        e = elems.long()
        w = (bn | bi)[e >> 5]
        return ((w >> (e & 31).to(torch.int32)) & 1).bool()

    def active_dofs(self, part, cls, l_neg, l_if):
        ops, ctx = self.ops, self.ctx
        bn, bp, bi, cr = cls
        has = ops.bitset_op(ops.BIT_OR, bn, bi, part.el_vtx.shape[0], ctx=ctx)
        act = ops.dofs_of_elements(part.el_dofs, has, part.ndof_loc, ctx=ctx)
        return ops.bitset_rank(act, part.ndof_loc, ctx=ctx)

    def gather_rows(self, el_dofs, lst, dofmap):
        return self.ops.gather_rows(el_dofs, lst.contiguous(), dofmap, ctx=self.ctx)

    def elmats(self, part, phi, l_neg, l_if):
        from .pipeline import TERMS_IF, TERMS_VOL
        ops, ctx, dim, p = self.ops, self.ctx, part.dim, part.fe_order
        els, xyz = part.el_vtx, part.vtx_xyz
        order = 2 * p
        l_neg, l_if = l_neg.contiguous(), l_if.contiguous()
        if p <= 2 and self.fused:   # procedural batch: 7 coefficients per uncut element instead of its matrix
            m_neg = ops.affine_coefs(dim, p, l_neg, els, xyz, order, TERMS_VOL, ctx=ctx)
        else:
            m_neg = ops.elmat_std(dim, p, l_neg, els, xyz, phi, order, TERMS_VOL, ctx=ctx)
        ra = ops.cutquad(dim, l_if, els, xyz, phi, ops.NEG, order, ctx=ctx)
        rb = ops.cutquad(dim, l_if, els, xyz, phi, ops.IF, order, ctx=ctx)
        m_if = ops.elmat_rule_pair(dim, p, l_if, els, xyz, phi, ra, TERMS_VOL, rb, TERMS_IF, ctx=ctx)
        ctx.raise_on_geometry_error()
        return m_neg, m_if

    def rows_offsets(self, rowptr, rows):
        return self.ops.rows_offsets(rowptr, rows, ctx=self.ctx)

    def rows_pack(self, rowptr, rows, off, vals, total):
        return self.ops.rows_pack(rowptr, rows, off, vals, total, ctx=self.ctx)

    def rows_add(self, rowptr, rows, off, buf, vals):
        return self.ops.rows_add(rowptr, rows, off, buf, vals, ctx=self.ctx)

    def csr(self, ndof, batches):
        ops = self.ops
        tables = [b[0] for b in batches]
        if self.fused:
            from ._lib import XgpuError
            try:
                plan = ops.csr_plan_rows(ndof, tables, ctx=self.ctx)
                return ops.csr_assemble_fused(plan, batches)
            except XgpuError as e:
                if getattr(e, "code", 0) != ops.XGPU_ERR_UNSUPPORTED or any(isinstance(b[1], ops.AffineCoefs) for b in batches):
                    raise
        plan = ops.csr_plan(ndof, tables, ctx=self.ctx)
        return ops.csr_assemble(plan, batches)


class MlsetGpuBackend(GpuBackend):
    """config 4 of BASELINE.json: a multi-level-set domain (cutint/mlsetintegration.cpp:6-186) assembled element-partitioned.
    `phi` of DistributedCutAssembly.step is the M x nv_local matrix of nodal level-set values.  Active elements = elements
    with contribution (xfem/cutinfo.cpp:306-334: positive weight sum of the order -1 multi-level-set rule); every active
    element gets its rule from xgpu_cutquad_ml_* (uncut relevant elements: the standard rule), forms: grad.grad + u v."""

    kernel_path = False

    def __init__(self, dts, ctx=None):
        super().__init__(ctx)
        self.dts = [tuple(int(d) for d in t) for t in dts]

    def classify(self, el_vtx, phis):
        ops, ctx = self.ops, self.ctx
        dim, ne = el_vtx.shape[1] - 1, el_vtx.shape[0]
        off, pts, wts = ops.cutquad_ml(dim, None, el_vtx, self._xyz, phis, self.dts, -1, ctx=ctx)
        ctx.raise_on_geometry_error()
        seg = ops.rule_element_sums(off, wts, ctx=ctx)           # weight sums in rule order (cutinfo.cpp:322-325)
        words, zero, zero2 = (ops.new_bits(ne, el_vtx.device) for _ in range(3))
        ops.classify_volumes(seg, torch.zeros_like(seg), (words, zero, zero2, None), ctx=ctx)   # "NEG" bit = (sum > 0)
        zero.zero_()                                              # no POS / IF sets in this backend: activity = the one bitset
        l_act = ops.bitset_to_list(words, ne, ctx=ctx)
        return l_act, l_act[:0], (words, zero, zero.clone(), None)

    def elmats(self, part, phis, l_act, l_none):
        from .pipeline import TERMS_VOL
        ops, ctx, dim, p = self.ops, self.ctx, part.dim, part.fe_order
        els, xyz = part.el_vtx, part.vtx_xyz
        l_act = l_act.contiguous()
        nd = ops.h1_ndof(dim, p)
        m = torch.zeros((l_act.numel(), nd, nd), dtype=torch.float64, device=els.device)
        if l_act.numel():
            off, q, w = ops.cutquad_ml(dim, l_act, els, xyz, phis, self.dts, 2 * p, ctx=ctx)
            ctx.raise_on_geometry_error()
            if w.numel():
                ops.elmat_rule(dim, p, l_act, els, xyz, phis[0].contiguous(), off, q, w, TERMS_VOL, out=m, ctx=ctx)
        return m, torch.zeros((0, nd, nd), dtype=torch.float64, device=els.device)

    def bind(self, part):
        self._xyz = part.vtx_xyz
        return self


class SpaceTimeGpuBackend(GpuBackend):
    """config 5 of BASELINE.json: one space-time slab (P1/P2 in space x nodal in time) assembled element-partitioned.
    `phi` of DistributedCutAssembly.step is the (k_l+1) x nv_local matrix of nodal level-set values (time-major).  The
    partition carries the SPACE-TIME dof table (make_slab_partition(..., time_fe_order=k): dof (j, i) = j * ndof_space + i,
    global key = space key * (k+1) + j, owner of the space dof).  Classification: CutInfo on the slab (part volumes of the
    order-0 space-time rules, xfem/cutinfo.cpp:90-124); NEG and IF elements get their space-time rule from
    xgpu_cutquad_st_* (uncut elements: the tensor-product standard rule), forms: u v + dt(u) v + grad u . grad v."""

    TERMS = [(0, 0, 1.0, 0), (3, 0, 1.0, 0), (1, 1, 1.0, 0)]   # (op_trial, op_test, coef, hpow); 3 = XGPU_OP_DT
    kernel_path = False

    def __init__(self, space_order, time_fe_order, lset_time_order, time_order, ctx=None):
        super().__init__(ctx)
        self.p, self.kt, self.kl, self.time_order = int(space_order), int(time_fe_order), int(lset_time_order), int(time_order)

    def bind(self, part):
        self._xyz = part.vtx_xyz
        return self

    def classify(self, el_vtx, phi_st):
        ops, ctx = self.ops, self.ctx
        dim, ne = el_vtx.shape[1] - 1, el_vtx.shape[0]
        bn, bp, bi, _ = ops.classify_st(dim, el_vtx, phi_st, self.kl, 0, want_ratio=False, ctx=ctx)
        ctx.raise_on_geometry_error()
        return ops.bitset_to_list(bn, ne, ctx=ctx), ops.bitset_to_list(bi, ne, ctx=ctx), (bn, bp, bi, None)

    def elmats(self, part, phi_st, l_neg, l_if):
        ops, ctx, dim = self.ops, self.ctx, part.dim
        els, xyz = part.el_vtx, part.vtx_xyz
        out = []
        nd = ops.h1_ndof(dim, self.p) * (self.kt + 1)
        for l in (l_neg.contiguous(), l_if.contiguous()):
            m = torch.zeros((l.numel(), nd, nd), dtype=torch.float64, device=els.device)
            if l.numel():
                off, q, tv, w = ops.cutquad_st(dim, l, els, xyz, phi_st, self.kl, ops.NEG, self.time_order, 2 * self.p, ctx=ctx)
                if w.numel():
                    ops.elmat_st(dim, self.p, self.kt, l, els, xyz, off, q, tv, w, self.TERMS, out=m, ctx=ctx)
            out.append(m)
        ctx.raise_on_geometry_error()
        return out[0], out[1]


# ------------------------------------------------------------------ slab partition of the structured chain (input producer)
def structured_dof_keys(mesh, vg, nv1, NV, p, el_gid, v_owner=None, el_owner=None):
    """Global keys (any total order all ranks agree on) of the H1(order p) dofs of a sub-mesh of the structured Kuhn
    chain, in the order fes.H1 numbers them (vertices, (p-1) per edge, (p-1)(p-2)/2 per face, cell bubbles), computed
    WITHOUT communication from the global vertex ids `vg`: the vertices of an edge / face lie in one cell, so their id
    differences come from a set of 7 offsets -> key = smallest vertex id * 8 (64) + offset code(s).  Owner of an entity:
    the largest owner of its vertices; cell dofs belong to the element's rank.  Returns (gid int64, owner int64 or None)."""
    dev = vg.device
    offs = torch.tensor(sorted([1, nv1, nv1 + 1, nv1 * nv1, nv1 * nv1 + 1, nv1 * nv1 + nv1, nv1 * nv1 + nv1 + 1]),
                        dtype=torch.int64, device=dev)

    def code(d):
        c = torch.searchsorted(offs, d)
        assert bool((offs[c.clamp(max=6)] == d).all()), "not an edge / face of the structured Kuhn mesh"
        return c

    gid, owner = [vg], [v_owner]
    base = NV
    if p >= 2 and mesh.dim >= 2:
        ev, _ = mesh.edges()
        g = torch.sort(vg[ev.long()], dim=1).values
        ekey = g[:, 0] * 8 + code(g[:, 1] - g[:, 0])
        k = torch.arange(p - 1, device=dev, dtype=torch.int64)
        gid.append((base + ekey.reshape(-1, 1) * (p - 1) + k).reshape(-1))
        if v_owner is not None:
            owner.append(v_owner[ev.long()].max(dim=1).values.repeat_interleave(p - 1))
        base += 8 * NV * (p - 1)
    if p >= 3 and mesh.dim == 3:
        nfd = (p - 1) * (p - 2) // 2
        fv, _ = mesh.faces()
        g = torch.sort(vg[fv.long()], dim=1).values
        fkey = g[:, 0] * 64 + code(g[:, 1] - g[:, 0]) * 8 + code(g[:, 2] - g[:, 0])
        k = torch.arange(nfd, device=dev, dtype=torch.int64)
        gid.append((base + fkey.reshape(-1, 1) * nfd + k).reshape(-1))
        if v_owner is not None:
            owner.append(v_owner[fv.long()].max(dim=1).values.repeat_interleave(nfd))
        base += 64 * NV * nfd
        ncd = (p - 1) * (p - 2) * (p - 3) // 6
        if ncd:
            k = torch.arange(ncd, device=dev, dtype=torch.int64)
            gid.append((base + el_gid.reshape(-1, 1) * ncd + k).reshape(-1))
            if v_owner is not None:
                owner.append(el_owner.repeat_interleave(ncd))
    gid = torch.cat(gid)
    assert int(gid.max().item()) < (1 << 52)
    return gid, (torch.cat(owner) if v_owner is not None else None)


def balanced_layers(n, world, radius=0.4, alpha=0.12):
    """layer bounds (world + 1 ascending layer indices 0..n) of the strong split of the unit cube with the centred sphere:
    contiguous z-slabs of (nearly) equal WORK, work of a layer = alpha * (all its elements: classification) + its elements
    inside the sphere (assembly), from the analytic cross-section — SURVEY.md §8(e): the cut elements cluster in few
    slabs, so the element ranges are weighted, not equal."""
    import math
    w = []
    for k in range(n):
        z = (k + 0.5) / n - 0.5
        rr = radius * radius - z * z
        w.append(alpha + (math.pi * rr if rr > 0 else 0.0))
    tot = sum(w)
    bounds, acc, nxt = [0], 0.0, 1
    for k in range(n):
        acc += w[k]
        while nxt < world and acc >= tot * nxt / world and len(bounds) < world:
            bounds.append(max(k + 1, bounds[-1] + 1))
            nxt += 1
    while len(bounds) < world:
        bounds.append(bounds[-1] + 1)
    bounds.append(n)
    for i in range(world, 0, -1):   # every rank keeps at least one layer
        bounds[i - 1] = min(bounds[i - 1], bounds[i] - 1)
    return bounds


def make_slab_partition(n, world, rank, device, fe_order=2, radius=0.4, time_fe_order=None, scaling="weak", bounds=None):
    """scaling="weak": chain [0,1]^2 x [0,world] of `world` unit cubes, n^3 x 6 Kuhn tets each; rank r owns cube r
    (a contiguous element range of the global mesh) plus one halo layer of cells per neighbour.
    Level set: spheres of `radius` centred at (1/2, 1/2, j), j = 0..world for world > 1 (centres ON
    the slab interfaces, so the interface rows are active and each rank sees two half spheres = the
    config-2 element counts), the centred sphere for world == 1.
    scaling="strong": THE unit cube (n^3 x 6 tets, centred sphere = BASELINE config 2) cut into `world` z-slabs of whole
    cell layers, `bounds` (default: balanced_layers) = the world + 1 layer indices; rank r owns layers bounds[r]..bounds[r+1]."""
    from . import meshes
    from .fes import H1
    from .pipeline import sphere_levelset
    if fe_order > 4:
        raise NotImplementedError("orders 1..4")
    lo, hi = (1 if rank > 0 else 0), (1 if rank < world - 1 else 0)
    strong = scaling == "strong" and world > 1
    if strong:
        bounds = list(bounds) if bounds is not None else balanced_layers(n, world, radius)
        assert len(bounds) == world + 1 and bounds[0] == 0 and bounds[-1] == n and all(b1 > b0 for b0, b1 in zip(bounds, bounds[1:]))
        nzo, kbase, ntot = bounds[rank + 1] - bounds[rank], bounds[rank], n      # own layers, first own layer, layers of the global mesh
    else:
        nzo, kbase, ntot = n, rank * n, n * world
    nzl = nzo + lo + hi
    mesh = meshes.MakeStructured3DMesh(nx=n, ny=n, nz=nzl, device=device)
    # physical coordinates: local layers cover the global cell layers [kbase - lo, kbase + nzo + hi)
    k0 = kbase - lo
    pts = mesh.points.clone()
    nv1 = n + 1
    # z = K / n from the integer layer index K: bit-identical on every rank that holds the vertex
    kk = torch.arange(mesh.nv, device=device, dtype=torch.int64) // (nv1 * nv1)
    pts[:, 2] = (kk + k0).to(torch.float64) / n
    # own elements first (cells lo .. lo+n-1), then halo
    ne_layer = 6 * n * n
    cell_k = torch.arange(mesh.ne, device=device) // ne_layer
    own = (cell_k >= lo) & (cell_k < lo + nzo)
    perm = torch.cat([torch.nonzero(own).reshape(-1), torch.nonzero(~own).reshape(-1)])
    el_vtx = mesh.elements[perm].contiguous()
    el_rank = torch.where(cell_k[perm] < lo, rank - 1, torch.where(cell_k[perm] >= lo + nzo, rank + 1, rank)).to(torch.int32)
    lmesh = meshes.Mesh(pts, el_vtx)
    V = H1(lmesh, order=fe_order)
    # global keys: vertex gid = i + (n+1)(j + (n+1)K); edge key from its two global vertex ids
    vloc = torch.arange(lmesh.nv, device=device, dtype=torch.int64)
    K = vloc // (nv1 * nv1) + k0
    vg = vloc % (nv1 * nv1) + nv1 * nv1 * K
    NV = nv1 * nv1 * (ntot + 1)
    if strong:   # a vertex layer K belongs to the rank whose slab it closes from above: bounds[r] < K <= bounds[r+1] (K = 0: rank 0)
        bt = torch.tensor(bounds[1:], device=device, dtype=torch.int64)
        v_owner = torch.searchsorted(bt, K.clamp(min=0)).clamp(max=world - 1)
    else:
        v_owner = torch.where(K <= 0, torch.zeros_like(K), (K + n - 1) // n - 1).clamp(max=world - 1)
    e_loc = perm.to(torch.int64)                                   # local element id of the unpermuted local mesh
    el_gid = e_loc + 6 * n * n * k0                                # global element id (6 * global cell + path)
    gid, ent_owner = structured_dof_keys(lmesh, vg, nv1, NV, fe_order, el_gid, v_owner, el_rank.to(torch.int64))
    owner = ent_owner.to(torch.int32)
    halo_for = {}
    ne_own = 6 * n * n * nzo
    own_cells_k = cell_k[perm][:ne_own]
    if rank > 0:
        halo_for[rank - 1] = torch.nonzero(own_cells_k == lo).reshape(-1).to(torch.int32)
    if rank < world - 1:
        halo_for[rank + 1] = torch.nonzero(own_cells_k == lo + nzo - 1).reshape(-1).to(torch.int32)
    centers = [(.5, .5, .5)] if (world == 1 or strong) else [(.5, .5, float(j)) for j in range(world + 1)]
    phi = sphere_levelset(pts, centers, radius)
    if time_fe_order is not None:
        # space-time slab (config 5): time-major dof table, key = space key * nt + time node; the level set is the
        # sphere chain moving in y, nodal in time on the same time nodes: (nt x nv_local)
        import math
        from .spacetime import ScalarTimeFE
        nt = time_fe_order + 1
        el_dofs = torch.cat([V.el_dofs + j * V.ndof for j in range(nt)], dim=1).to(torch.int32).contiguous()
        gid_st = torch.cat([gid * nt + j for j in range(nt)])
        owner_st = owner.repeat(nt)
        phis = []
        for t in ScalarTimeFE(time_fe_order).nodes():
            cy = 0.5 + 0.1 * math.sin(2 * math.pi * 0.25 * t)
            phis.append(sphere_levelset(pts, [(c[0], cy, c[2]) for c in centers], radius))
        part = LocalPartition(3, el_vtx, ne_own, el_rank, pts, el_dofs, V.ndof * nt, gid_st, owner_st, halo_for, fe_order, rank)
        return part, torch.stack(phis).contiguous()
    part = LocalPartition(3, el_vtx, ne_own, el_rank, pts, V.el_dofs, V.ndof, gid, owner, halo_for, fe_order, rank)
    return part, phi
