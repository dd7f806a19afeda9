"""Checker compute backend for the distributed host logic (CPU, gloo): the same three operations
as ngsxfem_b200.distributed.GpuBackend, served by the oracle.  Test infrastructure only."""
import numpy as np
import torch

from oracle import xoracle as xo
from oracle.pipeline_ref import TERMS_IF, TERMS_VOL, bits_to_list


class OracleBackend:
    def classify(self, el_vtx, phi):
        els, ph = el_vtx.numpy(), phi.numpy()
        dim, ne = els.shape[1] - 1, els.shape[0]
        bn, bp, bi, cr, err, _ = xo.cutinfo_update(dim, els, ph)
        assert err == 0
        return torch.from_numpy(bits_to_list(bn, ne)), torch.from_numpy(bits_to_list(bi, ne)), (bn, bp, bi, cr)

    def element_activity(self, cls, elems):
        bn, bp, bi, cr = cls
        bits = np.unpackbits(np.asarray(bn, np.uint8) | np.asarray(bi, np.uint8), bitorder="little")
        return torch.from_numpy(bits[elems.numpy()].astype(bool))

    def active_dofs(self, part, cls, l_neg, l_if):
        act = torch.zeros(part.ndof_loc, dtype=torch.bool)
        act[part.el_dofs[torch.cat([l_neg, l_if]).long()].reshape(-1).long()] = True
        dofmap = torch.full((part.ndof_loc,), -1, dtype=torch.int32)
        n = int(act.sum())
        dofmap[act] = torch.arange(n, dtype=torch.int32)
        return dofmap, n

    def gather_rows(self, el_dofs, lst, dofmap):
        return dofmap[el_dofs[lst.long()].long()].to(torch.int32)

    def elmats(self, part, phi, l_neg, l_if):
        els, xyz, ph = part.el_vtx.numpy(), part.vtx_xyz.numpy(), phi.numpy()
        dim, p = part.dim, part.fe_order
        order = 2 * p
        ln, li = l_neg.numpy().astype(np.int32), l_if.numpy().astype(np.int32)
        m_neg = xo.batch_cut_elmats(dim, p, ln, els, xyz, ph, xo.NEG, order, TERMS_VOL)
        m_if = xo.batch_cut_elmats(dim, p, li, els, xyz, ph, xo.NEG, order, TERMS_VOL)
        m_if = m_if + xo.batch_cut_elmats(dim, p, li, els, xyz, ph, xo.IF, order, TERMS_IF)
        return torch.from_numpy(m_neg), torch.from_numpy(m_if)

    @staticmethod
    def _positions(rowptr, rows, off):
        live = rows >= 0
        r = rows[live].long()
        ln = rowptr[r + 1] - rowptr[r]
        src = torch.repeat_interleave(rowptr[r], ln) + (torch.arange(int(ln.sum())) - torch.repeat_interleave(torch.cumsum(ln, 0) - ln, ln))
        dst = torch.repeat_interleave(off[:-1][live], ln) + (torch.arange(int(ln.sum())) - torch.repeat_interleave(torch.cumsum(ln, 0) - ln, ln))
        return src, dst

    def rows_offsets(self, rowptr, rows):
        ln = torch.zeros(rows.numel(), dtype=torch.int64)
        live = rows >= 0
        r = rows[live].long()
        ln[live] = rowptr[r + 1] - rowptr[r]
        return torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(ln, 0)])

    def rows_pack(self, rowptr, rows, off, vals, total):
        out = torch.zeros(total, dtype=vals.dtype)
        src, dst = self._positions(rowptr, rows, off)
        out[dst] = vals[src]
        return out

    def rows_add(self, rowptr, rows, off, buf, vals):
        src, dst = self._positions(rowptr, rows, off)
        vals[src] += buf[dst]
        return vals

    def csr(self, ndof, batches):
        W = max(t.shape[1] for t, _ in batches)
        tabs, mats = [], []
        for t, m in batches:
            tn = np.full((t.shape[0], W), -1, np.int32)
            tn[:, :t.shape[1]] = t.numpy()
            mn = np.zeros((t.shape[0], W, W))
            if m is not None:
                mn[:, :t.shape[1], :t.shape[1]] = m.numpy()
            tabs.append(tn); mats.append(mn)
        tab, mt = np.concatenate(tabs), np.concatenate(mats)
        rowptr, col = xo.graph(ndof, tab)
        vals = xo.scatter(ndof, tab, mt, rowptr, col)
        return torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(vals)


class SpaceTimeOracleBackend(OracleBackend):
    """config 5 (one space-time slab) served by the oracle: slab classification from the part volumes of the order-0
    space-time rules (xfem/cutinfo.cpp:90-124), space-time element matrices u v + dt(u) v + grad u . grad v on NEG —
    the CPU counterpart of ngsxfem_b200.distributed.SpaceTimeGpuBackend"""
    TERMS = [(0, 0, 1.0, 0), (3, 0, 1.0, 0), (1, 1, 1.0, 0)]

    def __init__(self, space_order, time_fe_order, lset_time_order, time_order):
        self.p, self.kt, self.kl, self.time_order = space_order, time_fe_order, lset_time_order, time_order

    def bind(self, part):
        self._xyz = part.vtx_xyz
        return self

    @staticmethod
    def slab_classes(els, xyz, L, kl):
        ne = els.shape[0]
        neg = np.zeros(ne, bool); pos = np.zeros(ne, bool); cut = np.zeros(ne, bool)
        for el in range(ne):
            T = xo.make_trafo(xyz[els[el]], els[el])
            pv = {}
            for np_ in (xo.POS, xo.NEG):
                _, _, w = xo.spacetime_cut_rule(T, L[:, els[el]], kl, np_, 0, 0)
                s = 0.0
                for x in w:
                    s += x
                pv[np_] = s
            if pv[xo.NEG] > 0:
                (cut if pv[xo.POS] > 0 else neg)[el] = True
            else:
                pos[el] = True
        return neg, pos, cut

    def classify(self, el_vtx, phi_st):
        els, L = el_vtx.numpy(), phi_st.numpy()
        neg, pos, cut = self.slab_classes(els, self._xyz.numpy(), L, self.kl)
        pk = lambda m: np.packbits(m, bitorder="little")
        return (torch.from_numpy(np.flatnonzero(neg).astype(np.int32)), torch.from_numpy(np.flatnonzero(cut).astype(np.int32)),
                (pk(neg), pk(pos), pk(cut), None))

    @staticmethod
    def slab_elmats(els, xyz, L, lst, p, kt, kl, time_order, terms):
        nd = xo.h1_ndof(els.shape[1] - 1, p) * (kt + 1)
        out = np.zeros((len(lst), nd, nd))
        for i, el in enumerate(lst):
            T = xo.make_trafo(xyz[els[el]], els[el])
            q, tv, w = xo.spacetime_cut_rule(T, L[:, els[el]], kl, xo.NEG, time_order, 2 * p)
            if len(w):
                xo.elmat_st_from_rule(T, p, kt, q, tv, w, terms, out=out[i])
        return out

    def elmats(self, part, phi_st, l_neg, l_if):
        els, xyz, L = part.el_vtx.numpy(), part.vtx_xyz.numpy(), phi_st.numpy()
        f = lambda l: torch.from_numpy(self.slab_elmats(els, xyz, L, l.numpy(), self.p, self.kt, self.kl, self.time_order, self.TERMS))
        return f(l_neg), f(l_if)
