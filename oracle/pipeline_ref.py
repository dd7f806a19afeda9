"""ORACLE (test infrastructure): the config-2 hot path on the CPU, the per-element call sequence
of the reference over the C restatement in libxoracle.so.

Used only as the checker of the CUDA pipeline (tests/, __graft_entry__.smoke()) and as the timed
CPU arm of bench.py (`cpu_baseline`, `--impl reference`; kind "port": ngsxfem itself cannot be
built here, every source needs NGSolve headers).  Never imported by the product package.
"""
import time

import numpy as np

from . import xoracle as xo

TERMS_VOL = [(1, 1, 1.0, 0), (0, 0, 1.0, 0)]
TERMS_IF = [(0, 0, 1.0, 0)]


def bits_to_list(bits, n):
    return np.flatnonzero(np.unpackbits(bits, bitorder="little")[:n]).astype(np.int32)


def oracle_step(dim, el_vtx, vtx_xyz, el_dofs, ndof_full, fe_order, phi, threads=False):
    """returns (result dict, seconds spent inside the restated hot-path functions)."""
    ne = el_vtx.shape[0]
    t = {}
    t0 = time.perf_counter()
    bn, bp, bi, cr, err, _ = xo.cutinfo_update(dim, el_vtx, phi)          # cutinfo.cpp:61-133
    t["classify"] = time.perf_counter() - t0
    assert err == 0
    l_neg, l_if = bits_to_list(bn, ne), bits_to_list(bi, ne)
    l_has = np.sort(np.concatenate([l_neg, l_if])).astype(np.int32)
    # Restrict: active dofs of the HASNEG elements, compressed ascending (restrictedfespace.cpp:23-51)
    act = np.zeros(ndof_full, dtype=bool)
    act[el_dofs[l_has].reshape(-1)] = True
    dofmap = np.full(ndof_full, -1, dtype=np.int32)
    ndof = int(act.sum())
    dofmap[act] = np.arange(ndof, dtype=np.int32)
    t_neg, t_if = dofmap[el_dofs[l_neg]], dofmap[el_dofs[l_if]]
    order = 2 * fe_order
    t0 = time.perf_counter()
    m_neg = xo.batch_cut_elmats(dim, fe_order, l_neg, el_vtx, vtx_xyz, phi, xo.NEG, order, TERMS_VOL)
    m_if = xo.batch_cut_elmats(dim, fe_order, l_if, el_vtx, vtx_xyz, phi, xo.NEG, order, TERMS_VOL)
    m_if += xo.batch_cut_elmats(dim, fe_order, l_if, el_vtx, vtx_xyz, phi, xo.IF, order, TERMS_IF)
    t["elmat"] = time.perf_counter() - t0
    table = np.concatenate([t_neg, t_if]).astype(np.int32)
    mats = np.concatenate([m_neg, m_if])
    t0 = time.perf_counter()
    rowptr, col = xo.graph(ndof, table, once=True)                                     # restrictedblf.cpp:142-163
    t["graph"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    vals = xo.scatter(ndof, table, mats, rowptr, col, threads=threads)      # AddElementMatrix
    t["scatter"] = time.perf_counter() - t0
    res = dict(ndof=ndof, nnz=len(col), n_neg=len(l_neg), n_if=len(l_if), units=len(l_neg) + len(l_if),
               rowptr=rowptr, col=col, vals=vals, bits=(bn, bp, bi), cut_ratio=cr, phases=t)
    return res, sum(t.values())
