"""Interface-row kernels of the multi-GPU merge (xgpu_rows_offsets / _pack / _add) on ONE GPU against the
torch restatement used by the gloo tests (tests/dist_backend.py): ragged rows, inactive candidates (row id < 0), empty
lists; the host total is read back through the mapped staging block."""
import ctypes as C

import pytest
import torch

from ngsxfem_b200 import _lib, ops
from dist_backend import OracleBackend

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("nrows,nsel,seed", [(50, 17, 0), (1000, 333, 1), (7, 0, 2), (64, 64, 3)])
def test_rows_offsets_pack_add_match_the_restatement(nrows, nsel, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0, 40, (nrows,), generator=g)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lens, 0)])
    nnz = int(rowptr[-1])
    vals = torch.rand(nnz, dtype=torch.float64, generator=g)
    rows = torch.randint(0, nrows, (nsel,), generator=g).to(torch.int32)
    if nsel > 3:
        rows[::3] = -1                      # inactive candidates: length 0
    ob = OracleBackend()
    off_ref = ob.rows_offsets(rowptr, rows)
    total = int(off_ref[-1])
    pack_ref = ob.rows_pack(rowptr, rows, off_ref, vals, total)
    d_rowptr, d_rows, d_vals = rowptr.to(DEV), rows.to(DEV), vals.to(DEV)
    off = ops.rows_offsets(d_rowptr, d_rows)
    assert torch.equal(off.cpu(), off_ref)
    # the same call with the host total (synchronises; read back without the copy engine)
    ctx = _lib.default_ctx()
    off2 = torch.empty_like(off)
    tot = C.c_int64(-1)
    ctx.check(ctx.lib.xgpu_rows_offsets(ctx.h, d_rows.numel(), _lib.ptr(d_rowptr), _lib.ptr(d_rows), _lib.ptr(off2), C.byref(tot)))
    assert tot.value == total and torch.equal(off2, off)
    pack = ops.rows_pack(d_rowptr, d_rows, off, d_vals, total)
    assert torch.equal(pack.cpu(), pack_ref)
    # add: every listed row receives the buffer once per occurrence, in list order (duplicates allowed on the CPU side
    # only through index_put accumulate=False semantics, so use distinct rows here)
    uniq = torch.unique(rows[rows >= 0]).to(torch.int32)
    off_u = ob.rows_offsets(rowptr, uniq)
    buf = torch.rand(int(off_u[-1]), dtype=torch.float64, generator=g)
    want = ob.rows_add(rowptr, uniq, off_u, buf, vals.clone())
    got = ops.rows_add(d_rowptr, uniq.to(DEV), ops.rows_offsets(d_rowptr, uniq.to(DEV)), buf.to(DEV), d_vals.clone())
    assert torch.equal(got.cpu(), want)
