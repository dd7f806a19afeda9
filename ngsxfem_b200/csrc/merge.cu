// merge.cu — interface-row exchange of the element-partitioned assembly (the reference's analogue is
// NGSolve's ParallelDofs cumulation of a distributed matrix [third party]): the values of a list of CSR rows are
// packed into one contiguous message buffer, and a received buffer is added into the owner's rows.  Both sides of an
// interface hold IDENTICAL column orders for these rows (the numbering is (owner, gid) ordered), so only values travel.
// One warp per listed row; rows[i] < 0 marks an inactive candidate (length 0).  The NCCL send/recv itself is issued
// by the host layer (torch.distributed) on these buffers.
#include "common.cuh"

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__global__ void k_rows_len(int64_t n, const long long* __restrict__ rowptr, const int32_t* __restrict__ rows,
                           int32_t* __restrict__ len)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rows[i];
  len[i] = (r >= 0) ? (int32_t)(rowptr[r + 1] - rowptr[r]) : 0;
}

template <bool ADD>
__global__ void __launch_bounds__(256) k_rows_move(int64_t n, const long long* __restrict__ rowptr,
                                                   const int32_t* __restrict__ rows, const long long* __restrict__ offsets,
                                                   double* __restrict__ vals, double* __restrict__ buf)
{
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w0; i < n; i += nw) {
    const int r = rows[i];
    if (r < 0) continue;
    const long long a = rowptr[r], o = offsets[i];
    const int len = (int)(rowptr[r + 1] - a);
    for (int k = lane; k < len; k += 32) {
      if (ADD) vals[a + k] += buf[o + k];
      else buf[o + k] = vals[a + k];
    }
  }
}

// ---- static interface tables: activity flags, row lists, index messages, global numbering (no host round trips)
__global__ void k_iface_marks(int64_t npairs, const unsigned int* __restrict__ has, const int32_t* __restrict__ pair_elem,
                              const int32_t* __restrict__ pair_gpos, unsigned char* __restrict__ flags)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npairs) return;
  const int e = pair_elem[i];
  if ((has[e >> 5] >> (e & 31)) & 1u) flags[pair_gpos[i]] = 1;   // plain byte stores of the same value: no atomics needed
}

__global__ void k_iface_rows(int64_t n, const int32_t* __restrict__ cand, const unsigned char* __restrict__ flags,
                             const int32_t* __restrict__ dofmap, int32_t* __restrict__ rows)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rows[i] = flags[i] ? dofmap[cand[i]] : -1;
}

__global__ void k_iface_index(int64_t n, const int32_t* __restrict__ cand, const unsigned char* __restrict__ flags,
                              const int32_t* __restrict__ dofmap, long long n_lo, long long* __restrict__ idx)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  idx[i] = flags[i] ? (long long)dofmap[cand[i]] - n_lo : -1;
}

__global__ void k_l2g_own(int64_t ndof, long long n_lo, long long n_own, const long long* __restrict__ counts, int rank,
                          long long* __restrict__ l2g)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ndof) return;
  long long off = 0;
  for (int k = 0; k < rank; k++) off += counts[k];
  l2g[i] = (i >= n_lo && i < n_lo + n_own) ? off + (i - n_lo) : -1;
}

__global__ void k_l2g_ghosts(int64_t n, const int32_t* __restrict__ dofmap_range, const long long* __restrict__ rbuf,
                             const long long* __restrict__ counts, int q, long long* __restrict__ l2g)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int loc = dofmap_range[i];
  if (loc < 0) return;
  long long off = 0;
  for (int k = 0; k < q; k++) off += counts[k];
  const long long r = rbuf[i];
  l2g[loc] = r >= 0 ? r + off : r;
}

}  // namespace

extern "C" int xgpu_iface_marks(xgpu_ctx* ctx, const uint32_t* has_bits, int64_t npairs, const int32_t* pair_elem,
                                const int32_t* pair_gpos, int64_t nflags, uint8_t* flags)
{
  XGPU_NVTX(ctx, "xgpu_iface_marks");
  if (!ctx || npairs < 0 || nflags < 0 || (nflags > 0 && !flags)) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_iface_marks: bad argument");
  if (nflags == 0) return XGPU_OK;
  XGPU_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)nflags, ctx->stream));
  if (npairs == 0) return XGPU_OK;
  k_iface_marks<<<nblk(npairs, 256), 256, 0, ctx->stream>>>(npairs, has_bits, pair_elem, pair_gpos, flags);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_iface_rows(xgpu_ctx* ctx, int64_t n, const int32_t* cand, const uint8_t* flags, const int32_t* dofmap,
                               int32_t* rows)
{
  if (!ctx || n < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_iface_rows: bad argument");
  if (n == 0) return XGPU_OK;
  k_iface_rows<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, cand, flags, dofmap, rows);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_iface_index(xgpu_ctx* ctx, int64_t n, const int32_t* cand, const uint8_t* flags, const int32_t* dofmap,
                                int64_t n_lo, int64_t* idx)
{
  if (!ctx || n < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_iface_index: bad argument");
  if (n == 0) return XGPU_OK;
  k_iface_index<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, cand, flags, dofmap, (long long)n_lo, (long long*)idx);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_l2g_own(xgpu_ctx* ctx, int64_t ndof, int64_t n_lo, int64_t n_own, const int64_t* counts, int rank, int64_t* l2g)
{
  if (!ctx || ndof < 0 || rank < 0 || !counts) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_l2g_own: bad argument");
  if (ndof == 0) return XGPU_OK;
  k_l2g_own<<<nblk(ndof, 256), 256, 0, ctx->stream>>>(ndof, n_lo, n_own, (const long long*)counts, rank, (long long*)l2g);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_l2g_ghosts(xgpu_ctx* ctx, int64_t n, const int32_t* dofmap_range, const int64_t* rbuf, const int64_t* counts,
                               int q, int64_t* l2g)
{
  if (!ctx || n < 0 || q < 0 || !counts) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_l2g_ghosts: bad argument");
  if (n == 0) return XGPU_OK;
  k_l2g_ghosts<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, dofmap_range, (const long long*)rbuf, (const long long*)counts, q, (long long*)l2g);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_rows_offsets(xgpu_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* rows, int64_t* offsets,
                                 int64_t* total_host)
{
  if (!ctx || n < 0 || !offsets) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_rows_offsets: bad argument");
  if (total_host) *total_host = 0;
  void* p;
  int rc = xgpu_scratch(ctx, 0, sizeof(int32_t) * (n + 1), &p);
  if (rc) return rc;
  if (n > 0) {
    k_rows_len<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, (const long long*)rowptr, rows, (int32_t*)p);
    XGPU_LAUNCH_CHECK(ctx);
  }
  rc = xgpu_exclusive_scan_i32(ctx, (const int32_t*)p, n, offsets);
  if (rc) return rc;
  if (total_host) {
    XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, offsets + n, sizeof(int64_t), ctx->stream));
    XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *total_host = ctx->h_pinned[0];
  }
  return XGPU_OK;
}

extern "C" int xgpu_rows_pack(xgpu_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* rows, const int64_t* offsets,
                              const double* vals, double* out)
{
  XGPU_NVTX(ctx, "xgpu_rows_pack");
  if (!ctx || n < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_rows_pack: bad argument");
  if (n == 0) return XGPU_OK;
  unsigned grid = nblk(n, 8);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  k_rows_move<false><<<grid, 256, 0, ctx->stream>>>(n, (const long long*)rowptr, rows, (const long long*)offsets,
                                                    const_cast<double*>(vals), out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_rows_add(xgpu_ctx* ctx, int64_t n, const int64_t* rowptr, const int32_t* rows, const int64_t* offsets,
                             const double* in, double* vals)
{
  XGPU_NVTX(ctx, "xgpu_rows_add");
  if (!ctx || n < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_rows_add: bad argument");
  if (n == 0) return XGPU_OK;
  unsigned grid = nblk(n, 8);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  k_rows_move<true><<<grid, 256, 0, ctx->stream>>>(n, (const long long*)rowptr, rows, (const long long*)offsets, vals,
                                                   const_cast<double*>(in));
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}
