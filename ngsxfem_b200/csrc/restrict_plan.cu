// restrict_plan.cu — Restrict(fes, active elements) and the transposed table of the restricted dof rows in ONE sequence.
// (RestrictedFESpace utils/restrictedfespace.cpp:23-73: active dofs = union of the dofs of the active elements, compressed
//  numbering; RestrictedBilinearForm::GetGraph utils/restrictedblf.cpp:142-163: the table of dof rows handed to MatrixGraph.)
//
// The separate entry points (xgpu_dofs_of_elements -> xgpu_bitset_rank -> xgpu_gather_rows -> xgpu_adjacency_plan) walk the
// dof rows of the active elements four times with two atomic passes (mark, histogram).  Here:
//   1. ONE atomic pass over the dof rows: occurrences per dof of the full space (a dof is active iff its count is > 0 —
//      the count is also the length of its segment in the transposed table);
//   2. one scan: compressed dof number (rank among the active dofs) and the compressed counts;
//   3. scan of the compressed counts = segment offsets;
//   4. ONE pass that writes the compressed dof rows AND fills the transposed table (atomic cursor per dof);
//   5. segment sort + statistics (xg_plan_finish).
#include "csr.cuh"
#include "scan.cuh"
#include <algorithm>

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__global__ void k_count_dofs(int64_t nrows, const int32_t* __restrict__ list, int nd, const int32_t* __restrict__ el_dofs,
                             int32_t* __restrict__ cnt)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * nd) return;
  const int64_t r = i / nd;
  const int k = (int)(i - r * nd);
  const int64_t src = list ? (int64_t)list[r] : r;
  const int32_t d = __ldg(el_dofs + src * nd + k);
  if (d >= 0) atomicAdd(cnt + d, 1);
}

// number of active dofs (count > 0) per tile
__global__ void __launch_bounds__(SCAN_THREADS) k_active_reduce(const int32_t* __restrict__ cnt, int64_t n, long long* __restrict__ bsum)
{
  __shared__ long long tot;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += cnt[base + k] > 0;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_active_sums(long long* __restrict__ bsum, int64_t nb)
{
  __shared__ long long tot;
  long long carry = 0;
  for (int64_t base = 0; base < nb; base += SCAN_THREADS) {
    const int64_t i = base + threadIdx.x;
    const long long v = (i < nb) ? bsum[i] : 0;
    const long long ex = block_excl_scan(v, &tot);
    if (i < nb) bsum[i] = carry + ex;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}

// dofmap[d] = rank of d among the active dofs (or -1); cntc[rank] = count of d
__global__ void __launch_bounds__(SCAN_THREADS) k_active_down(const int32_t* __restrict__ cnt, int64_t n, const long long* __restrict__ bsum,
                                                              int32_t* __restrict__ dofmap, int32_t* __restrict__ cntc)
{
  __shared__ long long tot;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? cnt[base + k] : 0; s += v[k] > 0; }
  long long ex = block_excl_scan(s, &tot) + bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) {
      if (v[k] > 0) { dofmap[base + k] = (int32_t)ex; cntc[ex] = v[k]; ex++; }
      else dofmap[base + k] = -1;
    }
  }
}

__global__ void k_gather_fill(int64_t nrows, const int32_t* __restrict__ list, int nd, int bi, const int32_t* __restrict__ el_dofs,
                              const int32_t* __restrict__ dofmap, const long long* __restrict__ adj_ptr,
                              int32_t* __restrict__ cursor, int32_t* __restrict__ table, long long* __restrict__ adj)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows * nd) return;
  const int64_t r = i / nd;
  const int k = (int)(i - r * nd);
  const int64_t src = list ? (int64_t)list[r] : r;
  int32_t d = __ldg(el_dofs + src * nd + k);
  if (d >= 0) d = __ldg(dofmap + d);
  table[i] = d;
  if (d >= 0) {
    const int pos = atomicAdd(cursor + d, 1);
    adj[adj_ptr[d] + pos] = XG_ADJ(bi, r, k);
  }
}

}  // namespace

extern "C" int xgpu_restrict_plan(xgpu_ctx* ctx, int nd, const int32_t* el_dofs, int nlists, const int32_t* const* lists_host,
                                  const int64_t* counts_host, int64_t ndof_full, int32_t* dofmap, int64_t* ndof_host,
                                  int32_t* const* tables_host, xgpu_plan** plan_out)
{
  XGPU_NVTX(ctx, "xgpu_restrict_plan");
  if (!ctx || nd < 1 || nd > 255 || !el_dofs || nlists < 1 || nlists > XG_MAXBATCH || !counts_host || ndof_full < 0 ||
      ndof_full >= (int64_t)0x7fffffff || !dofmap || !ndof_host || !tables_host || !plan_out)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_restrict_plan: bad argument");
  *plan_out = nullptr;
  cudaStream_t st = ctx->stream;
  int64_t nent_total = 0;
  for (int i = 0; i < nlists; i++) {
    if (counts_host[i] < 0 || counts_host[i] * nd >= (int64_t)0x7fffffff) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "xgpu_restrict_plan: a list has 2^31 or more table entries; split it");
    if (counts_host[i] > 0 && !tables_host[i]) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_restrict_plan: missing output table");
    nent_total += counts_host[i] * nd;
  }
  cudaEvent_t ev0;
  XGPU_CUDA(ctx, cudaEventCreate(&ev0));
  XGPU_CUDA(ctx, cudaEventRecord(ev0, st));
  // 1. occurrences per dof of the full space
  void* sp;
  int rc = xgpu_scratch(ctx, 10, sizeof(int32_t) * (size_t)(2 * ndof_full + 4), &sp);
  if (rc) { cudaEventDestroy(ev0); return rc; }
  int32_t* cnt = (int32_t*)sp;
  int32_t* cntc = cnt + ndof_full + 2;
  XGPU_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)(ndof_full + 2), st));
  for (int i = 0; i < nlists; i++) {
    if (counts_host[i] == 0) continue;
    k_count_dofs<<<nblk(counts_host[i] * nd, 256), 256, 0, st>>>(counts_host[i], lists_host ? lists_host[i] : nullptr, nd, el_dofs, cnt);
    XGPU_LAUNCH_CHECK(ctx);
  }
  // 2. compressed numbering + compressed counts
  int64_t ndof = 0;
  if (ndof_full > 0) {
    const int64_t nb = (ndof_full + SCAN_TILE - 1) / SCAN_TILE;
    void* bp;
    rc = xgpu_scratch(ctx, 7, sizeof(long long) * (nb + 1), &bp);
    if (rc) { cudaEventDestroy(ev0); return rc; }
    long long* bsum = (long long*)bp;
    k_active_reduce<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(cnt, ndof_full, bsum);
    XGPU_LAUNCH_CHECK(ctx);
    k_active_sums<<<1, SCAN_THREADS, 0, st>>>(bsum, nb);
    XGPU_LAUNCH_CHECK(ctx);
    k_active_down<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(cnt, ndof_full, bsum, dofmap, cntc);
    XGPU_LAUNCH_CHECK(ctx);
    XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, bsum + nb, sizeof(long long), st));
    XGPU_CUDA(ctx, cudaStreamSynchronize(st));
    ndof = ctx->h_pinned[0];
  }
  *ndof_host = ndof;
  // 3.-5. the transposed table of the compressed dof rows
  xgpu_plan* P = new xgpu_plan();
  ctx->plans_created++;
  P->ndof = ndof;
  P->wmax = nd;
  P->stream = st;
  P->bd.nb = nlists;
  P->bd.rowoff[0] = 0;
  for (int i = 0; i < nlists; i++) {
    P->bd.table[i] = tables_host[i]; P->bd.mats[i] = nullptr; P->bd.coefs[i] = nullptr; P->bd.width[i] = nd;
    P->bd.rowoff[i + 1] = P->bd.rowoff[i] + counts_host[i];
  }
  P->ev[0] = ev0;
#define XR_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { xgpu_plan_destroy(P); return xgpu_fail(ctx, XGPU_ERR_CUDA, cudaGetErrorString(e_)); } } while (0)
#define XR_TRYRC(call) do { int r_ = (call); if (r_) { xgpu_plan_destroy(P); return r_; } } while (0)
  for (int i = 1; i < 3; i++) XR_TRY(cudaEventCreate(&P->ev[i]));
  XR_TRY(cudaMallocAsync(&P->adj_ptr, sizeof(long long) * (ndof + 1), st));
  XR_TRY(cudaMallocAsync(&P->rowptr, sizeof(long long) * (ndof + 1), st));
  XR_TRY(cudaMallocAsync(&P->tmp32, sizeof(int32_t) * (ndof + 2), st));
  XR_TRY(cudaMemsetAsync(P->tmp32, 0, sizeof(int32_t) * (ndof + 2), st));
  XR_TRYRC(xgpu_exclusive_scan_i32(ctx, cntc, ndof, (int64_t*)P->adj_ptr));
  P->nadj = nent_total;  // upper bound (entries -1 are not stored): no second host round trip
  XR_TRY(cudaMallocAsync(&P->adj, sizeof(long long) * (size_t)(nent_total > 0 ? nent_total : 1), st));
  for (int i = 0; i < nlists; i++) {
    if (counts_host[i] == 0) continue;
    k_gather_fill<<<nblk(counts_host[i] * nd, 256), 256, 0, st>>>(counts_host[i], lists_host ? lists_host[i] : nullptr, nd, i, el_dofs,
                                                                    dofmap, P->adj_ptr, P->tmp32, tables_host[i], P->adj);
    ctx->launches++;
    XR_TRY(cudaGetLastError());
  }
  XR_TRYRC(xg_plan_finish(ctx, P, false));
  XR_TRY(cudaEventRecord(P->ev[2], st));
#undef XR_TRY
#undef XR_TRYRC
  P->with_pattern = false;
  *plan_out = P;
  return XGPU_OK;
}
