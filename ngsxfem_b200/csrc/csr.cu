// csr.cu — sparsity graph + deterministic element-matrix assembly.
// (RestrictedBilinearForm::GetGraph utils/restrictedblf.cpp:142-163 -> NGSolve MatrixGraph;
//  SparseMatrix::AddElementMatrix.)
//
// The reference scatters element by element with a binary search per entry (and relies on
// element colouring for thread safety).  Here assembly is a ROW-WISE GATHER:
//   plan    : transposed table  dof -> [(table row, local index)], built by histogram + scan +
//             atomic-cursor fill, then each segment is sorted so the order is deterministic.
//   pattern : one warp per matrix row gathers the dofs of all adjacent table rows into shared
//             memory, bitonic-sorts, uniques: "union of cliques, ascending columns, explicit
//             zeros kept" — bit-exact pattern.  Count pass -> scan -> fill pass.
//   scatter : one warp per matrix row, each lane owns nonzeros of the row and sums the matching
//             element-matrix entries in (batch, row) order: no atomics, bit-reproducible.
// All kernels are HBM/L2-gather bound; algorithmic bytes are stated in DESIGN.md.
#include "common.cuh"
#include <vector>

#define XG_MAXBATCH 8

struct BatchDev {
  int nb;
  const int32_t* table[XG_MAXBATCH];
  const double* mats[XG_MAXBATCH];
  long long rowoff[XG_MAXBATCH + 1];
  int width[XG_MAXBATCH];
};

struct xgpu_plan {
  int64_t ndof = 0, nadj = 0, nnz = 0;
  int maxcand = 0;
  BatchDev bd;
  long long* adj_ptr = nullptr;   // ndof+1
  long long* adj = nullptr;       // nadj entries: (global table row << 8) | local index
  long long* rowptr = nullptr;    // ndof+1
  int32_t* tmp32 = nullptr;       // ndof+1 counters
};

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ int find_batch(const BatchDev& bd, long long grow)
{
  int b = 0;
  while (b + 1 < bd.nb && grow >= bd.rowoff[b + 1]) b++;
  return b;
}

__global__ void k_adj_count(const int32_t* __restrict__ table, int64_t nent, int32_t* __restrict__ cnt)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nent) return;
  const int32_t d = __ldg(table + i);
  if (d >= 0) atomicAdd(cnt + d, 1);
}

__global__ void k_adj_fill(const int32_t* __restrict__ table, int64_t nent, int width, long long rowoff,
                           const long long* __restrict__ adj_ptr, int32_t* __restrict__ cursor, long long* __restrict__ adj)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nent) return;
  const int32_t d = __ldg(table + i);
  if (d < 0) return;
  const int64_t r = i / width;
  const int loc = (int)(i - r * width);
  const int pos = atomicAdd(cursor + d, 1);
  adj[adj_ptr[d] + pos] = ((rowoff + r) << 8) | (long long)loc;
}

// sort every adjacency segment ascending (insertion sort, segments are short) and compute the
// candidate count of each row = sum of the widths of the adjacent table rows
__global__ void k_adj_sort(int64_t ndof, const long long* __restrict__ adj_ptr, long long* __restrict__ adj,
                           BatchDev bd, int32_t* __restrict__ cand, int* __restrict__ maxcand)
{
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  if (d < ndof) {
    const long long a = adj_ptr[d], b = adj_ptr[d + 1];
    for (long long i = a + 1; i < b; i++) {
      const long long key = adj[i];
      long long j = i - 1;
      while (j >= a && adj[j] > key) { adj[j + 1] = adj[j]; j--; }
      adj[j + 1] = key;
    }
    for (long long i = a; i < b; i++) c += bd.width[find_batch(bd, adj[i] >> 8)];
    cand[d] = c;
  }
  for (int s = 16; s > 0; s >>= 1) c = max(c, __shfl_down_sync(0xffffffffu, c, s));
  if ((threadIdx.x & 31) == 0 && c > 0) atomicMax(maxcand, c);
}

// gather + sort + unique of one row's candidate columns; returns #unique (valid in all lanes).
// buf: CAP ints of shared memory owned by this warp.
__device__ __forceinline__ int row_unique(int64_t d, const long long* __restrict__ adj_ptr,
                                          const long long* __restrict__ adj, const BatchDev& bd, int* buf, int lane)
{
  const long long a = adj_ptr[d], b = adj_ptr[d + 1];
  int m = 0;
  long long prev = -1;
  for (long long i = a; i < b; i++) {
    const long long grow = adj[i] >> 8;
    if (grow == prev) continue;  // the same table row twice (a dof repeated inside one row)
    prev = grow;
    const int bi = find_batch(bd, grow);
    const int w = bd.width[bi];
    const int32_t* row = bd.table[bi] + (grow - bd.rowoff[bi]) * w;
    for (int k = lane; k < w; k += 32) {
      const int32_t c = __ldg(row + k);
      buf[m + k] = (c >= 0) ? c : 0x7fffffff;
    }
    m += w;
  }
  int P = 32;
  while (P < m) P <<= 1;
  for (int k = m + lane; k < P; k += 32) buf[k] = 0x7fffffff;
  __syncwarp();
  for (int size = 2; size <= P; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const int x = buf[lo], y = buf[hi];
        if ((x > y) == up) { buf[lo] = y; buf[hi] = x; }
      }
      __syncwarp();
    }
  // unique (in place compaction, ascending): count and positions by ballot prefix
  int nuniq = 0;
  for (int base = 0; base < P; base += 32) {
    const int k = base + lane;
    const int x = buf[k];
    const bool keep = (x != 0x7fffffff) && (k == 0 || buf[k - 1] != x);
    __syncwarp();
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[nuniq + __popc(bal & ((1u << lane) - 1u))] = x;  // nuniq + rank <= k: safe
    nuniq += __popc(bal);
    __syncwarp();
  }
  return nuniq;
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_pattern(int64_t ndof, const long long* __restrict__ adj_ptr,
                                                 const long long* __restrict__ adj, BatchDev bd, int CAP,
                                                 int32_t* __restrict__ rowcnt, const long long* __restrict__ rowptr,
                                                 int32_t* __restrict__ colidx)
{
  extern __shared__ int sbuf[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int* buf = sbuf + (size_t)wib * CAP;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t d = warp0; d < ndof; d += nwarps) {
    const int nu = row_unique(d, adj_ptr, adj, bd, buf, lane);
    if (!FILL) { if (lane == 0) rowcnt[d] = nu; }
    else {
      const long long o = rowptr[d];
      for (int k = lane; k < nu; k += 32) colidx[o + k] = buf[k];
    }
    __syncwarp();
  }
}

// row-wise gather of the element matrices (see header comment)
__global__ void __launch_bounds__(256) k_scatter(int64_t ndof, const long long* __restrict__ adj_ptr,
                                                 const long long* __restrict__ adj, BatchDev bd,
                                                 const long long* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                 double* __restrict__ vals)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t d = warp0; d < ndof; d += nwarps) {
    const long long r0 = rowptr[d], r1 = rowptr[d + 1];
    const long long a = adj_ptr[d], b = adj_ptr[d + 1];
    for (long long cb = r0; cb < r1; cb += 32) {
      const long long k = cb + lane;
      const int32_t mycol = (k < r1) ? __ldg(colidx + k) : -2;
      double acc = 0.0;
      for (long long i = a; i < b; i++) {
        const long long e = adj[i];
        const long long grow = e >> 8;
        const int li = (int)(e & 255);
        const int bi = find_batch(bd, grow);
        const int w = bd.width[bi];
        const long long lr = grow - bd.rowoff[bi];
        const int32_t* row = bd.table[bi] + lr * w;
        const double* mrow = bd.mats[bi] + ((size_t)lr * w + li) * w;
        for (int lj = 0; lj < w; lj++)
          if (__ldg(row + lj) == mycol) acc += __ldg(mrow + lj);
      }
      if (k < r1) vals[k] += acc;
    }
  }
}

__global__ void k_spmv(int64_t nrows, const long long* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                       const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y)
{
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nrows) return;
  double s = 0.0;
  for (long long k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) s += vals[k] * __ldg(x + colidx[k]);
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if (lane == 0) y[r] = s;
}

int fill_bd(xgpu_ctx* ctx, const xgpu_batch* bh, int nb, BatchDev& bd)
{
  if (nb < 1 || nb > XG_MAXBATCH) return xgpu_fail(ctx, XGPU_ERR_ARG, "csr: 1..8 batches");
  bd.nb = nb;
  bd.rowoff[0] = 0;
  for (int i = 0; i < nb; i++) {
    if (bh[i].width < 1 || bh[i].width > 255 || bh[i].nrows < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "csr: bad batch");
    bd.table[i] = bh[i].table; bd.mats[i] = bh[i].mats; bd.width[i] = bh[i].width;
    bd.rowoff[i + 1] = bd.rowoff[i] + bh[i].nrows;
  }
  return XGPU_OK;
}

}  // namespace

extern "C" void xgpu_plan_destroy(xgpu_plan* p)
{
  if (!p) return;
  if (p->adj_ptr) cudaFree(p->adj_ptr);
  if (p->adj) cudaFree(p->adj);
  if (p->rowptr) cudaFree(p->rowptr);
  if (p->tmp32) cudaFree(p->tmp32);
  delete p;
}

extern "C" int xgpu_csr_plan(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                             xgpu_plan** plan_out, int64_t* nnz_host)
{
  if (!ctx || !plan_out || ndof < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_plan: bad argument");
  *plan_out = nullptr;
  xgpu_plan* P = new xgpu_plan();
  int rc = fill_bd(ctx, batches_host, nbatches, P->bd);
  if (rc) { delete P; return rc; }
  P->ndof = ndof;
  cudaStream_t st = ctx->stream;
#define XG_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { xgpu_plan_destroy(P); return xgpu_fail(ctx, XGPU_ERR_CUDA, cudaGetErrorString(e_)); } } while (0)
#define XG_TRYRC(call) do { int r_ = (call); if (r_) { xgpu_plan_destroy(P); return r_; } } while (0)
  XG_TRY(cudaMalloc(&P->adj_ptr, sizeof(long long) * (ndof + 1)));
  XG_TRY(cudaMalloc(&P->rowptr, sizeof(long long) * (ndof + 1)));
  XG_TRY(cudaMalloc(&P->tmp32, sizeof(int32_t) * (ndof + 2)));
  XG_TRY(cudaMemsetAsync(P->tmp32, 0, sizeof(int32_t) * (ndof + 2), st));
  for (int i = 0; i < nbatches; i++) {
    const int64_t nent = batches_host[i].nrows * batches_host[i].width;
    if (nent == 0) continue;
    k_adj_count<<<nblk(nent, 256), 256, 0, st>>>(batches_host[i].table, nent, P->tmp32);
    ctx->launches++;
  }
  XG_TRYRC(xgpu_exclusive_scan_i32(ctx, P->tmp32, ndof, (int64_t*)P->adj_ptr));
  XG_TRY(cudaMemcpyAsync(ctx->h_pinned, P->adj_ptr + ndof, sizeof(long long), cudaMemcpyDeviceToHost, st));
  XG_TRY(cudaStreamSynchronize(st));
  P->nadj = ctx->h_pinned[0];
  XG_TRY(cudaMalloc(&P->adj, sizeof(long long) * (P->nadj > 0 ? P->nadj : 1)));
  XG_TRY(cudaMemsetAsync(P->tmp32, 0, sizeof(int32_t) * (ndof + 2), st));
  for (int i = 0; i < nbatches; i++) {
    const int64_t nent = batches_host[i].nrows * batches_host[i].width;
    if (nent == 0) continue;
    k_adj_fill<<<nblk(nent, 256), 256, 0, st>>>(batches_host[i].table, nent, batches_host[i].width, P->bd.rowoff[i],
                                                 P->adj_ptr, P->tmp32, P->adj);
    ctx->launches++;
  }
  void* sp;
  XG_TRYRC(xgpu_scratch(ctx, 6, 64, &sp));
  int* d_max = (int*)sp;
  XG_TRY(cudaMemsetAsync(d_max, 0, sizeof(int), st));
  if (ndof > 0) {
    k_adj_sort<<<nblk(ndof, 128), 128, 0, st>>>(ndof, P->adj_ptr, P->adj, P->bd, P->tmp32, d_max);
    ctx->launches++;
  }
  XG_TRY(cudaMemcpyAsync(ctx->h_pinned, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
  XG_TRY(cudaStreamSynchronize(st));
  P->maxcand = *(int*)ctx->h_pinned;
  int CAP = 32;
  while (CAP < P->maxcand) CAP <<= 1;
  if (CAP > 8192) { xgpu_plan_destroy(P); return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "csr: a row has more than 8192 candidate columns"); }
  // count pass
  if (ndof > 0) {
    const size_t sm = sizeof(int) * 4 * CAP;
    XG_TRY(cudaFuncSetAttribute(k_pattern<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    unsigned grid = nblk(ndof, 4);
    if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
    k_pattern<false><<<grid, 128, sm, st>>>(ndof, P->adj_ptr, P->adj, P->bd, CAP, P->tmp32, nullptr, nullptr);
    ctx->launches++;
    XG_TRY(cudaGetLastError());
  }
  XG_TRYRC(xgpu_exclusive_scan_i32(ctx, P->tmp32, ndof, (int64_t*)P->rowptr));
  XG_TRY(cudaMemcpyAsync(ctx->h_pinned, P->rowptr + ndof, sizeof(long long), cudaMemcpyDeviceToHost, st));
  XG_TRY(cudaStreamSynchronize(st));
  P->nnz = ctx->h_pinned[0];
  if (nnz_host) *nnz_host = P->nnz;
  *plan_out = P;
  return XGPU_OK;
#undef XG_TRY
#undef XG_TRYRC
}

extern "C" int xgpu_csr_pattern(xgpu_ctx* ctx, xgpu_plan* P, int64_t* rowptr, int32_t* colidx)
{
  if (!ctx || !P) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_pattern: bad argument");
  cudaStream_t st = ctx->stream;
  XGPU_CUDA(ctx, cudaMemcpyAsync(rowptr, P->rowptr, sizeof(long long) * (P->ndof + 1), cudaMemcpyDeviceToDevice, st));
  if (P->ndof == 0 || P->nnz == 0) return XGPU_OK;
  int CAP = 32;
  while (CAP < P->maxcand) CAP <<= 1;
  const size_t sm = sizeof(int) * 4 * CAP;
  XGPU_CUDA(ctx, cudaFuncSetAttribute(k_pattern<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  unsigned grid = nblk(P->ndof, 4);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  k_pattern<true><<<grid, 128, sm, st>>>(P->ndof, P->adj_ptr, P->adj, P->bd, CAP, nullptr, P->rowptr, colidx);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_csr_scatter(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                                const int64_t* rowptr, const int32_t* colidx, double* vals)
{
  if (!ctx || !P || nbatches != P->bd.nb) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_scatter: bad argument");
  BatchDev bd = P->bd;
  for (int i = 0; i < nbatches; i++) {
    if (batches_host[i].nrows != bd.rowoff[i + 1] - bd.rowoff[i] || batches_host[i].width != bd.width[i])
      return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_scatter: batches do not match the plan");
    bd.table[i] = batches_host[i].table;
    bd.mats[i] = batches_host[i].mats;
    if (!bd.mats[i] && batches_host[i].nrows > 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_scatter: mats missing");
  }
  if (P->ndof == 0) return XGPU_OK;
  unsigned grid = nblk(P->ndof, 8);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  k_scatter<<<grid, 256, 0, ctx->stream>>>(P->ndof, P->adj_ptr, P->adj, bd, (const long long*)rowptr, colidx, vals);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_csr_spmv(xgpu_ctx* ctx, int64_t nrows, const int64_t* rowptr, const int32_t* colidx,
                             const double* vals, const double* x, double* y)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (nrows == 0) return XGPU_OK;
  k_spmv<<<nblk(nrows * 32, 256), 256, 0, ctx->stream>>>(nrows, (const long long*)rowptr, colidx, vals, x, y);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}
