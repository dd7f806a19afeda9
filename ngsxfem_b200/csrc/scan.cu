// scan.cu — device-wide exclusive prefix sum (reduce / scan-of-sums / downsweep), hand written.
// Used for ordered compaction (bitset -> list), CSR-of-points offsets and CSR row pointers.
#include "scan.cuh"

namespace {

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t* __restrict__ in, int64_t n, long long* __restrict__ bsum)
{
  __shared__ long long tot;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += in[base + k];
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(long long* __restrict__ bsum, int64_t nb)
{
  __shared__ long long tot;
  long long carry = 0;
  for (int64_t base = 0; base < nb; base += SCAN_THREADS) {
    int64_t i = base + threadIdx.x;
    long long v = (i < nb) ? bsum[i] : 0;
    long long ex = block_excl_scan(v, &tot);
    if (i < nb) bsum[i] = carry + ex;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}

template <typename OUT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const int32_t* __restrict__ in, int64_t n,
                                                            const long long* __restrict__ bsum, OUT* __restrict__ out)
{
  __shared__ long long tot;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
  long long ex = block_excl_scan(s, &tot) + bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = (OUT)ex;
    ex += v[k];
    if (base + k == n - 1) out[n] = (OUT)ex;
  }
}

template <typename OUT>
int scan_impl(xgpu_ctx* ctx, const int32_t* in, int64_t n, OUT* out)
{
  if (n <= 0) {
    XGPU_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(OUT), ctx->stream));
    return XGPU_OK;
  }
  int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  void* p;
  int rc = xgpu_scratch(ctx, 7, sizeof(long long) * (nb + 1), &p);
  if (rc) return rc;
  long long* bsum = (long long*)p;
  k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, n, bsum);
  XGPU_LAUNCH_CHECK(ctx);
  k_scan_sums<<<1, SCAN_THREADS, 0, ctx->stream>>>(bsum, nb);
  XGPU_LAUNCH_CHECK(ctx);
  k_scan_down<OUT><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, n, bsum, out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

int xgpu_exclusive_scan_i32(xgpu_ctx* ctx, const int32_t* in, int64_t n, int64_t* out)
{
  return scan_impl<long long>(ctx, in, n, (long long*)out);
}

int xgpu_exclusive_scan_i32_o32(xgpu_ctx* ctx, const int32_t* in, int64_t n, int32_t* out)
{
  return scan_impl<int32_t>(ctx, in, n, out);
}
