#pragma once
// scan.cuh — block-level building blocks of the device-wide prefix sums (scan.cu, restrict_plan.cu)
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long warp_incl_scan(long long v)
{
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (unsigned)d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ long long block_excl_scan(long long v, long long* total)
{
  __shared__ long long wsum[32];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long inc = warp_incl_scan(v);
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    long long w = (lane < (blockDim.x >> 5)) ? wsum[lane] : 0;
    long long winc = warp_incl_scan(w);
    wsum[lane] = winc - w;  // exclusive warp offsets
    __syncwarp();
    if (lane == (blockDim.x >> 5) - 1) *total = winc;
  }
  __syncthreads();
  long long res = wsum[wid] + inc - v;
  __syncthreads();
  return res;
}

}  // namespace
