// FP64 peaks of this GPU, measured: (1) a DFMA chain (8 independent accumulators per thread, every SM full),
// (2) a DMMA chain (mma.sync.aligned.m8n8k4.f64, 4 independent accumulator tiles per warp).  Prints one JSON object.
// Used as the denominator of the contraction kernels' roofline (bench.py "roofline_fp64").  tools/peak_fp64.py builds it.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_dmma(double* out, int iters, double a, double b)
{
  double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  const double av = a + threadIdx.x * 1e-9, bv = b;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int t = 0; t < 4; t++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(av), "d"(bv));
  }
  double s = 0;
  for (int t = 0; t < 4; t++) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> static float time_ms(F launch, int reps)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount, bpsm = 8, threads = 256, blocks = nsm * bpsm;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const int it_f = 1 << 15, it_m = 1 << 13;
  float ms_f = time_ms([&] { k_dfma<<<blocks, threads>>>(out, it_f, 0.999999, 1e-7); }, 5);
  float ms_m = time_ms([&] { k_dmma<<<blocks, threads>>>(out, it_m, 0.999999, 1e-7); }, 5);
  const double fl_f = 2.0 * 8 * it_f * (double)blocks * threads;
  const double fl_m = 512.0 * 4 * it_m * (double)blocks * (threads / 32);
  if (cudaGetLastError() != cudaSuccess) { printf("{\"error\": \"cuda\"}\n"); return 1; }
  printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f}\n",
         p.name, nsm, fl_f / (ms_f * 1e-3) / 1e12, fl_m / (ms_m * 1e-3) / 1e12, ms_f, ms_m);
  return 0;
}
