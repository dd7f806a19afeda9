// common.cuh — context, error handling, scratch arena, device-wide scan used by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/xgpu.h"

#define XGPU_NSM 148  // B200: 2 dies x 74 SMs
#define XGPU_NSCRATCH 12

struct xgpu_rule_host {
  int n = 0;
  std::vector<double> pts;  // n x 3
  std::vector<double> wts;
  double* d_pts = nullptr;  // device copies (n x 3, n)
  double* d_wts = nullptr;
};

#include <nvtx3/nvToolsExt.h>

struct xgpu_ctx;
// NVTX range around an entry point (visible in Nsight Systems / ncu --nvtx) + the context's call counter (xgpu_get_stats)
struct XgpuRange {
  XgpuRange(xgpu_ctx* ctx, const char* name);
  ~XgpuRange() { nvtxRangePop(); }
};
#define XGPU_NVTX(ctx, name) XgpuRange xgpu_nvtx_range_(ctx, name)

struct xgpu_ctx {
  int device = 0;
  cudaStream_t stream = 0;
  std::string err;
  int64_t launches = 0;
  int64_t api_calls = 0;      // entry points wrapped in XGPU_NVTX
  int64_t plans_created = 0;
  // device status: [0] = code, [1] = pad ; elem as unsigned long long min
  int* d_status = nullptr;               // 2 ints
  unsigned long long* d_status_elem = nullptr;
  // grow-only scratch arena (slots by index so independent temporaries do not alias)
  void* scratch[XGPU_NSCRATCH] = {};
  size_t scratch_bytes[XGPU_NSCRATCH] = {};
  // current mesh deformation (xgpu_set_deformation): order 0 = none
  int def_order = 0;
  const int32_t* def_el_dofs = nullptr;  // ne x nd(def_order)
  const double* def_vals = nullptr;      // ndof x dim
  int opt_dmma = 0;   // element matrices on FP64 tensor tiles: -1 never, 0 at order >= 3, 1 also at order 2
  int opt_csr_small = 1;  // one-pass CSR: thread-per-row kernels (register sorting network) for rows with few adjacent table rows
  int opt_elmat_lp = 1;  // cut-element matrices at order <= 2: lane-per-point kernel (elmat_lp.cu); 0 = staged half-warp kernel
  int opt_cut_bpsm = 2;  // k_elmat_cut: resident blocks per SM (1 leaves half of the register file to concurrent kernels)
  int opt_short_blocks = 0;  // cut quadrature / cut element matrices: one tile per block instead of a persistent grid (a
                             // low-priority side stream then yields the SMs to the main stream block by block)
  int opt_csr_warp = 1;   // one-pass CSR: warp-per-row kernels (sorting network over the warp) for rows with <= 256 candidates
  int opt_csr_warp_bpsm = 12;  // one-pass CSR, warp-per-row kernels: blocks per SM of the persistent grid
  int opt_csr_warp_half = 1;  // one-pass CSR, warp-per-row pattern pass: rows with <= 128 candidates sort 4 keys per lane instead of 8
  int opt_csr_trips = 0;  // one-pass CSR: columns per lane (0 policy = 1 for rows up to 32 wide; 2: half-width lane groups)
  int opt_csr_closed_form = 1;  // one-pass CSR: rows of affine P2 tets (mass + stiffness) in closed form instead of the reference tensors
  int opt_csr_fork = 1;   // one-pass CSR: the warp-per-row kernel runs on an internal second stream next to the thread-per-row one
  cudaStream_t aux_stream = nullptr;  // created on first use (same priority as `stream` at that time)
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  double fp_downscale = 1.0;  // dFacetPatch(downscale=): xgpu_facetpatch_set_downscale
  int opt_st_tp = 1;  // space-time element matrices: tensor-structured kernel where instantiated (0: always the generic one)
  int64_t epoch = 0;  // bumped whenever a plan takes over the shared pattern scratch (slot 8)
  int64_t* h_pinned = nullptr;  // 16 x int64 pinned staging for host scalars
  std::map<std::pair<int, int>, xgpu_rule_host> rules;  // (dim, order) -> rule
  std::map<int, void*> reftensors;                     // cached reference tensors (device)
};

inline XgpuRange::XgpuRange(xgpu_ctx* ctx, const char* name)
{
  if (ctx) ctx->api_calls++;
  nvtxRangePushA(name);
}

inline int xgpu_fail(xgpu_ctx* ctx, int code, const char* msg)
{
  if (ctx) ctx->err = msg;
  return code;
}

#define XGPU_CUDA(ctx, call)                                                          \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      char buf__[512];                                                                \
      snprintf(buf__, sizeof(buf__), "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return xgpu_fail(ctx, XGPU_ERR_CUDA, buf__);                                    \
    }                                                                                 \
  } while (0)

#define XGPU_LAUNCH_CHECK(ctx)        \
  do {                                \
    (ctx)->launches++;                \
    XGPU_CUDA(ctx, cudaGetLastError()); \
  } while (0)

// Device memory of the library comes from the device's stream-ordered pool (cudaMallocAsync):
// xgpu_create raises the pool's release threshold so freed blocks stay cached, which makes the
// per-assembly plan buffers (hundreds of MB, sizes known only after a device-side count) cost
// microseconds instead of the ~5 ms a cudaMalloc/cudaFree pair costs.
inline int xgpu_scratch(xgpu_ctx* ctx, int slot, size_t bytes, void** out)
{
  if (bytes > ctx->scratch_bytes[slot]) {
    if (ctx->scratch[slot]) {
      XGPU_CUDA(ctx, cudaFreeAsync(ctx->scratch[slot], ctx->stream));
      ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    XGPU_CUDA(ctx, cudaMallocAsync(&ctx->scratch[slot], want, ctx->stream));
    ctx->scratch_bytes[slot] = want;
  }
  *out = ctx->scratch[slot];
  return XGPU_OK;
}

// Host scalars (counts, totals, status words) are read back WITHOUT the copy engine: a one-warp kernel writes them into
// the context's mapped pinned staging block (cudaHostAlloc(..., Mapped): the host pointer is valid on the device under
// UVA).  A cudaMemcpyAsync of 8 bytes queues on the device-to-host DMA engine behind any bulk D2H copy running on
// another stream (the e2e leg ships 1.7 GB of results while the next step is being enqueued) and stalled the host for
// the whole transfer (profiles/r01_notes.md, v34).
static __global__ void k_to_host(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, int nbytes)
{
  for (int i = threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}

inline cudaError_t xgpu_d2h_small(void* dst_host_mapped, const void* src_dev, size_t bytes, cudaStream_t st)
{
  k_to_host<<<1, 32, 0, st>>>((unsigned char*)dst_host_mapped, (const unsigned char*)src_dev, (int)bytes);
  return cudaGetLastError();
}

// fork / join of the context's internal second stream: work launched between xgpu_aux_begin and xgpu_aux_end runs on
// ctx->aux_stream, ordered after everything enqueued on ctx->stream so far; xgpu_aux_end makes ctx->stream wait for it.
// Usage: cudaStream_t main = ctx->stream; xgpu_aux_begin(ctx); <launches use ctx->stream (= aux)>; xgpu_aux_end(ctx, main);
inline cudaError_t xgpu_aux_begin(xgpu_ctx* ctx)
{
  cudaError_t e;
  if (!ctx->aux_stream) {
    int prio = 0;
    if ((e = cudaStreamGetPriority(ctx->stream, &prio)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  if ((e = cudaEventRecord(ctx->aux_fork, ctx->stream)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0)) != cudaSuccess) return e;
  ctx->stream = ctx->aux_stream;
  return cudaSuccess;
}
inline cudaError_t xgpu_aux_end(xgpu_ctx* ctx, cudaStream_t main)
{
  ctx->stream = main;
  cudaError_t e;
  if ((e = cudaEventRecord(ctx->aux_join, ctx->aux_stream)) != cudaSuccess) return e;
  return cudaStreamWaitEvent(main, ctx->aux_join, 0);
}

// record a per-element geometry error: keep the smallest failing element TOGETHER with its code — one 64-bit word
// (element << 8 | -code) and ONE atomicMin, so concurrent failures cannot pair an element with another element's code.
// `status` only flags that something failed (xgpu_get_status decodes the word).
__device__ __forceinline__ void xgpu_record_error(int* status, unsigned long long* status_elem, int code, long long elem)
{
  atomicMin(status_elem, ((unsigned long long)elem << 8) | (unsigned long long)((-code) & 0xff));
  *(volatile int*)status = code;   // any failing code: non-zero flag
}

// ---- device-wide exclusive scan (int32 counts -> int64 offsets), out has n+1 entries ----
int xgpu_exclusive_scan_i32(xgpu_ctx* ctx, const int32_t* in, int64_t n, int64_t* out);
// the same with 32-bit output offsets (n+1 entries)
int xgpu_exclusive_scan_i32_o32(xgpu_ctx* ctx, const int32_t* in, int64_t n, int32_t* out);

// deform.cu: per-point interface weight factor / integrals on the deformed mesh (ctx->def_* set)
int xgpu_if_weights_def(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                        const double* vtx_xyz, const double* phi, const int64_t* offsets, const double* pts, double* wts);
int xgpu_integrate_def(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                       const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts,
                       int e0, int e1, int e2, double* elsum);
int xgpu_sum_f64(xgpu_ctx* ctx, double* buf_a, int64_t n, double* buf_b, double* result_host);

const xgpu_rule_host* xgpu_rule_lookup(xgpu_ctx* ctx, int dim, int order);
