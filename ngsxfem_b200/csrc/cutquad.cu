// cutquad.cu — StraightCutIntegrationRule on the GPU (cutint/straightcutrule.cpp:582-677).
//
//   k_cq_count<D>  one thread per listed element: #points from the sign pattern only
//                  (1 / 2 / 3 sub-simplices, Decompose :131-196) -> device scan -> CSR-of-points.
//   k_cq_fill<D>   one WARP per element: the decomposition (<= 4 cut points, <= 3 sub-simplices)
//                  is computed redundantly per lane in registers/local memory, the reference rule
//                  is staged in shared memory once per block, lanes stride over the
//                  (sub-simplex, rule point) pairs so that the 32 B/point output is written
//                  coalesced.  HBM bound: 16 B + 32 B (+ vertex coords for IF) in, 32 B/point out.
//   k_integrate<D> sum_q w_q |det F| x^e over a rule batch (cutint/cutintegral.cpp:126-212),
//                  fixed-order reductions (bit reproducible).
#include "common.cuh"
#include "cutgeom.cuh"

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <int D>
__device__ __forceinline__ void load_el(const int32_t* __restrict__ el_vtx, const double* __restrict__ phi, int64_t el,
                                        int* v, double* l)
{
#pragma unroll
  for (int k = 0; k < D + 1; k++) { v[k] = __ldg(el_vtx + el * (D + 1) + k); l[k] = __ldg(phi + v[k]); }
}

template <int D>
__global__ void __launch_bounds__(256) k_cq_count(int64_t n, const int32_t* __restrict__ elems,
                                                  const int32_t* __restrict__ el_vtx, const double* __restrict__ phi,
                                                  int dt, int np_vol, int np_sub, int32_t* __restrict__ counts)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t el = elems ? (int64_t)elems[i] : i;
  int v[D + 1]; double l[D + 1];
  load_el<D>(el_vtx, phi, el, v, l);
  const int dom = xg_check_cut(l, D + 1);
  int cnt;
  if (dom != XGPU_IF) cnt = (dom == dt) ? np_vol : 0;
  else {
    int nneg = 0;
#pragma unroll
    for (int k = 0; k < D + 1; k++) nneg += (l[k] < 0.0);
    cnt = xg_nsub(D, nneg, dt) * np_sub;
  }
  counts[i] = cnt;
}

// Fill: a block owns 128 consecutive listed elements.  Phase 1, one THREAD per element: the
// decomposition (<= 4 cut points, <= 3 sub-simplices) and the weight factors go to shared memory
// (no redundant per-lane geometry).  Phase 2, one WARP per element in turn: lanes stride over the
// (sub-simplex, rule point) pairs and write the 32 B/point output coalesced.
struct CqDesc {
  double p[3][4][3];  // sub-simplex vertices
  double vol[3];      // |det| of every sub-simplex
  double iffac;       // interface weight factor (1 for NEG/POS)
  int ns;             // >0 sub-simplices, 0 = standard rule of an uncut matching element, <0 = error / nothing
  int D;              // dimension of the sub-simplices
};

template <int D>
__global__ void __launch_bounds__(128) k_cq_fill(int64_t n, const int32_t* __restrict__ elems,
                                                 const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                                 const double* __restrict__ phi, int dt,
                                                 const double* __restrict__ rv_p, const double* __restrict__ rv_w, int np_vol,
                                                 const double* __restrict__ rs_p, const double* __restrict__ rs_w, int np_sub,
                                                 const int64_t* __restrict__ offsets, double* __restrict__ pts,
                                                 double* __restrict__ wts, int* status, unsigned long long* status_elem,
                                                 int untransformed)
{
  extern __shared__ double smem[];  // sub rule: np_sub x 4 (xi0, xi1, xi2, w), then 128 descriptors
  CqDesc* desc = reinterpret_cast<CqDesc*>(smem + 4 * ((np_sub + 1) & ~1));
  for (int t = threadIdx.x; t < np_sub; t += blockDim.x) {
    smem[4 * t + 0] = rs_p[3 * t + 0]; smem[4 * t + 1] = rs_p[3 * t + 1]; smem[4 * t + 2] = rs_p[3 * t + 2];
    smem[4 * t + 3] = rs_w[t];
  }
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 128; base < n; base += (int64_t)gridDim.x * 128) {
    __syncthreads();
    {  // phase 1
      const int64_t i = base + threadIdx.x;
      CqDesc& Dd = desc[threadIdx.x];
      Dd.ns = -1;
      if (i < n && offsets[i + 1] > offsets[i]) {
        const int64_t el = elems ? (int64_t)elems[i] : i;
        int v[D + 1]; double l[D + 1];
        load_el<D>(el_vtx, phi, el, v, l);
        const int dom = xg_check_cut(l, D + 1);
        if (dom != XGPU_IF) Dd.ns = 0;  // uncut, matching: the standard rule (:669-674)
        else {
          XgSimplex ref, sub[3];
          xg_ref_simplex(D, ref);
          const int ns = xg_decompose(ref, l, dt, sub);
          if (ns < 0) xgpu_record_error(status, status_elem, ns, el);
          else {
            double iffac = 1.0;
            // untransformed (straightcutrule.cpp:814-879): the factor is applied per point afterwards (deformed mesh)
            if (dt == XGPU_IF && !untransformed) {  // w <- w * ||F^{-T} n_ref||  (:558-578); affine F from the vertices
              double vx[D + 1][3], F[D * D], Fi[D * D];
#pragma unroll
              for (int k = 0; k < D + 1; k++)
#pragma unroll
                for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[k] * D + c) : 0.0;
              xg_affine_jac<D>(vx, F);
              xg_inv<D>(F, Fi);
              iffac = xg_if_weight_factor<D>(Fi, l);
            }
            for (int s = 0; s < ns; s++) {
              Dd.vol[s] = xg_volume(sub[s]);
              for (int a = 0; a < 4; a++) for (int c = 0; c < 3; c++) Dd.p[s][a][c] = sub[s].p[a][c];
            }
            Dd.D = sub[0].D;
            Dd.iffac = iffac;
          }
          Dd.ns = ns;
        }
      }
    }
    __syncthreads();
    // phase 2
    for (int k = wib; k < 128; k += 4) {
      const int64_t i = base + k;
      if (i >= n) break;
      const CqDesc& Dd = desc[k];
      const int64_t off = offsets[i];
      const int cnt = (int)(offsets[i + 1] - off);
      if (cnt == 0) continue;
      if (Dd.ns == 0) {
        for (int q = lane; q < cnt; q += 32) {
          pts[3 * (off + q) + 0] = rv_p[3 * q + 0]; pts[3 * (off + q) + 1] = rv_p[3 * q + 1]; pts[3 * (off + q) + 2] = rv_p[3 * q + 2];
          wts[off + q] = rv_w[q];
        }
      } else if (Dd.ns < 0) {
        for (int q = lane; q < cnt; q += 32) {
          pts[3 * (off + q) + 0] = 0.0; pts[3 * (off + q) + 1] = 0.0; pts[3 * (off + q) + 2] = 0.0; wts[off + q] = 0.0;
        }
      } else {
        const int sd = Dd.D;
        for (int idx = lane; idx < cnt; idx += 32) {
          const int s = idx / np_sub, q = idx - s * np_sub;
          const double* xi = &smem[4 * q];
          double ow = 1.0;
          for (int m = 0; m < sd; m++) ow -= xi[m];
          double x[3];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            double t = ow * Dd.p[s][0][c];
            for (int m = 0; m < sd; m++) t += xi[m] * Dd.p[s][m + 1][c];
            x[c] = t;
          }
          pts[3 * (off + idx) + 0] = x[0]; pts[3 * (off + idx) + 1] = x[1]; pts[3 * (off + idx) + 2] = x[2];
          wts[off + idx] = xi[3] * Dd.vol[s] * Dd.iffac;
        }
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(128) k_integrate(int64_t n, const int32_t* __restrict__ elems,
                                                   const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                                   const int64_t* __restrict__ offsets, const double* __restrict__ pts,
                                                   const double* __restrict__ wts, int e0, int e1, int e2,
                                                   double* __restrict__ elsum,
                                                   const double* __restrict__ tvals = nullptr, int et = 0)
{
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int64_t el = elems ? (int64_t)elems[i] : i;
  double vx[D + 1][3], F[D * D], Fi[D * D];
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
    const int v = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
    for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v * D + c) : 0.0;
  }
  xg_affine_jac<D>(vx, F);
  const double meas = fabs(xg_inv<D>(F, Fi));
  const int ex[3] = {e0, e1, e2};
  double s = 0.0;
  for (int64_t q = offsets[i] + lane; q < offsets[i + 1]; q += 32) {
    double lam[D + 1], sl = 1.0;
#pragma unroll
    for (int k = 0; k < D; k++) { lam[k] = pts[3 * q + k]; sl -= lam[k]; }
    lam[D] = sl;
    double f = 1.0;
#pragma unroll
    for (int c = 0; c < D; c++) {
      double xc = 0.0;
#pragma unroll
      for (int k = 0; k < D + 1; k++) xc += lam[k] * vx[k][c];
      for (int e = 0; e < ex[c]; e++) f *= xc;
    }
    if (tvals) { const double t = tvals[q]; for (int e = 0; e < et; e++) f *= t; }  // tref^et (space-time rules)
    s += wts[q] * meas * f;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if (lane == 0) elsum[i] = s;
}

// the same for simplices of dimension D embedded in R^SD (boundary elements, dCut(..., vb=BND)): the measure is
// sqrt(det F^T F) of the D x SD Jacobian (segment length in 2D / 3D, triangle: |cross product|)
template <int D, int SD>
__global__ void __launch_bounds__(128) k_integrate_emb(int64_t n, const int32_t* __restrict__ elems,
                                                       const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                                       const int64_t* __restrict__ offsets, const double* __restrict__ pts,
                                                       const double* __restrict__ wts, int e0, int e1, int e2,
                                                       double* __restrict__ elsum)
{
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int64_t el = elems ? (int64_t)elems[i] : i;
  double vx[D + 1][3];
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
    const int v = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
    for (int c = 0; c < 3; c++) vx[k][c] = (c < SD) ? __ldg(vtx_xyz + (int64_t)v * SD + c) : 0.0;
  }
  double meas;
  {
    double a[3], b[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { a[c] = vx[0][c] - vx[D][c]; b[c] = (D == 2) ? vx[1][c] - vx[D][c] : 0.0; }
    if (D == 1) meas = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    else {
      const double c0 = a[1] * b[2] - a[2] * b[1], c1 = a[2] * b[0] - a[0] * b[2], c2 = a[0] * b[1] - a[1] * b[0];
      meas = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    }
  }
  const int ex[3] = {e0, e1, e2};
  double s = 0.0;
  for (int64_t q = offsets[i] + lane; q < offsets[i + 1]; q += 32) {
    double lam[D + 1], sl = 1.0;
#pragma unroll
    for (int k = 0; k < D; k++) { lam[k] = pts[3 * q + k]; sl -= lam[k]; }
    lam[D] = sl;
    double f = 1.0;
#pragma unroll
    for (int c = 0; c < SD; c++) {
      double xc = 0.0;
#pragma unroll
      for (int k = 0; k < D + 1; k++) xc += lam[k] * vx[k][c];
      for (int e = 0; e < ex[c]; e++) f *= xc;
    }
    s += wts[q] * meas * f;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if (lane == 0) elsum[i] = s;
}

// fixed-order sum: each block reduces a contiguous tile with a shared-memory tree
__global__ void __launch_bounds__(256) k_sum_tiles(const double* __restrict__ in, int64_t n, double* __restrict__ out)
{
  __shared__ double sm[256];
  const int64_t tile = 256 * 16;
  const int64_t base = (int64_t)blockIdx.x * tile;
  double s = 0.0;
  for (int k = 0; k < 16; k++) {
    const int64_t i = base + (int64_t)k * 256 + threadIdx.x;
    if (i < n) s += in[i];
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) sm[threadIdx.x] += sm[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

}  // namespace

int xgpu_sum_f64(xgpu_ctx* ctx, double* buf_a, int64_t n, double* buf_b, double* result_host)
{
  // ping-pong tiles until one value is left; buf_b must hold ceil(n/4096) doubles
  double* in = buf_a; double* out = buf_b;
  if (n == 0) { *result_host = 0.0; return XGPU_OK; }
  while (n > 1) {
    const int64_t nb = (n + 4095) / 4096;
    k_sum_tiles<<<(unsigned)nb, 256, 0, ctx->stream>>>(in, n, out);
    XGPU_LAUNCH_CHECK(ctx);
    double* t = in; in = out; out = t;
    n = nb;
  }
  XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, in, sizeof(double), ctx->stream));
  XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *result_host = *(double*)ctx->h_pinned;
  return XGPU_OK;
}

extern "C" int xgpu_reduce_sum(xgpu_ctx* ctx, int64_t n, const double* vals, double* result_host)
{
  if (!ctx || n < 0 || !result_host || (n > 0 && !vals)) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_reduce_sum: bad argument");
  *result_host = 0.0;
  if (n == 0) return XGPU_OK;
  void* p;
  const int64_t nb = (n + 4095) / 4096;
  int rc = xgpu_scratch(ctx, 1, sizeof(double) * (n + nb + 2), &p);
  if (rc) return rc;
  double* buf = (double*)p;
  XGPU_CUDA(ctx, cudaMemcpyAsync(buf, vals, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  return xgpu_sum_f64(ctx, buf, n, buf + n, result_host);
}

extern "C" int xgpu_cutquad_count(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                  const double* phi, int dt, int order, int64_t* offsets, int64_t* total_host)
{
  XGPU_NVTX(ctx, "xgpu_cutquad_count");
  if (!ctx || dim < 1 || dim > 3 || dt < 0 || dt > 2) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_count: bad argument");
  if (total_host) *total_host = 0;
  const xgpu_rule_host* rv = xgpu_rule_lookup(ctx, dim, order);
  const xgpu_rule_host* rs = (dt == XGPU_IF) ? xgpu_rule_lookup(ctx, dim - 1, order) : rv;
  if (!rv || !rs) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_count: no rule");
  void* p;
  int rc = xgpu_scratch(ctx, 0, sizeof(int32_t) * (n + 1), &p);
  if (rc) return rc;
  int32_t* counts = (int32_t*)p;
  if (n > 0) {
    if (dim == 3) k_cq_count<3><<<nblk(n, 256), 256, 0, ctx->stream>>>(n, elems, el_vtx, phi, dt, rv->n, rs->n, counts);
    else if (dim == 2) k_cq_count<2><<<nblk(n, 256), 256, 0, ctx->stream>>>(n, elems, el_vtx, phi, dt, rv->n, rs->n, counts);
    else k_cq_count<1><<<nblk(n, 256), 256, 0, ctx->stream>>>(n, elems, el_vtx, phi, dt, rv->n, rs->n, counts);
    XGPU_LAUNCH_CHECK(ctx);
  }
  rc = xgpu_exclusive_scan_i32(ctx, counts, n, offsets);
  if (rc) return rc;
  if (total_host) {
    XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, offsets + n, sizeof(int64_t), ctx->stream));
    XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *total_host = ctx->h_pinned[0];
  }
  return XGPU_OK;
}

extern "C" int xgpu_cutquad_fill(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                 const double* vtx_xyz, const double* phi, int dt, int order,
                                 const int64_t* offsets, double* pts, double* wts)
{
  XGPU_NVTX(ctx, "xgpu_cutquad_fill");
  if (!ctx || dim < 1 || dim > 3 || dt < 0 || dt > 2) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_fill: bad argument");
  if (n == 0) return XGPU_OK;
  const xgpu_rule_host* rv = xgpu_rule_lookup(ctx, dim, order);
  const xgpu_rule_host* rs = (dt == XGPU_IF) ? xgpu_rule_lookup(ctx, dim - 1, order) : rv;
  if (!rv || !rs) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_fill: no rule");
  const int deformed_if = (ctx->def_order > 0 && dt == XGPU_IF) ? 1 : 0;
  if (deformed_if && dim < 2) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_cutquad_fill: deformation needs dim >= 2");
  const size_t sm = sizeof(double) * 4 * ((rs->n + 1) & ~1) + 128 * sizeof(CqDesc);
  unsigned grid = nblk(n, 128);  // 128 listed elements per block and pass
  if (grid > XGPU_NSM * 8 && !ctx->opt_short_blocks) grid = XGPU_NSM * 8;
#define XG_FILL(DD)                                                                                                  \
  {                                                                                                                  \
    if (sm > 48 * 1024) XGPU_CUDA(ctx, cudaFuncSetAttribute(k_cq_fill<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_cq_fill<DD><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phi, dt, rv->d_pts, rv->d_wts, rv->n,   \
                                                  rs->d_pts, rs->d_wts, rs->n, offsets, pts, wts, ctx->d_status,     \
                                                  ctx->d_status_elem, deformed_if);                                  \
  }
  if (dim == 3) XG_FILL(3) else if (dim == 2) XG_FILL(2) else XG_FILL(1)
#undef XG_FILL
  XGPU_LAUNCH_CHECK(ctx);
  // deformed mesh: the interface factor ||F^{-T} n_ref|| depends on the point (straightcutrule.cpp:558-578)
  if (deformed_if) return xgpu_if_weights_def(ctx, dim, n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts);
  return XGPU_OK;
}

extern "C" int xgpu_integrate_monomial(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                       const double* vtx_xyz, const int64_t* offsets, const double* pts,
                                       const double* wts, const int* expo_host, double* result_host)
{
  if (!ctx || dim < 1 || dim > 3 || !result_host) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_integrate_monomial: bad argument");
  *result_host = 0.0;
  if (n == 0) return XGPU_OK;
  void* p;
  const int64_t nb = (n + 4095) / 4096;
  int rc = xgpu_scratch(ctx, 1, sizeof(double) * (n + nb + 2), &p);
  if (rc) return rc;
  double* elsum = (double*)p;
  const int e0 = expo_host[0], e1 = dim > 1 ? expo_host[1] : 0, e2 = dim > 2 ? expo_host[2] : 0;
  if (ctx->def_order > 0) {
    rc = xgpu_integrate_def(ctx, dim, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum);
    if (rc) return rc;
    return xgpu_sum_f64(ctx, elsum, n, elsum + n, result_host);
  }
  const unsigned grid = nblk(n * 32, 128);
  if (dim == 3) k_integrate<3><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum);
  else if (dim == 2) k_integrate<2><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum);
  else k_integrate<1><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum);
  XGPU_LAUNCH_CHECK(ctx);
  return xgpu_sum_f64(ctx, elsum, n, elsum + n, result_host);
}

extern "C" int xgpu_integrate_monomial_st(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                          const double* vtx_xyz, const int64_t* offsets, const double* pts,
                                          const double* tvals, const double* wts, const int* expo_host, double* result_host)
{
  if (!ctx || dim < 1 || dim > 3 || !result_host || !expo_host)  // tvals may be NULL when the batch has no points at all
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_integrate_monomial_st: bad argument");
  if (ctx->def_order > 0) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "space-time integrals do not support a mesh deformation");
  *result_host = 0.0;
  if (n == 0) return XGPU_OK;
  void* p;
  const int64_t nb = (n + 4095) / 4096;
  int rc = xgpu_scratch(ctx, 1, sizeof(double) * (n + nb + 2), &p);
  if (rc) return rc;
  double* elsum = (double*)p;
  const int e0 = expo_host[0], e1 = dim > 1 ? expo_host[1] : 0, e2 = dim > 2 ? expo_host[2] : 0, et = expo_host[3];
  const unsigned grid = nblk(n * 32, 128);
  if (dim == 3) k_integrate<3><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum, tvals, et);
  else if (dim == 2) k_integrate<2><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum, tvals, et);
  else k_integrate<1><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum, tvals, et);
  XGPU_LAUNCH_CHECK(ctx);
  return xgpu_sum_f64(ctx, elsum, n, elsum + n, result_host);
}

extern "C" int xgpu_integrate_monomial_bnd(xgpu_ctx* ctx, int dim, int sdim, int64_t n, const int32_t* elems,
                                           const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets,
                                           const double* pts, const double* wts, const int* expo_host, double* result_host)
{
  if (!ctx || !result_host || !((dim == 1 && (sdim == 2 || sdim == 3)) || (dim == 2 && sdim == 3)))
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_integrate_monomial_bnd: segments in 2D/3D or triangles in 3D");
  if (ctx->def_order > 0) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_integrate_monomial_bnd: no deformation support");
  *result_host = 0.0;
  if (n == 0) return XGPU_OK;
  void* p;
  const int64_t nb = (n + 4095) / 4096;
  int rc = xgpu_scratch(ctx, 1, sizeof(double) * (n + nb + 2), &p);
  if (rc) return rc;
  double* elsum = (double*)p;
  const int e0 = expo_host[0], e1 = expo_host[1], e2 = sdim > 2 ? expo_host[2] : 0;
  const unsigned grid = nblk(n * 32, 128);
  if (dim == 1 && sdim == 2) k_integrate_emb<1, 2><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum);
  else if (dim == 1) k_integrate_emb<1, 3><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum);
  else k_integrate_emb<2, 3><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum);
  XGPU_LAUNCH_CHECK(ctx);
  return xgpu_sum_f64(ctx, elsum, n, elsum + n, result_host);
}
