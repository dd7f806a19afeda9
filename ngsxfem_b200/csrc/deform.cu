// deform.cu — geometry passes that depend on the isoparametric deformation (xgpu_set_deformation):
//   k_if_weights_def<D>  interface rules on a deformed mesh: w_q <- w_q ||F(xhat_q)^{-T} n_ref||
//                        (TransformQuadUntrafoToIRInterface, cutint/straightcutrule.cpp:558-578; the
//                        affine part of the factor is skipped in k_cq_fill, `untransformed`)
//   k_integrate_def<D>   sum_q w_q |det F(xhat_q)| x(xhat_q)^e  (cutint/cutintegral.cpp:126-212)
// One warp per listed element: the element's deformation coefficients are gathered into shared
// memory once, lanes stride over the element's points.  Compiled without FMA contraction like the
// other geometry kernels (build.py NO_FMA).
#include "common.cuh"
#include "cutgeom.cuh"
#include "deform.cuh"

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <int D>
__device__ __forceinline__ void load_vertices(const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                              int64_t el, int* v, double (*vx)[3])
{
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
    v[k] = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
    for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[k] * D + c) : 0.0;
  }
}

template <int D>
__global__ void __launch_bounds__(128) k_if_weights_def(int64_t n, const int32_t* __restrict__ elems,
                                                        const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                                        const double* __restrict__ phi, const int64_t* __restrict__ offsets,
                                                        const double* __restrict__ pts, double* __restrict__ wts, DefDev df)
{
  __shared__ double sdefs[4][XG_DEFS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* defs = sdefs[wib];
  const int64_t warp0 = (int64_t)blockIdx.x * 4 + wib, nwarps = (int64_t)gridDim.x * 4;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t q0 = offsets[i], q1 = offsets[i + 1];
    if (q1 == q0) continue;
    const int64_t el = elems ? (int64_t)elems[i] : i;
    int v[D + 1];
    double vx[D + 1][3], l[D + 1], F[D * D];
    load_vertices<D>(el_vtx, vtx_xyz, el, v, vx);
#pragma unroll
    for (int k = 0; k < D + 1; k++) l[k] = __ldg(phi + v[k]);
    xg_affine_jac<D>(vx, F);
    __syncwarp();
    xg_def_gather<D>(df, el, lane, defs);
    __syncwarp();
    for (int64_t q = q0 + lane; q < q1; q += 32) {
      const double xh[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
      double Fq[D * D], Fiq[D * D], xd[3] = {0, 0, 0};
#pragma unroll
      for (int k = 0; k < D * D; k++) Fq[k] = F[k];
      xg_def_apply<D>(df.order, v, xh, defs, Fq, xd);
      xg_inv<D>(Fq, Fiq);
      wts[q] *= xg_if_weight_factor<D>(Fiq, l);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(128) k_integrate_def(int64_t n, const int32_t* __restrict__ elems,
                                                       const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                                       const int64_t* __restrict__ offsets, const double* __restrict__ pts,
                                                       const double* __restrict__ wts, int e0, int e1, int e2,
                                                       double* __restrict__ elsum, DefDev df)
{
  __shared__ double sdefs[4][XG_DEFS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* defs = sdefs[wib];
  const int64_t i = (int64_t)blockIdx.x * 4 + wib;
  if (i >= n) return;
  const int64_t el = elems ? (int64_t)elems[i] : i;
  int v[D + 1];
  double vx[D + 1][3], F[D * D];
  load_vertices<D>(el_vtx, vtx_xyz, el, v, vx);
  xg_affine_jac<D>(vx, F);
  xg_def_gather<D>(df, el, lane, defs);
  __syncwarp();
  const int ex[3] = {e0, e1, e2};
  double s = 0.0;
  for (int64_t q = offsets[i] + lane; q < offsets[i + 1]; q += 32) {
    const double xh[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
    double lam[D + 1], sl = 1.0, x[3] = {0, 0, 0}, Fq[D * D], Fiq[D * D];
#pragma unroll
    for (int k = 0; k < D; k++) { lam[k] = xh[k]; sl -= lam[k]; }
    lam[D] = sl;
#pragma unroll
    for (int c = 0; c < D; c++) {
      double xc = 0.0;
#pragma unroll
      for (int k = 0; k < D + 1; k++) xc += lam[k] * vx[k][c];
      x[c] = xc;
    }
#pragma unroll
    for (int k = 0; k < D * D; k++) Fq[k] = F[k];
    xg_def_apply<D>(df.order, v, xh, defs, Fq, x);
    const double meas = fabs(xg_inv<D>(Fq, Fiq));
    double f = 1.0;
#pragma unroll
    for (int c = 0; c < D; c++)
      for (int e = 0; e < ex[c]; e++) f *= x[c];
    s += wts[q] * meas * f;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if (lane == 0) elsum[i] = s;
}

// physical coordinates and measures of the points of per-element rules — where NGSolve evaluates a CoefficientFunction
// (mip.GetPoint(), mip.GetMeasure(); xfem/symboliccutbfi.cpp:174-193); deformation-aware (df.order > 0)
template <int D>
__global__ void __launch_bounds__(128) k_rule_phys(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                                   const double* __restrict__ vtx_xyz, const int64_t* __restrict__ offsets,
                                                   const double* __restrict__ pts, double* __restrict__ xyz,
                                                   double* __restrict__ meas_out, DefDev df)
{
  __shared__ double sdefs[4][XG_DEFS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* defs = sdefs[wib];
  const int64_t warp0 = (int64_t)blockIdx.x * 4 + wib, nwarps = (int64_t)gridDim.x * 4;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t q0 = offsets[i], q1 = offsets[i + 1];
    if (q1 == q0) continue;
    const int64_t el = elems ? (int64_t)elems[i] : i;
    int v[D + 1];
    double vx[D + 1][3], F[D * D];
    load_vertices<D>(el_vtx, vtx_xyz, el, v, vx);
    xg_affine_jac<D>(vx, F);
    if (df.order > 0) {
      __syncwarp();
      xg_def_gather<D>(df, el, lane, defs);
      __syncwarp();
    }
    for (int64_t q = q0 + lane; q < q1; q += 32) {
      const double xh[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
      double lam[D + 1], sl = 1.0, x[3] = {0, 0, 0}, Fq[D * D], Fiq[D * D];
#pragma unroll
      for (int k = 0; k < D; k++) { lam[k] = xh[k]; sl -= lam[k]; }
      lam[D] = sl;
#pragma unroll
      for (int c = 0; c < D; c++) {
        double xc = 0.0;
#pragma unroll
        for (int k = 0; k < D + 1; k++) xc += lam[k] * vx[k][c];
        x[c] = xc;
      }
#pragma unroll
      for (int k = 0; k < D * D; k++) Fq[k] = F[k];
      if (df.order > 0) xg_def_apply<D>(df.order, v, xh, defs, Fq, x);
#pragma unroll
      for (int c = 0; c < D; c++) xyz[q * D + c] = x[c];
      if (meas_out) meas_out[q] = fabs(xg_inv<D>(Fq, Fiq));
    }
  }
}

template <int D>
struct SmemDotSink {
  const double* c;
  double v;
  __device__ __forceinline__ void operator()(int i, const AD<D>& s) { v += c[i] * s.v; }
};

// values of an H1(order) function at the points of per-element rules (GridFunction evaluation at integration points)
template <int D>
__global__ void __launch_bounds__(128) k_h1_eval_rule(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                                      int order, int nd, const int32_t* __restrict__ el_dofs,
                                                      const double* __restrict__ vals, const int64_t* __restrict__ offsets,
                                                      const double* __restrict__ pts, double* __restrict__ out)
{
  __shared__ double sc[4][36];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * 4 + wib, nwarps = (int64_t)gridDim.x * 4;
  for (int64_t i = warp0; i < n; i += nwarps) {
    const int64_t q0 = offsets[i], q1 = offsets[i + 1];
    if (q1 == q0) continue;
    const int64_t el = elems ? (int64_t)elems[i] : i;
    int v[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; k++) v[k] = __ldg(el_vtx + el * (D + 1) + k);
    __syncwarp();
    for (int k = lane; k < nd; k += 32) {
      const int32_t d = __ldg(el_dofs + el * nd + k);
      sc[wib][k] = d >= 0 ? __ldg(vals + d) : 0.0;
    }
    __syncwarp();
    for (int64_t q = q0 + lane; q < q1; q += 32) {
      const double xh[3] = {pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]};
      SmemDotSink<D> sk{sc[wib], 0.0};
      xg_h1_shapes_rt<D>(order, v, xh, sk);
      out[q] = sk.v;
    }
  }
}

}  // namespace

extern "C" int xgpu_rule_phys(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
                              const int64_t* offsets, const double* pts, double* xyz_out, double* meas_out)
{
  XGPU_NVTX(ctx, "xgpu_rule_phys");
  if (!ctx || dim < 1 || dim > 3 || n < 0 || !offsets) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_rule_phys: bad argument");
  if (n == 0) return XGPU_OK;
  const DefDev df = xg_defdev(ctx, dim);
  if (df.order > 0 && dim < 2) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "deformation needs dim >= 2");
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  if (dim == 3) k_rule_phys<3><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, xyz_out, meas_out, df);
  else if (dim == 2) k_rule_phys<2><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, xyz_out, meas_out, df);
  else k_rule_phys<1><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, xyz_out, meas_out, df);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_h1_eval_rule(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                 const int32_t* el_dofs, const double* vals, const int64_t* offsets, const double* pts, double* out)
{
  XGPU_NVTX(ctx, "xgpu_h1_eval_rule");
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || n < 0 || !offsets || !el_dofs || !vals)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_h1_eval_rule: bad argument (orders 1..4)");
  if (n == 0) return XGPU_OK;
  const int nd = xg_h1_ndof(dim, fe_order);
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  if (dim == 3) k_h1_eval_rule<3><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, fe_order, nd, el_dofs, vals, offsets, pts, out);
  else if (dim == 2) k_h1_eval_rule<2><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, fe_order, nd, el_dofs, vals, offsets, pts, out);
  else k_h1_eval_rule<1><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, fe_order, nd, el_dofs, vals, offsets, pts, out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

int xgpu_if_weights_def(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                        const double* vtx_xyz, const double* phi, const int64_t* offsets, const double* pts, double* wts)
{
  if (n == 0) return XGPU_OK;
  const DefDev df = xg_defdev(ctx, dim);
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  if (dim == 3) k_if_weights_def<3><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, df);
  else if (dim == 2) k_if_weights_def<2><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, df);
  else return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "deformation needs dim >= 2");
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

int xgpu_integrate_def(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                       const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts,
                       int e0, int e1, int e2, double* elsum)
{
  if (n == 0) return XGPU_OK;
  const DefDev df = xg_defdev(ctx, dim);
  const unsigned grid = nblk(n, 4);
  if (dim == 3) k_integrate_def<3><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum, df);
  else if (dim == 2) k_integrate_def<2><<<grid, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, e0, e1, e2, elsum, df);
  else return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "deformation needs dim >= 2");
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}
