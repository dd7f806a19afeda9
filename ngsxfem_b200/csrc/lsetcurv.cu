// lsetcurv.cu — the isoparametric deformation producer: LevelSetMeshAdaptation.CalcDeformation
// (python/lsetcurv.py:281-356) = lset_ho.Set(levelset), qn.Set(grad lset_ho), InterpolateToP1, ProjectShift
// (lsetcurving/projshift.cpp:9-201) with ShiftIntegrator::CalcElementVector (lsetcurving/shiftintegrators.cpp:39-124),
// SearchCorrespondingPoint (lsetcurving/calcpointshift.cpp:145-206) and ElementInRelevantBand (:120-141).
//
//   k_band<D>            thread per element, ballot -> BitArray words
//   k_rule_points<D>     thread per (element, rule point): physical coordinates (where a user level-set function is sampled)
//   k_h1_grad_points<D>  thread per (element, rule point): physical gradient of an H1(order) function (qn.Set(lset_ho.Deriv()))
//   k_shift_points<D>    thread per (element, rule point): the Newton search along the search direction for the point where
//                        lset_ho takes the value of lset_p1 (<= 20 steps, |defect| < 1e-14), threshold clamp in reference
//                        coordinates, physical shift F * (xhat_final - xhat)
//   k_l2_project<D>      WARP per element: element-local L2 projection M^{-1} int f phi_i (shapes of all rule points staged in
//                        shared memory, mass matrix + right-hand sides by the lanes, Cholesky in shared memory) — both halves
//                        of GridFunction.Set and the projection of ProjectShift (vertex rows zeroed there, :108-109).
// The sums / means over the elements sharing a dof go through the transposed table (xgpu_adjacency_plan +
// xgpu_vec_scatter_mean): fixed order, bit-reproducible.
#include "common.cuh"
#include "cutgeom.cuh"
#include "deform.cuh"

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

#define XL_MAXND 35
#define XL_MAXNP 64

template <int D>
__global__ void __launch_bounds__(256) k_band(int64_t ne, const int32_t* __restrict__ el_vtx, const double* __restrict__ lp1,
                                              double lower, double upper, uint32_t* __restrict__ bits)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool in = false;
  if (e < ne) {
    bool hn = false, hp = false;
#pragma unroll
    for (int k = 0; k < D + 1; k++) {
      const double v = __ldg(lp1 + __ldg(el_vtx + e * (D + 1) + k));
      if (v == 0.0) hn = hp = true;
      if (v > lower) hp = true;
      if (v < upper) hn = true;
    }
    in = hn && hp;
  }
  const unsigned w = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && e < ne) bits[e >> 5] = w;
}

template <int D>
__device__ __forceinline__ void load_geo(const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz, int64_t el,
                                         int* v, double (*vx)[3])
{
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
    v[k] = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
    for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[k] * D + c) : 0.0;
  }
}

template <int D>
__global__ void __launch_bounds__(256) k_rule_points(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                                     const double* __restrict__ vtx_xyz, const double* __restrict__ rp, int np,
                                                     double* __restrict__ out)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * np) return;
  const int64_t i = t / np;
  const int q = (int)(t - i * np);
  const int64_t el = elems ? (int64_t)elems[i] : i;
  int v[D + 1]; double vx[D + 1][3];
  load_geo<D>(el_vtx, vtx_xyz, el, v, vx);
  double lam[D + 1], s = 1.0;
#pragma unroll
  for (int k = 0; k < D; k++) { lam[k] = rp[3 * q + k]; s -= lam[k]; }
  lam[D] = s;
#pragma unroll
  for (int c = 0; c < D; c++) {
    double x = 0.0;
#pragma unroll
    for (int k = 0; k < D + 1; k++) x += lam[k] * vx[k][c];
    out[t * D + c] = x;
  }
}

// value and reference gradient of sum_i vals[dofs[i]] phi_i
template <int D>
struct DotSink {
  const int32_t* dofs;
  const double* vals;
  double v;
  double g[D];
  __device__ __forceinline__ void operator()(int i, const AD<D>& s)
  {
    const double c = __ldg(vals + __ldg(dofs + i));
    v += c * s.v;
#pragma unroll
    for (int k = 0; k < D; k++) g[k] += c * s.d[k];
  }
};

// value of a vector field sum_i q[dofs[i]][:] phi_i
template <int D>
struct VecSink {
  const int32_t* dofs;
  const double* q;
  double v[D];
  __device__ __forceinline__ void operator()(int i, const AD<D>& s)
  {
    const int64_t d = __ldg(dofs + i);
#pragma unroll
    for (int k = 0; k < D; k++) v[k] += __ldg(q + d * D + k) * s.v;
  }
};

template <int D>
__device__ __forceinline__ void eval_ho(int order, const int* v, const int32_t* dofs, const double* vals, const double* xh,
                                        double& val, double* gref)
{
  DotSink<D> sk{dofs, vals, 0.0, {}};
#pragma unroll
  for (int k = 0; k < D; k++) sk.g[k] = 0.0;
  xg_h1_shapes_rt<D>(order, v, xh, sk);
  val = sk.v;
#pragma unroll
  for (int k = 0; k < D; k++) gref[k] = sk.g[k];
}

template <int D>
__global__ void __launch_bounds__(128) k_h1_grad_points(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                                        const double* __restrict__ vtx_xyz, int order, int nd,
                                                        const int32_t* __restrict__ el_dofs, const double* __restrict__ vals,
                                                        const double* __restrict__ rp, int np, double* __restrict__ out)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * np) return;
  const int64_t i = t / np;
  const int q = (int)(t - i * np);
  const int64_t el = elems ? (int64_t)elems[i] : i;
  int v[D + 1]; double vx[D + 1][3], F[D * D], Fi[D * D];
  load_geo<D>(el_vtx, vtx_xyz, el, v, vx);
  xg_affine_jac<D>(vx, F);
  xg_inv<D>(F, Fi);
  const double xh[3] = {rp[3 * q], rp[3 * q + 1], rp[3 * q + 2]};
  double val, gr[D];
  eval_ho<D>(order, v, el_dofs + el * nd, vals, xh, val, gr);
#pragma unroll
  for (int c = 0; c < D; c++) {   // F^{-T} grad_ref
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) s += Fi[j * D + c] * gr[j];
    out[t * D + c] = s;
  }
}

template <int D>
__global__ void __launch_bounds__(128) k_shift_points(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                                      const double* __restrict__ vtx_xyz, int order, int nd,
                                                      const int32_t* __restrict__ el_dofs, const double* __restrict__ lset_ho,
                                                      const double* __restrict__ lset_p1, int qn_mode, const double* __restrict__ qn,
                                                      const double* __restrict__ blending, double threshold,
                                                      const double* __restrict__ rp, int np, double* __restrict__ out,
                                                      unsigned long long* __restrict__ nfail)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * np) return;
  const int64_t i = t / np;
  const int q = (int)(t - i * np);
  const int64_t el = elems ? (int64_t)elems[i] : i;
  int v[D + 1]; double vx[D + 1][3], F[D * D], Fi[D * D];
  load_geo<D>(el_vtx, vtx_xyz, el, v, vx);
  xg_affine_jac<D>(vx, F);
  xg_inv<D>(F, Fi);
  const int32_t* dofs = el_dofs + el * nd;
  const double xh[3] = {rp[3 * q], rp[3 * q + 1], rp[3 * q + 2]};
  double lp1[D + 1];
#pragma unroll
  for (int k = 0; k < D + 1; k++) lp1[k] = __ldg(lset_p1 + v[k]);
  // search direction: physical vector `grad` (qn), then normal = F^{-1} grad in reference coordinates (shiftintegrators.cpp:77-80)
  double v0, g0[D], grad[D];
  eval_ho<D>(order, v, dofs, lset_ho, xh, v0, g0);
  if (qn_mode == 2) {
    VecSink<D> vs{dofs, qn, {}};
#pragma unroll
    for (int k = 0; k < D; k++) vs.v[k] = 0.0;
    xg_h1_shapes_rt<D>(order, v, xh, vs);
#pragma unroll
    for (int k = 0; k < D; k++) grad[k] = vs.v[k];
  } else {
    double gr[D];
#pragma unroll
    for (int j = 0; j < D; j++) gr[j] = (qn_mode == 0) ? lp1[j] - lp1[D] : g0[j];
#pragma unroll
    for (int c = 0; c < D; c++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) s += Fi[j * D + c] * gr[j];
      grad[c] = s;
    }
  }
  double dir[D];
#pragma unroll
  for (int a = 0; a < D; a++) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) s += Fi[a * D + j] * grad[j];
    dir[a] = s;
  }
  // goal value (:84-98): (1 - alpha) lset_p1(x) + alpha lset_ho(x)
  double ll = 1.0, p1v = 0.0;
#pragma unroll
  for (int j = 0; j < D; j++) { p1v += xh[j] * lp1[j]; ll -= xh[j]; }
  p1v += ll * lp1[D];
  const double alpha = blending ? blending[t] : 0.0;
  const double goal = (1.0 - alpha) * p1v + alpha * v0;
  // SearchCorrespondingPoint, fixed search direction (calcpointshift.cpp:145-206)
  double cur[3] = {xh[0], xh[1], xh[2]};
  int it;
  for (it = 0; it < 20; it++) {
    double cv, cg[D];
    eval_ho<D>(order, v, dofs, lset_ho, cur, cv, cg);
    const double defect = goal - cv;
    if (fabs(defect) < 1e-14) break;
    double dphidn = 0.0;
#pragma unroll
    for (int c = 0; c < D; c++) dphidn += cg[c] * dir[c];
#pragma unroll
    for (int c = 0; c < D; c++) cur[c] += defect / dphidn * dir[c];
  }
  double rd[D];
  if (it == 20) {
    atomicAdd(nfail, 1ull);
#pragma unroll
    for (int c = 0; c < D; c++) rd[c] = 0.0;
  } else {
#pragma unroll
    for (int c = 0; c < D; c++) rd[c] = cur[c] - xh[c];
  }
  double len = 0.0;
#pragma unroll
  for (int c = 0; c < D; c++) len += rd[c] * rd[c];
  len = sqrt(len);
  if (threshold >= 0.0 && len > threshold) {   // max_deform in REFERENCE coordinates (:108-111)
#pragma unroll
    for (int c = 0; c < D; c++) rd[c] *= threshold / len;
  }
#pragma unroll
  for (int a = 0; a < D; a++) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) s += F[a * D + j] * rd[j];
    out[t * D + a] = s;
  }
}

template <int D>
struct ShapeRowSink {
  double* row;
  __device__ __forceinline__ void operator()(int i, const AD<D>& s) { row[i] = s.v; }
};

// element-local L2 projection: warp per element.  shared per warp: S[np][nd] | M[nd][nd] | R[nd][ncomp]
template <int D>
__global__ void __launch_bounds__(64) k_l2_project(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                                   int order, int nd, const double* __restrict__ rp, const double* __restrict__ rw, int np,
                                                   int ncomp, const double* __restrict__ fvals, int zero_vertex,
                                                   double* __restrict__ elres)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int per = np * nd + nd * nd + nd * ncomp;
  double* S = smem + (size_t)wib * per;
  double* M = S + np * nd;
  double* R = M + nd * nd;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (i >= n) return;
  const int64_t el = elems ? (int64_t)elems[i] : i;
  int v[D + 1];
#pragma unroll
  for (int k = 0; k < D + 1; k++) v[k] = __ldg(el_vtx + el * (D + 1) + k);
  for (int q = lane; q < np; q += 32) {
    const double xh[3] = {rp[3 * q], rp[3 * q + 1], rp[3 * q + 2]};
    ShapeRowSink<D> sk{S + q * nd};
    xg_h1_shapes_rt<D>(order, v, xh, sk);
  }
  __syncwarp();
  for (int k = lane; k < nd * nd; k += 32) {
    const int a = k / nd, b = k - a * nd;
    double s = 0.0;
    for (int q = 0; q < np; q++) s += rw[q] * S[q * nd + a] * S[q * nd + b];
    M[k] = s;
  }
  for (int k = lane; k < nd * ncomp; k += 32) {
    const int a = k / ncomp, c = k - a * ncomp;
    double s = 0.0;
    for (int q = 0; q < np; q++) s += rw[q] * S[q * nd + a] * fvals[((int64_t)i * np + q) * ncomp + c];
    R[k] = s;
  }
  __syncwarp();
  // Cholesky M = L L^T (lower triangle in place), right-looking
  for (int j = 0; j < nd; j++) {
    if (lane == 0) M[j * nd + j] = sqrt(M[j * nd + j]);
    __syncwarp();
    const double ljj = M[j * nd + j];
    for (int a = j + 1 + lane; a < nd; a += 32) M[a * nd + j] /= ljj;
    __syncwarp();
    for (int a = j + 1 + lane; a < nd; a += 32) {
      const double laj = M[a * nd + j];
      for (int b = j + 1; b <= a; b++) M[a * nd + b] -= laj * M[b * nd + j];
    }
    __syncwarp();
  }
  // forward / backward substitution: lane = component
  if (lane < ncomp) {
    const int c = lane;
    for (int a = 0; a < nd; a++) {
      double t = R[a * ncomp + c];
      for (int k = 0; k < a; k++) t -= M[a * nd + k] * R[k * ncomp + c];
      R[a * ncomp + c] = t / M[a * nd + a];
    }
    for (int a = nd - 1; a >= 0; a--) {
      double t = R[a * ncomp + c];
      for (int k = a + 1; k < nd; k++) t -= M[k * nd + a] * R[k * ncomp + c];
      R[a * ncomp + c] = t / M[a * nd + a];
    }
  }
  __syncwarp();
  for (int k = lane; k < nd * ncomp; k += 32) {
    const int a = k / ncomp, c = k - a * ncomp;
    const double r = (zero_vertex && a < D + 1) ? 0.0 : R[k];
    elres[((int64_t)c * n + i) * nd + a] = r;   // component-major: one n x nd batch per component (xgpu_vec_scatter_mean)
  }
}

}  // namespace

extern "C" int xgpu_band_elements(xgpu_ctx* ctx, int dim, int64_t ne, const int32_t* el_vtx, const double* lset_p1,
                                  double lower, double upper, uint32_t* bits_out)
{
  XGPU_NVTX(ctx, "xgpu_band_elements");
  if (!ctx || dim < 1 || dim > 3 || ne < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_band_elements: bad argument");
  if (ne == 0) return XGPU_OK;
  if (dim == 3) k_band<3><<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, el_vtx, lset_p1, lower, upper, bits_out);
  else if (dim == 2) k_band<2><<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, el_vtx, lset_p1, lower, upper, bits_out);
  else k_band<1><<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, el_vtx, lset_p1, lower, upper, bits_out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_rule_points_phys(xgpu_ctx* ctx, int dim, int order_rule, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                     const double* vtx_xyz, double* out)
{
  XGPU_NVTX(ctx, "xgpu_rule_points_phys");
  if (!ctx || dim < 1 || dim > 3 || n < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_rule_points_phys: bad argument");
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order_rule);
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_rule_points_phys: no rule of that order");
  if (n == 0) return XGPU_OK;
  const unsigned g = nblk(n * r->n, 256);
  if (dim == 3) k_rule_points<3><<<g, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, r->d_pts, r->n, out);
  else if (dim == 2) k_rule_points<2><<<g, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, r->d_pts, r->n, out);
  else k_rule_points<1><<<g, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, r->d_pts, r->n, out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_h1_grad_points(xgpu_ctx* ctx, int dim, int fe_order, int order_rule, int64_t n, const int32_t* elems,
                                   const int32_t* el_vtx, const double* vtx_xyz, const int32_t* el_dofs, const double* vals,
                                   double* out)
{
  XGPU_NVTX(ctx, "xgpu_h1_grad_points");
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || n < 0 || !el_dofs || !vals)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_h1_grad_points: bad argument (orders 1..4)");
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order_rule);
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_h1_grad_points: no rule of that order");
  if (n == 0) return XGPU_OK;
  const int nd = xg_h1_ndof(dim, fe_order);
  const unsigned g = nblk(n * r->n, 128);
  if (dim == 3) k_h1_grad_points<3><<<g, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, fe_order, nd, el_dofs, vals, r->d_pts, r->n, out);
  else if (dim == 2) k_h1_grad_points<2><<<g, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, fe_order, nd, el_dofs, vals, r->d_pts, r->n, out);
  else k_h1_grad_points<1><<<g, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, fe_order, nd, el_dofs, vals, r->d_pts, r->n, out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_shift_points(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                 const double* vtx_xyz, const int32_t* el_dofs, const double* lset_ho, const double* lset_p1,
                                 int qn_mode, const double* qn, const double* blending, double threshold, double* shift_out,
                                 int64_t* nfail_host)
{
  XGPU_NVTX(ctx, "xgpu_shift_points");
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || n < 0 || !el_dofs || !lset_ho || !lset_p1 ||
      qn_mode < 0 || qn_mode > 2 || (qn_mode == 2 && !qn))
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_shift_points: bad argument (orders 1..4, qn_mode 0..2)");
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, 2 * fe_order);   // shiftintegrators.cpp:69
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_shift_points: no rule of that order");
  if (nfail_host) *nfail_host = 0;
  if (n == 0) return XGPU_OK;
  void* sp;
  int rc = xgpu_scratch(ctx, 0, sizeof(unsigned long long), &sp);
  if (rc) return rc;
  unsigned long long* nfail = (unsigned long long*)sp;
  XGPU_CUDA(ctx, cudaMemsetAsync(nfail, 0, sizeof(unsigned long long), ctx->stream));
  const int nd = xg_h1_ndof(dim, fe_order);
  const unsigned g = nblk(n * r->n, 128);
#define XL_SHIFT(DD) k_shift_points<DD><<<g, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, fe_order, nd, el_dofs, lset_ho, lset_p1, \
                                                                   qn_mode, qn, blending, threshold, r->d_pts, r->n, shift_out, nfail)
  if (dim == 3) XL_SHIFT(3); else if (dim == 2) XL_SHIFT(2); else XL_SHIFT(1);
#undef XL_SHIFT
  XGPU_LAUNCH_CHECK(ctx);
  if (nfail_host) {
    XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, nfail, sizeof(unsigned long long), ctx->stream));
    XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *nfail_host = ctx->h_pinned[0];
  }
  return XGPU_OK;
}

extern "C" int xgpu_l2_project(xgpu_ctx* ctx, int dim, int fe_order, int order_rule, int64_t n, const int32_t* elems,
                               const int32_t* el_vtx, int ncomp, const double* fvals, int zero_vertex, double* elres)
{
  XGPU_NVTX(ctx, "xgpu_l2_project");
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || n < 0 || ncomp < 1 || ncomp > 32 || (n > 0 && (!fvals || !elres)))
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_l2_project: bad argument (orders 1..4, 1..32 components)");
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order_rule);
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_l2_project: no rule of that order");
  if (order_rule < 2 * fe_order) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_l2_project: the rule must integrate the mass matrix (order >= 2 fe_order)");
  if (n == 0) return XGPU_OK;
  const int nd = xg_h1_ndof(dim, fe_order);
  const size_t sm = sizeof(double) * 2 * ((size_t)r->n * nd + (size_t)nd * nd + (size_t)nd * ncomp);
  if (sm > 220 * 1024) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "xgpu_l2_project: rule too large for the shared-memory staging");
  const unsigned g = nblk(n, 2);
#define XL_PROJ(DD)                                                                                                        \
  {                                                                                                                        \
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_l2_project<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));           \
    k_l2_project<DD><<<g, 64, sm, ctx->stream>>>(n, elems, el_vtx, fe_order, nd, r->d_pts, r->d_wts, r->n, ncomp, fvals,   \
                                                 zero_vertex, elres);                                                      \
  }
  if (dim == 3) XL_PROJ(3) else if (dim == 2) XL_PROJ(2) else XL_PROJ(1)
#undef XL_PROJ
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}
