// elvec.cu — batched cut / uncut element VECTORS  sum_q B_q^T (f(x_q) w_q |det F|)
// (SymbolicCutLinearFormIntegrator::T_CalcElementVector, xfem/symboliccutlfi.cpp:51-140) for the
// closed menu  f(x) * op_test(v),  f = coef * h^hpow * x^e0 y^e1 z^e2,  op_test in {ID, DN}.
// One WARP per element, same staging as the element-matrix kernels: every lane evaluates the H1
// basis (values + normal derivatives) and the folded coefficients at one point of the current
// chunk into shared memory, then lane i sums dof i (i, i+32, ...) over the staged points in point
// order (fixed summation order: bit reproducible).  HBM: 16 B + vertex gathers + 32 B/point in,
// nd * 8 B out per element.
#include "common.cuh"
#include "cutgeom.cuh"
#include "deform.cuh"

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

#define XG_MAXLTERMS 8
struct LTermsDev {
  int n;
  int op_test[XG_MAXLTERMS], hpow[XG_MAXLTERMS], expo[XG_MAXLTERMS][3];
  double coef[XG_MAXLTERMS];
};

__device__ __forceinline__ double ipow_d(double h, int p)
{
  double r = 1.0;
  const int ap = p < 0 ? -p : p;
  for (int i = 0; i < ap; i++) r *= h;
  return p < 0 ? 1.0 / r : r;
}

template <int D>
struct VecSink {
  double* B;         // this point's slab: [val | dn], each ndp long
  int ndp;
  const double* Fi;
  const double* nrm;
  __device__ __forceinline__ void operator()(int i, const AD<D>& a)
  {
    B[i] = a.v;
    double dn = 0.0;
#pragma unroll
    for (int c = 0; c < D; c++) {
      double g = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) g += Fi[j * D + c] * a.d[j];
      dn += g * nrm[c];
    }
    B[ndp + i] = dn;
  }
};

template <int D, bool DEF>
__global__ void __launch_bounds__(128) k_elvec_points(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
    const double* __restrict__ vtx_xyz, const double* __restrict__ phi,
    const int64_t* __restrict__ offsets, const double* __restrict__ pts, const double* __restrict__ wts, int np_shared,
    int p, int nd, LTermsDev T, int accumulate, const int32_t* __restrict__ slots, double* __restrict__ vecs, DefDev df)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int ndp = (nd + 1) & ~1;
  const int PS = 2 * ndp + 2;  // per point: val, dn, (ID weight, DN weight)
  double* Bw = smem + (size_t)wib * (32 * PS + (DEF ? XG_DEFS : 0));
  double* defs = Bw + (size_t)32 * PS;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = warp0; e < n; e += nwarps) {
    const int64_t el = elems ? (int64_t)elems[e] : e;
    int v[D + 1];
    double vx[D + 1][3], l[D + 1], F[D * D], Fi[D * D], nrm[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < D + 1; k++) {
      v[k] = __ldg(el_vtx + el * (D + 1) + k);
      l[k] = phi ? __ldg(phi + v[k]) : 0.0;
#pragma unroll
      for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[k] * D + c) : 0.0;
    }
    xg_affine_jac<D>(vx, F);
    const double meas = fabs(xg_inv<D>(F, Fi));
    const double h = xg_mesh_size<D>(meas);
    if (phi) xg_phys_normal<D>(Fi, l, nrm);
    if (DEF) { __syncwarp(); xg_def_gather<D>(df, el, lane, defs); __syncwarp(); }
    const int64_t q0 = offsets ? offsets[e] : 0;
    const int npe = offsets ? (int)(offsets[e + 1] - q0) : np_shared;
    double acc[2] = {0.0, 0.0};  // dofs lane, lane + 32
    for (int c0 = 0; c0 < npe; c0 += 32) {
      const int cs = min(32, npe - c0);
      __syncwarp();
      if (lane < cs) {
        double* B = Bw + (size_t)lane * PS;
        const double* xq = pts + 3 * (q0 + c0 + lane);
        const double xh[3] = {xq[0], xq[1], xq[2]};
        double Fq[D * D], Fiq[D * D], nq[3] = {0, 0, 0}, x[3] = {0, 0, 0}, measq = meas, hq = h;
        {  // physical point (coefficient evaluation)
          double s = 1.0;
#pragma unroll
          for (int k = 0; k < D; k++) s -= xh[k];
#pragma unroll
          for (int c = 0; c < D; c++) {
            double t = s * vx[D][c];
#pragma unroll
            for (int k = 0; k < D; k++) t += xh[k] * vx[k][c];
            x[c] = t;
          }
        }
        if (DEF) {
#pragma unroll
          for (int k = 0; k < D * D; k++) Fq[k] = F[k];
          xg_def_apply<D>(df.order, v, xh, defs, Fq, x);
          measq = fabs(xg_inv<D>(Fq, Fiq));
          hq = xg_mesh_size<D>(measq);
          if (phi) xg_phys_normal<D>(Fiq, l, nq);
        }
        for (int i = nd; i < ndp; i++) { B[i] = 0.0; B[ndp + i] = 0.0; }
        VecSink<D> sink{B, ndp, DEF ? Fiq : Fi, DEF ? nq : nrm};
        xg_h1_shapes_rt<D>(p, v, xh, sink);
        const double fac = wts[q0 + c0 + lane] * measq;
        double cid = 0.0, cdn = 0.0;
        for (int t = 0; t < T.n; t++) {
          double c = T.coef[t] * fac;
          if (T.hpow[t] != 0) c *= ipow_d(hq, T.hpow[t]);
#pragma unroll
          for (int k = 0; k < D; k++) for (int ee = 0; ee < T.expo[t][k]; ee++) c *= x[k];
          if (T.op_test[t] == 0) cid += c; else cdn += c;
        }
        B[2 * ndp] = cid; B[2 * ndp + 1] = cdn;
      }
      __syncwarp();
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int i = lane + 32 * b;
        if (i < nd)
          for (int q = 0; q < cs; q++) {
            const double* B = Bw + (size_t)q * PS;
            acc[b] += B[2 * ndp] * B[i] + B[2 * ndp + 1] * B[ndp + i];
          }
      }
    }
    double* V = vecs + (size_t)(slots ? (int64_t)slots[e] : e) * nd;
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int i = lane + 32 * b;
      if (i < nd) { if (accumulate) V[i] += acc[b]; else V[i] = acc[b]; }
    }
  }
}

int make_lterms(xgpu_ctx* ctx, const xgpu_lterm* th, int nt, int dim, LTermsDev& T)
{
  if (nt < 0 || nt > XG_MAXLTERMS) return xgpu_fail(ctx, XGPU_ERR_ARG, "elvec: too many terms (max 8)");
  T.n = nt;
  for (int t = 0; t < nt; t++) {
    if (th[t].op_test != 0 && th[t].op_test != 2)
      return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elvec: unsupported test operator (menu: ID, DN)");
    T.op_test[t] = th[t].op_test; T.hpow[t] = th[t].hpow; T.coef[t] = th[t].coef;
    for (int k = 0; k < 3; k++) {
      if (th[t].expo[k] < 0 || (k >= dim && th[t].expo[k] != 0)) return xgpu_fail(ctx, XGPU_ERR_ARG, "elvec: bad monomial exponent");
      T.expo[t][k] = th[t].expo[k];
    }
  }
  return XGPU_OK;
}

template <int D>
int launch_vec(xgpu_ctx* ctx, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
               const double* phi, const int64_t* offsets, const double* pts, const double* wts, int np_shared,
               int p, const LTermsDev& T, int accumulate, const int32_t* slots, double* vecs)
{
  const int nd = xg_h1_ndof(D, p), ndp = (nd + 1) & ~1, PS = 2 * ndp + 2;
  const DefDev df = xg_defdev(ctx, D);
  const size_t sm = (size_t)4 * (32 * PS + (df.order > 0 ? XG_DEFS : 0)) * sizeof(double);
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  if (df.order > 0) {
    if constexpr (D >= 2) {
      XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elvec_points<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_elvec_points<D, true><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, np_shared, p, nd, T,
                                                              accumulate, slots, vecs, df);
    } else return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elvec: deformation needs dim >= 2");
  } else {
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elvec_points<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_elvec_points<D, false><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, np_shared, p, nd, T,
                                                             accumulate, slots, vecs, df);
  }
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

extern "C" int xgpu_elvec_rule(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                               const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
                               const int64_t* offsets, const double* pts, const double* wts,
                               const xgpu_lterm* terms_host, int nterms, int accumulate, const int32_t* slots, double* vecs)
{
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || !offsets)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elvec_rule: bad argument (orders 1..4)");
  if (n == 0) return XGPU_OK;
  LTermsDev T;
  int rc = make_lterms(ctx, terms_host, nterms, dim, T);
  if (rc) return rc;
  if (dim == 3) return launch_vec<3>(ctx, n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, 0, fe_order, T, accumulate, slots, vecs);
  if (dim == 2) return launch_vec<2>(ctx, n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, 0, fe_order, T, accumulate, slots, vecs);
  return launch_vec<1>(ctx, n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, 0, fe_order, T, accumulate, slots, vecs);
}

extern "C" int xgpu_elvec_std(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                              const int32_t* el_vtx, const double* vtx_xyz, const double* phi, int order,
                              const xgpu_lterm* terms_host, int nterms, int accumulate, const int32_t* slots, double* vecs)
{
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elvec_std: bad argument (orders 1..4)");
  if (n == 0) return XGPU_OK;
  LTermsDev T;
  int rc = make_lterms(ctx, terms_host, nterms, dim, T);
  if (rc) return rc;
  if (order < 0) order = 2 * fe_order;  // symboliccutlfi.cpp:86-88
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order);
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elvec_std: no rule");
  if (dim == 3) return launch_vec<3>(ctx, n, elems, el_vtx, vtx_xyz, phi, nullptr, r->d_pts, r->d_wts, r->n, fe_order, T, accumulate, slots, vecs);
  if (dim == 2) return launch_vec<2>(ctx, n, elems, el_vtx, vtx_xyz, phi, nullptr, r->d_pts, r->d_wts, r->n, fe_order, T, accumulate, slots, vecs);
  return launch_vec<1>(ctx, n, elems, el_vtx, vtx_xyz, phi, nullptr, r->d_pts, r->d_wts, r->n, fe_order, T, accumulate, slots, vecs);
}
