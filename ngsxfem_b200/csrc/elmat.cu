// elmat.cu — C ABI of the batched element matrices; kernels and launch templates in elmat.cuh, the instantiations of the
// mass/stiffness and tensor-tile kernels in elmat_ms_*.cu / elmat_dmma_*.cu (separate translation units: build time)
#include "elmat.cuh"

int xg_launch_ms_1(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats, const Rule2* r2);
int xg_launch_ms_2(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats, const Rule2* r2);
int xg_launch_ms_3(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats, const Rule2* r2);
int xg_launch_dmma_2(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats);
int xg_launch_dmma_3(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats);

int xg_launch_lp(xgpu_ctx* ctx, int dim, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, const TermsDev& T, int accumulate, const int32_t* slots, double* mats, const Rule2* r2);

namespace {

template <int D>
int xg_launch_ms(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats, const Rule2* r2 = nullptr)
{
  // order <= 2 with per-element rules (cut elements): one lane per quadrature point, matrix in registers (elmat_lp.cu);
  // ctx->opt_elmat_lp = 0 keeps the staged half-warp kernel
  if (p <= 2 && offsets && ctx->opt_elmat_lp) return xg_launch_lp(ctx, D, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, T, accumulate, slots, mats, r2);
  if (D == 1) return xg_launch_ms_1(ctx, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats, r2);
  if (D == 2) return xg_launch_ms_2(ctx, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats, r2);
  return xg_launch_ms_3(ctx, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats, r2);
}

template <int D>
int xg_launch_dmma(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats)
{
  if (D == 2) return xg_launch_dmma_2(ctx, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
  return xg_launch_dmma_3(ctx, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
}

template <int D>
int launch_points(xgpu_ctx* ctx, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
                  const double* phi, const int64_t* offsets, const double* pts, const double* wts, int np_shared,
                  int p, const TermsDev& T, int accumulate, const int32_t* slots, double* mats, bool only_ms = false)
{
  // FP64 tensor tiles at order >= 3 (dense contraction); ctx->opt_dmma: -1 never, 0 policy, 1 whenever compiled (p >= 2)
  if constexpr (D >= 2) if (only_ms && ctx->opt_dmma >= 0 && (p >= 3 || (ctx->opt_dmma > 0 && p == 2)))
    return xg_launch_dmma<D>(ctx, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
  if (only_ms) return xg_launch_ms<D>(ctx, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
  const int nd = xg_h1_ndof(D, p);
  const int ndp = (nd + 1) & ~1;
  const int PS = ndp * (2 + D) + XG_MAXTERMS;
  int CH = 32;
  while (CH > 4 && (size_t)4 * CH * PS * sizeof(double) > 56 * 1024) CH >>= 1;
  const DefDev df = xg_defdev(ctx, D);
  const size_t sm = (size_t)4 * (CH * PS + (df.order > 0 ? XG_DEFS : 0)) * sizeof(double);
  const int nblocks = (ndp / 2) * (ndp / 2);
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 8) grid = XGPU_NSM * 8;
#define XG_LP2(MB, DF)                                                                                              \
  {                                                                                                                 \
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_points<D, MB, DF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_elmat_points<D, MB, DF><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, np_shared, \
                                                              p, nd, CH, T, accumulate, slots, mats, df);           \
  }
#define XG_LP(MB)                                                                                                   \
  {                                                                                                                 \
    if (df.order > 0) {                                                                                             \
      if constexpr (D >= 2) XG_LP2(MB, true)                                                                        \
      else return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elmat: deformation needs dim >= 2");                        \
    } else XG_LP2(MB, false)                                                                                        \
  }
  if (nblocks <= 32) XG_LP(1) else if (nblocks <= 128) XG_LP(4) else if (nblocks <= 352) XG_LP(11)
  else return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elmat: element too large");
#undef XG_LP
#undef XG_LP2
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

// reference tensors  M_ref, K_ref^{ab}+K_ref^{ba} (a<b), K_ref^{aa}  on the device, cached per (dim,p,order)
int get_reftensor(xgpu_ctx* ctx, int dim, int p, int order, const double** out)
{
  const int key = dim * 10000 + p * 100 + (order < 0 ? 0 : order);
  auto it = ctx->reftensors.find(key);
  if (it != ctx->reftensors.end()) { *out = (const double*)it->second; return XGPU_OK; }
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order);
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "elmat: no rule");
  const int nd = xg_h1_ndof(dim, p), nn = nd * nd, ng = dim * (dim + 1) / 2;
  std::vector<double> Tm((size_t)(1 + ng) * nn, 0.0), sh(nd), dsh(nd * dim);
  for (int q = 0; q < r->n; q++) {
    xgpu_h1_shapes_host(dim, p, nullptr, &r->pts[3 * q], sh.data(), dsh.data());
    const double w = r->wts[q];
    for (int i = 0; i < nd; i++)
      for (int j = 0; j < nd; j++) {
        Tm[i * nd + j] += w * sh[i] * sh[j];
        int m = 1;
        for (int a = 0; a < dim; a++)
          for (int b = a; b < dim; b++, m++) {
            double k = dsh[i * dim + a] * dsh[j * dim + b];
            if (a != b) k += dsh[i * dim + b] * dsh[j * dim + a];
            Tm[(size_t)m * nn + i * nd + j] += w * k;
          }
      }
  }
  void* d;
  XGPU_CUDA(ctx, cudaMalloc(&d, sizeof(double) * Tm.size()));
  XGPU_CUDA(ctx, cudaMemcpyAsync(d, Tm.data(), sizeof(double) * Tm.size(), cudaMemcpyHostToDevice, ctx->stream));
  XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->reftensors[key] = d;
  *out = (const double*)d;
  return XGPU_OK;
}

}  // namespace

bool xg_reftensor_is(xgpu_ctx* ctx, const double* tref, int dim, int p)
{
  for (const auto& kv : ctx->reftensors)
    if (kv.second == (const void*)tref && kv.first / 10000 == dim && (kv.first / 100) % 100 == p) return true;
  return false;
}

extern "C" int xgpu_elmat_rule(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                               const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
                               const int64_t* offsets, const double* pts, const double* wts,
                               const xgpu_term* terms_host, int nterms, int accumulate, const int32_t* slots, double* mats)
{
  XGPU_NVTX(ctx, "xgpu_elmat_rule");
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || !offsets)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_rule: bad argument (orders 1..4)");
  if (n == 0) return XGPU_OK;
  TermsDev T; bool ms;
  int rc = make_terms(ctx, terms_host, nterms, T, &ms);
  if (rc) return rc;
  if (dim == 3) return launch_points<3>(ctx, n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, 0, fe_order, T, accumulate, slots, mats, ms);
  if (dim == 2) return launch_points<2>(ctx, n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, 0, fe_order, T, accumulate, slots, mats, ms);
  return launch_points<1>(ctx, n, elems, el_vtx, vtx_xyz, phi, offsets, pts, wts, 0, fe_order, T, accumulate, slots, mats, ms);
}

namespace {
template <int D>
int launch_pair(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
                const int64_t* offsets, const double* pts, const double* wts, const TermsDev& T, const Rule2& R2,
                int accumulate, const int32_t* slots, double* mats)
{
  if (p >= 1 && p <= 4) return xg_launch_ms<D>(ctx, p, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, 0, T, accumulate, slots, mats, &R2);
  return xgpu_fail(ctx, XGPU_ERR_ARG, "elmat: order");
}
}  // namespace

extern "C" int xgpu_elmat_rule_pair(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                                    const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
                                    const int64_t* offsets_a, const double* pts_a, const double* wts_a,
                                    const xgpu_term* terms_a_host, int nterms_a,
                                    const int64_t* offsets_b, const double* pts_b, const double* wts_b,
                                    const xgpu_term* terms_b_host, int nterms_b,
                                    int accumulate, const int32_t* slots, double* mats)
{
  XGPU_NVTX(ctx, "xgpu_elmat_rule_pair");
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || !offsets_a || !offsets_b)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_rule_pair: bad argument (orders 1..4)");
  if (n == 0) return XGPU_OK;
  TermsDev Ta; Rule2 R2; bool msa, msb;
  int rc = make_terms(ctx, terms_a_host, nterms_a, Ta, &msa);
  if (rc) return rc;
  rc = make_terms(ctx, terms_b_host, nterms_b, R2.T, &msb);
  if (rc) return rc;
  const bool tensor = dim >= 2 && ctx->opt_dmma >= 0 && (fe_order >= 3 || (ctx->opt_dmma > 0 && fe_order == 2));
  if (!(msa && msb) || tensor) {  // not the mass/stiffness CUDA-core kernel: two passes, the second accumulating
    rc = xgpu_elmat_rule(ctx, dim, fe_order, n, elems, el_vtx, vtx_xyz, phi, offsets_a, pts_a, wts_a, terms_a_host, nterms_a,
                         accumulate, slots, mats);
    if (rc) return rc;
    return xgpu_elmat_rule(ctx, dim, fe_order, n, elems, el_vtx, vtx_xyz, phi, offsets_b, pts_b, wts_b, terms_b_host, nterms_b,
                           1, slots, mats);
  }
  R2.offsets = offsets_b; R2.pts = pts_b; R2.wts = wts_b;
  if (dim == 3) return launch_pair<3>(ctx, fe_order, n, elems, el_vtx, vtx_xyz, offsets_a, pts_a, wts_a, Ta, R2, accumulate, slots, mats);
  if (dim == 2) return launch_pair<2>(ctx, fe_order, n, elems, el_vtx, vtx_xyz, offsets_a, pts_a, wts_a, Ta, R2, accumulate, slots, mats);
  return launch_pair<1>(ctx, fe_order, n, elems, el_vtx, vtx_xyz, offsets_a, pts_a, wts_a, Ta, R2, accumulate, slots, mats);
}

namespace {
// the geometry half of k_elmat_affine: same operations in the same order, so that  sum_m coefs[m] * tref[m]  evaluated
// elsewhere (csr_fused.cu) reproduces its matrices bit for bit
template <int D>
__global__ void __launch_bounds__(256) k_affine_coefs(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                                      const double* __restrict__ vtx_xyz, TermsDev T, double* __restrict__ coefs)
{
  constexpr int NG = D * (D + 1) / 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int64_t el = elems ? (int64_t)elems[e] : e;
  double vx[D + 1][3], F[D * D], Fi[D * D], coef[8];
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
    const int v = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
    for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v * D + c) : 0.0;
  }
  xg_affine_jac<D>(vx, F);
  const double meas = fabs(xg_inv<D>(F, Fi));
  const double h = (D == 1) ? meas : ((D == 2) ? sqrt(meas) : cbrt(meas));
  double cm = 0.0, cs = 0.0;
  for (int t = 0; t < T.n; t++) {
    const double c = T.coef[t] * ipow(h, T.hpow[t]) * meas;
    if (T.op_test[t] == 0) cm += c; else cs += c;
  }
  coef[0] = cm;
  int m = 1;
#pragma unroll
  for (int a = 0; a < D; a++)
#pragma unroll
    for (int b = a; b < D; b++) {
      double gab = 0.0;
#pragma unroll
      for (int c = 0; c < D; c++) gab += Fi[a * D + c] * Fi[b * D + c];
      coef[m++] = cs * gab;
    }
#pragma unroll
  for (int k = 1 + NG; k < 8; k++) coef[k] = 0.0;
  double2* out = reinterpret_cast<double2*>(coefs + e * 8);
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = make_double2(coef[2 * k], coef[2 * k + 1]);
}
}  // namespace

extern "C" int xgpu_affine_coefs(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                 const double* vtx_xyz, const xgpu_term* terms_host, int nterms, double* coefs)
{
  XGPU_NVTX(ctx, "xgpu_affine_coefs");
  if (!ctx || dim < 1 || dim > 3 || n < 0 || (n > 0 && !coefs)) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_affine_coefs: bad argument");
  if (ctx->def_order != 0) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_affine_coefs: the mesh is deformed (elements are not affine)");
  if (n == 0) return XGPU_OK;
  TermsDev T; bool ms;
  int rc = make_terms(ctx, terms_host, nterms, T, &ms);
  if (rc) return rc;
  if (!ms) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_affine_coefs: ID*ID and GRAD.GRAD terms only");
  const unsigned grid = nblk(n, 256);
  if (dim == 3) k_affine_coefs<3><<<grid, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, T, coefs);
  else if (dim == 2) k_affine_coefs<2><<<grid, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, T, coefs);
  else k_affine_coefs<1><<<grid, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, T, coefs);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_reftensor(xgpu_ctx* ctx, int dim, int fe_order, int order, const double** tref_dev_out, int* ncoef_host)
{
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > 2 || !tref_dev_out)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_reftensor: bad argument (orientation-independent bases: orders 1, 2)");
  if (ncoef_host) *ncoef_host = 1 + dim * (dim + 1) / 2;
  return get_reftensor(ctx, dim, fe_order, order, tref_dev_out);
}

extern "C" int xgpu_elmat_std(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                              const int32_t* el_vtx, const double* vtx_xyz, const double* phi, int order,
                              const xgpu_term* terms_host, int nterms, int accumulate, const int32_t* slots, double* mats)
{
  XGPU_NVTX(ctx, "xgpu_elmat_std");
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_std: bad argument (orders 1..4)");
  if (n == 0) return XGPU_OK;
  TermsDev T; bool ms;
  int rc = make_terms(ctx, terms_host, nterms, T, &ms);
  if (rc) return rc;
  if (order < 0) {  // symboliccutbfi.cpp:100-118
    int dtr = 99, dte = 99;
    for (int t = 0; t < nterms; t++) {
      dtr = std::min(dtr, terms_host[t].op_trial == 0 ? 0 : 1);
      dte = std::min(dte, terms_host[t].op_test == 0 ? 0 : 1);
    }
    if (nterms == 0) dtr = dte = 0;
    order = 2 * fe_order - (dtr + dte);
  }
  // fast path: affine elements (no deformation), orientation independent bases (P1, P2), mass / stiffness terms only
  if (ms && fe_order <= 2 && ctx->def_order == 0) {
    const double* Tref;
    rc = get_reftensor(ctx, dim, fe_order, order, &Tref);
    if (rc) return rc;
    const int nd = xg_h1_ndof(dim, fe_order);
    unsigned grid = nblk((n + 31) / 32, 8);
    if (grid > XGPU_NSM * 8) grid = XGPU_NSM * 8;
    if (dim == 3) k_elmat_affine<3><<<grid, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, nd, Tref, T, accumulate, slots, mats);
    else if (dim == 2) k_elmat_affine<2><<<grid, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, nd, Tref, T, accumulate, slots, mats);
    else k_elmat_affine<1><<<grid, 256, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, nd, Tref, T, accumulate, slots, mats);
    XGPU_LAUNCH_CHECK(ctx);
    return XGPU_OK;
  }
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order);
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_std: no rule");
  if (dim == 3) return launch_points<3>(ctx, n, elems, el_vtx, vtx_xyz, phi, nullptr, r->d_pts, r->d_wts, r->n, fe_order, T, accumulate, slots, mats, ms);
  if (dim == 2) return launch_points<2>(ctx, n, elems, el_vtx, vtx_xyz, phi, nullptr, r->d_pts, r->d_wts, r->n, fe_order, T, accumulate, slots, mats, ms);
  return launch_points<1>(ctx, n, elems, el_vtx, vtx_xyz, phi, nullptr, r->d_pts, r->d_wts, r->n, fe_order, T, accumulate, slots, mats, ms);
}
