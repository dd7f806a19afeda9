// elmat_st.cu — space-time element matrices / vectors on space-time cut rules (config 5 of BASELINE.json):
// SymbolicCutBilinearFormIntegrator::T_CalcElementMatrixAdd (xfem/symboliccutbfi.cpp:76-486) called with a
// SpaceTimeFE (spacetime/SpaceTimeFE.cpp:57-135): tensor-product basis  shape[j*nd_s + i] = L_j(t) N_i(xhat)
// (time-major dofs, SpaceTimeFE.hpp:185-218), L_j = Lagrange polynomials on the time nodes of NodalTimeFE
// (SpaceTimeFE.cpp:258-282), grad = L_j grad N_i (space only), dt = L_j' N_i (REFERENCE time derivative, DiffOpDt,
// spacetime/diffopDt.hpp).  The rule comes from xgpu_cutquad_st_* (spacetimecutrule.cpp:104-244): reference point,
// reference time (the reference keeps it in the IP's weight slot, :192-196) and weight w_space*w_time*(t1-t0); the
// measure is the SPATIAL |det F|.
//
//   k_elmat_st<D,MAXB>  one warp per element: lanes stage [val | grad_0..grad_{D-1} | dt] of all nd_s*(k_t+1)
//        functions at one point each in shared memory, then own 2x2 register blocks of the matrix and contract over
//        the staged points in point order (fixed summation order => bit-reproducible).
//   k_elvec_st<D>       the same staging, lane per dof sums  f(x_q, t_q) w_q |det F| op(v)  over the points.
// (both in elmat_st_impl.cuh, instantiated per dimension in elmat_st_d1/2/3.cu; the tensor-structured kernel for the common
// orders is elmat_st_tp.cu.)  This file: the C ABI, argument checks, dispatch, and the space-time facet-patch expansion.
#include "common.cuh"
#include "cutgeom.cuh"
#include "deform.cuh"
#include "h1.cuh"
#include "elmat_st.cuh"


#define XG_ST_MAXT 4  // time orders 0..3

#define XG_ST_DECL(DD)                                                                                                 \
  int xg_launch_elmat_st_d##DD(xgpu_ctx* ctx, int p, int kt, int64_t n, const int32_t* elems, const int32_t* el_vtx,    \
                               const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,   \
                               const double* wts, const StTermsDev& T, int accumulate, const int32_t* slots, double* mats); \
  int xg_launch_elvec_st_d##DD(xgpu_ctx* ctx, int p, int kt, int64_t n, const int32_t* elems, const int32_t* el_vtx,    \
                               const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,   \
                               const double* wts, const StTermsDev& T, int accumulate, const int32_t* slots, double* vecs);
XG_ST_DECL(1) XG_ST_DECL(2) XG_ST_DECL(3)
#undef XG_ST_DECL

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// space-time facet patch = (time matrix) kron (spatial facet-patch matrix): out[(h,j,i),(g,l,k)] = T[j][l] in[(h,i),(g,k)],
// h, g = element of the patch, time-major dofs per element.  One thread per output entry, coalesced stores.
struct TimeMat { int nt; double t[XG_ST_MAXT * XG_ST_MAXT]; };
__global__ void __launch_bounds__(256) k_kron_time(int64_t n, int nds, TimeMat T, const double* __restrict__ in,
                                                   double* __restrict__ out)
{
  const int nt = T.nt, ws = 2 * nds, wo = ws * nt;
  const int64_t per = (int64_t)wo * wo;
  const int64_t total = n * per;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = idx / per;
    const int rc = (int)(idx - f * per);
    const int r = rc / wo, c = rc - r * wo;
    const int h = r / (nds * nt), rj = (r - h * nds * nt) / nds, ri = r - h * nds * nt - rj * nds;
    const int g = c / (nds * nt), cl = (c - g * nds * nt) / nds, ck = c - g * nds * nt - cl * nds;
    out[idx] = T.t[rj * nt + cl] * __ldg(in + f * ws * ws + (int64_t)(h * nds + ri) * ws + (g * nds + ck));
  }
}

int st_check(xgpu_ctx* ctx, int dim, int fe_order, int time_fe_order, int64_t n, const void* offsets, const void* tvals, const char* who)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || time_fe_order < 0 || time_fe_order >= XG_ST_MAXT ||
      n < 0 || (n > 0 && (!offsets || !tvals)))  // an empty batch carries no arrays
    return xgpu_fail(ctx, XGPU_ERR_ARG, who);
  if (ctx->def_order > 0)
    return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "space-time forms do not support a mesh deformation");
  return XGPU_OK;
}

}  // namespace

extern "C" int xgpu_elmat_st(xgpu_ctx* ctx, int dim, int fe_order, int time_fe_order, int64_t n, const int32_t* elems,
                             const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts,
                             const double* tvals, const double* wts, const xgpu_term* terms_host, int nterms,
                             int accumulate, const int32_t* slots, double* mats)
{
  XGPU_NVTX(ctx, "xgpu_elmat_st");
  int rc = st_check(ctx, dim, fe_order, time_fe_order, n, offsets, tvals, "xgpu_elmat_st: bad argument (space orders 1..4, time orders 0..3)");
  if (rc) return rc;
  if (nterms < 0 || nterms > XG_ST_MAXTERMS) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_st: too many terms (max 8)");
  if (n == 0) return XGPU_OK;
  StTermsDev T = {};
  T.n = nterms;
  for (int t = 0; t < nterms; t++) {
    const int ou = terms_host[t].op_trial, ot = terms_host[t].op_test;
    const bool ok = (ou == 1 && ot == 1) || ((ou == 0 || ou == XGPU_OP_DT) && (ot == 0 || ot == XGPU_OP_DT));
    if (!ok) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_elmat_st: unsupported operator pair (menu: ID/DT x ID/DT, GRAD.GRAD)");
    T.op_trial[t] = ou; T.op_test[t] = ot; T.hpow[t] = terms_host[t].hpow; T.coef[t] = terms_host[t].coef;
  }
  if (ctx->opt_st_tp) {  // tensor-structured kernel for the common orders (elmat_st_tp.cu)
    rc = xg_launch_elmat_st_tp(ctx, dim, fe_order, time_fe_order, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, mats);
    if (rc != 1) return rc;
  }
  if (dim == 3) return xg_launch_elmat_st_d3(ctx, fe_order, time_fe_order, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, mats);
  if (dim == 2) return xg_launch_elmat_st_d2(ctx, fe_order, time_fe_order, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, mats);
  return xg_launch_elmat_st_d1(ctx, fe_order, time_fe_order, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, mats);
}

extern "C" int xgpu_elvec_st(xgpu_ctx* ctx, int dim, int fe_order, int time_fe_order, int64_t n, const int32_t* elems,
                             const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts,
                             const double* tvals, const double* wts, const xgpu_stlterm* terms_host, int nterms,
                             int accumulate, const int32_t* slots, double* vecs)
{
  int rc = st_check(ctx, dim, fe_order, time_fe_order, n, offsets, tvals, "xgpu_elvec_st: bad argument (space orders 1..4, time orders 0..3)");
  if (rc) return rc;
  if (nterms < 0 || nterms > XG_ST_MAXTERMS) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elvec_st: too many terms (max 8)");
  if (n == 0) return XGPU_OK;
  StTermsDev T = {};
  T.n = nterms;
  for (int t = 0; t < nterms; t++) {
    const int ot = terms_host[t].op_test;
    if (ot != 0 && ot != XGPU_OP_DT) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_elvec_st: test operator must be ID or DT");
    T.op_test[t] = ot; T.op_trial[t] = 0; T.hpow[t] = terms_host[t].hpow; T.coef[t] = terms_host[t].coef;
    for (int c = 0; c < 4; c++) {
      if (terms_host[t].expo[c] < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elvec_st: negative exponent");
      T.expo[t][c] = terms_host[t].expo[c];
    }
  }
  if (dim == 3) return xg_launch_elvec_st_d3(ctx, fe_order, time_fe_order, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, vecs);
  if (dim == 2) return xg_launch_elvec_st_d2(ctx, fe_order, time_fe_order, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, vecs);
  return xg_launch_elvec_st_d1(ctx, fe_order, time_fe_order, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, vecs);
}

// Lagrange values on NodalTimeFE's nodes (host)
static void st_time_shapes_host(int order, double t, double* L)
{
  double nd[XG_ST_MAXT] = {0.0, 1.0, 0.0, 0.0};
  if (order == 2) { nd[1] = 0.5; nd[2] = 1.0; }
  if (order == 3) { nd[1] = 0.5 * (1.0 - 1.0 / sqrt(5.0)); nd[2] = 0.5 * (1.0 + 1.0 / sqrt(5.0)); nd[3] = 1.0; }
  for (int i = 0; i <= order; i++) {
    double s = 1.0;
    for (int j = 0; j <= order; j++) if (j != i) s *= (t - nd[j]) / (nd[i] - nd[j]);
    L[i] = s;
  }
}

extern "C" int xgpu_facetpatch_st(xgpu_ctx* ctx, int dim, int fe_order, int time_fe_order, int64_t n, const int32_t* facets,
                                  const int32_t* facet_els, const int32_t* el_vtx, const double* vtx_xyz, int order,
                                  int time_order, double tref, double coef, int hpow, double* mats)
{
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER || time_fe_order < 0 || time_fe_order >= XG_ST_MAXT)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_facetpatch_st: bad argument (space orders 1..4, time orders 0..3)");
  if (ctx->def_order > 0) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "space-time forms do not support a mesh deformation");
  if (n == 0) return XGPU_OK;
  const int nds = xg_h1_ndof(dim, fe_order), nt = time_fe_order + 1, ws = 2 * nds;
  TimeMat T;
  T.nt = nt;
  for (int i = 0; i < XG_ST_MAXT * XG_ST_MAXT; i++) T.t[i] = 0.0;
  double L[XG_ST_MAXT];
  if (time_order >= 0) {  // tensor-product rule in time (symboliccutbfi.cpp:2095-2160): SelectIntegrationRule(ET_SEGM, time_order)
    const xgpu_rule_host* rt = xgpu_rule_lookup(ctx, 1, time_order);
    if (!rt) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_facetpatch_st: no time rule");
    for (int q = 0; q < rt->n; q++) {
      st_time_shapes_host(time_fe_order, rt->pts[3 * q], L);
      for (int j = 0; j < nt; j++) for (int l = 0; l < nt; l++) T.t[j * nt + l] += rt->wts[q] * L[j] * L[l];
    }
  } else {                // fixed reference time (:2162-2177)
    st_time_shapes_host(time_fe_order, tref, L);
    for (int j = 0; j < nt; j++) for (int l = 0; l < nt; l++) T.t[j * nt + l] = L[j] * L[l];
  }
  void* p;
  int rc = xgpu_scratch(ctx, 9, sizeof(double) * (size_t)n * ws * ws, &p);
  if (rc) return rc;
  rc = xgpu_facetpatch(ctx, dim, fe_order, n, facets, facet_els, el_vtx, vtx_xyz, order, coef, hpow, (double*)p);
  if (rc) return rc;
  const int64_t total = n * (int64_t)(ws * nt) * (ws * nt);
  unsigned grid = nblk(total, 256);
  if (grid > XGPU_NSM * 32) grid = XGPU_NSM * 32;
  k_kron_time<<<grid, 256, 0, ctx->stream>>>(n, nds, T, (const double*)p, mats);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}
