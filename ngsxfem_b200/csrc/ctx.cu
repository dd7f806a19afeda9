// ctx.cu — context life cycle, status, default quadrature tables (host), host basis evaluation.
#include "common.cuh"
#include "h1.cuh"
#include <math.h>
#include <string.h>
#include <limits.h>

#include "rule_tables.inc"

namespace {

// ---- Gauss-Jacobi (alpha, 0) on [0,1] with weight (1-x)^alpha: deflated Newton in long double
void jacobi_eval(int n, long double a, long double x, long double* P, long double* dP)
{
  long double p0 = 1.0L, p1 = 0.5L * (a + (a + 2.0L) * x);
  long double d0 = 0.0L, d1 = 0.5L * (a + 2.0L);
  if (n == 0) { *P = p0; *dP = d0; return; }
  for (int k = 2; k <= n; k++) {
    long double kk = k;
    long double c1 = 2.0L * kk * (kk + a) * (2.0L * kk + a - 2.0L);
    long double c2 = (2.0L * kk + a - 1.0L) * a * a;
    long double c3 = (2.0L * kk + a - 1.0L) * (2.0L * kk + a) * (2.0L * kk + a - 2.0L);
    long double c4 = 2.0L * (kk + a - 1.0L) * (kk - 1.0L) * (2.0L * kk + a);
    long double p2 = ((c2 + c3 * x) * p1 - c4 * p0) / c1;
    long double d2 = ((c2 + c3 * x) * d1 + c3 * p1 - c4 * d0) / c1;
    p0 = p1; p1 = p2; d0 = d1; d1 = d2;
  }
  *P = p1; *dP = d1;
}

void gauss_jacobi01(int n, int alpha, std::vector<double>& x, std::vector<double>& w)
{
  const long double a = alpha;
  const long double PI = 3.14159265358979323846264338327950288L;
  std::vector<long double> r(n);
  for (int i = 0; i < n; i++) {
    long double xi = -cosl((2.0L * i + 1.0L + 0.5L * a) * PI / (2.0L * n + a + 1.0L));
    for (int it = 0; it < 100; it++) {
      long double P, dP, s = 0.0L;
      jacobi_eval(n, a, xi, &P, &dP);
      for (int j = 0; j < i; j++) s += 1.0L / (xi - r[j]);
      long double dx = P / (dP - s * P);
      xi -= dx;
      if (fabsl(dx) < 1e-19L) break;
    }
    r[i] = xi;
  }
  // sort ascending (deflation finds them in order for these guesses, but be safe)
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (r[j] < r[i]) { long double t = r[i]; r[i] = r[j]; r[j] = t; }
  x.resize(n); w.resize(n);
  for (int i = 0; i < n; i++) {
    long double P, dP;
    jacobi_eval(n, a, r[i], &P, &dP);
    x[i] = (double)(0.5L * (1.0L + r[i]));
    w[i] = (double)(1.0L / ((1.0L - r[i] * r[i]) * dP * dP));
  }
}

void sym_rule(xgpu_rule_host& R, const double (*P)[3], const double* W, int n)
{
  R.n = n; R.pts.resize(3 * n); R.wts.resize(n);
  for (int i = 0; i < n; i++) { for (int k = 0; k < 3; k++) R.pts[3 * i + k] = P[i][k]; R.wts[i] = W[i]; }
}

// NGSolve's SelectIntegrationRule construction (third party, restated; see include/xgpu.h)
int default_rule(int dim, int order, xgpu_rule_host& R)
{
  if (order < 0) order = 0;
  const int n = order / 2 + 1;
  if (dim == 0) { R.n = 1; R.pts = {0, 0, 0}; R.wts = {1.0}; return 0; }
  std::vector<double> x0, w0, x1, w1, x2, w2;
  if (dim == 1) {
    gauss_jacobi01(n, 0, x0, w0);
    R.n = n; R.pts.assign(3 * n, 0.0); R.wts = w0;
    for (int i = 0; i < n; i++) R.pts[3 * i] = x0[i];
    return 0;
  }
  if (dim == 2) {
    if (order <= 1) { sym_rule(R, SYM_TRIG1_P, SYM_TRIG1_W, SYM_TRIG1_N); return 0; }
    if (order == 2) { sym_rule(R, SYM_TRIG3_P, SYM_TRIG3_W, SYM_TRIG3_N); return 0; }
    if (order <= 4) { sym_rule(R, SYM_TRIG6_P, SYM_TRIG6_W, SYM_TRIG6_N); return 0; }
    if (order <= 6) { sym_rule(R, SYM_TRIG12_P, SYM_TRIG12_W, SYM_TRIG12_N); return 0; }
    gauss_jacobi01(n, 1, x0, w0); gauss_jacobi01(n, 0, x1, w1);
    R.n = n * n; R.pts.assign(3 * R.n, 0.0); R.wts.resize(R.n);
    int k = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++, k++) {
      R.pts[3 * k] = x0[i]; R.pts[3 * k + 1] = x1[j] * (1.0 - x0[i]); R.wts[k] = w0[i] * w1[j];
    }
    return 0;
  }
  if (dim == 3) {
    if (order <= 1) { sym_rule(R, SYM_TET1_P, SYM_TET1_W, SYM_TET1_N); return 0; }
    if (order == 2) { sym_rule(R, SYM_TET4_P, SYM_TET4_W, SYM_TET4_N); return 0; }
    if (order <= 5) { sym_rule(R, SYM_TET14_P, SYM_TET14_W, SYM_TET14_N); return 0; }
    gauss_jacobi01(n, 2, x0, w0); gauss_jacobi01(n, 1, x1, w1); gauss_jacobi01(n, 0, x2, w2);
    R.n = n * n * n; R.pts.assign(3 * R.n, 0.0); R.wts.resize(R.n);
    int k = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int l = 0; l < n; l++, k++) {
      R.pts[3 * k] = x0[i];
      R.pts[3 * k + 1] = x1[j] * (1.0 - x0[i]);
      R.pts[3 * k + 2] = x2[l] * (1.0 - x0[i]) * (1.0 - x1[j]);
      R.wts[k] = w0[i] * w1[j] * w2[l];
    }
    return 0;
  }
  return -1;
}

int upload_rule(xgpu_ctx* ctx, xgpu_rule_host& R)
{
  if (R.d_pts) { cudaFree(R.d_pts); R.d_pts = nullptr; }
  if (R.d_wts) { cudaFree(R.d_wts); R.d_wts = nullptr; }
  XGPU_CUDA(ctx, cudaMalloc(&R.d_pts, sizeof(double) * 3 * R.n));
  XGPU_CUDA(ctx, cudaMalloc(&R.d_wts, sizeof(double) * R.n));
  XGPU_CUDA(ctx, cudaMemcpy(R.d_pts, R.pts.data(), sizeof(double) * 3 * R.n, cudaMemcpyHostToDevice));
  XGPU_CUDA(ctx, cudaMemcpy(R.d_wts, R.wts.data(), sizeof(double) * R.n, cudaMemcpyHostToDevice));
  return XGPU_OK;
}

struct HostSink {
  double* sh; double* dsh; int D;
  template <int DD> void operator()(int i, const AD<DD>& a) {
    if (sh) sh[i] = a.v;
    if (dsh) for (int k = 0; k < DD; k++) dsh[i * DD + k] = a.d[k];
  }
};

}  // namespace

const xgpu_rule_host* xgpu_rule_lookup(xgpu_ctx* ctx, int dim, int order)
{
  if (order < 0) order = 0;
  auto key = std::make_pair(dim, order);
  auto it = ctx->rules.find(key);
  if (it != ctx->rules.end()) return &it->second;
  xgpu_rule_host R;
  if (default_rule(dim, order, R)) return nullptr;
  auto& slot = ctx->rules[key];
  slot = R;
  if (upload_rule(ctx, slot) != XGPU_OK) { ctx->rules.erase(key); return nullptr; }
  return &slot;
}

extern "C" const char* xgpu_version(void) { return "xgpu 0.1 (sm_100a)"; }

extern "C" int xgpu_create(int device, xgpu_ctx** out)
{
  if (!out) return XGPU_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return XGPU_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return XGPU_ERR_CUDA;
  xgpu_ctx* ctx = new xgpu_ctx();
  ctx->device = device;
  {  // keep freed blocks of the stream-ordered pool cached (see xgpu_scratch)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  if (cudaMalloc(&ctx->d_status, 2 * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&ctx->d_status_elem, sizeof(unsigned long long)) != cudaSuccess ||
      cudaHostAlloc(&ctx->h_pinned, 16 * sizeof(int64_t), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    delete ctx;
    return XGPU_ERR_CUDA;
  }
  *out = ctx;
  return xgpu_clear_status(ctx);
}

extern "C" void xgpu_destroy(xgpu_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < XGPU_NSCRATCH; i++) if (ctx->scratch[i]) cudaFreeAsync(ctx->scratch[i], ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->rules) { if (kv.second.d_pts) cudaFree(kv.second.d_pts); if (kv.second.d_wts) cudaFree(kv.second.d_wts); }
  for (auto& kv : ctx->reftensors) if (kv.second) cudaFree(kv.second);
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
  if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
  if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
  if (ctx->d_status) cudaFree(ctx->d_status);
  if (ctx->d_status_elem) cudaFree(ctx->d_status_elem);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  delete ctx;
}

extern "C" int xgpu_set_stream(xgpu_ctx* ctx, void* s)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (ctx->stream != (cudaStream_t)s) {
    // scratch / plan buffers are ordered on the old stream: drain it before switching
    XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->stream = (cudaStream_t)s;
  return XGPU_OK;
}

extern "C" int xgpu_set_option(xgpu_ctx* ctx, const char* key, int value)
{
  if (!ctx || !key) return XGPU_ERR_ARG;
  if (std::string(key) == "elmat_dmma") { ctx->opt_dmma = value < 0 ? -1 : (value > 0 ? 1 : 0); return XGPU_OK; }
  if (std::string(key) == "csr_small") { ctx->opt_csr_small = value != 0; return XGPU_OK; }
  if (std::string(key) == "elmat_lp") { ctx->opt_elmat_lp = value != 0; return XGPU_OK; }
  if (std::string(key) == "cut_blocks_per_sm") { ctx->opt_cut_bpsm = value <= 1 ? 1 : 2; return XGPU_OK; }
  if (std::string(key) == "short_blocks") { ctx->opt_short_blocks = value != 0; return XGPU_OK; }
  if (std::string(key) == "csr_warp") { ctx->opt_csr_warp = value != 0; return XGPU_OK; }
  if (std::string(key) == "csr_warp_bpsm") { ctx->opt_csr_warp_bpsm = value; return XGPU_OK; }
  if (std::string(key) == "csr_warp_half") { ctx->opt_csr_warp_half = value != 0; return XGPU_OK; }
  if (std::string(key) == "csr_closed_form") { ctx->opt_csr_closed_form = value != 0; return XGPU_OK; }
  if (std::string(key) == "csr_fork") { ctx->opt_csr_fork = value != 0; return XGPU_OK; }
  if (std::string(key) == "csr_trips") { ctx->opt_csr_trips = value == 2 ? 2 : 0; return XGPU_OK; }
  if (std::string(key) == "elmat_st_tp") { ctx->opt_st_tp = value != 0; return XGPU_OK; }
  return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_set_option: unknown key");
}

extern "C" int xgpu_set_deformation(xgpu_ctx* ctx, int order, const int32_t* el_dofs, const double* vals)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (order <= 0 || !el_dofs || !vals) {
    if (order > 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_set_deformation: el_dofs / vals missing");
    ctx->def_order = 0; ctx->def_el_dofs = nullptr; ctx->def_vals = nullptr;
    return XGPU_OK;
  }
  if (order > XG_MAX_ORDER) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_set_deformation: orders 1..4");
  ctx->def_order = order; ctx->def_el_dofs = el_dofs; ctx->def_vals = vals;
  return XGPU_OK;
}

extern "C" int xgpu_sync(xgpu_ctx* ctx)
{
  if (!ctx) return XGPU_ERR_ARG;
  XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XGPU_OK;
}

extern "C" const char* xgpu_last_error(xgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int64_t xgpu_launch_count(xgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int xgpu_clear_status(xgpu_ctx* ctx)
{
  if (!ctx) return XGPU_ERR_ARG;
  XGPU_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, 2 * sizeof(int), ctx->stream));
  XGPU_CUDA(ctx, cudaMemsetAsync(ctx->d_status_elem, 0xff, sizeof(unsigned long long), ctx->stream));
  return XGPU_OK;
}

extern "C" int xgpu_get_status(xgpu_ctx* ctx, int* code_host, int64_t* elem_host)
{
  if (!ctx) return XGPU_ERR_ARG;
  XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, ctx->d_status, sizeof(int), ctx->stream));
  XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned + 1, ctx->d_status_elem, sizeof(unsigned long long), ctx->stream));
  XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int code = *(int*)ctx->h_pinned;
  const unsigned long long word = ctx->h_pinned[1];   // (element << 8 | -code) of the smallest failing element
  if (code && word != ~0ull) code = -(int)(word & 0xff);
  if (code_host) *code_host = code;
  if (elem_host) *elem_host = (code && word != ~0ull) ? (int64_t)(word >> 8) : -1;
  return XGPU_OK;
}

extern "C" int xgpu_get_stats(xgpu_ctx* ctx, xgpu_stats* out)
{
  if (!ctx || !out) return XGPU_ERR_ARG;
  int code = 0;
  int64_t elem = -1;
  int rc = xgpu_get_status(ctx, &code, &elem);
  if (rc) return rc;
  out->kernel_launches = ctx->launches;
  out->api_calls = ctx->api_calls;
  out->plans_created = ctx->plans_created;
  size_t sb = 0;
  for (int i = 0; i < XGPU_NSCRATCH; i++) sb += ctx->scratch_bytes[i];
  out->scratch_bytes = (int64_t)sb;
  out->status_code = code;
  out->status_element = elem;
  return XGPU_OK;
}

extern "C" int xgpu_rule_size(xgpu_ctx* ctx, int dim, int order)
{
  if (!ctx) return XGPU_ERR_ARG;
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order);
  return r ? r->n : XGPU_ERR_ARG;
}

extern "C" int xgpu_get_rule(xgpu_ctx* ctx, int dim, int order, double* pts_host, double* wts_host)
{
  if (!ctx) return XGPU_ERR_ARG;
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order);
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_get_rule: unsupported element dimension");
  memcpy(pts_host, r->pts.data(), sizeof(double) * 3 * r->n);
  memcpy(wts_host, r->wts.data(), sizeof(double) * r->n);
  return r->n;
}

extern "C" int xgpu_set_rule(xgpu_ctx* ctx, int dim, int order, int n, const double* pts_host, const double* wts_host)
{
  if (!ctx || n <= 0 || dim < 0 || dim > 3) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_set_rule: bad argument");
  if (order < 0) order = 0;
  xgpu_rule_host& R = ctx->rules[std::make_pair(dim, order)];
  R.n = n;
  R.pts.assign(pts_host, pts_host + 3 * n);
  R.wts.assign(wts_host, wts_host + n);
  // cached reference tensors depend on the rules
  XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& kv : ctx->reftensors) if (kv.second) cudaFree(kv.second);
  ctx->reftensors.clear();
  return upload_rule(ctx, R);
}

extern "C" int xgpu_h1_ndof(int dim, int order) { return xg_h1_ndof(dim, order); }

extern "C" int xgpu_h1_shapes_host(int dim, int order, const int32_t* vnum, const double* xhat, double* shape, double* dshape)
{
  if (order < 1 || order > XG_MAX_ORDER || dim < 1 || dim > 3) return XGPU_ERR_UNSUPPORTED;
  HostSink sink{shape, dshape, dim};
  int vn[4] = {0, 1, 2, 3};
  for (int i = 0; i <= dim; i++) vn[i] = vnum ? vnum[i] : i;
  if (dim == 1) return xg_h1_shapes_rt<1>(order, vn, xhat, sink);
  if (dim == 2) return xg_h1_shapes_rt<2>(order, vn, xhat, sink);
  return xg_h1_shapes_rt<3>(order, vn, xhat, sink);
}
