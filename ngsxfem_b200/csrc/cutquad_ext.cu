// cutquad_ext.cu — multi-level-set and space-time cut rules on the GPU.
//
//   k_ml<D,FILL>  CreateMultiLevelsetCutIntegrationRule (cutint/mlsetintegration.cpp:6-186) +
//                 StraightCutsIntegrationRule (cutint/straightcutrule.cpp:679-812): one THREAD per
//                 element.  The successive decomposition (simplex list after level-set i is cut by
//                 level set i+1, ...) is walked depth first with an explicit stack whose children
//                 are pushed in reverse order — this emits the final sub-simplices in exactly the
//                 order of the reference's level-by-level lists, with a stack of <= 2*M+1 entries
//                 instead of 3^M simplices.  Weight fix-ups by the number of IF entries (:759-808)
//                 are constant per (element, tuple) on affine elements and applied at emission.
//   k_st<D,FILL>  SpaceTimeCutIntegrationRule (cutint/spacetimecutrule.cpp:8-244): per element the
//                 topology-change times (roots of every vertex' time polynomial), then for every time
//                 sub-interval and Gauss point the spatial straight-cut rule at that time.
//   FILL=false counts the points (-> scan -> CSR-of-points), FILL=true writes them.  The count
//   needs the full geometry here (clamped re-evaluation decides which sub-simplices survive).
// Cut elements are a few per cent of a mesh and every thread works alone, so these kernels are
// latency/occupancy bound by design; they are sized by the list of relevant elements.
#include "common.cuh"
#include "cutgeom.cuh"

#define XG_ML_MAXM 16
#define XG_ML_MAXT 64
#define XG_ST_MAXORD 8
#define XG_ST_MAXCP 40

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

struct MlTuples { int M, nt; signed char dts[XG_ML_MAXT * XG_ML_MAXM]; };
struct RuleSet { const double* p[4]; const double* w[4]; int n[4]; };  // reference rules by dimension 0..3

struct Emit {
  double* pts; double* wts; double* tv;
  long long base;
  int n;
};

// append the plain rule of sub-simplex S (GetPlainIntegrationRule :80-100) scaled by fac
template <bool FILL>
__device__ __forceinline__ void emit_simplex(const XgSimplex& S, const RuleSet& R, double fac, double t, Emit& E)
{
  const int np = R.n[S.D];
  if (FILL) {
    const double vol = xg_volume(S) * fac;
    for (int q = 0; q < np; q++) {
      double xi[3] = {R.p[S.D][3 * q], R.p[S.D][3 * q + 1], R.p[S.D][3 * q + 2]}, x[3];
      xg_map_point(S, xi, x);
      const long long o = E.base + E.n + q;
      E.pts[3 * o] = x[0]; E.pts[3 * o + 1] = x[1]; E.pts[3 * o + 2] = x[2];
      E.wts[o] = R.w[S.D][q] * vol;
      if (E.tv) E.tv[o] = t;
    }
  }
  E.n += np;
}

// the shared standard rule of an uncut element (straightcutrule.cpp:669-674)
template <int D, bool FILL>
__device__ __forceinline__ void emit_std(const RuleSet& R, double fac, double t, Emit& E)
{
  const int np = R.n[D];
  if (FILL) {
    for (int q = 0; q < np; q++) {
      const long long o = E.base + E.n + q;
      E.pts[3 * o] = R.p[D][3 * q]; E.pts[3 * o + 1] = R.p[D][3 * q + 1]; E.pts[3 * o + 2] = R.p[D][3 * q + 2];
      E.wts[o] = R.w[D][q] * fac;
      if (E.tv) E.tv[o] = t;
    }
  }
  E.n += np;
}

template <int D>
__device__ __forceinline__ void load_geom(const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                          int64_t el, double* F, double* Fi, double* det)
{
  double vx[D + 1][3];
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
    const int v = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
    for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v * D + c) : 0.0;
  }
  xg_affine_jac<D>(vx, F);
  *det = xg_inv<D>(F, Fi);
}

// linear level set through its vertex values (LevelsetWrapper :503-526) evaluated at a reference point
template <int D>
__device__ __forceinline__ double lset_eval(const double* v, const double* p)
{
  double s = v[D];
#pragma unroll
  for (int i = 0; i < D; i++) s += (v[i] - v[D]) * p[i];
  return s;
}

struct StackEnt { XgSimplex S; int lev; };

template <int D, bool FILL>
__global__ void __launch_bounds__(128) k_ml(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                            const double* __restrict__ vtx_xyz, const double* __restrict__ phis, int64_t nv,
                                            MlTuples tp, RuleSet R, const int64_t* __restrict__ offsets,
                                            int32_t* __restrict__ counts, double* __restrict__ pts, double* __restrict__ wts,
                                            int* status, unsigned long long* status_elem)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t el = elems ? (int64_t)elems[i] : i;
  const int M = tp.M;
  int v[D + 1];
#pragma unroll
  for (int k = 0; k < D + 1; k++) v[k] = __ldg(el_vtx + el * (D + 1) + k);
  double lv[XG_ML_MAXM][D + 1];
  int ldt[XG_ML_MAXM];
  for (int j = 0; j < M; j++) {
#pragma unroll
    for (int k = 0; k < D + 1; k++) lv[j][k] = __ldg(phis + (int64_t)j * nv + v[k]);
    ldt[j] = xg_check_cut(lv[j], D + 1);
  }
  double F[D * D], Fi[D * D], det = 1.0;
  if (FILL) load_geom<D>(el_vtx, vtx_xyz, el, F, Fi, &det);
  Emit E{pts, wts, nullptr, FILL ? offsets[i] : 0, 0};
  int err = 0;
  StackEnt st[2 * XG_ML_MAXM + 2];
  for (int t = 0; t < tp.nt && !err; t++) {
    const signed char* dts = tp.dts + t * M;
    bool compatible = true, cut_element = false;
    for (int j = 0; j < M; j++) {
      if (ldt[j] != XGPU_IF && ldt[j] != dts[j]) compatible = false;
      else if (ldt[j] == XGPU_IF) cut_element = true;
    }
    if (!compatible) continue;
    if (!cut_element) { emit_std<D, FILL>(R, 1.0, 0.0, E); continue; }
    int cid[XG_ML_MAXM], nc = 0, nif = 0, ifi[3] = {0, 0, 0};
    for (int j = 0; j < M; j++)
      if (ldt[j] == XGPU_IF || ldt[j] != dts[j]) {
        if (dts[j] == XGPU_IF) { if (nif < 3) ifi[nif] = j; nif++; }
        cid[nc++] = j;
      }
    // weight fix-up factor (straightcutrule.cpp:759-808), constant on an affine element
    double fac = 1.0;
    // the reference's "not implemented" codimension cases only throw when points were produced
    const bool bad_codim = (nif > 3) || (nif == 3 && D != 3) || (nif == 2 && D == 1);
    const int n_before = E.n;
    if (FILL && !bad_codim) {
      if (nif == 1) fac = xg_if_weight_factor<D>(Fi, lv[ifi[0]]);
      else if (nif == 2 && D == 2) fac = 1.0 / fabs(det);
      else if (nif == 2 && D == 3) {
        double n0[3], n1[3], cr[3], a0 = 0, a1 = 0;
        for (int c = 0; c < 3; c++) { n0[c] = lv[ifi[0]][c] - lv[ifi[0]][3]; n1[c] = lv[ifi[1]][c] - lv[ifi[1]][3]; a0 += n0[c] * n0[c]; a1 += n1[c] * n1[c]; }
        a0 = sqrt(a0); a1 = sqrt(a1);
        for (int c = 0; c < 3; c++) { n0[c] /= a0; n1[c] /= a1; }
        cr[0] = n0[1] * n1[2] - n0[2] * n1[1]; cr[1] = n0[2] * n1[0] - n0[0] * n1[2]; cr[2] = n0[0] * n1[1] - n0[1] * n1[0];
        const double cpf = 1.0 / sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
        double s = 0.0;
        for (int r = 0; r < 3; r++) {
          double tr = 0.0;
          for (int c = 0; c < 3; c++) tr += cpf * F[r * 3 + c] * cr[c];
          s += tr * tr;
        }
        fac = sqrt(s) / fabs(det);
      } else if (nif == 3) fac = 1.0 / fabs(det);
    }
    // depth-first successive decomposition
    int sp = 0;
    xg_ref_simplex(D, st[0].S);
    st[0].lev = 0;
    sp = 1;
    while (sp > 0 && !err) {
      sp--;
      const int lev = st[sp].lev;
      XgSimplex S = st[sp].S;
      if (lev == nc) { emit_simplex<FILL>(S, R, fac, 0.0, E); continue; }
      const int j = cid[lev];
      const int want = dts[j];
      double val[4];
      const int nps = S.D + 1;
      if (lev == 0) { for (int k = 0; k < nps; k++) val[k] = lv[j][k]; }
      else {
        for (int k = 0; k < nps; k++) {  // update_initial_coefs (:548-556): |d| <= 1e-14 -> +1e-14
          const double d = lset_eval<D>(lv[j], S.p[k]);
          val[k] = (fabs(d) > 1e-14) ? d : 1e-14;
        }
      }
      const int ds = xg_check_cut(val, nps);
      if (ds == XGPU_IF || lev == 0) {
        XgSimplex sub[3];
        const int k = xg_decompose(S, val, want, sub);
        if (k < 0) { err = k; break; }
        for (int c = k - 1; c >= 0; c--) { st[sp].S = sub[c]; st[sp].lev = lev + 1; sp++; }
      } else if (ds == want) {
        st[sp].S = S; st[sp].lev = lev + 1; sp++;
      }
    }
    if (!err && bad_codim && E.n > n_before) err = XGPU_ERR_GEOM_NOTIMPL;
  }
  if (err) {
    // the count pass reports 0 points for a failing element, so the fill pass has nothing to write
    xgpu_record_error(status, status_elem, err, el);
    if (!FILL) counts[i] = 0;
    return;
  }
  if (!FILL) counts[i] = E.n;
}

// ---------------------------------------------------------------- space-time
__device__ __forceinline__ void time_nodes(int order, double* nd)
{
  if (order == 0) { nd[0] = 0.0; return; }
  if (order == 1) { nd[0] = 0.0; nd[1] = 1.0; return; }
  if (order == 2) { nd[0] = 0.0; nd[1] = 0.5; nd[2] = 1.0; return; }
  if (order == 3) { nd[0] = 0.0; nd[1] = 0.5 * (1.0 - 1.0 / sqrt(5.0)); nd[2] = 0.5 * (1.0 + 1.0 / sqrt(5.0)); nd[3] = 1.0; return; }
  if (order == 4) {  // SpaceTimeFE.cpp:271-272 (Gauss-Lobatto)
    nd[0] = 0.0; nd[1] = 0.5 * (1.0 - sqrt(3.0 / 7.0)); nd[2] = 0.5; nd[3] = 0.5 * (1.0 + sqrt(3.0 / 7.0)); nd[4] = 1.0; return;
  }
  if (order == 5) {  // SpaceTimeFE.cpp:273-278
    const double a = sqrt(1.0 / 3.0 + 2.0 * sqrt(7.0) / 21.0), b = sqrt(1.0 / 3.0 - 2.0 * sqrt(7.0) / 21.0);
    nd[0] = 0.0; nd[1] = 0.5 * (1.0 - a); nd[2] = 0.5 * (1.0 - b); nd[3] = 0.5 * (1.0 + b); nd[4] = 0.5 * (1.0 + a); nd[5] = 1.0; return;
  }
  // order >= 6: Gauss-Lobatto-Legendre points, Newton on x P_N - P_{N-1} from the Chebyshev-Lobatto start
  // (CalcLobattoPointsRec, SpaceTimeFE.cpp:229-256); rare orders, every thread recomputes them
  const int N = order;
  for (int j = 0; j <= N; j++) {
    double x = cos(3.14159265358979323846 * (double)j / N);
    for (int it = 0; it < 100; it++) {
      double pm = 1.0, p = x;
      for (int k = 1; k < N; k++) { const double pn = ((2.0 * k + 1.0) * (x * p) - (double)k * pm) / ((double)k + 1.0); pm = p; p = pn; }
      const double xn = x - (x * p - pm) / (N * p);
      const bool done = fabs(xn - x) <= 1e-20;
      x = xn;
      if (done) break;
    }
    nd[j] = 0.5 * (1.0 - x);
  }
  nd[0] = 0.0; nd[N] = 1.0;
}

__device__ __forceinline__ void time_shapes(int order, const double* nd, double t, double* sh)
{
  for (int i = 0; i <= order; i++) {
    double s = 1.0;
    for (int j = 0; j <= order; j++) if (j != i) s *= (t - nd[j]) / (nd[i] - nd[j]);
    sh[i] = s;
  }
}

__device__ __forceinline__ double eval_time(const double* li, int order, const double* nd, double t)
{
  double sh[XG_ST_MAXORD + 1], s = 0.0;
  time_shapes(order, nd, t, sh);
  for (int i = 0; i <= order; i++) s += li[i] * sh[i];
  return s;
}

// root_finding (spacetimecutrule.cpp:8-102); appends to roots, returns count or error
__device__ int root_finding(const double* li, int order, const double* nd, double* roots, int cap)
{
  int nr = 0;
  if (order == 0) return 0;
  if (order == 1) {
    if ((li[0] >= 0) != (li[1] >= 0)) roots[nr++] = -li[0] / (li[1] - li[0]);
    return nr;
  }
  if (order == 2) {
    const double c = li[0], a = 2 * li[0] + 2 * li[2] - 4. * li[1], b = li[2] - a - c;
    if (fabs(a) < 1e-12) {
      const double x = -c / b;
      if (x < 1 && x > 0) roots[nr++] = x;
    } else {
      const double w = b * b - 4 * a * c;
      if (fabs(w) < 1e-12) {
        const double x = -b / (2 * a);
        if (x < 1 && x > 0) roots[nr++] = x;
      } else {
        const double x1 = (b + sqrt(w)) / (-2.0 * a);
        const double x2 = (b - sqrt(w)) / (-2.0 * a);
        if (x1 < 1 && x1 > 0) roots[nr++] = x1;
        if (x2 < 1 && x2 > 0) roots[nr++] = x2;
      }
    }
    return nr;
  }
  const int subdivs = 50, bis_it = 70;
  const double EPS = 1e-15;
  const double dx = 1. / subdivs;
  double prev = 0.0;
  for (int i = 0; i < subdivs + 1; i++) {
    const double xi = dx * i;
    double vi = eval_time(li, order, nd, xi);
    if (vi == 0) vi = EPS;
    if (i >= 1 && prev * vi < 0) {
      double a = xi - dx, b = xi;
      double aval = eval_time(li, order, nd, a), bval = eval_time(li, order, nd, b);
      for (int j = 0; j < bis_it; j++) {
        const double xl = a - aval * (b - a) / (bval - aval);
        double val = eval_time(li, order, nd, xl);
        if (2 * fabs(val) < EPS) { a = b = xl; break; }
        const double xm = 0.5 * (a + b);
        val = eval_time(li, order, nd, xm);
        if (val == 0) break;
        if (val * aval < 0) { b = xm; bval = val; }
        else if (val * bval < 0) { a = xm; aval = val; }
        else return XGPU_ERR_GEOM_NOTIMPL;
      }
      if (nr >= cap) return XGPU_ERR_CAPACITY;
      roots[nr++] = 0.5 * (a + b);
    }
    prev = vi;
  }
  return nr;
}

template <int D, bool FILL>
__global__ void __launch_bounds__(128) k_st(int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
                                            const double* __restrict__ vtx_xyz, const double* __restrict__ phi_st, int64_t nv,
                                            int tord, int dt, RuleSet R, const double* __restrict__ tq_p,
                                            const double* __restrict__ tq_w, int ntq, const int64_t* __restrict__ offsets,
                                            int32_t* __restrict__ counts, double* __restrict__ pts, double* __restrict__ tvals,
                                            double* __restrict__ wts, int* status, unsigned long long* status_elem)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t el = elems ? (int64_t)elems[i] : i;
  const int nt = tord + 1;
  int v[D + 1];
#pragma unroll
  for (int k = 0; k < D + 1; k++) v[k] = __ldg(el_vtx + el * (D + 1) + k);
  double nd[XG_ST_MAXORD + 1];
  time_nodes(tord, nd);
  double ls[XG_ST_MAXORD + 1][D + 1];  // time-major nodal values of the element (:116-119)
  for (int j = 0; j < nt; j++)
#pragma unroll
    for (int k = 0; k < D + 1; k++) ls[j][k] = __ldg(phi_st + (int64_t)j * nv + v[k]);
  double cutp[XG_ST_MAXCP];
  int ncp = 2, err = 0;
  cutp[0] = 0.0; cutp[1] = 1.0;
  for (int k = 0; k < D + 1 && !err; k++) {
    double li[XG_ST_MAXORD + 1];
    for (int j = 0; j < nt; j++) li[j] = ls[j][k];
    const int nr = root_finding(li, tord, nd, cutp + ncp, XG_ST_MAXCP - ncp);
    if (nr < 0) { err = nr; break; }
    ncp += nr;
  }
  double F[D * D], Fi[D * D], det = 1.0;
  if (FILL) load_geom<D>(el_vtx, vtx_xyz, el, F, Fi, &det);
  Emit E{pts, wts, tvals, FILL ? offsets[i] : 0, 0};
  if (!err) {
    for (int a = 1; a < ncp; a++) {  // ascending sort of the cut times
      const double key = cutp[a];
      int b = a - 1;
      while (b >= 0 && cutp[b] > key) { cutp[b + 1] = cutp[b]; b--; }
      cutp[b + 1] = key;
    }
    for (int c = 0; c < ncp - 1 && !err; c++) {
      const double t0 = cutp[c], t1 = cutp[c + 1];
      for (int q = 0; q < ntq && !err; q++) {
        const double t = t0 + tq_p[3 * q] * (t1 - t0);
        double sh[XG_ST_MAXORD + 1], lt[D + 1];
        time_shapes(tord, nd, t, sh);
#pragma unroll
        for (int k = 0; k < D + 1; k++) {
          double d = 0.0;
          for (int j = 0; j < nt; j++) d += ls[j][k] * sh[j];
          if (fabs(d) < 1e-14) d = (d >= 0) ? 1e-14 : -1e-14;  // EPS_STCR_LSET_PERTUBATION (:167-175)
          lt[k] = d;
        }
        const double wfac = tq_w[q] * (t1 - t0);
        const int dom = xg_check_cut(lt, D + 1);
        if (dom == XGPU_IF) {
          XgSimplex ref, sub[3];
          xg_ref_simplex(D, ref);
          const int ns = xg_decompose(ref, lt, dt, sub);
          if (ns < 0) { err = ns; break; }
          double fac = wfac;
          if (FILL && dt == XGPU_IF) fac *= xg_if_weight_factor<D>(Fi, lt);
          for (int s = 0; s < ns; s++) emit_simplex<FILL>(sub[s], R, fac, t, E);
        } else if (dom == dt) {
          emit_std<D, FILL>(R, wfac, t, E);
        }
      }
    }
  }
  if (err) {
    xgpu_record_error(status, status_elem, err, el);
    if (!FILL) counts[i] = 0;
    return;
  }
  if (!FILL) counts[i] = E.n;
}

// CutInformation::Update on a time slab in ONE pass (xfem/cutinfo.cpp:90-124 with time_order >= 0): the part volumes of
// the POS and NEG space-time rules are summed in rule order — the same additions, in the same order, as k_st<D,true> of
// each domain followed by k_rule_sums, without materialising the points — then the element is classified by ballot.
template <int D>
__global__ void __launch_bounds__(128) k_st_classify(int64_t n, const int32_t* __restrict__ el_vtx,
                                                     const double* __restrict__ phi_st, int64_t nv, int tord, RuleSet R,
                                                     const double* __restrict__ tq_p, const double* __restrict__ tq_w, int ntq,
                                                     uint32_t* __restrict__ bneg, uint32_t* __restrict__ bpos,
                                                     uint32_t* __restrict__ bif, double* __restrict__ ratio, int* status,
                                                     unsigned long long* status_elem)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < n;
  double pv[2] = {0.0, 0.0};  // [XGPU_NEG], [XGPU_POS]
  if (in) {
    const int nt = tord + 1;
    int v[D + 1];
#pragma unroll
    for (int k = 0; k < D + 1; k++) v[k] = __ldg(el_vtx + i * (D + 1) + k);
    double nd[XG_ST_MAXORD + 1];
    time_nodes(tord, nd);
    double ls[XG_ST_MAXORD + 1][D + 1];
    for (int j = 0; j < nt; j++)
#pragma unroll
      for (int k = 0; k < D + 1; k++) ls[j][k] = __ldg(phi_st + (int64_t)j * nv + v[k]);
    double cutp[XG_ST_MAXCP];
    int ncp = 2, err = 0;
    cutp[0] = 0.0; cutp[1] = 1.0;
    for (int k = 0; k < D + 1 && !err; k++) {
      double li[XG_ST_MAXORD + 1];
      for (int j = 0; j < nt; j++) li[j] = ls[j][k];
      const int nr = root_finding(li, tord, nd, cutp + ncp, XG_ST_MAXCP - ncp);
      if (nr < 0) { err = nr; break; }
      ncp += nr;
    }
    if (!err) {
      for (int a = 1; a < ncp; a++) {
        const double key = cutp[a];
        int b = a - 1;
        while (b >= 0 && cutp[b] > key) { cutp[b + 1] = cutp[b]; b--; }
        cutp[b + 1] = key;
      }
      for (int c = 0; c < ncp - 1 && !err; c++) {
        const double t0 = cutp[c], t1 = cutp[c + 1];
        for (int q = 0; q < ntq && !err; q++) {
          const double t = t0 + tq_p[3 * q] * (t1 - t0);
          double sh[XG_ST_MAXORD + 1], lt[D + 1];
          time_shapes(tord, nd, t, sh);
#pragma unroll
          for (int k = 0; k < D + 1; k++) {
            double d = 0.0;
            for (int j = 0; j < nt; j++) d += ls[j][k] * sh[j];
            if (fabs(d) < 1e-14) d = (d >= 0) ? 1e-14 : -1e-14;
            lt[k] = d;
          }
          const double wfac = tq_w[q] * (t1 - t0);
          const int dom = xg_check_cut(lt, D + 1);
          if (dom == XGPU_IF) {
            XgSimplex ref, sub[3];
            xg_ref_simplex(D, ref);
            for (int np = 0; np < 2 && !err; np++) {
              const int ns = xg_decompose(ref, lt, np, sub);
              if (ns < 0) { err = ns; break; }
              for (int s = 0; s < ns; s++) {
                const double vol = xg_volume(sub[s]) * wfac;
                const int npt = R.n[sub[s].D];
                for (int qq = 0; qq < npt; qq++) pv[np] += R.w[sub[s].D][qq] * vol;
              }
            }
          } else {
            for (int qq = 0; qq < R.n[D]; qq++) pv[dom] += R.w[D][qq] * wfac;
          }
        }
      }
    }
    if (err) { xgpu_record_error(status, status_elem, err, i); pv[0] = pv[1] = 0.0; }
  }
  const bool hn = in && pv[XGPU_NEG] > 0.0, hp = pv[XGPU_POS] > 0.0;
  const unsigned mn = __ballot_sync(0xffffffffu, hn && !hp);
  const unsigned mi = __ballot_sync(0xffffffffu, hn && hp);
  const unsigned mp = __ballot_sync(0xffffffffu, in && !hn);
  if (in && ratio) ratio[i] = pv[XGPU_NEG] / (pv[XGPU_NEG] + pv[XGPU_POS]);
  if ((threadIdx.x & 31) == 0 && in) { bneg[i >> 5] = mn; bif[i >> 5] = mi; bpos[i >> 5] = mp; }
}

int make_ruleset(xgpu_ctx* ctx, int dim, int order, RuleSet& R)
{
  for (int d = 0; d < 4; d++) { R.p[d] = nullptr; R.w[d] = nullptr; R.n[d] = 0; }
  for (int d = 0; d <= dim; d++) {
    const xgpu_rule_host* r = xgpu_rule_lookup(ctx, d, order);
    if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "cutquad: no rule");
    R.p[d] = r->d_pts; R.w[d] = r->d_wts; R.n[d] = r->n;
  }
  return XGPU_OK;
}

int make_tuples(xgpu_ctx* ctx, int M, const int* dts_host, int ntuples, MlTuples& tp)
{
  if (M < 1 || M > XG_ML_MAXM || ntuples < 0 || ntuples > XG_ML_MAXT || (ntuples > 0 && !dts_host))
    return xgpu_fail(ctx, XGPU_ERR_ARG, "cutquad_ml: bad argument (M<=16, ntuples<=64)");
  tp.M = M; tp.nt = ntuples;
  for (int i = 0; i < M * ntuples; i++) {
    if (dts_host[i] < 0 || dts_host[i] > 2) return xgpu_fail(ctx, XGPU_ERR_ARG, "cutquad_ml: domain types must be NEG/POS/IF");
    tp.dts[i] = (signed char)dts_host[i];
  }
  return XGPU_OK;
}

int finish_count(xgpu_ctx* ctx, const int32_t* counts, int64_t n, int64_t* offsets, int64_t* total_host)
{
  int rc = xgpu_exclusive_scan_i32(ctx, counts, n, offsets);
  if (rc) return rc;
  if (total_host) {
    XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, offsets + n, sizeof(int64_t), ctx->stream));
    XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *total_host = ctx->h_pinned[0];
  }
  return XGPU_OK;
}

template <bool FILL>
int launch_ml(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
              const double* phis, int64_t nv, const MlTuples& tp, const RuleSet& R, const int64_t* offsets, int32_t* counts,
              double* pts, double* wts)
{
  const unsigned g = nblk(n, 128);
  if (dim == 3) k_ml<3, FILL><<<g, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phis, nv, tp, R, offsets, counts, pts, wts, ctx->d_status, ctx->d_status_elem);
  else if (dim == 2) k_ml<2, FILL><<<g, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phis, nv, tp, R, offsets, counts, pts, wts, ctx->d_status, ctx->d_status_elem);
  else k_ml<1, FILL><<<g, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phis, nv, tp, R, offsets, counts, pts, wts, ctx->d_status, ctx->d_status_elem);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

template <bool FILL>
int launch_st(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
              const double* phi_st, int64_t nv, int tord, int dt, const RuleSet& R, const xgpu_rule_host* tr,
              const int64_t* offsets, int32_t* counts, double* pts, double* tvals, double* wts)
{
  const unsigned g = nblk(n, 128);
#define XG_ST(DD) k_st<DD, FILL><<<g, 128, 0, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phi_st, nv, tord, dt, R, tr->d_pts, tr->d_wts, tr->n, offsets, counts, pts, tvals, wts, ctx->d_status, ctx->d_status_elem)
  if (dim == 3) XG_ST(3); else if (dim == 2) XG_ST(2); else XG_ST(1);
#undef XG_ST
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

extern "C" int xgpu_cutquad_ml_count(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                     const double* phis, int64_t nv, int M, const int* dts_host, int ntuples, int order,
                                     int64_t* offsets, int64_t* total_host)
{
  if (!ctx || dim < 1 || dim > 3 || n < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_ml_count: bad argument");
  if (total_host) *total_host = 0;
  MlTuples tp; RuleSet R;
  int rc = make_tuples(ctx, M, dts_host, ntuples, tp);
  if (rc) return rc;
  rc = make_ruleset(ctx, dim, order, R);
  if (rc) return rc;
  void* p;
  rc = xgpu_scratch(ctx, 0, sizeof(int32_t) * (n + 1), &p);
  if (rc) return rc;
  if (n > 0) { rc = launch_ml<false>(ctx, dim, n, elems, el_vtx, nullptr, phis, nv, tp, R, nullptr, (int32_t*)p, nullptr, nullptr); if (rc) return rc; }
  return finish_count(ctx, (int32_t*)p, n, offsets, total_host);
}

extern "C" int xgpu_cutquad_ml_fill(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                    const double* vtx_xyz, const double* phis, int64_t nv, int M, const int* dts_host,
                                    int ntuples, int order, const int64_t* offsets, double* pts, double* wts)
{
  XGPU_NVTX(ctx, "xgpu_cutquad_ml_fill");
  if (!ctx || dim < 1 || dim > 3 || n < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_ml_fill: bad argument");
  if (ctx->def_order > 0) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_cutquad_ml_fill: no deformation support (unset it)");
  if (n == 0) return XGPU_OK;
  MlTuples tp; RuleSet R;
  int rc = make_tuples(ctx, M, dts_host, ntuples, tp);
  if (rc) return rc;
  rc = make_ruleset(ctx, dim, order, R);
  if (rc) return rc;
  return launch_ml<true>(ctx, dim, n, elems, el_vtx, vtx_xyz, phis, nv, tp, R, offsets, nullptr, pts, wts);
}

extern "C" int xgpu_cutquad_st_count(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                     const double* phi_st, int64_t nv, int time_fe_order, int dt, int order_time,
                                     int order_space, int64_t* offsets, int64_t* total_host)
{
  if (!ctx || dim < 1 || dim > 3 || n < 0 || dt < 0 || dt > 2 || time_fe_order < 0 || time_fe_order > XG_ST_MAXORD)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_st_count: bad argument (time order <= 8)");
  if (total_host) *total_host = 0;
  if (order_space == -1) order_space = 5;  // spacetimecutrule.cpp:150-151
  RuleSet R;
  int rc = make_ruleset(ctx, dim, order_space, R);
  if (rc) return rc;
  const xgpu_rule_host* tr = xgpu_rule_lookup(ctx, 1, order_time);
  if (!tr) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_st_count: no time rule");
  void* p;
  rc = xgpu_scratch(ctx, 0, sizeof(int32_t) * (n + 1), &p);
  if (rc) return rc;
  if (n > 0) { rc = launch_st<false>(ctx, dim, n, elems, el_vtx, nullptr, phi_st, nv, time_fe_order, dt, R, tr, nullptr, (int32_t*)p, nullptr, nullptr, nullptr); if (rc) return rc; }
  return finish_count(ctx, (int32_t*)p, n, offsets, total_host);
}

extern "C" int xgpu_cutquad_st_fill(xgpu_ctx* ctx, int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                                    const double* vtx_xyz, const double* phi_st, int64_t nv, int time_fe_order, int dt,
                                    int order_time, int order_space, const int64_t* offsets, double* pts, double* tvals,
                                    double* wts)
{
  XGPU_NVTX(ctx, "xgpu_cutquad_st_fill");
  if (ctx && ctx->def_order > 0) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_cutquad_st_fill: no deformation support (unset it)");
  if (!ctx || dim < 1 || dim > 3 || n < 0 || dt < 0 || dt > 2 || time_fe_order < 0 || time_fe_order > XG_ST_MAXORD)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_st_fill: bad argument (time order <= 8)");
  if (n == 0) return XGPU_OK;
  if (order_space == -1) order_space = 5;
  RuleSet R;
  int rc = make_ruleset(ctx, dim, order_space, R);
  if (rc) return rc;
  const xgpu_rule_host* tr = xgpu_rule_lookup(ctx, 1, order_time);
  if (!tr) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_cutquad_st_fill: no time rule");
  return launch_st<true>(ctx, dim, n, elems, el_vtx, vtx_xyz, phi_st, nv, time_fe_order, dt, R, tr, offsets, nullptr, pts, tvals, wts);
}

extern "C" int xgpu_classify_st(xgpu_ctx* ctx, int dim, int64_t ne, const int32_t* el_vtx, const double* phi_st, int64_t nv,
                                int time_fe_order, int time_order, uint32_t* bits_neg, uint32_t* bits_pos, uint32_t* bits_if,
                                double* cut_ratio)
{
  XGPU_NVTX(ctx, "xgpu_classify_st");
  if (!ctx || dim < 1 || dim > 3 || ne < 0 || time_fe_order < 0 || time_fe_order > XG_ST_MAXORD || !bits_neg || !bits_pos || !bits_if)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_classify_st: bad argument (time order <= 8)");
  if (ne == 0) return XGPU_OK;
  RuleSet R;
  int rc = make_ruleset(ctx, dim, 0, R);  // order 0 (cutinfo.cpp:98)
  if (rc) return rc;
  const xgpu_rule_host* tr = xgpu_rule_lookup(ctx, 1, time_order < 0 ? 0 : time_order);
  if (!tr) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_classify_st: no time rule");
  const unsigned g = nblk(ne, 128);
#define XG_STC(DD) k_st_classify<DD><<<g, 128, 0, ctx->stream>>>(ne, el_vtx, phi_st, nv, time_fe_order, R, tr->d_pts, tr->d_wts, tr->n, bits_neg, bits_pos, bits_if, cut_ratio, ctx->d_status, ctx->d_status_elem)
  if (dim == 3) XG_STC(3); else if (dim == 2) XG_STC(2); else XG_STC(1);
#undef XG_STC
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}
