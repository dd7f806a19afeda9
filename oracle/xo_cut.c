/* xo_cut.c — ORACLE (test infrastructure): straight-cut quadrature on simplices.
 * Restates cutint/straightcutrule.cpp, cutint/mlsetintegration.cpp, cutint/spacetimecutrule.cpp
 * (line numbers cited per function).  Quads/hexes (straightcutrule.cpp:206-501) are not restated.
 */
#include "xoracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- CheckIfStraightCut: straightcutrule.cpp:11-24 (and :26-39) ---- */
int xo_check_if_straight_cut(const double* v, int n, double eps)
{
  int haspos = 0, hasneg = 0;
  for (int i = 0; i < n; i++) {
    if (!haspos && (v[i] > eps)) haspos = 1;
    if (!hasneg && (v[i] < -eps)) hasneg = 1;
    if (haspos && hasneg) break;
  }
  if ((hasneg && haspos) || (!hasneg && !haspos)) return XO_IF;
  else if (hasneg) return XO_NEG;
  else return XO_POS;
}

/* ---- SimpleX(ELEMENT_TYPE): straightcutrule.hpp:56-60 ---- */
void xo_ref_simplex(int dim, xo_simplex* s)
{
  memset(s, 0, sizeof(*s));
  s->D = dim;
  if (dim == 1) { s->p[0][0] = 1; }
  else if (dim == 2) { s->p[0][0] = 1; s->p[1][1] = 1; }
  else if (dim == 3) { s->p[0][0] = 1; s->p[1][1] = 1; s->p[2][2] = 1; }
}

static void sub3(const double* a, const double* b, double* c) { for (int i = 0; i < 3; i++) c[i] = a[i] - b[i]; }
static void cross3(const double* a, const double* b, double* c)
{
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static double det3(const double* a, const double* b, const double* c)
{ /* determinant of the matrix with columns a, b, c */
  return a[0] * (b[1] * c[2] - b[2] * c[1]) - b[0] * (a[1] * c[2] - a[2] * c[1]) + c[0] * (a[1] * b[2] - a[2] * b[1]);
}

/* ---- SimpleX::GetVolume: straightcutrule.cpp:66-72 ---- */
double xo_simplex_volume(const xo_simplex* s)
{
  double a[3], b[3], c[3];
  if (s->D == 0) return 1.0;
  if (s->D == 1) { sub3(s->p[1], s->p[0], a); return norm3(a); }
  if (s->D == 2) { sub3(s->p[2], s->p[0], a); sub3(s->p[1], s->p[0], b); cross3(a, b, c); return norm3(c); }
  sub3(s->p[3], s->p[0], a); sub3(s->p[2], s->p[0], b); sub3(s->p[1], s->p[0], c);
  return fabs(det3(a, b, c));
}

/* ---- SimpleX::CalcIFPolytopEUsingLset: straightcutrule.cpp:41-64 ---- */
static int calc_if_polytope(const xo_simplex* s, const double* l, double cut[][3], int* ncut)
{
  int np = s->D + 1;
  *ncut = 0;
  if (xo_check_if_straight_cut(l, np, 0.0) != XO_IF) return XO_ERR_PLAIN_GEOM;
  if (s->D == 1) {
    double t = l[0] / (l[0] - l[1]);
    for (int k = 0; k < 3; k++) cut[0][k] = s->p[0][k] + t * (s->p[1][k] - s->p[0][k]);
    *ncut = 1;
    return XO_OK;
  }
  for (int i = 0; i < np; i++)
    for (int j = i + 1; j < np; j++)
      if ((l[i] >= 0) != (l[j] >= 0)) {
        double pair[2] = { l[i], l[j] };
        /* the recursive 1D call re-checks the pair (:45-49): (0, negative) throws */
        if (xo_check_if_straight_cut(pair, 2, 0.0) != XO_IF) return XO_ERR_PLAIN_GEOM;
        double t = l[i] / (l[i] - l[j]);
        for (int k = 0; k < 3; k++) cut[*ncut][k] = s->p[i][k] + t * (s->p[j][k] - s->p[i][k]);
        (*ncut)++;
      }
  return XO_OK;
}

static void set_simplex(xo_simplex* o, int D, const double* q0, const double* q1, const double* q2, const double* q3)
{
  const double* q[4] = { q0, q1, q2, q3 };
  memset(o, 0, sizeof(*o));
  o->D = D;
  for (int i = 0; i <= D; i++) for (int k = 0; k < 3; k++) o->p[i][k] = q[i][k];
}

/* ---- LevelsetCutSimplex::Decompose: straightcutrule.cpp:131-196 ---- */
int xo_decompose(const xo_simplex* s, const double* l, int dt, xo_simplex* out)
{
  double c[6][3];
  int nc, D = s->D;
  int err = calc_if_polytope(s, l, c, &nc);
  if (err) return err;
  if (dt == XO_IF) {
    if (nc == D) { set_simplex(&out[0], D - 1, c[0], c[1], c[2], c[3]); return 1; }
    else if (nc == 4 && D == 3) {
      set_simplex(&out[0], 2, c[0], c[1], c[3], 0);
      set_simplex(&out[1], 2, c[0], c[2], c[3], 0);
      return 2;
    }
    return XO_ERR_BAD_SCUT;
  }
  int rel[4], nr = 0;
  for (int i = 0; i < D + 1; i++)
    if (((dt == XO_POS) && (l[i] >= 0)) || ((dt == XO_NEG) && (l[i] < 0))) rel[nr++] = i;
  if (nr == 1) { /* cut points ..., r0 */
    const double* q[4] = { 0, 0, 0, 0 };
    for (int i = 0; i < nc; i++) q[i] = c[i];
    if (nc > 3) return XO_ERR_NOT_IMPL;
    q[nc] = s->p[rel[0]];
    set_simplex(&out[0], nc, q[0], q[1], q[2], q[3]);
    return 1;
  }
  else if (nr == 2 && D == 2) {
    set_simplex(&out[0], 2, s->p[rel[0]], s->p[rel[1]], c[1], 0);
    set_simplex(&out[1], 2, c[0], c[1], s->p[rel[0]], 0);
    return 2;
  }
  else if (nr == 2 && D == 3) {
    set_simplex(&out[0], 3, s->p[rel[1]], c[1], c[2], c[3]);
    set_simplex(&out[1], 3, s->p[rel[0]], s->p[rel[1]], c[1], c[2]);
    set_simplex(&out[2], 3, c[0], c[1], c[2], s->p[rel[0]]);
    return 3;
  }
  else if (nr == 3 && D == 3) {
    set_simplex(&out[0], 3, c[0], c[1], c[2], s->p[rel[2]]);
    set_simplex(&out[1], 3, s->p[rel[0]], s->p[rel[1]], s->p[rel[2]], c[1]);
    set_simplex(&out[2], 3, c[0], c[1], s->p[rel[0]], s->p[rel[2]]);
    return 3;
  }
  return XO_ERR_NOT_IMPL;
}

/* ---- SimpleX::GetPlainIntegrationRule: straightcutrule.cpp:80-100 ---- */
static int plain_rule(const xo_simplex* s, int order, double* pts, double* wts, int at, int cap)
{
  static __thread double rp[3 * XO_MAXPTS_RULE], rw[XO_MAXPTS_RULE];
  double trafofac = xo_simplex_volume(s);
  int n = xo_select_rule(s->D, order, rp, rw);
  if (at + n > cap) return XO_ERR_CAPACITY;
  for (int q = 0; q < n; q++) {
    double ow = 1.0, pt[3];
    for (int m = 0; m < s->D; m++) ow -= rp[3 * q + m];
    for (int k = 0; k < 3; k++) pt[k] = ow * s->p[0][k];
    for (int m = 0; m < s->D; m++) for (int k = 0; k < 3; k++) pt[k] += rp[3 * q + m] * s->p[m + 1][k];
    for (int k = 0; k < 3; k++) pts[3 * (at + q) + k] = pt[k];
    wts[at + q] = rw[q] * trafofac;
  }
  return n;
}

/* ---- LevelsetWrapper for simplices: straightcutrule.cpp:503-556 ---- */
typedef struct { double c0, cx, cy, cz; } lsetw;
static lsetw lset_wrap(int dim, const double* v)
{
  lsetw L = { 0, 0, 0, 0 };
  if (dim == 1) { L.c0 = v[1]; L.cx = v[0] - v[1]; }
  else if (dim == 2) { L.c0 = v[2]; L.cx = v[0] - v[2]; L.cy = v[1] - v[2]; }
  else { L.c0 = v[3]; L.cx = v[0] - v[3]; L.cy = v[1] - v[3]; L.cz = v[2] - v[3]; }
  return L;
}
static double lset_eval(const lsetw* L, const double* p) { return L->c0 + L->cx * p[0] + L->cy * p[1] + L->cz * p[2]; }
static void lset_normal(const lsetw* L, double* n)
{
  n[0] = L->cx; n[1] = L->cy; n[2] = L->cz;
  double nn = norm3(n);
  n[0] /= nn; n[1] /= nn; n[2] /= nn;
}

static int inv_mat(int D, const double* F, double* Fi, double* det)
{
  if (D == 1) { *det = F[0]; Fi[0] = 1.0 / F[0]; return 0; }
  if (D == 2) {
    double d = F[0] * F[3] - F[1] * F[2];
    *det = d;
    Fi[0] = F[3] / d; Fi[1] = -F[1] / d; Fi[2] = -F[2] / d; Fi[3] = F[0] / d;
    return 0;
  }
  double d = F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
  *det = d;
  Fi[0] = (F[4] * F[8] - F[5] * F[7]) / d; Fi[1] = (F[2] * F[7] - F[1] * F[8]) / d; Fi[2] = (F[1] * F[5] - F[2] * F[4]) / d;
  Fi[3] = (F[5] * F[6] - F[3] * F[8]) / d; Fi[4] = (F[0] * F[8] - F[2] * F[6]) / d; Fi[5] = (F[2] * F[3] - F[0] * F[5]) / d;
  Fi[6] = (F[3] * F[7] - F[4] * F[6]) / d; Fi[7] = (F[1] * F[6] - F[0] * F[7]) / d; Fi[8] = (F[0] * F[4] - F[1] * F[3]) / d;
  return 0;
}

/* ---- TransformQuadUntrafoToIRInterface<D>: straightcutrule.cpp:558-578 ---- */
static void transform_if_weights(const xo_trafo* T, const lsetw* L, const double* pts, double* wts, int n)
{
  int D = T->dim;
  double nref[3];
  lset_normal(L, nref);
  for (int q = 0; q < n; q++) {
    double F[9], Fi[9], x[3], det;
    xo_trafo_eval(T, &pts[3 * q], F, x);
    inv_mat(D, F, Fi, &det);
    double nrm[3] = { 0, 0, 0 };
    for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) nrm[i] += Fi[j * D + i] * nref[j]; /* Finv^T n */
    wts[q] *= norm3(nrm);
  }
}

/* ---- StraightCutIntegrationRule: straightcutrule.cpp:582-677 (simplices) ---- */
int xo_straight_cut_rule(const xo_trafo* T, const double* lsetvals, int dt, int order,
                         double* pts, double* wts, int cap, int* is_std)
{
  int D = T->dim;
  if (is_std) *is_std = 0;
  int element_domain = xo_check_if_straight_cut(lsetvals, D + 1, 0.0);
  if (element_domain != XO_IF) {
    if (element_domain != dt) return 0;
    int n = xo_rule_size(D, order);
    if (n > cap) return XO_ERR_CAPACITY;
    xo_select_rule(D, order, pts, wts);
    if (is_std) *is_std = 1;
    return n;
  }
  xo_simplex ref, sub[3];
  xo_ref_simplex(D, &ref);
  int ns = xo_decompose(&ref, lsetvals, dt, sub);
  if (ns < 0) return ns;
  int n = 0;
  for (int i = 0; i < ns; i++) {
    int k = plain_rule(&sub[i], order, pts, wts, n, cap);
    if (k < 0) return k;
    n += k;
  }
  if (dt == XO_IF) {
    lsetw L = lset_wrap(D, lsetvals);
    transform_if_weights(T, &L, pts, wts, n);
  }
  return n;
}

/* ---- StraightCutsIntegrationRule: straightcutrule.cpp:679-812 ---- */
#define XO_MAXSIMP 729
static int straight_cuts_rule(const xo_trafo* T, const double* lsets /*(D+1) x M*/, int M, const int* dts,
                              int order, double* pts, double* wts, int at, int cap)
{
  int D = T->dim, np = D + 1;
  double v[4];
  xo_simplex* last = (xo_simplex*)malloc(sizeof(xo_simplex) * XO_MAXSIMP);
  xo_simplex* cur = (xo_simplex*)malloc(sizeof(xo_simplex) * XO_MAXSIMP);
  int nlast = 0, ncur = 0, ret = 0;
  xo_simplex ref;
  xo_ref_simplex(D, &ref);
  for (int k = 0; k < np; k++) v[k] = lsets[k * M + 0];
  nlast = xo_decompose(&ref, v, dts[0], last);
  if (nlast < 0) { ret = nlast; goto done; }
  for (int i = 1; i < M; i++) {
    for (int k = 0; k < np; k++) v[k] = lsets[k * M + i];
    lsetw L = lset_wrap(D, v);
    ncur = 0;
    for (int s = 0; s < nlast; s++) {
      double lv[4];
      int nps = last[s].D + 1;
      for (int k = 0; k < nps; k++) { /* update_initial_coefs :548-556 */
        double d = lset_eval(&L, last[s].p[k]);
        lv[k] = (fabs(d) > 1e-14) ? d : 1e-14;
      }
      int dts_s = xo_check_if_straight_cut(lv, nps, 0.0);
      if (dts_s == XO_IF) {
        if (ncur + 3 > XO_MAXSIMP) { ret = XO_ERR_CAPACITY; goto done; }
        int k = xo_decompose(&last[s], lv, dts[i], &cur[ncur]);
        if (k < 0) { ret = k; goto done; }
        ncur += k;
      }
      else if (dts_s == dts[i]) {
        if (ncur + 1 > XO_MAXSIMP) { ret = XO_ERR_CAPACITY; goto done; }
        cur[ncur++] = last[s];
      }
    }
    xo_simplex* t = last; last = cur; cur = t;
    nlast = ncur;
  }
  {
    int n = 0;
    for (int s = 0; s < nlast; s++) {
      int k = plain_rule(&last[s], order, pts, wts, at + n, cap);
      if (k < 0) { ret = k; goto done; }
      n += k;
    }
    if (n == 0) { ret = 0; goto done; }
    int ifidx[8], nif = 0;
    for (int i = 0; i < M; i++) if (dts[i] == XO_IF) { if (nif < 8) ifidx[nif] = i; nif++; }
    double* P = pts + 3 * at; double* W = wts + at;
    if (nif == 0) { }
    else if (nif == 1) {
      for (int k = 0; k < np; k++) v[k] = lsets[k * M + ifidx[0]];
      lsetw L = lset_wrap(D, v);
      transform_if_weights(T, &L, P, W, n);
    }
    else if (nif == 2) {
      if (D != 2 && D != 3) { ret = XO_ERR_NOT_IMPL; goto done; }
      if (D == 2) {
        for (int q = 0; q < n; q++) {
          double F[9], Fi[9], x[3], det;
          xo_trafo_eval(T, &P[3 * q], F, x); inv_mat(2, F, Fi, &det);
          W[q] /= fabs(det);
        }
      } else {
        double n0[3], n1[3], cr[3];
        for (int k = 0; k < np; k++) v[k] = lsets[k * M + ifidx[0]];
        lsetw L0 = lset_wrap(D, v);
        for (int k = 0; k < np; k++) v[k] = lsets[k * M + ifidx[1]];
        lsetw L1 = lset_wrap(D, v);
        lset_normal(&L0, n0); lset_normal(&L1, n1);
        cross3(n0, n1, cr);
        double cp_fac = 1.0 / norm3(cr);
        for (int q = 0; q < n; q++) {
          double F[9], Fi[9], x[3], det, nrm[3] = { 0, 0, 0 };
          xo_trafo_eval(T, &P[3 * q], F, x); inv_mat(3, F, Fi, &det);
          for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) nrm[i] += cp_fac * F[i * 3 + j] * cr[j];
          W[q] = W[q] * norm3(nrm) / fabs(det);
        }
      }
    }
    else if (nif == 3) {
      if (D != 3) { ret = XO_ERR_NOT_IMPL; goto done; }
      for (int q = 0; q < n; q++) {
        double F[9], Fi[9], x[3], det;
        xo_trafo_eval(T, &P[3 * q], F, x); inv_mat(3, F, Fi, &det);
        W[q] /= fabs(det);
      }
    }
    else { ret = XO_ERR_NOT_IMPL; goto done; }
    ret = n;
  }
done:
  free(last); free(cur);
  return ret;
}

/* ---- CreateMultiLevelsetCutIntegrationRule: mlsetintegration.cpp:6-186 ---- */
int xo_mlset_cut_rule(const xo_trafo* T, const double* lsets, int M, const int* dts_all, int ntuples,
                      int order, double* pts, double* wts, int cap)
{
  int D = T->dim, np = D + 1;
  int lset_dts[16];
  double v[4];
  if (M > 16) return XO_ERR_ARG;
  for (int i = 0; i < M; i++) {
    for (int k = 0; k < np; k++) v[k] = lsets[k * M + i];
    lset_dts[i] = xo_check_if_straight_cut(v, np, 0.0);
  }
  int n = 0;
  for (int t = 0; t < ntuples; t++) {
    const int* dts = dts_all + t * M;
    int compatible = 1, cut_element = 0;
    for (int i = 0; i < M; i++) {
      if ((lset_dts[i] != XO_IF) && (lset_dts[i] != dts[i])) compatible = 0;
      else if (lset_dts[i] == XO_IF) cut_element = 1;
    }
    if (!compatible) continue;
    if (!cut_element) {
      int k = xo_rule_size(D, order);
      if (n + k > cap) return XO_ERR_CAPACITY;
      xo_select_rule(D, order, pts + 3 * n, wts + n);
      n += k;
      continue;
    }
    int cid[16], nc = 0, ctd[16];
    for (int i = 0; i < M; i++)
      if ((lset_dts[i] == XO_IF) || (lset_dts[i] != dts[i])) { cid[nc] = i; ctd[nc] = dts[i]; nc++; }
    if (nc == 1) {
      int j = cid[0];
      for (int k = 0; k < np; k++) v[k] = lsets[k * M + j];
      int k = xo_straight_cut_rule(T, v, dts[j], order, pts + 3 * n, wts + n, cap - n, 0);
      if (k < 0) return k;
      n += k;
    } else {
      double cond[4 * 16];
      for (int i = 0; i < nc; i++) for (int k = 0; k < np; k++) cond[k * nc + i] = lsets[k * M + cid[i]];
      int k = straight_cuts_rule(T, cond, nc, ctd, order, pts, wts, n, cap);
      if (k < 0) return k;
      n += k;
    }
  }
  return n;
}

/* ---- time basis: spacetime/SpaceTimeFE.cpp:260-270 nodes; Lagrange on them.
 *      order 1: {0,1}; order 2: {0,1/2,1} (spacetimecutrule.cpp:13-24 states these bases) ---- */
static void time_nodes(int order, double* nodes)
{
  if (order == 0) { nodes[0] = 0.0; return; }
  if (order == 1) { nodes[0] = 0; nodes[1] = 1; return; }
  if (order == 2) { nodes[0] = 0; nodes[1] = 0.5; nodes[2] = 1; return; }
  if (order == 3) { /* Gauss-Lobatto on [0,1] */
    nodes[0] = 0; nodes[1] = 0.5 * (1.0 - 1.0 / sqrt(5.0)); nodes[2] = 0.5 * (1.0 + 1.0 / sqrt(5.0)); nodes[3] = 1; return;
  }
  if (order == 4) { /* SpaceTimeFE.cpp:271-272 */
    nodes[0] = 0; nodes[1] = 0.5 * (1.0 - sqrt(3.0 / 7.0)); nodes[2] = 0.5; nodes[3] = 0.5 * (1.0 + sqrt(3.0 / 7.0)); nodes[4] = 1; return;
  }
  if (order == 5) { /* SpaceTimeFE.cpp:273-278 */
    const double a = sqrt(1.0 / 3.0 + 2.0 * sqrt(7.0) / 21.0), b = sqrt(1.0 / 3.0 - 2.0 * sqrt(7.0) / 21.0);
    nodes[0] = 0; nodes[1] = 0.5 * (1.0 - a); nodes[2] = 0.5 * (1.0 - b); nodes[3] = 0.5 * (1.0 + b); nodes[4] = 0.5 * (1.0 + a); nodes[5] = 1; return;
  }
  /* order >= 6: Gauss-Lobatto-Legendre points by Newton's method on (x P_N - P_{N-1}), started from the Chebyshev-Lobatto
   * points (CalcLobattoPointsRec, SpaceTimeFE.cpp:229-256); every node iterates on its own */
  const int N = order;
  for (int j = 0; j <= N; j++) {
    double x = cos(M_PI * (double)j / N);
    for (int it = 0; it < 100; it++) {
      double pm = 1.0, p = x;                       /* P_0, P_1 */
      for (int k = 1; k < N; k++) { const double pn = ((2.0 * k + 1.0) * (x * p) - (double)k * pm) / ((double)k + 1.0); pm = p; p = pn; }
      const double xn = x - (x * p - pm) / (N * p);
      const int done = fabs(xn - x) <= 1e-20;
      x = xn;
      if (done) break;
    }
    nodes[j] = 0.5 * (1.0 - x);
  }
  nodes[0] = 0.0; nodes[N] = 1.0;
}
void xo_time_shapes(int order, double t, double* shape)
{
  double nd[32];
  time_nodes(order, nd);
  for (int i = 0; i <= order; i++) {
    double s = 1.0;
    for (int j = 0; j <= order; j++) if (j != i) s *= (t - nd[j]) / (nd[i] - nd[j]);
    shape[i] = s;
  }
}

/* ---- root_finding: spacetimecutrule.cpp:8-102 ---- */
static double eval_time(const double* li, int order, double t)
{
  double sh[32], s = 0;
  xo_time_shapes(order, t, sh);
  for (int i = 0; i <= order; i++) s += li[i] * sh[i];
  return s;
}
static int root_finding(const double* li, int order, double* roots)
{
  int nr = 0;
  if (order == 0) return 0;
  if (order == 1) {
    if ((li[0] >= 0) != (li[1] >= 0)) roots[nr++] = -li[0] / (li[1] - li[0]);
    return nr;
  }
  if (order == 2) {
    double c = li[0], a = 2 * li[0] + 2 * li[2] - 4. * li[1], b = li[2] - a - c;
    if (fabs(a) < 1e-12) {
      double x = -c / b;
      if (x < 1 && x > 0) roots[nr++] = x;
    } else {
      const double w = b * b - 4 * a * c;
      if (fabs(w) < 1e-12) {
        double x = -b / (2 * a);
        if (x < 1 && x > 0) roots[nr++] = x;
      } else {
        double x1 = (b + sqrt(w)) / (-2.0 * a);
        double x2 = (b - sqrt(w)) / (-2.0 * a);
        if (x1 < 1 && x1 > 0) roots[nr++] = x1;
        if (x2 < 1 && x2 > 0) roots[nr++] = x2;
      }
    }
    return nr;
  }
  {
    const int subdivs = 50, bis_it = 70;
    const double EPS = 1e-15; /* EPS_STCR_ROOT_SEARCH_BISECTION utils/ngsxstd.cpp:66-82 */
    double vals[51], ia[50], ib[50];
    int ni = 0;
    double dx = 1. / subdivs;
    for (int i = 0; i < subdivs + 1; i++) {
      double xi = dx * i;
      vals[i] = eval_time(li, order, xi);
      if (vals[i] == 0) vals[i] = EPS;
      if (i >= 1) if (vals[i - 1] * vals[i] < 0) { ia[ni] = xi - dx; ib[ni] = xi; ni++; }
    }
    for (int k = 0; k < ni; k++) {
      double a = ia[k], b = ib[k], xm;
      double aval = eval_time(li, order, a), bval = eval_time(li, order, b);
      for (int j = 0; j < bis_it; j++) {
        const double xl = a - aval * (b - a) / (bval - aval);
        double val = eval_time(li, order, xl);
        if (2 * fabs(val) < EPS) { a = b = xl; break; }
        xm = 0.5 * (a + b);
        val = eval_time(li, order, xm);
        if (val == 0) break;
        if (val * aval < 0) { b = xm; bval = val; }
        else if (val * bval < 0) { a = xm; aval = val; }
        else return XO_ERR_NOT_IMPL;
      }
      roots[nr++] = 0.5 * (a + b);
    }
  }
  return nr;
}

static int cmp_double(const void* a, const void* b)
{
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* ---- SpaceTimeCutIntegrationRule: spacetimecutrule.cpp:104-244 (DO_NAIVE_TIMEINT=false) ---- */
int xo_spacetime_cut_rule(const xo_trafo* T, const double* lset_st, int time_fe_order, int dt,
                          int order_time, int order_space, double* pts, double* tvals,
                          double* wts, int cap)
{
  int D = T->dim, ns = D + 1, nt = time_fe_order + 1;
  double cutp[2 + 4 * 64];
  int ncp = 2;
  cutp[0] = 0; cutp[1] = 1;
  for (int i = 0; i < ns; i++) {
    double li[32], r[64];
    for (int j = 0; j < nt; j++) li[j] = lset_st[j * ns + i];
    int nr = root_finding(li, time_fe_order, r);
    if (nr < 0) return nr;
    for (int k = 0; k < nr; k++) cutp[ncp++] = r[k];
  }
  qsort(cutp, ncp, sizeof(double), cmp_double);
  double tp[3 * 64], tw[64];
  int ntq = xo_select_rule(1, order_time, tp, tw);
  if (order_space == -1) order_space = 5;
  int n = 0;
  for (int i = 0; i < ncp - 1; i++) {
    double t0 = cutp[i], t1 = cutp[i + 1];
    for (int q = 0; q < ntq; q++) {
      double t = t0 + tp[3 * q] * (t1 - t0);
      double sh[32], lt[4];
      xo_time_shapes(time_fe_order, t, sh);
      for (int k = 0; k < ns; k++) {
        double d = 0;
        for (int j = 0; j < nt; j++) d += lset_st[j * ns + k] * sh[j];
        if (fabs(d) < 1e-14) d = (d >= 0) ? 1e-14 : -1e-14; /* EPS_STCR_LSET_PERTUBATION */
        lt[k] = d;
      }
      int element_domain = xo_check_if_straight_cut(lt, ns, 0.0);
      int k = 0;
      if (element_domain == XO_IF) {
        k = xo_straight_cut_rule(T, lt, dt, order_space, pts + 3 * n, wts + n, cap - n, 0);
        if (k < 0) return k;
      } else if (element_domain == dt) {
        k = xo_rule_size(D, order_space);
        if (n + k > cap) return XO_ERR_CAPACITY;
        xo_select_rule(D, order_space, pts + 3 * n, wts + n);
      }
      for (int m = n; m < n + k; m++) { wts[m] = wts[m] * tw[q] * (t1 - t0); tvals[m] = t; }
      n += k;
    }
  }
  return n;
}
