/* xo_rules.c — ORACLE (test infrastructure): reference quadrature rules.
 *
 * Restates the *published algorithm* of NGSolve's SelectIntegrationRule (third party, pip
 * `ngsolve>=6.2.2606`, source not vendored in the reference; call sites:
 * cutint/straightcutrule.cpp:87-90,111,118,263,673,875; cutint/mlsetintegration.cpp:91;
 * cutint/spacetimecutrule.cpp:147-149; xfem/symboliccutbfi.cpp:2009-2012):
 *   SEGM : Gauss-Legendre on [0,1], n = order/2+1
 *   TRIG : hard-coded symmetric rules up to order 6 (1/3/6/12 points), else Duffy
 *          Gauss-Jacobi(1,0) x Gauss-Legendre with n = order/2+1 per direction
 *   TET  : hard-coded symmetric rules up to order 5 (1/4/14 points), else Duffy
 *          Gauss-Jacobi(2,0) x Gauss-Jacobi(1,0) x Gauss-Legendre
 * PARITY UNPINNED for the point ORDER inside a rule (no NGSolve here); every rule is checked
 * for exactness on all monomials up to its degree (tests/test_oracle_rules.py).
 */
#include "xoracle.h"
#include <math.h>
#include <string.h>
#include "xo_rule_tables.inc"

/* Jacobi polynomial P_n^{(a,0)}(x) and derivative by the three-term recurrence.
 * Extended precision: the weight formula 1/((1-x^2) P'^2) amplifies root errors near the
 * end points, and rule weights enter every parity comparison at 1e-13. */
typedef long double xo_ld;
static void jacobi_a0(int n, xo_ld a, xo_ld x, xo_ld* P, xo_ld* dP)
{
  xo_ld p0 = 1.0L, p1 = 0.5L * (a + (a + 2.0L) * x);
  xo_ld d0 = 0.0L, d1 = 0.5L * (a + 2.0L);
  if (n == 0) { *P = p0; *dP = d0; return; }
  for (int k = 2; k <= n; k++) {
    xo_ld kk = k, b = 0.0L;
    xo_ld c1 = 2.0L * kk * (kk + a + b) * (2.0L * kk + a + b - 2.0L);
    xo_ld c2 = (2.0L * kk + a + b - 1.0L) * (a * a - b * b);
    xo_ld c3 = (2.0L * kk + a + b - 1.0L) * (2.0L * kk + a + b) * (2.0L * kk + a + b - 2.0L);
    xo_ld c4 = 2.0L * (kk + a - 1.0L) * (kk + b - 1.0L) * (2.0L * kk + a + b);
    xo_ld p2 = ((c2 + c3 * x) * p1 - c4 * p0) / c1;
    xo_ld d2 = ((c2 + c3 * x) * d1 + c3 * p1 - c4 * d0) / c1;
    p0 = p1; p1 = p2; d0 = d1; d1 = d2;
  }
  *P = p1; *dP = d1;
}

/* roots of P_n^{(a,0)} by interlacing brackets + safeguarded Newton */
static void jacobi_roots(int n, xo_ld a, xo_ld* r)
{
  xo_ld prev[64], cur[64];
  int np = 0;
  for (int m = 1; m <= n; m++) {
    for (int i = 0; i < m; i++) {
      xo_ld lo = (i == 0) ? -1.0L : prev[i - 1];
      xo_ld hi = (i == m - 1) ? 1.0L : prev[i];
      xo_ld flo, fhi, d;
      jacobi_a0(m, a, lo, &flo, &d);
      jacobi_a0(m, a, hi, &fhi, &d);
      xo_ld x = 0.5L * (lo + hi);
      for (int it = 0; it < 200; it++) {
        xo_ld f, df;
        jacobi_a0(m, a, x, &f, &df);
        if (f == 0.0L) break;
        if ((f < 0) == (flo < 0)) { lo = x; flo = f; } else { hi = x; fhi = f; }
        xo_ld xn = x - f / df;
        if (!(xn > lo && xn < hi)) xn = 0.5L * (lo + hi);
        if (fabsl(xn - x) < 1e-20L * (1.0L + fabsl(x))) { x = xn; break; }
        x = xn;
      }
      /* two polishing Newton steps */
      for (int it = 0; it < 2; it++) { xo_ld f, df; jacobi_a0(m, a, x, &f, &df); x -= f / df; }
      cur[i] = x;
    }
    np = m;
    memcpy(prev, cur, sizeof(xo_ld) * np);
  }
  memcpy(r, prev, sizeof(xo_ld) * n);
}

void xo_gauss_jacobi01(int n, int alpha, double* x, double* w)
{
  xo_ld r[64];
  xo_ld a = (xo_ld)alpha;
  jacobi_roots(n, a, r);
  for (int i = 0; i < n; i++) {
    xo_ld P, dP;
    jacobi_a0(n, a, r[i], &P, &dP);
    /* weight on [-1,1] for (1-x)^a: 2^{a+1}/((1-x^2) P'^2) (b = 0, integer a); mapped to
     * [0,1] with weight (1-x)^a the factor 2^{a+1} cancels */
    x[i] = (double)(0.5L * (1.0L + r[i]));
    w[i] = (double)(1.0L / ((1.0L - r[i] * r[i]) * dP * dP));
  }
}

static int npts1d(int order) { if (order < 0) order = 0; return order / 2 + 1; }

int xo_rule_size(int dim, int order)
{
  if (order < 0) order = 0;
  int n = npts1d(order);
  switch (dim) {
    case 0: return 1;
    case 1: return n;
    case 2:
      if (order <= 1) return 1;
      if (order == 2) return 3;
      if (order <= 4) return 6;
      if (order <= 6) return 12;
      return n * n;
    case 3:
      if (order <= 1) return 1;
      if (order == 2) return 4;
      if (order <= 5) return 14;
      return n * n * n;
  }
  return 0;
}

static int copy_sym(const double (*P)[3], const double* W, int n, double* pts, double* wts)
{
  for (int i = 0; i < n; i++) {
    pts[3 * i] = P[i][0]; pts[3 * i + 1] = P[i][1]; pts[3 * i + 2] = P[i][2];
    wts[i] = W[i];
  }
  return n;
}

int xo_select_rule(int dim, int order, double* pts, double* wts)
{
  if (order < 0) order = 0;
  int n = npts1d(order);
  double x0[64], w0[64], x1[64], w1[64], x2[64], w2[64];
  if (dim == 0) { pts[0] = pts[1] = pts[2] = 0.0; wts[0] = 1.0; return 1; }
  if (dim == 1) {
    xo_gauss_jacobi01(n, 0, x0, w0);
    for (int i = 0; i < n; i++) { pts[3 * i] = x0[i]; pts[3 * i + 1] = 0; pts[3 * i + 2] = 0; wts[i] = w0[i]; }
    return n;
  }
  if (dim == 2) {
    if (order <= 1) return copy_sym(SYM_TRIG1_P, SYM_TRIG1_W, SYM_TRIG1_N, pts, wts);
    if (order == 2) return copy_sym(SYM_TRIG3_P, SYM_TRIG3_W, SYM_TRIG3_N, pts, wts);
    if (order <= 4) return copy_sym(SYM_TRIG6_P, SYM_TRIG6_W, SYM_TRIG6_N, pts, wts);
    if (order <= 6) return copy_sym(SYM_TRIG12_P, SYM_TRIG12_W, SYM_TRIG12_N, pts, wts);
    xo_gauss_jacobi01(n, 1, x0, w0);
    xo_gauss_jacobi01(n, 0, x1, w1);
    int k = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++, k++) {
        pts[3 * k] = x0[i]; pts[3 * k + 1] = x1[j] * (1.0 - x0[i]); pts[3 * k + 2] = 0.0;
        wts[k] = w0[i] * w1[j];
      }
    return k;
  }
  if (dim == 3) {
    if (order <= 1) return copy_sym(SYM_TET1_P, SYM_TET1_W, SYM_TET1_N, pts, wts);
    if (order == 2) return copy_sym(SYM_TET4_P, SYM_TET4_W, SYM_TET4_N, pts, wts);
    if (order <= 5) return copy_sym(SYM_TET14_P, SYM_TET14_W, SYM_TET14_N, pts, wts);
    xo_gauss_jacobi01(n, 2, x0, w0);
    xo_gauss_jacobi01(n, 1, x1, w1);
    xo_gauss_jacobi01(n, 0, x2, w2);
    int k = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
        for (int l = 0; l < n; l++, k++) {
          pts[3 * k] = x0[i];
          pts[3 * k + 1] = x1[j] * (1.0 - x0[i]);
          pts[3 * k + 2] = x2[l] * (1.0 - x0[i]) * (1.0 - x1[j]);
          wts[k] = w0[i] * w1[j] * w2[l];
        }
    return k;
  }
  return 0;
}
