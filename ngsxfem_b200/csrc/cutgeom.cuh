// cutgeom.cuh — per-element straight-cut geometry on simplices (host + device).
// Semantics follow cutint/straightcutrule.cpp: CheckIfStraightCut :11-24, cut points :41-64
// (note the `>= 0` sign convention there vs. the strict one in CheckIfStraightCut),
// Decompose :131-196, GetVolume :66-72, the point map of GetPlainIntegrationRule :80-100 and
// the interface weight transform :558-578.
#pragma once
#include <math.h>
#include "h1.cuh"
#include "../../include/xgpu.h"

struct XgSimplex {
  int D;
  double p[4][3];
};

XG_HD int xg_check_cut(const double* v, int n)
{
  bool haspos = false, hasneg = false;
  for (int i = 0; i < n; i++) { haspos |= (v[i] > 0.0); hasneg |= (v[i] < 0.0); }
  if ((hasneg && haspos) || (!hasneg && !haspos)) return XGPU_IF;
  return hasneg ? XGPU_NEG : XGPU_POS;
}

XG_HD void xg_ref_simplex(int dim, XgSimplex& s)
{
  s.D = dim;
  for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) s.p[i][k] = 0.0;
  for (int i = 0; i < dim; i++) s.p[i][i] = 1.0;
}

XG_HD double xg_volume(const XgSimplex& s)
{
  if (s.D == 0) return 1.0;
  double a[3], b[3], c[3];
  if (s.D == 1) {
    for (int k = 0; k < 3; k++) a[k] = s.p[1][k] - s.p[0][k];
    return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  }
  if (s.D == 2) {
    for (int k = 0; k < 3; k++) { a[k] = s.p[2][k] - s.p[0][k]; b[k] = s.p[1][k] - s.p[0][k]; }
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
    return sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  }
  for (int k = 0; k < 3; k++) { a[k] = s.p[3][k] - s.p[0][k]; b[k] = s.p[2][k] - s.p[0][k]; c[k] = s.p[1][k] - s.p[0][k]; }
  double det = a[0] * (b[1] * c[2] - b[2] * c[1]) - b[0] * (a[1] * c[2] - a[2] * c[1]) + c[0] * (a[1] * b[2] - a[2] * b[1]);
  return fabs(det);
}

XG_HD void xg_set(XgSimplex& o, int D, const double* q0, const double* q1, const double* q2, const double* q3)
{
  o.D = D;
  for (int k = 0; k < 3; k++) {
    o.p[0][k] = q0[k];
    o.p[1][k] = (D >= 1) ? q1[k] : 0.0;
    o.p[2][k] = (D >= 2) ? q2[k] : 0.0;
    o.p[3][k] = (D >= 3) ? q3[k] : 0.0;
  }
}

// returns number of sub-simplices (1..3) or a negative XGPU_ERR_GEOM_* code
XG_HD int xg_decompose(const XgSimplex& s, const double* l, int dt, XgSimplex* out)
{
  const int D = s.D, np = D + 1;
  double c[6][3];
  int nc = 0;
  if (xg_check_cut(l, np) != XGPU_IF) return XGPU_ERR_GEOM_PLAIN;
  if (D == 1) {
    double t = l[0] / (l[0] - l[1]);
    for (int k = 0; k < 3; k++) c[0][k] = s.p[0][k] + t * (s.p[1][k] - s.p[0][k]);
    nc = 1;
  } else {
    for (int i = 0; i < np; i++)
      for (int j = i + 1; j < np; j++)
        if ((l[i] >= 0) != (l[j] >= 0)) {
          // the reference re-checks the vertex pair (:45-49): (0, negative) is "plain geometry"
          bool hp = (l[i] > 0) || (l[j] > 0), hn = (l[i] < 0) || (l[j] < 0);
          if (!(hp && hn)) return XGPU_ERR_GEOM_PLAIN;
          double t = l[i] / (l[i] - l[j]);
          for (int k = 0; k < 3; k++) c[nc][k] = s.p[i][k] + t * (s.p[j][k] - s.p[i][k]);
          nc++;
        }
  }
  if (dt == XGPU_IF) {
    if (nc == D) { xg_set(out[0], D - 1, c[0], c[1], c[2], c[3]); return 1; }
    if (nc == 4 && D == 3) {
      xg_set(out[0], 2, c[0], c[1], c[3], c[3]);
      xg_set(out[1], 2, c[0], c[2], c[3], c[3]);
      return 2;
    }
    return XGPU_ERR_GEOM_SCUT;
  }
  int rel[4], nr = 0;
  for (int i = 0; i < np; i++)
    if (((dt == XGPU_POS) && (l[i] >= 0)) || ((dt == XGPU_NEG) && (l[i] < 0))) rel[nr++] = i;
  if (nr == 1) {
    if (nc > 3) return XGPU_ERR_GEOM_NOTIMPL;
    const double* r0 = s.p[rel[0]];
    if (nc == 1) xg_set(out[0], 1, c[0], r0, r0, r0);
    else if (nc == 2) xg_set(out[0], 2, c[0], c[1], r0, r0);
    else xg_set(out[0], 3, c[0], c[1], c[2], r0);
    return 1;
  }
  if (nr == 2 && D == 2) {
    xg_set(out[0], 2, s.p[rel[0]], s.p[rel[1]], c[1], c[1]);
    xg_set(out[1], 2, c[0], c[1], s.p[rel[0]], c[1]);
    return 2;
  }
  if (nr == 2 && D == 3) {
    xg_set(out[0], 3, s.p[rel[1]], c[1], c[2], c[3]);
    xg_set(out[1], 3, s.p[rel[0]], s.p[rel[1]], c[1], c[2]);
    xg_set(out[2], 3, c[0], c[1], c[2], s.p[rel[0]]);
    return 3;
  }
  if (nr == 3 && D == 3) {
    xg_set(out[0], 3, c[0], c[1], c[2], s.p[rel[2]]);
    xg_set(out[1], 3, s.p[rel[0]], s.p[rel[1]], s.p[rel[2]], c[1]);
    xg_set(out[2], 3, c[0], c[1], s.p[rel[0]], s.p[rel[2]]);
    return 3;
  }
  return XGPU_ERR_GEOM_NOTIMPL;
}

// number of sub-simplices Decompose yields for a cut simplex of dimension D with `nneg`
// vertices having l < 0 (out of D+1) — no geometry needed (used by the count kernels)
XG_HD int xg_nsub(int D, int nneg, int dt)
{
  const int npos = D + 1 - nneg;  // vertices with l >= 0
  if (dt == XGPU_IF) return (D == 3 && nneg == 2) ? 2 : 1;
  const int nr = (dt == XGPU_NEG) ? nneg : npos;
  if (nr == 1) return 1;
  if (D == 2) return 2;
  return 3;
}

// map a reference-rule point into the sub-simplex (GetPlainIntegrationRule :80-100)
XG_HD void xg_map_point(const XgSimplex& s, const double* xi, double* x)
{
  double ow = 1.0;
  for (int m = 0; m < s.D; m++) ow -= xi[m];
  for (int k = 0; k < 3; k++) x[k] = ow * s.p[0][k];
  for (int m = 0; m < s.D; m++)
    for (int k = 0; k < 3; k++) x[k] += xi[m] * s.p[m + 1][k];
}

// affine Jacobian of a simplex element: F[i][j] = v_j[i] - v_D[i]  (row-major D x D)
template <int D>
XG_HD void xg_affine_jac(const double (*v)[3], double* F)
{
  for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) F[i * D + j] = v[j][i] - v[D][i];
}

template <int D>
XG_HD double xg_inv(const double* F, double* Fi)
{
  if (D == 1) { Fi[0] = 1.0 / F[0]; return F[0]; }
  if (D == 2) {
    double d = F[0] * F[3] - F[1] * F[2];
    Fi[0] = F[3] / d; Fi[1] = -F[1] / d; Fi[2] = -F[2] / d; Fi[3] = F[0] / d;
    return d;
  }
  double d = F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
  Fi[0] = (F[4] * F[8] - F[5] * F[7]) / d; Fi[1] = (F[2] * F[7] - F[1] * F[8]) / d; Fi[2] = (F[1] * F[5] - F[2] * F[4]) / d;
  Fi[3] = (F[5] * F[6] - F[3] * F[8]) / d; Fi[4] = (F[0] * F[8] - F[2] * F[6]) / d; Fi[5] = (F[2] * F[3] - F[0] * F[5]) / d;
  Fi[6] = (F[3] * F[7] - F[4] * F[6]) / d; Fi[7] = (F[1] * F[6] - F[0] * F[7]) / d; Fi[8] = (F[0] * F[4] - F[1] * F[3]) / d;
  return d;
}

// || F^{-T} n_ref ||, n_ref = normalised reference gradient of the P1 level set (:558-578)
template <int D>
XG_HD double xg_if_weight_factor(const double* Fi, const double* l)
{
  double g[3] = {0, 0, 0}, nn = 0;
  for (int i = 0; i < D; i++) { g[i] = l[i] - l[D]; nn += g[i] * g[i]; }
  nn = sqrt(nn);
  for (int i = 0; i < D; i++) g[i] /= nn;
  double s = 0;
  for (int i = 0; i < D; i++) {
    double t = 0;
    for (int j = 0; j < D; j++) t += Fi[j * D + i] * g[j];
    s += t * t;
  }
  return sqrt(s);
}
