/* xo_facet.c — ORACLE (test infrastructure only): cut integrals over interior facets and element boundaries.
 *
 * Restates, for affine simplices and one P1 level set,
 *   SymbolicCutFacetBilinearFormIntegrator::T_CalcFacetMatrix   xfem/symboliccutbfi.cpp:886-1290  (dCut(..., skeleton=True))
 *   SymbolicCutBilinearFormIntegrator::T_CalcElementMatrixEBAdd xfem/symboliccutbfi.cpp:488-883   (dCut(..., element_boundary=True))
 * : the cut rule of the FACET (a (dim-1)-simplex; level-set values at its vertices, StraightCutIntegrationRuleUntransformed
 * of order 2*max(order), :978 / :1029), carried into the reference coordinates of the neighbouring element(s) — both
 * neighbours see the same physical points (Facet2ElementTrafo on the globally sorted facet vertices) — weights times the
 * facet measure (:1129), traces of the shape functions of both elements, the unit normal nF of the FIRST element.
 * Closed integrand menu: sum of  coef * h^hpow * Op_test(v | side) * Op_trial(u | side),  Op in {value, nF . grad},
 * side in {0: this element, 1: .Other()}; h = mesh size of the first element.  The DG forms of demos/fictdom_dg.py:108-125
 * (jump, mean flux) and derivative-jump ghost penalties (xfem/ghostpenalty.cpp, first normal derivative) expand into it.
 */
#include "xoracle.h"
#include <math.h>
#include <string.h>

static double inv_det_(int D, const double* F, double* Fi)
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

/* facet_vnum: the dim global vertex numbers of the facet (the order the level-set values lset_facet are given in);
 * T2 == NULL: one-sided (element boundary), mat is nd x nd; else mat is (2 nd) x (2 nd), rows = test, [side 0 | side 1].
 * dt = XO_NEG / XO_POS; order < 0: 2 * fe_order.  ADDS into mat.  returns #points or an error code (< 0). */
int xo_cut_facet_elmat(const xo_trafo* T1, const xo_trafo* T2, int p, const int* facet_vnum, const double* lset_facet,
                       int dt, int order, const xo_fterm* terms, int nterms, double* mat)
{
  const int D = T1->dim, nd = xo_h1_ndof(D, p), ns = T2 ? 2 : 1, N = ns * nd;
  if (D < 2 || dt == XO_IF) return XO_ERR_ARG;
  if (order < 0) order = 2 * p;
  /* the facet as a (D-1)-simplex: vertices in the given order, reference convention vertex k <-> coordinate k,
   * last vertex <-> origin */
  xo_trafo TF;
  memset(&TF, 0, sizeof(TF));
  TF.dim = D - 1;
  int loc1[3];
  for (int i = 0; i < D; i++) {
    loc1[i] = -1;
    for (int k = 0; k < D + 1; k++) if (T1->vnum[k] == facet_vnum[i]) loc1[i] = k;
    if (loc1[i] < 0) return XO_ERR_ARG;
    for (int c = 0; c < 3; c++) TF.vtx[i][c] = T1->vtx[loc1[i]][c];
    TF.vnum[i] = facet_vnum[i];
  }
  static __thread double pts[3 * 4 * XO_MAXPTS_RULE], wts[4 * XO_MAXPTS_RULE];
  int is_std = 0;
  const int np = xo_straight_cut_rule(&TF, lset_facet, dt, order, pts, wts, 4 * XO_MAXPTS_RULE, &is_std);
  if (np <= 0) return np;
  /* facet measure relative to the reference facet: segment length / |cross product| */
  double a[3], b[3], JF;
  for (int c = 0; c < 3; c++) { a[c] = TF.vtx[0][c] - TF.vtx[D - 1][c]; b[c] = (D == 3) ? TF.vtx[1][c] - TF.vtx[D - 1][c] : 0.0; }
  if (D == 2) JF = sqrt(a[0] * a[0] + a[1] * a[1]);
  else {
    const double c0 = a[1] * b[2] - a[2] * b[1], c1 = a[2] * b[0] - a[0] * b[2], c2 = a[0] * b[1] - a[1] * b[0];
    JF = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
  }
  /* first element: Jacobian (affine), mesh size, outward unit normal on the facet = -grad(lam_o)/|grad(lam_o)|, o = the
   * vertex of T1 opposite to the facet */
  const xo_trafo* TT[2] = { T1, T2 };
  double Fi[2][9], nF[3] = { 0, 0, 0 }, h = 0;
  for (int s = 0; s < ns; s++) {
    double F[9], x[3], xh[3] = { 0, 0, 0 };
    xo_trafo_eval(TT[s], xh, F, x);
    const double det = inv_det_(D, F, Fi[s]);
    if (s == 0) h = pow(fabs(det), 1.0 / D);
  }
  {
    int o = -1;
    for (int k = 0; k < D + 1; k++) {
      int on = 0;
      for (int i = 0; i < D; i++) on |= (loc1[i] == k);
      if (!on) o = k;
    }
    double gr[3] = { 0, 0, 0 };
    for (int j = 0; j < D; j++) gr[j] = (o < D) ? (j == o ? 1.0 : 0.0) : -1.0;   /* reference gradient of lam_o */
    double nn = 0;
    for (int i = 0; i < D; i++) { double t = 0; for (int j = 0; j < D; j++) t += Fi[0][j * D + i] * gr[j]; nF[i] = -t; nn += t * t; }
    nn = sqrt(nn);
    for (int i = 0; i < D; i++) nF[i] /= nn;
  }
  double X[2][2][64];   /* [kind][side][dof]: 0 value, 1 normal derivative */
  for (int q = 0; q < np; q++) {
    double mu[3], sm = 1.0;
    for (int i = 0; i < D - 1; i++) { mu[i] = pts[3 * q + i]; sm -= mu[i]; }
    mu[D - 1] = sm;
    for (int s = 0; s < ns; s++) {
      const xo_trafo* T = TT[s];
      double xh[3] = { 0, 0, 0 };
      for (int k = 0; k < D; k++)
        for (int i = 0; i < D; i++) if (T->vnum[k] == facet_vnum[i]) xh[k] = mu[i];
      double sh[64], dsh[64 * 3];
      xo_h1_shapes(D, p, T->vnum, xh, sh, dsh);
      for (int aa = 0; aa < nd; aa++) {
        double dn = 0;
        for (int i = 0; i < D; i++) {
          double g = 0;
          for (int j = 0; j < D; j++) g += Fi[s][j * D + i] * dsh[aa * D + j];
          dn += g * nF[i];
        }
        X[0][s][aa] = sh[aa]; X[1][s][aa] = dn;
      }
    }
    const double fac = wts[q] * JF;
    for (int t = 0; t < nterms; t++) {
      const xo_fterm* tm = &terms[t];
      if (tm->trial_side >= ns || tm->test_side >= ns) return XO_ERR_ARG;
      double c = tm->coef * fac;
      if (tm->hpow != 0) c *= pow(h, (double)tm->hpow);
      for (int i = 0; i < nd; i++)
        for (int j = 0; j < nd; j++)
          mat[(tm->test_side * nd + i) * N + tm->trial_side * nd + j] += c * X[tm->test_kind][tm->test_side][i] * X[tm->trial_kind][tm->trial_side][j];
    }
  }
  return np;
}
