/* xo_shift.c — ORACLE (test infrastructure only, see xoracle.h): the isoparametric deformation producer.
 *
 * Restates  lsetcurving/projshift.cpp:9-201 (ProjectShift), lsetcurving/shiftintegrators.cpp:39-124
 * (ShiftIntegrator::CalcElementVector), lsetcurving/calcpointshift.cpp:81-141 (ElementInRelevantBand), :145-206
 * (SearchCorrespondingPoint) and the element-wise L2 projection + averaging NGSolve's GridFunction.Set performs
 * (python/lsetcurv.py:309-310: lset_ho.Set(levelset), qn.Set(lset_ho.Deriv())).
 *
 * UNPINNED third-party arithmetic: GridFunction.Set lives in NGSolve (not in /root/reference); it is restated here as
 * "per element M^{-1} int f phi_i with the rule of order 2p + given bonus, then the arithmetic mean over the elements
 * sharing a dof" — the published behaviour for H1 spaces.  With qn == NULL the reference differentiates lset_p1 by a
 * central difference with eps = 1e-7 (calcpointshift.cpp:24-41); the gradient of a P1 function is constant, so the exact
 * value is used (difference: rounding of the finite difference, ~1e-9 relative in a search DIRECTION only).
 * Pinned by the reference's own known answer: tests/pytests/test_intcurved.py:17-91 (errors and convergence orders of
 * cut integrals on the deformed mesh), restated in tests/test_oracle_lsetcurv.py. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "xoracle.h"

#define XS_MAXND 35
#define XS_MAXNP 128

/* calcpointshift.cpp:120-141 */
int xo_element_in_band(const double* vals, int n, double lower, double upper)
{
  int has_neg = 0, has_pos = 0;
  for (int s = 0; s < n; s++) {
    const double v = vals[s];
    if (v == 0.0) has_neg = has_pos = 1;
    if (v > lower) has_pos = 1;
    if (v < upper) has_neg = 1;
  }
  return has_neg && has_pos;
}

static void inv_small(int D, const double* F, double* Fi, double* det)
{
  if (D == 1) { *det = F[0]; Fi[0] = 1.0 / F[0]; return; }
  if (D == 2) {
    const double d = F[0] * F[3] - F[1] * F[2];
    *det = d;
    Fi[0] = F[3] / d; Fi[1] = -F[1] / d; Fi[2] = -F[2] / d; Fi[3] = F[0] / d;
    return;
  }
  const double d = F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
  *det = d;
  Fi[0] = (F[4] * F[8] - F[5] * F[7]) / d; Fi[1] = (F[2] * F[7] - F[1] * F[8]) / d; Fi[2] = (F[1] * F[5] - F[2] * F[4]) / d;
  Fi[3] = (F[5] * F[6] - F[3] * F[8]) / d; Fi[4] = (F[0] * F[8] - F[2] * F[6]) / d; Fi[5] = (F[2] * F[3] - F[0] * F[5]) / d;
  Fi[6] = (F[3] * F[7] - F[4] * F[6]) / d; Fi[7] = (F[1] * F[6] - F[0] * F[7]) / d; Fi[8] = (F[0] * F[4] - F[1] * F[3]) / d;
}

static void make_trafo(int dim, const int32_t* ev, const double* vtx_xyz, xo_trafo* T)
{
  memset(T, 0, sizeof(*T));
  T->dim = dim;
  for (int k = 0; k < dim + 1; k++) {
    T->vnum[k] = ev[k];
    for (int c = 0; c < dim; c++) T->vtx[k][c] = vtx_xyz[(int64_t)ev[k] * dim + c];
  }
  T->def_order = 0; T->def = NULL;
}

/* solve (sum_q w_q phi phi^T) X = B for the nd x ncomp right-hand sides (Cholesky); the |det F| of an affine element
 * cancels between the mass matrix and the right-hand side */
static void local_project(int dim, int fe_order, const int* vnum, int np, const double* pts, const double* wts, int ncomp,
                          const double* fvals /* np x ncomp */, double* X /* nd x ncomp */)
{
  const int nd = xo_h1_ndof(dim, fe_order);
  double M[XS_MAXND * XS_MAXND], shape[XS_MAXND], dshape[XS_MAXND * 3];
  memset(M, 0, sizeof(double) * nd * nd);
  for (int i = 0; i < nd * ncomp; i++) X[i] = 0.0;
  for (int q = 0; q < np; q++) {
    xo_h1_shapes(dim, fe_order, vnum, pts + 3 * q, shape, dshape);
    for (int i = 0; i < nd; i++) {
      for (int j = 0; j < nd; j++) M[i * nd + j] += wts[q] * shape[i] * shape[j];
      for (int c = 0; c < ncomp; c++) X[i * ncomp + c] += wts[q] * shape[i] * fvals[q * ncomp + c];
    }
  }
  /* Cholesky M = L L^T in place (lower) */
  for (int j = 0; j < nd; j++) {
    double s = M[j * nd + j];
    for (int k = 0; k < j; k++) s -= M[j * nd + k] * M[j * nd + k];
    const double l = sqrt(s);
    M[j * nd + j] = l;
    for (int i = j + 1; i < nd; i++) {
      double t = M[i * nd + j];
      for (int k = 0; k < j; k++) t -= M[i * nd + k] * M[j * nd + k];
      M[i * nd + j] = t / l;
    }
  }
  for (int c = 0; c < ncomp; c++) {
    for (int i = 0; i < nd; i++) {
      double t = X[i * ncomp + c];
      for (int k = 0; k < i; k++) t -= M[i * nd + k] * X[k * ncomp + c];
      X[i * ncomp + c] = t / M[i * nd + i];
    }
    for (int i = nd - 1; i >= 0; i--) {
      double t = X[i * ncomp + c];
      for (int k = i + 1; k < nd; k++) t -= M[k * nd + i] * X[k * ncomp + c];
      X[i * ncomp + c] = t / M[i * nd + i];
    }
  }
}

/* physical coordinates of the rule points of order `order_rule` on every element: out ne x np x dim */
int xo_rule_points_phys(int dim, int order_rule, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz, double* out)
{
  double pts[3 * XS_MAXNP], wts[XS_MAXNP];
  const int np = xo_select_rule(dim, order_rule, pts, wts);
  for (int64_t e = 0; e < ne; e++) {
    xo_trafo T;
    make_trafo(dim, el_vtx + e * (dim + 1), vtx_xyz, &T);
    for (int q = 0; q < np; q++) {
      double F[9], x[3];
      xo_trafo_eval(&T, pts + 3 * q, F, x);
      for (int c = 0; c < dim; c++) out[((int64_t)e * np + q) * dim + c] = x[c];
    }
  }
  return np;
}

/* GridFunction.Set of point values given at the rule points of every element: element-wise L2 projection, then the mean
 * over the elements sharing a dof.  fvals: ne x np x ncomp, out: ndof x ncomp */
void xo_set_from_points(int dim, int fe_order, int order_rule, int64_t ne, const int32_t* el_vtx, const int32_t* el_dofs,
                        int64_t ndof, int ncomp, const double* fvals, double* out)
{
  const int nd = xo_h1_ndof(dim, fe_order);
  double pts[3 * XS_MAXNP], wts[XS_MAXNP];
  const int np = xo_select_rule(dim, order_rule, pts, wts);
  double* cnt = (double*)calloc((size_t)ndof, sizeof(double));
  memset(out, 0, sizeof(double) * (size_t)ndof * ncomp);
  for (int64_t e = 0; e < ne; e++) {
    int vnum[4];
    for (int k = 0; k < dim + 1; k++) vnum[k] = el_vtx[e * (dim + 1) + k];
    double X[XS_MAXND * 3];
    local_project(dim, fe_order, vnum, np, pts, wts, ncomp, fvals + (int64_t)e * np * ncomp, X);
    for (int i = 0; i < nd; i++) {
      const int32_t d = el_dofs[e * nd + i];
      if (d < 0) continue;
      for (int c = 0; c < ncomp; c++) out[(int64_t)d * ncomp + c] += X[i * ncomp + c];
      cnt[d] += 1.0;
    }
  }
  for (int64_t d = 0; d < ndof; d++)
    if (cnt[d] > 0) for (int c = 0; c < ncomp; c++) out[d * ncomp + c] *= 1.0 / cnt[d];
  free(cnt);
}

static void eval_ho(int dim, int fe_order, const int* vnum, const double* coefs, const double* xhat, double* val, double* gref)
{
  const int nd = xo_h1_ndof(dim, fe_order);
  double shape[XS_MAXND], dshape[XS_MAXND * 3];
  xo_h1_shapes(dim, fe_order, vnum, xhat, shape, dshape);
  double v = 0.0, g[3] = {0, 0, 0};
  for (int i = 0; i < nd; i++) {
    v += coefs[i] * shape[i];
    for (int c = 0; c < dim; c++) g[c] += coefs[i] * dshape[i * dim + c];
  }
  *val = v;
  for (int c = 0; c < dim; c++) gref[c] = g[c];
}

/* qn.Set(grad lset_ho) on a continuous H1(order)^dim space: physical gradient at the rule points, projection, mean */
void xo_set_gradient(int dim, int fe_order, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz, const int32_t* el_dofs,
                     int64_t ndof, const double* lset_ho, double* qn_out)
{
  const int nd = xo_h1_ndof(dim, fe_order);
  double pts[3 * XS_MAXNP], wts[XS_MAXNP];
  const int order_rule = 2 * fe_order;
  const int np = xo_select_rule(dim, order_rule, pts, wts);
  double* fv = (double*)malloc(sizeof(double) * (size_t)ne * np * dim);
  for (int64_t e = 0; e < ne; e++) {
    xo_trafo T;
    make_trafo(dim, el_vtx + e * (dim + 1), vtx_xyz, &T);
    double coefs[XS_MAXND];
    for (int i = 0; i < nd; i++) coefs[i] = lset_ho[el_dofs[e * nd + i]];
    for (int q = 0; q < np; q++) {
      double F[9], Fi[9], x[3], det, v, gr[3];
      xo_trafo_eval(&T, pts + 3 * q, F, x);
      inv_small(dim, F, Fi, &det);
      eval_ho(dim, fe_order, T.vnum, coefs, pts + 3 * q, &v, gr);
      for (int c = 0; c < dim; c++) {   /* F^{-T} grad_ref */
        double s = 0.0;
        for (int j = 0; j < dim; j++) s += Fi[j * dim + c] * gr[j];
        fv[((int64_t)e * np + q) * dim + c] = s;
      }
    }
  }
  xo_set_from_points(dim, fe_order, order_rule, ne, el_vtx, el_dofs, ndof, dim, fv, qn_out);
  free(fv);
}

/* ProjectShift.  qn_mode: 0 = grad lset_p1 (qn == nullptr in the reference), 1 = grad lset_ho evaluated at the point (an L2
 * qn field of the same order reproduces it exactly), 2 = the H1(order)^dim coefficient vector `qn` (ndof x dim).
 * ba: optional element marker bytes (bit i = byte i>>3, bit i&7), blending: optional ne x np values of the blending
 * function at the rule points (order 2 fe_order).  deform: ndof x dim out.  returns the number of searches that did not
 * converge in 20 steps (the reference prints a message and keeps the initial point). */
int64_t xo_project_shift(int dim, int fe_order, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz,
                         const int32_t* el_dofs, int64_t ndof, const double* lset_ho, const double* lset_p1, int qn_mode,
                         const double* qn, const uint8_t* ba, const double* blending, double lower, double upper,
                         double threshold, double* deform)
{
  const int nd = xo_h1_ndof(dim, fe_order);
  double pts[3 * XS_MAXNP], wts[XS_MAXNP];
  const int np = xo_select_rule(dim, 2 * fe_order, pts, wts);   /* shiftintegrators.cpp:69 */
  double* cnt = (double*)calloc((size_t)ndof, sizeof(double));
  memset(deform, 0, sizeof(double) * (size_t)ndof * dim);
  int64_t nfail = 0;
  for (int64_t e = 0; e < ne; e++) {
    const int32_t* ev = el_vtx + e * (dim + 1);
    double lp1[4];
    for (int k = 0; k < dim + 1; k++) lp1[k] = lset_p1[ev[k]];
    if (ba) { if (!((ba[e >> 3] >> (e & 7)) & 1)) continue; }               /* projshift.cpp:73-74 */
    else if (!xo_element_in_band(lp1, dim + 1, lower, upper)) continue;        /* :75-76 */
    xo_trafo T;
    make_trafo(dim, ev, vtx_xyz, &T);
    double coefs[XS_MAXND], qc[XS_MAXND * 3];
    for (int i = 0; i < nd; i++) {
      const int32_t d = el_dofs[e * nd + i];
      coefs[i] = lset_ho[d];
      if (qn_mode == 2) for (int c = 0; c < dim; c++) qc[i * dim + c] = qn[(int64_t)d * dim + c];
    }
    double fv[XS_MAXNP * 3];
    for (int q = 0; q < np; q++) {
      const double* xh = pts + 3 * q;
      double F[9], Fi[9], x[3], det;
      xo_trafo_eval(&T, xh, F, x);
      inv_small(dim, F, Fi, &det);
      /* search direction (physical vector `grad`, then normal = F^{-1} grad, shiftintegrators.cpp:77-80) */
      double grad[3] = {0, 0, 0}, v0, g0[3];
      eval_ho(dim, fe_order, T.vnum, coefs, xh, &v0, g0);
      if (qn_mode == 0) {
        double gr[3];
        for (int j = 0; j < dim; j++) gr[j] = lp1[j] - lp1[dim];   /* reference gradient of the P1 function */
        for (int c = 0; c < dim; c++) { double s = 0; for (int j = 0; j < dim; j++) s += Fi[j * dim + c] * gr[j]; grad[c] = s; }
      } else if (qn_mode == 1) {
        for (int c = 0; c < dim; c++) { double s = 0; for (int j = 0; j < dim; j++) s += Fi[j * dim + c] * g0[j]; grad[c] = s; }
      } else {
        double shape[XS_MAXND], dshape[XS_MAXND * 3];
        xo_h1_shapes(dim, fe_order, T.vnum, xh, shape, dshape);
        for (int i = 0; i < nd; i++) for (int c = 0; c < dim; c++) grad[c] += qc[i * dim + c] * shape[i];
      }
      double dir[3] = {0, 0, 0};
      for (int i = 0; i < dim; i++) { double s = 0; for (int j = 0; j < dim; j++) s += Fi[i * dim + j] * grad[j]; dir[i] = s; }
      /* goal value (:84-98) */
      double lam_last = 1.0, p1v = 0.0;
      for (int j = 0; j < dim; j++) { p1v += xh[j] * lp1[j]; lam_last -= xh[j]; }
      p1v += lam_last * lp1[dim];
      const double alpha = blending ? blending[(int64_t)e * np + q] : 0.0;
      const double goal = (1.0 - alpha) * p1v + alpha * v0;
      /* SearchCorrespondingPoint, dynamic_search_dir = false (calcpointshift.cpp:145-206) */
      double cur[3] = {xh[0], xh[1], xh[2]};
      int it;
      for (it = 0; it < 20; it++) {
        double cv, cg[3];
        eval_ho(dim, fe_order, T.vnum, coefs, cur, &cv, cg);
        const double defect = goal - cv;
        if (fabs(defect) < 1e-14) break;
        double dphidn = 0.0;
        for (int c = 0; c < dim; c++) dphidn += cg[c] * dir[c];
        for (int c = 0; c < dim; c++) cur[c] += defect / dphidn * dir[c];
      }
      double rd[3] = {0, 0, 0};
      if (it == 20) nfail++;
      else for (int c = 0; c < dim; c++) rd[c] = cur[c] - xh[c];
      double len = 0.0;
      for (int c = 0; c < dim; c++) len += rd[c] * rd[c];
      len = sqrt(len);
      if (threshold >= 0.0 && len > threshold) for (int c = 0; c < dim; c++) rd[c] *= threshold / len;   /* :108-111 */
      for (int i = 0; i < dim; i++) { double s = 0; for (int j = 0; j < dim; j++) s += F[i * dim + j] * rd[j]; fv[q * dim + i] = s; }
    }
    double X[XS_MAXND * 3];
    local_project(dim, fe_order, T.vnum, np, pts, wts, dim, fv, X);
    for (int l = 0; l < dim + 1; l++) for (int c = 0; c < dim; c++) X[l * dim + c] = 0.0;   /* vertex values to zero (projshift.cpp:108-109) */
    for (int i = 0; i < nd; i++) {
      const int32_t d = el_dofs[e * nd + i];
      for (int c = 0; c < dim; c++) deform[(int64_t)d * dim + c] += X[i * dim + c];
      cnt[d] += 1.0;
    }
  }
  for (int64_t d = 0; d < ndof; d++)
    if (cnt[d] > 0) for (int c = 0; c < dim; c++) deform[d * dim + c] *= 1.0 / cnt[d];   /* :165-197 */
  free(cnt);
  return nfail;
}

/* cut integral of 1 over the (deformed) mesh: sum_q w_q |det F(xhat_q)| with the rules of xo_straight_cut_rule
 * (cutint/cutintegral.cpp:126-212 with mesh.SetDeformation): def_vals ndof x dim in the H1(def_order) space given by
 * el_dofs_def (def_order = 0: affine).  returns the integral, *err = first rule error */
double xo_integrate_one_deformed(int dim, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz, const double* phi, int dt,
                                 int order, int def_order, const int32_t* el_dofs_def, const double* def_vals, int* err)
{
  const int ndd = def_order > 0 ? xo_h1_ndof(dim, def_order) : 0;
  double tot = 0.0;
  *err = 0;
  enum { CAP = 4096 };
  double* pts = (double*)malloc(sizeof(double) * 3 * CAP);
  double* wts = (double*)malloc(sizeof(double) * CAP);
  for (int64_t e = 0; e < ne; e++) {
    const int32_t* ev = el_vtx + e * (dim + 1);
    xo_trafo T;
    make_trafo(dim, ev, vtx_xyz, &T);
    double loc[XS_MAXND * 3], l[4];
    if (def_order > 0) {
      for (int i = 0; i < ndd; i++)
        for (int c = 0; c < dim; c++) loc[i * dim + c] = def_vals[(int64_t)el_dofs_def[e * ndd + i] * dim + c];
      T.def_order = def_order; T.def = loc;
    }
    for (int k = 0; k < dim + 1; k++) l[k] = phi[ev[k]];
    int is_std = 0;
    const int n = xo_straight_cut_rule(&T, l, dt, order, pts, wts, CAP, &is_std);
    if (n < 0) { if (!*err) *err = n; continue; }
    for (int q = 0; q < n; q++) {
      double F[9], Fi[9], x[3], det;
      xo_trafo_eval(&T, pts + 3 * q, F, x);
      inv_small(dim, F, Fi, &det);
      tot += wts[q] * fabs(det);
    }
  }
  free(pts); free(wts);
  return tot;
}
