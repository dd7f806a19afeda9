/* xo_fe.c — ORACLE (test infrastructure): H1 shapes, element transformation, cut element
 * matrices and facet-patch matrices.
 *
 * Restates xfem/symboliccutbfi.cpp:76-486 (T_CalcElementMatrixAdd) and :1747-1846, :1982-2377
 * (facet patch).  The shape functions / element transformation live in NGSolve (third party,
 * not vendored): hierarchical H1HighOrderFE restated from its published construction
 * (vertex hats; edge functions = scaled integrated Legendre in (lam_e - lam_s) oriented from the
 * smaller to the larger global vertex number; face/cell bubbles x scaled Legendre products) —
 * PARITY UNPINNED at entry level, basis independent invariants are tested instead.
 */
#include "xoracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

typedef struct { double v, d[3]; } ad;
static ad ad_c(double c) { ad r = { c, { 0, 0, 0 } }; return r; }
static ad ad_add(ad a, ad b) { ad r = { a.v + b.v, { a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2] } }; return r; }
static ad ad_sub(ad a, ad b) { ad r = { a.v - b.v, { a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2] } }; return r; }
static ad ad_mul(ad a, ad b)
{
  ad r = { a.v * b.v, { a.v * b.d[0] + a.d[0] * b.v, a.v * b.d[1] + a.d[1] * b.v, a.v * b.d[2] + a.d[2] * b.v } };
  return r;
}
static ad ad_scale(double s, ad a) { ad r = { s * a.v, { s * a.d[0], s * a.d[1], s * a.d[2] } }; return r; }

/* scaled Legendre P_k^S(x,t) = t^k P_k(x/t), k = 0..n */
static void leg_scaled(int n, ad x, ad t, ad* P)
{
  P[0] = ad_c(1.0);
  if (n >= 1) P[1] = x;
  ad tt = ad_mul(t, t);
  for (int k = 1; k < n; k++)
    P[k + 1] = ad_scale(1.0 / (k + 1), ad_sub(ad_scale(2.0 * k + 1, ad_mul(x, P[k])), ad_scale((double)k, ad_mul(tt, P[k - 1]))));
}
/* scaled integrated Legendre L_k^S(x,t), k = 2..n stored at L[k] (L[1] = x) */
static void intleg_scaled(int n, ad x, ad t, ad* L)
{
  ad tt = ad_mul(t, t);
  L[1] = x;
  if (n >= 2) L[2] = ad_scale(0.5, ad_sub(ad_mul(x, x), tt));
  for (int k = 2; k < n; k++)
    L[k + 1] = ad_scale(1.0 / (k + 1), ad_sub(ad_scale(2.0 * k - 1, ad_mul(x, L[k])), ad_scale((double)(k - 2), ad_mul(tt, L[k - 1]))));
}

/* NGSolve ElementTopology (published tables, third party) */
static const int TET_EDGES[6][2] = { { 3, 0 }, { 3, 1 }, { 3, 2 }, { 0, 1 }, { 0, 2 }, { 1, 2 } };
static const int TET_FACES[4][3] = { { 3, 1, 2 }, { 3, 2, 0 }, { 3, 0, 1 }, { 0, 2, 1 } };
static const int TRIG_EDGES[3][2] = { { 2, 0 }, { 1, 2 }, { 0, 1 } };

int xo_h1_ndof(int dim, int p)
{
  if (dim == 1) return p + 1;
  if (dim == 2) return 3 + 3 * (p - 1) + (p - 1) * (p - 2) / 2;
  return 4 + 6 * (p - 1) + 4 * ((p - 1) * (p - 2) / 2) + (p - 1) * (p - 2) * (p - 3) / 6;
}

static void sort3(int* f, const int* vnum)
{
  for (int a = 0; a < 3; a++)
    for (int b = a + 1; b < 3; b++)
      if (vnum[f[a]] > vnum[f[b]]) { int t = f[a]; f[a] = f[b]; f[b] = t; }
}

static int face_shapes(int p, const int* fin, const int* vnum, const ad* lam, ad* out)
{
  int f[3] = { fin[0], fin[1], fin[2] }, n = 0;
  ad P[16], Q[16];
  if (p < 3) return 0;
  sort3(f, vnum);
  ad bub = ad_mul(ad_mul(lam[f[0]], lam[f[1]]), lam[f[2]]);
  leg_scaled(p - 3, ad_sub(lam[f[1]], lam[f[0]]), ad_add(lam[f[0]], lam[f[1]]), P);
  leg_scaled(p - 3, ad_sub(ad_scale(2.0, lam[f[2]]), ad_c(1.0)), ad_c(1.0), Q);
  for (int i = 0; i <= p - 3; i++)
    for (int j = 0; j <= p - 3 - i; j++)
      out[n++] = ad_mul(bub, ad_mul(P[i], Q[j]));
  return n;
}

void xo_h1_shapes(int dim, int p, const int* vnum, const double* xhat, double* shape, double* dshape)
{
  ad lam[4], sh[64], L[16];
  int n = 0, nv = dim + 1;
  double s = 1.0;
  for (int i = 0; i < dim; i++) { lam[i] = ad_c(xhat[i]); lam[i].d[i] = 1.0; s -= xhat[i]; }
  lam[dim] = ad_c(s);
  for (int i = 0; i < dim; i++) lam[dim].d[i] = -1.0;
  for (int i = 0; i < nv; i++) sh[n++] = lam[i];
  if (dim == 1) {
    int es = 1, ee = 0; /* segment as one "edge" {1,0} */
    if (vnum[es] > vnum[ee]) { int t = es; es = ee; ee = t; }
    if (p >= 2) {
      intleg_scaled(p, ad_sub(lam[ee], lam[es]), ad_add(lam[ee], lam[es]), L);
      for (int k = 2; k <= p; k++) sh[n++] = L[k];
    }
  } else {
    int ned = (dim == 2) ? 3 : 6;
    for (int e = 0; e < ned; e++) {
      int es = (dim == 2) ? TRIG_EDGES[e][0] : TET_EDGES[e][0];
      int ee = (dim == 2) ? TRIG_EDGES[e][1] : TET_EDGES[e][1];
      if (vnum[es] > vnum[ee]) { int t = es; es = ee; ee = t; }
      if (p >= 2) {
        intleg_scaled(p, ad_sub(lam[ee], lam[es]), ad_add(lam[ee], lam[es]), L);
        for (int k = 2; k <= p; k++) sh[n++] = L[k];
      }
    }
    if (dim == 2) {
      int f[3] = { 0, 1, 2 };
      n += face_shapes(p, f, vnum, lam, sh + n);
    } else {
      for (int fc = 0; fc < 4; fc++) n += face_shapes(p, TET_FACES[fc], vnum, lam, sh + n);
      if (p >= 4) {
        ad bub = ad_mul(ad_mul(lam[0], lam[1]), ad_mul(lam[2], lam[3]));
        ad P[16], Q[16], R[16];
        leg_scaled(p - 4, ad_sub(lam[1], lam[0]), ad_add(lam[0], lam[1]), P);
        leg_scaled(p - 4, ad_sub(ad_scale(2.0, lam[2]), ad_c(1.0)), ad_c(1.0), Q);
        leg_scaled(p - 4, ad_sub(ad_scale(2.0, lam[3]), ad_c(1.0)), ad_c(1.0), R);
        for (int i = 0; i <= p - 4; i++)
          for (int j = 0; j <= p - 4 - i; j++)
            for (int k = 0; k <= p - 4 - i - j; k++)
              sh[n++] = ad_mul(bub, ad_mul(P[i], ad_mul(Q[j], R[k])));
      }
    }
  }
  for (int i = 0; i < n; i++) {
    if (shape) shape[i] = sh[i].v;
    if (dshape) for (int k = 0; k < dim; k++) dshape[i * dim + k] = sh[i].d[k];
  }
}

/* ElementTransformation (+AddDeformation) [NGSolve]: x = sum lam_i v_i + sum_j def_j phi_j */
void xo_trafo_eval(const xo_trafo* T, const double* xhat, double* F, double* x)
{
  int D = T->dim;
  double lam[4], s = 1.0;
  for (int i = 0; i < D; i++) { lam[i] = xhat[i]; s -= xhat[i]; }
  lam[D] = s;
  for (int i = 0; i < D; i++) {
    x[i] = 0;
    for (int k = 0; k <= D; k++) x[i] += lam[k] * T->vtx[k][i];
    for (int j = 0; j < D; j++) F[i * D + j] = T->vtx[j][i] - T->vtx[D][i];
  }
  if (T->def_order > 0 && T->def) {
    double sh[64], dsh[64 * 3];
    int nd = xo_h1_ndof(D, T->def_order);
    xo_h1_shapes(D, T->def_order, T->vnum, xhat, sh, dsh);
    for (int a = 0; a < nd; a++)
      for (int i = 0; i < D; i++) {
        x[i] += T->def[a * D + i] * sh[a];
        for (int j = 0; j < D; j++) F[i * D + j] += T->def[a * D + i] * dsh[a * D + j];
      }
  }
}

static double inv_det(int D, const double* F, double* Fi)
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

static int difforder(int op) { return op == XO_OP_ID ? 0 : 1; }

/* the integrand evaluation of symboliccutbfi.cpp:133-304 for the closed menu of terms:
 * elmat(test i, trial j) += sum_q coef h^hpow (w_q |det F|) op_test(phi_i) op_trial(phi_j) */
void xo_elmat_from_rule(const xo_trafo* T, int p, const double* lsetvals,
                        const double* pts, const double* wts, int np,
                        const xo_term* terms, int nterms, double* elmat)
{
  int D = T->dim, nd = xo_h1_ndof(D, p);
  double sh[64], dsh[64 * 3], g[64 * 3], dn[64];
  double lgr[3] = { 0, 0, 0 };
  if (lsetvals) for (int i = 0; i < D; i++) lgr[i] = lsetvals[i] - lsetvals[D]; /* LevelsetWrapper coefs */
  for (int q = 0; q < np; q++) {
    double F[9], Fi[9], x[3];
    xo_trafo_eval(T, &pts[3 * q], F, x);
    double det = inv_det(D, F, Fi);
    double meas = fabs(det);
    double h = pow(meas, 1.0 / D); /* specialcf.mesh_size [NGSolve]: |det F|^(1/D) on volume elements */
    xo_h1_shapes(D, p, T->vnum, &pts[3 * q], sh, dsh);
    double n[3] = { 0, 0, 0 }, nn = 0;
    for (int i = 0; i < D; i++) { for (int j = 0; j < D; j++) n[i] += Fi[j * D + i] * lgr[j]; nn += n[i] * n[i]; }
    nn = sqrt(nn);
    if (nn > 0) for (int i = 0; i < D; i++) n[i] /= nn;
    for (int a = 0; a < nd; a++) {
      dn[a] = 0;
      for (int i = 0; i < D; i++) {
        double s = 0;
        for (int j = 0; j < D; j++) s += Fi[j * D + i] * dsh[a * D + j];
        g[a * D + i] = s;
        dn[a] += s * n[i];
      }
    }
    double fac = wts[q] * meas;
    for (int t = 0; t < nterms; t++) {
      double c = terms[t].coef * fac;
      if (terms[t].hpow != 0) c *= pow(h, (double)terms[t].hpow);
      int ot = terms[t].op_test, ou = terms[t].op_trial;
      for (int i = 0; i < nd; i++)
        for (int j = 0; j < nd; j++) {
          double v;
          if (ot == XO_OP_GRAD && ou == XO_OP_GRAD) {
            v = 0;
            for (int k = 0; k < D; k++) v += g[i * D + k] * g[j * D + k];
          } else {
            double a = (ot == XO_OP_ID) ? sh[i] : dn[i];
            double b = (ou == XO_OP_ID) ? sh[j] : dn[j];
            v = a * b;
          }
          elmat[i * nd + j] += c * v;
        }
    }
  }
}

int xo_cut_elmat(const xo_trafo* T, int p, const double* lsetvals, int dt, int order,
                 const xo_term* terms, int nterms, double* elmat)
{
  static __thread double pts[3 * 4 * XO_MAXPTS_RULE], wts[4 * XO_MAXPTS_RULE];
  if (order < 0) { /* symboliccutbfi.cpp:100-118 */
    int dtr = 99, dte = 99;
    for (int t = 0; t < nterms; t++) {
      if (difforder(terms[t].op_trial) < dtr) dtr = difforder(terms[t].op_trial);
      if (difforder(terms[t].op_test) < dte) dte = difforder(terms[t].op_test);
    }
    if (nterms == 0) dtr = dte = 0;
    order = 2 * p - (dtr + dte);
  }
  int np = xo_straight_cut_rule(T, lsetvals, dt, order, pts, wts, 4 * XO_MAXPTS_RULE, 0);
  if (np <= 0) return np;
  xo_elmat_from_rule(T, p, lsetvals, pts, wts, np, terms, nterms, elmat);
  return np;
}

/* ---- space-time FE (spacetime/SpaceTimeFE.cpp:57-135): shape[j*nd_s+i] = L_j(t) N_i(xhat), time-major;
 *      DiffOpDt = L_j'(t) N_i (reference-time derivative); nodes of NodalTimeFE (SpaceTimeFE.cpp:258-282) ---- */
void xo_time_dshapes(int order, double t, double* dshape)
{
  double nodes[8];
  if (order == 0) { dshape[0] = 0.0; return; }
  if (order == 1) { nodes[0] = 0; nodes[1] = 1; }
  else if (order == 2) { nodes[0] = 0; nodes[1] = 0.5; nodes[2] = 1; }
  else { nodes[0] = 0; nodes[1] = 0.5 * (1.0 - 1.0 / sqrt(5.0)); nodes[2] = 0.5 * (1.0 + 1.0 / sqrt(5.0)); nodes[3] = 1; }
  for (int i = 0; i <= order; i++) {
    double ds = 0.0;
    for (int k = 0; k <= order; k++) if (k != i) {
      double pr = 1.0 / (nodes[i] - nodes[k]);
      for (int j = 0; j <= order; j++) if (j != i && j != k) pr *= (t - nodes[j]) / (nodes[i] - nodes[j]);
      ds += pr;
    }
    dshape[i] = ds;
  }
}

/* values / physical space gradients / reference-time derivatives of all space-time functions at (xhat, t) */
static int st_basis(const xo_trafo* T, int p, int kt, const double* xhat, double t, const double* Fi,
                    double* val, double* g, double* dtv)
{
  int D = T->dim, nds = xo_h1_ndof(D, p), nt = kt + 1;
  double sh[64], dsh[64 * 3], L[8], dL[8];
  xo_h1_shapes(D, p, T->vnum, xhat, sh, dsh);
  xo_time_shapes(kt, t, L);
  xo_time_dshapes(kt, t, dL);
  for (int j = 0; j < nt; j++)
    for (int i = 0; i < nds; i++) {
      int k = j * nds + i;
      val[k] = L[j] * sh[i];
      dtv[k] = dL[j] * sh[i];
      for (int c = 0; c < D; c++) {
        double s = 0;
        for (int m = 0; m < D; m++) s += Fi[m * D + c] * dsh[i * D + m];
        g[k * D + c] = L[j] * s;
      }
    }
  return nds * nt;
}

/* symboliccutbfi.cpp:133-304 with a SpaceTimeFE and a space-time rule (spacetimecutrule.cpp:104-244):
 * elmat(test i, trial j) += sum_q coef h^hpow (w_q |det F|) op_test(phi_i) op_trial(phi_j), ops ID / GRAD.GRAD / DT */
void xo_elmat_st_from_rule(const xo_trafo* T, int p, int kt, const double* pts, const double* tvals,
                           const double* wts, int np, const xo_term* terms, int nterms, double* elmat)
{
  int D = T->dim;
  static __thread double val[256], g[256 * 3], dtv[256];
  for (int q = 0; q < np; q++) {
    double F[9], Fi[9], x[3];
    xo_trafo_eval(T, &pts[3 * q], F, x);
    double meas = fabs(inv_det(D, F, Fi));
    double h = pow(meas, 1.0 / D);
    int nd = st_basis(T, p, kt, &pts[3 * q], tvals[q], Fi, val, g, dtv);
    double fac = wts[q] * meas;
    for (int t = 0; t < nterms; t++) {
      double c = terms[t].coef * fac;
      if (terms[t].hpow != 0) c *= pow(h, (double)terms[t].hpow);
      int ot = terms[t].op_test, ou = terms[t].op_trial;
      for (int i = 0; i < nd; i++)
        for (int j = 0; j < nd; j++) {
          double v;
          if (ot == XO_OP_GRAD && ou == XO_OP_GRAD) {
            v = 0;
            for (int k = 0; k < D; k++) v += g[i * D + k] * g[j * D + k];
          } else {
            double a = (ot == XO_OP_ID) ? val[i] : dtv[i];
            double b = (ou == XO_OP_ID) ? val[j] : dtv[j];
            v = a * b;
          }
          elmat[i * nd + j] += c * v;
        }
    }
  }
}

/* symboliccutlfi.cpp:51-140 with a SpaceTimeFE: elvec_i += sum_q (w_q |det F|) f(x_q, t_q) op_test(phi_i),
 * f = coef h^hpow x^e0 y^e1 z^e2 tref^e3 */
void xo_elvec_st_from_rule(const xo_trafo* T, int p, int kt, const double* pts, const double* tvals,
                           const double* wts, int np, const xo_stlterm* terms, int nterms, double* elvec)
{
  int D = T->dim;
  static __thread double val[256], g[256 * 3], dtv[256];
  for (int q = 0; q < np; q++) {
    double F[9], Fi[9], x[3] = { 0, 0, 0 };
    xo_trafo_eval(T, &pts[3 * q], F, x);
    double meas = fabs(inv_det(D, F, Fi));
    double h = pow(meas, 1.0 / D);
    int nd = st_basis(T, p, kt, &pts[3 * q], tvals[q], Fi, val, g, dtv);
    for (int t = 0; t < nterms; t++) {
      double f = terms[t].coef * wts[q] * meas;
      if (terms[t].hpow != 0) f *= pow(h, (double)terms[t].hpow);
      for (int c = 0; c < D; c++) for (int k = 0; k < terms[t].expo[c]; k++) f *= x[c];
      for (int k = 0; k < terms[t].expo[3]; k++) f *= tvals[q];
      for (int i = 0; i < nd; i++) elvec[i] += f * (terms[t].op_test == XO_OP_ID ? val[i] : dtv[i]);
    }
  }
}

/* symboliccutlfi.cpp:51-140 (scalar evaluation branch :141-175): elvec += sum_q B_q^T (f(x_q) w_q |det F|) */
void xo_elvec_from_rule(const xo_trafo* T, int p, const double* lsetvals,
                        const double* pts, const double* wts, int np,
                        const xo_lterm* terms, int nterms, double* elvec)
{
  int D = T->dim, nd = xo_h1_ndof(D, p);
  double sh[64], dsh[64 * 3], dn[64];
  double lgr[3] = { 0, 0, 0 };
  if (lsetvals) for (int i = 0; i < D; i++) lgr[i] = lsetvals[i] - lsetvals[D];
  for (int q = 0; q < np; q++) {
    double F[9], Fi[9], x[3] = { 0, 0, 0 };
    xo_trafo_eval(T, &pts[3 * q], F, x);
    double det = inv_det(D, F, Fi);
    double meas = fabs(det);
    double h = pow(meas, 1.0 / D);
    xo_h1_shapes(D, p, T->vnum, &pts[3 * q], sh, dsh);
    double n[3] = { 0, 0, 0 }, nn = 0;
    for (int i = 0; i < D; i++) { for (int j = 0; j < D; j++) n[i] += Fi[j * D + i] * lgr[j]; nn += n[i] * n[i]; }
    nn = sqrt(nn);
    if (nn > 0) for (int i = 0; i < D; i++) n[i] /= nn;
    for (int a = 0; a < nd; a++) {
      dn[a] = 0;
      for (int i = 0; i < D; i++) {
        double s = 0;
        for (int j = 0; j < D; j++) s += Fi[j * D + i] * dsh[a * D + j];
        dn[a] += s * n[i];
      }
    }
    double fac = wts[q] * meas;
    for (int t = 0; t < nterms; t++) {
      double c = terms[t].coef * fac;
      if (terms[t].hpow != 0) c *= pow(h, (double)terms[t].hpow);
      for (int i = 0; i < D; i++) for (int e = 0; e < terms[t].expo[i]; e++) c *= x[i];
      const double* b = (terms[t].op_test == XO_OP_ID) ? sh : dn;
      for (int i = 0; i < nd; i++) elvec[i] += c * b[i];
    }
  }
}

int xo_cut_elvec(const xo_trafo* T, int p, const double* lsetvals, int dt, int order,
                 const xo_lterm* terms, int nterms, double* elvec)
{
  static __thread double pts[3 * 4 * XO_MAXPTS_RULE], wts[4 * XO_MAXPTS_RULE];
  if (order < 0) order = 2 * p; /* symboliccutlfi.cpp:86-88 */
  int np = xo_straight_cut_rule(T, lsetvals, dt, order, pts, wts, 4 * XO_MAXPTS_RULE, 0);
  if (np <= 0) return np;
  xo_elvec_from_rule(T, p, lsetvals, pts, wts, np, terms, nterms, elvec);
  return np;
}

/* MapPatchIntegrationPoint<D>: symboliccutbfi.cpp:1747-1846 (non space-time).
 * from (xhat_from, weight w_from) on T_from  ->  xhat_to, weight w_to on T_to.  returns 1 on
 * Newton failure (low order guess taken). */
static int map_patch_point(const xo_trafo* Tf, const xo_trafo* Tt, const double* xf, double wf,
                           double* xt, double* wt)
{
  int D = Tf->dim, failed = 0;
  double F[9], Fi[9], vec[3], zero[3] = { 0, 0, 0 };
  xo_trafo_eval(Tf, xf, F, vec);
  double detf = inv_det(D, F, Fi);
  const double h = sqrt(detf);
  double mipw = wf * fabs(detf); /* mip.GetWeight() */
  double a0[3], A[9], Ainv[9], x00[3], x0[3];
  xo_trafo_eval(Tt, zero, F, a0);
  for (int d = 0; d < D; d++) {
    double e[3] = { 0, 0, 0 }, ai[3];
    e[d] = 1;
    xo_trafo_eval(Tt, e, F, ai);
    for (int i = 0; i < D; i++) A[i * D + d] = ai[i] - a0[i];
  }
  double w00 = fabs(inv_det(D, A, Ainv));
  for (int i = 0; i < D; i++) {
    x00[i] = 0;
    for (int j = 0; j < D; j++) x00[i] += Ainv[i * D + j] * (vec[j] - a0[j]);
    x0[i] = x00[i];
  }
  for (int i = D; i < 3; i++) x00[i] = x0[i] = 0;
  int its = 0;
  double w = 0, diffn = 0;
  const double EPS = 1e-12; const int NIT = 15; const double MAXDIST = 10; /* ngsxstd.cpp:66-82 */
  while (its == 0 || (diffn > EPS * h && its < NIT)) {
    double xm[3], diff[3];
    xo_trafo_eval(Tt, x0, F, xm);
    double det = inv_det(D, F, Fi);
    diffn = 0;
    for (int i = 0; i < D; i++) { diff[i] = vec[i] - xm[i]; diffn += diff[i] * diff[i]; }
    diffn = sqrt(diffn);
    for (int i = 0; i < D; i++) { double u = 0; for (int j = 0; j < D; j++) u += Fi[i * D + j] * diff[j]; x0[i] += u; }
    its++;
    w = fabs(det);
  }
  double dist = 0;
  for (int i = 0; i < D; i++) dist += (x0[i] - x00[i]) * (x0[i] - x00[i]);
  dist = sqrt(dist);
  if (its >= NIT || diffn > EPS * h) { failed = 1; for (int i = 0; i < 3; i++) xt[i] = x00[i]; *wt = mipw / w00; }
  else if (dist > MAXDIST) { failed = 1; for (int i = 0; i < 3; i++) xt[i] = x00[i]; *wt = mipw / w00; }
  else { for (int i = 0; i < 3; i++) xt[i] = x0[i]; *wt = mipw / w; }
  return failed;
}

/* ScaleAndMapIntegrationRule<D>: symboliccutbfi.cpp:1849-1885.  The rule of element T is scaled towards the midpoint of the
 * reference facet LocalFacetNr: shift = (1 - s) / nv * sum of the reference vertices of the facet, point <- s * point + shift,
 * weight <- weight * s * s (in every dimension).  The facet's vertices (ElementTopology::GetFaces / GetEdges / GetVertices
 * of LocalFacetNr) are the vertices T shares with the neighbour To; reference vertex k < D is e_k, vertex D the origin. */
static void scale_and_map_rule(const xo_trafo* T, const xo_trafo* To, double s, int n, const double* rp, const double* rw,
                               double* sp, double* sw)
{
  int D = T->dim, nv = 0;
  double shift[3] = { 0, 0, 0 };
  for (int k = 0; k < D + 1; k++) {
    int shared = 0;
    for (int j = 0; j < D + 1; j++) shared |= (T->vnum[k] == To->vnum[j]);
    if (!shared) continue;
    nv++;
    if (k < D) shift[k] += 1.0;
  }
  for (int l = 0; l < D; l++) shift[l] *= (1 - s) / nv;
  for (int q = 0; q < n; q++) {
    for (int l = 0; l < 3; l++) sp[3 * q + l] = l < D ? rp[3 * q + l] * s + shift[l] : rp[3 * q + l];
    sw[q] = rw[q] * s * s;
  }
}

/* T_CalcFacetMatrix: symboliccutbfi.cpp:1982-2377, time_order=-1,
 * form coef*h^hpow*(u-u.Other())*(v-v.Other()).  order<0: 2*fe_order (:2004-2012).  downscale: :2019-2033. */
int xo_facet_patch_elmat_scaled(const xo_trafo* T1, const xo_trafo* T2, int p, int order,
                                double coef, int hpow, double downscale, double* elmat)
{
  int D = T1->dim, nd = xo_h1_ndof(D, p), N = 2 * nd, nfail = 0;
  static __thread double rp[3 * XO_MAXPTS_RULE], rw[XO_MAXPTS_RULE];
  static __thread double rp1[3 * XO_MAXPTS_RULE], rw1[XO_MAXPTS_RULE], rp2[3 * XO_MAXPTS_RULE], rw2[XO_MAXPTS_RULE];
  if (order < 0) order = 2 * p;
  if (!(downscale > 0 && downscale <= 1)) return -1; /* "Only 0 < downscale <= 1 valid" (:2031-2033) */
  int n = xo_select_rule(D, order, rp, rw);
  if (downscale < 1) {
    scale_and_map_rule(T1, T2, downscale, n, rp, rw, rp1, rw1);
    scale_and_map_rule(T2, T1, downscale, n, rp, rw, rp2, rw2);
  } else {
    memcpy(rp1, rp, sizeof(double) * 3 * n); memcpy(rw1, rw, sizeof(double) * n);
    memcpy(rp2, rp, sizeof(double) * 3 * n); memcpy(rw2, rw, sizeof(double) * n);
  }
  memset(elmat, 0, sizeof(double) * N * N);
  for (int l = 0; l < 2 * n; l++) {
    double x1[3], x2[3], w1;
    if (l < n) {
      double wt;
      for (int k = 0; k < 3; k++) x1[k] = rp1[3 * l + k];
      w1 = rw1[l];
      nfail += map_patch_point(T1, T2, x1, w1, x2, &wt);
    } else {
      for (int k = 0; k < 3; k++) x2[k] = rp2[3 * (l - n) + k];
      nfail += map_patch_point(T2, T1, x2, rw2[l - n], x1, &w1);
    }
    double F[9], Fi[9], x[3], J[128], sh[64];
    xo_trafo_eval(T1, x1, F, x);
    double meas = fabs(inv_det(D, F, Fi));
    double h = pow(meas, 1.0 / D);
    double c = coef * w1 * meas; /* mir1[l].GetWeight() */
    if (hpow != 0) c *= pow(h, (double)hpow);
    xo_h1_shapes(D, p, T1->vnum, x1, sh, 0);
    for (int i = 0; i < nd; i++) J[i] = sh[i];
    xo_h1_shapes(D, p, T2->vnum, x2, sh, 0);
    for (int i = 0; i < nd; i++) J[nd + i] = -sh[i];
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) elmat[i * N + j] += c * J[i] * J[j];
  }
  return nfail;
}

int xo_facet_patch_elmat(const xo_trafo* T1, const xo_trafo* T2, int p, int order,
                         double coef, int hpow, double* elmat)
{
  return xo_facet_patch_elmat_scaled(T1, T2, p, order, coef, hpow, 1.0, elmat);
}
