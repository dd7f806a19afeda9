/* xo_pipeline.c — ORACLE (test infrastructure): batched drivers over many elements, used as the
 * checker for the GPU batches and as the CPU baseline ("port", OpenMP over the host cores).
 * Each loop body is the per-element call sequence of the reference.
 */
#include "xoracle.h"
#include <string.h>
#include <stdlib.h>
#include <math.h>

static void fill_trafo(xo_trafo* T, int dim, const int32_t* ev, const double* vtx_xyz,
                       int def_order, const double* def_el)
{
  memset(T, 0, sizeof(*T));
  T->dim = dim;
  for (int k = 0; k <= dim; k++) {
    T->vnum[k] = ev[k];
    for (int i = 0; i < dim; i++) T->vtx[k][i] = vtx_xyz[(int64_t)ev[k] * dim + i];
  }
  T->def_order = def_order;
  T->def = def_el;
}

/* cut rules for a list of elements; output CSR-of-points (offsets must hold n+1).
 * pass pts==NULL to only count.  returns total or error. */
int64_t xo_batch_cut_rules(int dim, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                           const double* vtx_xyz, const double* phi, int dt, int order,
                           int64_t* offsets, double* pts, double* wts)
{
  int nv = dim + 1;
  int64_t err = 0;
  /* pass 1: counts */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    int64_t el = elems ? elems[i] : i;
    double v[4], p[3 * 3 * XO_MAXPTS_RULE / 8], w[3 * XO_MAXPTS_RULE / 8];
    xo_trafo T;
    fill_trafo(&T, dim, el_vtx + el * nv, vtx_xyz, 0, 0);
    for (int k = 0; k < nv; k++) v[k] = phi[el_vtx[el * nv + k]];
    int c = xo_straight_cut_rule(&T, v, dt, order, p, w, 3 * XO_MAXPTS_RULE / 8, 0);
    if (c < 0) {
#pragma omp critical
      err = c;
      c = 0;
    }
    offsets[i + 1] = c;
  }
  if (err) return err;
  offsets[0] = 0;
  for (int64_t i = 0; i < n; i++) offsets[i + 1] += offsets[i];
  if (!pts) return offsets[n];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    int64_t el = elems ? elems[i] : i;
    double v[4];
    xo_trafo T;
    fill_trafo(&T, dim, el_vtx + el * nv, vtx_xyz, 0, 0);
    for (int k = 0; k < nv; k++) v[k] = phi[el_vtx[el * nv + k]];
    int cap = (int)(offsets[i + 1] - offsets[i]);
    if (cap > 0) xo_straight_cut_rule(&T, v, dt, order, pts + 3 * offsets[i], wts + offsets[i], cap, 0);
  }
  return offsets[n];
}

/* element matrices for a list of elements (SymbolicCutBFI over dCut(lset, dt, order)):
 * mats[i] (nd x nd) is zeroed then filled.  def (optional): per element nd_def x dim. */
int xo_batch_cut_elmats(int dim, int fe_order, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                        const double* vtx_xyz, const double* phi, int dt, int order,
                        const xo_term* terms, int nterms, int def_order, const double* def,
                        double* mats)
{
  int nv = dim + 1, nd = xo_h1_ndof(dim, fe_order);
  int ndd = def_order > 0 ? xo_h1_ndof(dim, def_order) : 0;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n; i++) {
    int64_t el = elems ? elems[i] : i;
    double v[4];
    xo_trafo T;
    fill_trafo(&T, dim, el_vtx + el * nv, vtx_xyz, def_order, def ? def + el * (int64_t)ndd * dim : 0);
    for (int k = 0; k < nv; k++) v[k] = phi[el_vtx[el * nv + k]];
    double* m = mats + i * (int64_t)nd * nd;
    memset(m, 0, sizeof(double) * nd * nd);
    int c = xo_cut_elmat(&T, fe_order, v, dt, order, terms, nterms, m);
    if (c < 0) {
#pragma omp critical
      err = c;
    }
  }
  return err;
}

/* element vectors for a list of elements (SymbolicCutLFI over dCut(lset, dt, order)): vecs[i] (nd) zeroed then filled */
int xo_batch_cut_elvecs(int dim, int fe_order, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                        const double* vtx_xyz, const double* phi, int dt, int order,
                        const xo_lterm* terms, int nterms, int def_order, const double* def,
                        double* vecs)
{
  int nv = dim + 1, nd = xo_h1_ndof(dim, fe_order);
  int ndd = def_order > 0 ? xo_h1_ndof(dim, def_order) : 0;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n; i++) {
    int64_t el = elems ? elems[i] : i;
    double v[4];
    xo_trafo T;
    fill_trafo(&T, dim, el_vtx + el * nv, vtx_xyz, def_order, def ? def + el * (int64_t)ndd * dim : 0);
    for (int k = 0; k < nv; k++) v[k] = phi[el_vtx[el * nv + k]];
    double* m = vecs + i * (int64_t)nd;
    memset(m, 0, sizeof(double) * nd);
    int c = xo_cut_elvec(&T, fe_order, v, dt, order, terms, nterms, m);
    if (c < 0) {
#pragma omp critical
      err = c;
    }
  }
  return err;
}

/* facet patch matrices for a list of facets */
int xo_batch_facet_patch_scaled(int dim, int fe_order, int64_t n, const int32_t* facets, const int32_t* facet_els,
                                const int32_t* el_vtx, const double* vtx_xyz, int order, double coef, int hpow,
                                int def_order, const double* def, double downscale, double* mats);
int xo_batch_facet_patch(int dim, int fe_order, int64_t n, const int32_t* facets, const int32_t* facet_els,
                         const int32_t* el_vtx, const double* vtx_xyz, int order, double coef, int hpow,
                         int def_order, const double* def, double* mats)
{
  return xo_batch_facet_patch_scaled(dim, fe_order, n, facets, facet_els, el_vtx, vtx_xyz, order, coef, hpow, def_order, def, 1.0, mats);
}

/* the same with dFacetPatch(downscale=) (symboliccutbfi.cpp:2019-2033); -1: downscale outside (0, 1] */
int xo_batch_facet_patch_scaled(int dim, int fe_order, int64_t n, const int32_t* facets, const int32_t* facet_els,
                                const int32_t* el_vtx, const double* vtx_xyz, int order, double coef, int hpow,
                                int def_order, const double* def, double downscale, double* mats)
{
  if (!(downscale > 0 && downscale <= 1)) return -1;
  int nv = dim + 1, nd = xo_h1_ndof(dim, fe_order), N = 2 * nd;
  int ndd = def_order > 0 ? xo_h1_ndof(dim, def_order) : 0;
  int nfail = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nfail)
  for (int64_t i = 0; i < n; i++) {
    int64_t f = facets ? facets[i] : i;
    int32_t e1 = facet_els[2 * f], e2 = facet_els[2 * f + 1];
    xo_trafo T1, T2;
    fill_trafo(&T1, dim, el_vtx + (int64_t)e1 * nv, vtx_xyz, def_order, def ? def + e1 * (int64_t)ndd * dim : 0);
    fill_trafo(&T2, dim, el_vtx + (int64_t)e2 * nv, vtx_xyz, def_order, def ? def + e2 * (int64_t)ndd * dim : 0);
    nfail += xo_facet_patch_elmat_scaled(&T1, &T2, fe_order, order, coef, hpow, downscale, mats + i * (int64_t)N * N);
  }
  return nfail;
}

/* Integrate: sum_q w_q |det F| f(x_q) with f = prod x_i^{e_i} (cutint/cutintegral.cpp:126-212) */
double xo_integrate_monomial(int dim, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz,
                             const double* phi, int dt, int order, const int* expo, int* err_out)
{
  int nv = dim + 1, err = 0;
  double total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int64_t el = 0; el < ne; el++) {
    double v[4], p[3 * 3 * XO_MAXPTS_RULE / 8], w[3 * XO_MAXPTS_RULE / 8];
    xo_trafo T;
    fill_trafo(&T, dim, el_vtx + el * nv, vtx_xyz, 0, 0);
    for (int k = 0; k < nv; k++) v[k] = phi[el_vtx[el * nv + k]];
    int c = xo_straight_cut_rule(&T, v, dt, order, p, w, 3 * XO_MAXPTS_RULE / 8, 0);
    if (c < 0) {
#pragma omp critical
      err = c;
      continue;
    }
    for (int q = 0; q < c; q++) {
      double F[9], x[3], f = 1.0, det;
      xo_trafo_eval(&T, &p[3 * q], F, x);
      if (dim == 1) det = F[0];
      else if (dim == 2) det = F[0] * F[3] - F[1] * F[2];
      else det = F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
      for (int i = 0; i < dim; i++) for (int e = 0; e < expo[i]; e++) f *= x[i];
      total += w[q] * fabs(det) * f;
    }
  }
  if (err_out) *err_out = err;
  return total;
}
