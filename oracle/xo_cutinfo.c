/* xo_cutinfo.c — ORACLE (test infrastructure): CutInformation / MultiLevelsetCutInformation and
 * the marker helpers of xfem/cutinfo.cpp, restated over plain arrays.
 */
#include "xoracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

static inline void setbit(uint8_t* b, int64_t i)
{
#ifdef _OPENMP
  __atomic_fetch_or(&b[i >> 3], (uint8_t)(1u << (i & 7)), __ATOMIC_RELAXED); /* SetBitAtomic */
#else
  b[i >> 3] |= (uint8_t)(1u << (i & 7));
#endif
}
static inline int testbit(const uint8_t* b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }

/* CutInformation::Update, VOL loop: xfem/cutinfo.cpp:80-126.  As the reference, TWO order-0
 * cut rules (POS then NEG) are constructed for EVERY element and their weights summed. */
int xo_cutinfo_update(int dim, int64_t ne, const int32_t* el_vtx, const double* phi,
                      uint8_t* bits_neg, uint8_t* bits_pos, uint8_t* bits_if,
                      double* cut_ratio, int64_t* err_elem)
{
  int nv = dim + 1;
  int64_t nbytes = (ne + 7) / 8;
  int first_err = 0;
  int64_t first_el = -1;
  memset(bits_neg, 0, nbytes); memset(bits_pos, 0, nbytes); memset(bits_if, 0, nbytes);
#pragma omp parallel for schedule(static)
  for (int64_t el = 0; el < ne; el++) {
    double v[4], pts[3 * 64], wts[64];
    xo_trafo T;
    memset(&T, 0, sizeof(T));
    T.dim = dim;
    for (int k = 0; k < nv; k++) v[k] = phi[el_vtx[el * nv + k]];
    double part_vol[2] = { 0.0, 0.0 };
    int e = 0;
    const int order_np[2] = { XO_POS, XO_NEG };
    for (int a = 0; a < 2; a++) {
      int np = order_np[a];
      int n = xo_straight_cut_rule(&T, v, np, 0, pts, wts, 64, 0);
      if (n < 0) { e = n; break; }
      for (int q = 0; q < n; q++) part_vol[np] += wts[q];
    }
    if (e) {
#pragma omp critical
      { if (first_el < 0 || el < first_el) { first_el = el; first_err = e; } }
      if (cut_ratio) cut_ratio[el] = NAN;
      continue;
    }
    if (cut_ratio) cut_ratio[el] = part_vol[XO_NEG] / (part_vol[XO_NEG] + part_vol[XO_POS]);
    if (part_vol[XO_NEG] > 0.0) {
      if (part_vol[XO_POS] > 0.0) setbit(bits_if, el);
      else setbit(bits_neg, el);
    } else setbit(bits_pos, el);
  }
  if (err_elem) *err_elem = first_el;
  return first_err;
}

/* MultiLevelsetCutInformation::UpdateElementsOfDomainType: cutinfo.cpp:356-398 */
void xo_ml_elements_of_type(int dim, int64_t ne, const int32_t* el_vtx, const double* phis,
                            int64_t nvtx, int M, const int* dts, int ntuples, uint8_t* bits)
{
  int nv = dim + 1;
  memset(bits, 0, (ne + 7) / 8);
#pragma omp parallel for schedule(static)
  for (int64_t el = 0; el < ne; el++) {
    int dts_el[16];
    for (int j = 0; j < M; j++) {
      double v[4];
      for (int k = 0; k < nv; k++) v[k] = phis[(int64_t)j * nvtx + el_vtx[el * nv + k]];
      dts_el[j] = xo_check_if_straight_cut(v, nv, 0.0);
    }
    for (int i = 0; i < ntuples; i++) {
      int matching = 1;
      for (int j = 0; j < M; j++) if (dts[i * M + j] != dts_el[j]) matching = 0;
      if (matching) setbit(bits, el);
    }
  }
}

/* ...UpdateElementsWithContribution: cutinfo.cpp:306-334 (domain built with intorder=-1) */
int xo_ml_elements_with_contribution(int dim, int64_t ne, const int32_t* el_vtx, const double* vtx_xyz,
                                     const double* phis, int64_t nvtx, int M, const int* dts,
                                     int ntuples, uint8_t* bits)
{
  int nv = dim + 1, err = 0;
  memset(bits, 0, (ne + 7) / 8);
#pragma omp parallel for schedule(static)
  for (int64_t el = 0; el < ne; el++) {
    double lsets[4 * 16];
    xo_trafo T;
    memset(&T, 0, sizeof(T));
    T.dim = dim;
    for (int k = 0; k < nv; k++) {
      int32_t vi = el_vtx[el * nv + k];
      T.vnum[k] = vi;
      for (int i = 0; i < dim; i++) T.vtx[k][i] = vtx_xyz[(int64_t)vi * dim + i];
      for (int j = 0; j < M; j++) lsets[k * M + j] = phis[(int64_t)j * nvtx + vi];
    }
    int cap = 4096;
    double* pts = (double*)malloc(sizeof(double) * 4 * cap);
    double* wts = pts + 3 * cap;
    int n = xo_mlset_cut_rule(&T, lsets, M, dts, ntuples, -1, pts, wts, cap);
    if (n < 0) {
#pragma omp critical
      err = n;
    } else {
      double part_vol = 0;
      for (int q = 0; q < n; q++) part_vol += wts[q];
      if (part_vol > 0) setbit(bits, el);
    }
    free(pts);
  }
  return err;
}

/* GetFacetsWithNeighborTypes: cutinfo.cpp:476-538 (every facet here belongs to an element, so
 * fine_facet is all-true; periodic identification: xo_facets_with_neighbor_types_periodic) */
void xo_facets_with_neighbor_types(int64_t nf, const int32_t* facet_els, const uint8_t* a,
                                   const uint8_t* b, int bnd_val_a, int bnd_val_b, int use_and,
                                   uint8_t* out)
{
  memset(out, 0, (nf + 7) / 8);
  for (int64_t f = 0; f < nf; f++) {
    int32_t e0 = facet_els[2 * f], e1 = facet_els[2 * f + 1];
    int a_left = testbit(a, e0);
    int a_right = e1 >= 0 ? testbit(a, e1) : bnd_val_a;
    int b_left = testbit(b, e0);
    int b_right = e1 >= 0 ? testbit(b, e1) : bnd_val_b;
    if (use_and) { if ((a_left && b_right) || (a_right && b_left)) setbit(out, f); }
    else { if ((a_left || b_right) || (a_right || b_left)) setbit(out, f); }
  }
}

/* the same on a mesh with periodic identification (cutinfo.cpp:509-518): periodic_facet[f] = MeshAccess::GetPeriodicFacet(f)
 * (f itself where there is none).  A facet with ONE element whose identified facet has the larger number takes that facet's
 * first element as its second one; the other facet of the pair keeps the boundary values. */
void xo_facets_with_neighbor_types_periodic(int64_t nf, const int32_t* facet_els, const int32_t* periodic_facet,
                                            const uint8_t* a, const uint8_t* b, int bnd_val_a, int bnd_val_b, int use_and,
                                            uint8_t* out)
{
  memset(out, 0, (nf + 7) / 8);
  for (int64_t f = 0; f < nf; f++) {
    int32_t elnums[2] = { facet_els[2 * f], facet_els[2 * f + 1] };
    int size = elnums[1] >= 0 ? 2 : 1;
    if (size < 2) {
      int32_t facet2 = periodic_facet[f];
      if (facet2 > f) elnums[size++] = facet_els[2 * (int64_t)facet2];
    }
    int a_left = testbit(a, elnums[0]);
    int a_right = size > 1 ? testbit(a, elnums[1]) : bnd_val_a;
    int b_left = testbit(b, elnums[0]);
    int b_right = size > 1 ? testbit(b, elnums[1]) : bnd_val_b;
    if (use_and) { if ((a_left && b_right) || (a_right && b_left)) setbit(out, f); }
    else { if ((a_left || b_right) || (a_right || b_left)) setbit(out, f); }
  }
}

/* GetElementsWithNeighborFacets: cutinfo.cpp:540-566 */
void xo_elements_with_neighbor_facets(int64_t nf, const int32_t* facet_els, const uint8_t* facets,
                                      int64_t ne, uint8_t* out)
{
  memset(out, 0, (ne + 7) / 8);
  for (int64_t f = 0; f < nf; f++)
    if (testbit(facets, f))
      for (int k = 0; k < 2; k++) if (facet_els[2 * f + k] >= 0) setbit(out, facet_els[2 * f + k]);
}

/* GetDofsOfElements: cutinfo.cpp:599-622 */
void xo_dofs_of_elements(int64_t ne, int nd, const int32_t* el_dofs, const uint8_t* elems,
                         int64_t ndof, uint8_t* out)
{
  memset(out, 0, (ndof + 7) / 8);
  for (int64_t e = 0; e < ne; e++)
    if (testbit(elems, e))
      for (int k = 0; k < nd; k++) if (el_dofs[e * nd + k] >= 0) setbit(out, el_dofs[e * nd + k]);
}
