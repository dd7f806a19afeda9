// h1.cuh — hierarchical H1 simplex basis (values + reference gradients), host + device.
//
// NGSolve's H1HighOrderFE (third party, call sites xfem/symboliccutbfi.cpp:221-223, 431-434,
// 2257, 2289) restated from its published construction: vertex hats lam_i; per edge (oriented
// from the smaller to the larger global vertex number) the scaled integrated Legendre
// polynomials L_k^S(lam_e - lam_s, lam_e + lam_s), k = 2..p; per face (vertices sorted by global
// number) lam_f0 lam_f1 lam_f2 * P_i^S(lam_f1 - lam_f0, lam_f0 + lam_f1) * P_j(2 lam_f2 - 1);
// cell bubbles analogously.  dof order: vertices, edges, faces, cell.
// Reference element: lam = (x, y, z, 1-x-y-z) (straightcutrule.cpp:506-514).
#pragma once
#include <cuda_runtime.h>

#define XG_HD __host__ __device__ __forceinline__

template <int D>
struct AD {
  double v;
  double d[D];
};
template <int D> XG_HD AD<D> ad_const(double c) { AD<D> r; r.v = c; for (int i = 0; i < D; i++) r.d[i] = 0.0; return r; }
template <int D> XG_HD AD<D> operator+(const AD<D>& a, const AD<D>& b) { AD<D> r; r.v = a.v + b.v; for (int i = 0; i < D; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int D> XG_HD AD<D> operator-(const AD<D>& a, const AD<D>& b) { AD<D> r; r.v = a.v - b.v; for (int i = 0; i < D; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int D> XG_HD AD<D> operator*(const AD<D>& a, const AD<D>& b) { AD<D> r; r.v = a.v * b.v; for (int i = 0; i < D; i++) r.d[i] = a.v * b.d[i] + a.d[i] * b.v; return r; }
template <int D> XG_HD AD<D> operator*(double s, const AD<D>& a) { AD<D> r; r.v = s * a.v; for (int i = 0; i < D; i++) r.d[i] = s * a.d[i]; return r; }

XG_HD int xg_h1_ndof(int dim, int p)
{
  if (dim == 1) return p + 1;
  if (dim == 2) return 3 + 3 * (p - 1) + (p - 1) * (p - 2) / 2;
  return 4 + 6 * (p - 1) + 4 * ((p - 1) * (p - 2) / 2) + (p - 1) * (p - 2) * (p - 3) / 6;
}

// local topology (NGSolve ElementTopology, published tables)
#define XG_TET_EDGE_A(e) ((e) < 3 ? 3 : ((e) < 5 ? 0 : 1))
#define XG_TET_EDGE_B(e) ((e) < 3 ? (e) : ((e) == 3 ? 1 : 2))
#define XG_TRIG_EDGE_A(e) ((e) == 0 ? 2 : ((e) == 1 ? 1 : 0))
#define XG_TRIG_EDGE_B(e) ((e) == 0 ? 0 : ((e) == 1 ? 2 : 1))

// Sink receives (index, AD) pairs; P is a compile-time order so that everything unrolls.
template <int D, int P, class Sink>
XG_HD void xg_edge_shapes(const AD<D>& ls, const AD<D>& le, int& n, Sink& sink)
{
  if (P < 2) return;
  AD<D> x = le - ls, t = le + ls;
  AD<D> tt = t * t;
  AD<D> Lm = x;                          // L_1
  AD<D> L = 0.5 * (x * x - tt);          // L_2
  sink(n++, L);
#pragma unroll
  for (int k = 2; k < P; k++) {          // (k+1) L_{k+1} = (2k-1) x L_k - (k-2) t^2 L_{k-1}
    AD<D> Ln = (1.0 / (k + 1)) * ((2.0 * k - 1.0) * (x * L) - (double)(k - 2) * (tt * Lm));
    Lm = L; L = Ln;
    sink(n++, L);
  }
}

// scaled Legendre P_k^S(x,t), k=0..N into arr
template <int D, int N>
XG_HD void xg_leg_scaled(const AD<D>& x, const AD<D>& t, AD<D>* P)
{
  P[0] = ad_const<D>(1.0);
  if (N >= 1) P[1] = x;
  AD<D> tt = t * t;
#pragma unroll
  for (int k = 1; k < N; k++) P[k + 1] = (1.0 / (k + 1)) * ((2.0 * k + 1.0) * (x * P[k]) - (double)k * (tt * P[k - 1]));
}

template <int D, int P, class Sink>
XG_HD void xg_face_shapes(const AD<D>* lam, int a, int b, int c, const int* vnum, int& n, Sink& sink)
{
  if (P < 3) return;
  int f0 = a, f1 = b, f2 = c, t;
  if (vnum[f0] > vnum[f1]) { t = f0; f0 = f1; f1 = t; }
  if (vnum[f0] > vnum[f2]) { t = f0; f0 = f2; f2 = t; }
  if (vnum[f1] > vnum[f2]) { t = f1; f1 = f2; f2 = t; }
  AD<D> l0 = lam[0], l1 = lam[0], l2 = lam[0];
#pragma unroll
  for (int i = 0; i < D + 1; i++) { if (i == f0) l0 = lam[i]; if (i == f1) l1 = lam[i]; if (i == f2) l2 = lam[i]; }
  AD<D> bub = l0 * l1 * l2;
  constexpr int N = (P >= 3) ? P - 3 : 0;
  AD<D> Pa[N + 1], Qa[N + 1];
  xg_leg_scaled<D, N>(l1 - l0, l0 + l1, Pa);
  xg_leg_scaled<D, N>(2.0 * l2 - ad_const<D>(1.0), ad_const<D>(1.0), Qa);
#pragma unroll
  for (int i = 0; i <= N; i++)
#pragma unroll
    for (int j = 0; j <= N - i; j++) sink(n++, bub * (Pa[i] * Qa[j]));
}

// all shapes of order P from prepared barycentric coordinates lam[0..D] (value + derivatives w.r.t. ANY coordinates:
// reference ones, or physical ones when lam[i].d = F^{-T} grad_ref lam_i — the shapes are polynomials in lam, so the
// chain rule carried by AD yields the physical gradients without a per-function F^{-T} product); returns ndof
template <int D, int P, class Sink>
XG_HD int xg_h1_shapes_lam(const int* vnum, const AD<D>* lam, Sink& sink);

// all shapes of order P on the D-simplex at xhat (reference gradients); returns ndof
template <int D, int P, class Sink>
XG_HD int xg_h1_shapes(const int* vnum, const double* xhat, Sink& sink)
{
  AD<D> lam[D + 1];
  double s = 1.0;
#pragma unroll
  for (int i = 0; i < D; i++) {
    lam[i] = ad_const<D>(xhat[i]); lam[i].d[i] = 1.0; s -= xhat[i];
  }
  lam[D] = ad_const<D>(s);
#pragma unroll
  for (int i = 0; i < D; i++) lam[D].d[i] = -1.0;
  return xg_h1_shapes_lam<D, P>(vnum, lam, sink);
}

template <int D, int P, class Sink>
XG_HD int xg_h1_shapes_lam(const int* vnum, const AD<D>* lam, Sink& sink)
{
  int n = 0;
#pragma unroll
  for (int i = 0; i < D + 1; i++) sink(n++, lam[i]);
  if (P < 2) return n;
  if (D == 1) {
    bool sw = vnum[1] > vnum[0];
    xg_edge_shapes<D, P>(sw ? lam[0] : lam[1], sw ? lam[1] : lam[0], n, sink);
    return n;
  }
  constexpr int NED = (D == 2) ? 3 : 6;
#pragma unroll
  for (int e = 0; e < NED; e++) {
    const int ea = (D == 2) ? XG_TRIG_EDGE_A(e) : XG_TET_EDGE_A(e);
    const int eb = (D == 2) ? XG_TRIG_EDGE_B(e) : XG_TET_EDGE_B(e);
    const bool sw = (P > 2) && vnum[ea] > vnum[eb];   // P == 2: 0.5 ((le-ls)^2 - (le+ls)^2) is swap-invariant bit for bit
    AD<D> la = lam[0], lb = lam[0];
#pragma unroll
    for (int i = 0; i < D + 1; i++) { if (i == ea) la = lam[i]; if (i == eb) lb = lam[i]; }
    xg_edge_shapes<D, P>(sw ? lb : la, sw ? la : lb, n, sink);
  }
  if (P < 3) return n;
  if (D == 2) {
    xg_face_shapes<D, P>(lam, 0, 1, 2, vnum, n, sink);
  } else if (D == 3) {
    xg_face_shapes<D, P>(lam, 3, 1, 2, vnum, n, sink);
    xg_face_shapes<D, P>(lam, 3, 2, 0, vnum, n, sink);
    xg_face_shapes<D, P>(lam, 3, 0, 1, vnum, n, sink);
    xg_face_shapes<D, P>(lam, 0, 2, 1, vnum, n, sink);
    if (P >= 4) {
      constexpr int N = (P >= 4) ? P - 4 : 0;
      AD<D> bub = (lam[0] * lam[1]) * (lam[2] * lam[D]);
      AD<D> Pa[N + 1], Qa[N + 1], Ra[N + 1];
      xg_leg_scaled<D, N>(lam[1] - lam[0], lam[0] + lam[1], Pa);
      xg_leg_scaled<D, N>(2.0 * lam[2] - ad_const<D>(1.0), ad_const<D>(1.0), Qa);
      xg_leg_scaled<D, N>(2.0 * lam[D] - ad_const<D>(1.0), ad_const<D>(1.0), Ra);
#pragma unroll
      for (int i = 0; i <= N; i++)
#pragma unroll
        for (int j = 0; j <= N - i; j++)
#pragma unroll
          for (int k = 0; k <= N - i - j; k++) sink(n++, bub * (Pa[i] * (Qa[j] * Ra[k])));
    }
  }
  return n;
}

// runtime-order dispatch (orders 1..XG_MAX_ORDER)
#define XG_MAX_ORDER 4
template <int D, class Sink>
XG_HD int xg_h1_shapes_rt(int p, const int* vnum, const double* xhat, Sink& sink)
{
  switch (p) {
    case 1: return xg_h1_shapes<D, 1>(vnum, xhat, sink);
    case 2: return xg_h1_shapes<D, 2>(vnum, xhat, sink);
    case 3: return xg_h1_shapes<D, 3>(vnum, xhat, sink);
    case 4: return xg_h1_shapes<D, 4>(vnum, xhat, sink);
  }
  return 0;
}
