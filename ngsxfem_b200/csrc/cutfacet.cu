// cutfacet.cu — matrices of cut integrals over interior facets and element boundaries:
//   SymbolicCutFacetBilinearFormIntegrator::T_CalcFacetMatrix   xfem/symboliccutbfi.cpp:886-1290  dCut(..., skeleton=True)
//   SymbolicCutBilinearFormIntegrator::T_CalcElementMatrixEBAdd xfem/symboliccutbfi.cpp:488-883   dCut(..., element_boundary=True)
// for affine simplices.  The cut rule of every listed FACET (a (dim-1)-simplex; xgpu_cutquad_* on the facet-vertex table)
// is an input: points in the facet's reference coordinates, weights of the reference facet.  One warp per facet:
//   * every lane rebuilds the (tiny) geometry: Jacobians of the neighbour element(s), outward unit normal nF and mesh
//     size of the FIRST element, facet measure, where the facet's vertices sit in each element;
//   * lanes take one rule point each: barycentric coordinates of the point in each element (both neighbours see the same
//     physical point: no inversion), hierarchical H1 shapes with the element's own vertex numbers, traces and normal
//     derivatives staged in shared memory;
//   * the (ns nd) x (ns nd) matrix entries are strided over the lanes; the closed integrand menu — sum of
//     coef h^k Op_test(v|side) Op_trial(u|side), Op in {value, nF.grad}, side in {this element, .Other()} — is folded into
//     one 2x2x2x2 coefficient table per facet, an entry costs 4 FMAs per point.
// Not a hot path of BASELINE config 2 (facet terms appear in the DG / ghost-penalty variants, demos/fictdom_dg.py).
#include "common.cuh"
#include "cutgeom.cuh"

namespace {

#define XF_MAXT 16
struct FTermsDev {
  int n;
  int tk[XF_MAXT], ts[XF_MAXT], ek[XF_MAXT], es[XF_MAXT], hpow[XF_MAXT];   // trial kind / side, test kind / side
  double coef[XF_MAXT];
};

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ double ipow_(double h, int p)
{
  if (p == 0) return 1.0;
  double r = 1.0;
  const int ap = p < 0 ? -p : p;
  for (int i = 0; i < ap; i++) r *= h;
  return p < 0 ? 1.0 / r : r;
}

template <int D>
struct FacetSink {
  double* val;        // nd values
  double* dn;         // nd normal derivatives
  const double* m;    // m_j = sum_i Fi[j][i] nF[i]: dn = d . m
  __device__ __forceinline__ void operator()(int i, const AD<D>& a)
  {
    val[i] = a.v;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) s += a.d[j] * m[j];
    dn[i] = s;
  }
};

template <int D, int P>
__global__ void __launch_bounds__(128) k_facet_elmat(int64_t n, const int32_t* __restrict__ facet_vtx, const int32_t* __restrict__ facet_els,
                                                     const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                                     const int64_t* __restrict__ offsets, const double* __restrict__ pts,
                                                     const double* __restrict__ wts, FTermsDev T, int ns, int CH, double* __restrict__ mats)
{
  constexpr int ND = (D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2
                              : 4 + 6 * (P - 1) + 4 * ((P - 1) * (P - 2) / 2) + (P - 1) * (P - 2) * (P - 3) / 6;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per point: [side][kind][ND] + weight
  constexpr int PS = 4 * ND + 1;
  double* Bw = smem + (size_t)wib * CH * PS;
  const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (f >= n) return;
  const int N = ns * ND;
  double* M = mats + (size_t)f * N * N;
  for (int k = lane; k < N * N; k += 32) M[k] = 0.0;
  int fv[D];
#pragma unroll
  for (int i = 0; i < D; i++) fv[i] = __ldg(facet_vtx + f * D + i);
  int v[2][D + 1], loc[2][D + 1];
  double Fi[2][D * D], mvec[2][D], nF[3] = {0, 0, 0}, h = 0.0, JF = 0.0;
  for (int s = 0; s < ns; s++) {
    const int64_t el = facet_els[2 * f + s];
    double vx[D + 1][3], F[D * D];
#pragma unroll
    for (int k = 0; k < D + 1; k++) {
      v[s][k] = __ldg(el_vtx + el * (D + 1) + k);
      loc[s][k] = -1;
#pragma unroll
      for (int i = 0; i < D; i++) if (v[s][k] == fv[i]) loc[s][k] = i;
#pragma unroll
      for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[s][k] * D + c) : 0.0;
    }
    xg_affine_jac<D>(vx, F);
    const double det = xg_inv<D>(F, Fi[s]);
    if (s == 0) {
      h = (D == 2) ? sqrt(fabs(det)) : cbrt(fabs(det));
      // outward normal: -grad(lam_o), o = the vertex opposite to the facet
      int o = 0;
#pragma unroll
      for (int k = 0; k < D + 1; k++) if (loc[0][k] < 0) o = k;
      double nn = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) t += Fi[0][j * D + i] * ((o < D) ? (j == o ? 1.0 : 0.0) : -1.0);
        nF[i] = -t; nn += t * t;
      }
      nn = sqrt(nn);
#pragma unroll
      for (int i = 0; i < D; i++) nF[i] /= nn;
      // facet measure relative to the reference facet, from the facet's vertices in table order
      double pf[D][3];
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int k = 0; k < D + 1; k++) if (loc[0][k] == i) { pf[i][0] = vx[k][0]; pf[i][1] = vx[k][1]; pf[i][2] = vx[k][2]; }
      double a[3], b[3];
#pragma unroll
      for (int c = 0; c < 3; c++) { a[c] = pf[0][c] - pf[D - 1][c]; b[c] = (D == 3) ? pf[1][c] - pf[D - 1][c] : 0.0; }
      if (D == 2) JF = sqrt(a[0] * a[0] + a[1] * a[1]);
      else {
        const double c0 = a[1] * b[2] - a[2] * b[1], c1 = a[2] * b[0] - a[0] * b[2], c2 = a[0] * b[1] - a[1] * b[0];
        JF = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
      }
    }
  }
  for (int s = 0; s < ns; s++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) t += Fi[s][j * D + i] * nF[i];
      mvec[s][j] = t;
    }
  // coefficient table K[test side][trial side][test kind][trial kind]
  double K[2][2][2][2];
#pragma unroll
  for (int a = 0; a < 16; a++) (&K[0][0][0][0])[a] = 0.0;
  for (int t = 0; t < T.n; t++) K[T.es[t]][T.ts[t]][T.ek[t]][T.tk[t]] += T.coef[t] * ipow_(h, T.hpow[t]);
  const int64_t q0 = offsets[f];
  const int npe = (int)(offsets[f + 1] - q0);
  for (int c0 = 0; c0 < npe; c0 += CH) {
    const int cnt = min(CH, npe - c0);
    __syncwarp();
    for (int ql = lane; ql < cnt; ql += 32) {
      double* B = Bw + (size_t)ql * PS;
      const double* xq = pts + 3 * (q0 + c0 + ql);
      double mu[D], sm = 1.0;
#pragma unroll
      for (int i = 0; i < D - 1; i++) { mu[i] = xq[i]; sm -= mu[i]; }
      mu[D - 1] = sm;
      for (int s = 0; s < ns; s++) {
        double xh[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < D; k++)
#pragma unroll
          for (int i = 0; i < D; i++) if (loc[s][k] == i) xh[k] = mu[i];
        FacetSink<D> sink{B + (2 * s) * ND, B + (2 * s + 1) * ND, mvec[s]};
        xg_h1_shapes<D, P>(v[s], xh, sink);
      }
      B[4 * ND] = wts[q0 + c0 + ql] * JF;
    }
    __syncwarp();
    for (int k = lane; k < N * N; k += 32) {
      const int r = k / N, c = k - r * N;
      const int es = r / ND, i = r - es * ND, us = c / ND, j = c - us * ND;
      const double k00 = K[es][us][0][0], k01 = K[es][us][0][1], k10 = K[es][us][1][0], k11 = K[es][us][1][1];
      double acc = 0.0;
      for (int q = 0; q < cnt; q++) {
        const double* B = Bw + (size_t)q * PS;
        const double tv = B[(2 * es) * ND + i], td = B[(2 * es + 1) * ND + i];
        const double uv = B[(2 * us) * ND + j], ud = B[(2 * us + 1) * ND + j];
        acc += B[4 * ND] * (tv * (k00 * uv + k01 * ud) + td * (k10 * uv + k11 * ud));
      }
      M[k] += acc;
    }
  }
}

template <int D, int P>
int launch_facet(xgpu_ctx* ctx, int64_t n, const int32_t* facet_vtx, const int32_t* facet_els, const int32_t* el_vtx,
                 const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, const FTermsDev& T, int ns,
                 double* mats)
{
  constexpr int ND = (D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2
                              : 4 + 6 * (P - 1) + 4 * ((P - 1) * (P - 2) / 2) + (P - 1) * (P - 2) * (P - 3) / 6;
  const int CH = ND <= 10 ? 32 : 8;
  const size_t smem = sizeof(double) * 4 * (size_t)CH * (4 * ND + 1);
  XGPU_CUDA(ctx, cudaFuncSetAttribute(k_facet_elmat<D, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_facet_elmat<D, P><<<nblk(n, 4), 128, smem, ctx->stream>>>(n, facet_vtx, facet_els, el_vtx, vtx_xyz, offsets, pts, wts, T, ns, CH, mats);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

extern "C" int xgpu_facet_elmat(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* facet_vtx, const int32_t* facet_els,
                                const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts,
                                const double* wts, const xgpu_fterm* terms_host, int nterms, int nsides, double* mats)
{
  XGPU_NVTX(ctx, "xgpu_facet_elmat");
  if (!ctx || dim < 2 || dim > 3 || fe_order < 1 || fe_order > 4 || n < 0 || (nsides != 1 && nsides != 2) || nterms < 0 ||
      nterms > XF_MAXT || (nterms > 0 && !terms_host) || (n > 0 && (!facet_vtx || !facet_els || !offsets || !mats)))
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_facet_elmat: bad argument (dim 2, 3; orders 1..4; at most 16 terms)");
  if (ctx->def_order != 0) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_facet_elmat: the mesh is deformed (affine elements only)");
  if (n == 0) return XGPU_OK;
  FTermsDev T;
  T.n = nterms;
  for (int t = 0; t < nterms; t++) {
    const xgpu_fterm& a = terms_host[t];
    if ((a.trial_kind | 1) != 1 || (a.test_kind | 1) != 1 || a.trial_side < 0 || a.trial_side >= nsides || a.test_side < 0 || a.test_side >= nsides)
      return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_facet_elmat: term kinds are 0 (value) / 1 (normal derivative), sides 0 .. nsides-1");
    T.tk[t] = a.trial_kind; T.ts[t] = a.trial_side; T.ek[t] = a.test_kind; T.es[t] = a.test_side; T.hpow[t] = a.hpow; T.coef[t] = a.coef;
  }
#define XF_CASE(DD, PP) if (dim == DD && fe_order == PP) return launch_facet<DD, PP>(ctx, n, facet_vtx, facet_els, el_vtx, vtx_xyz, offsets, pts, wts, T, nsides, mats);
  XF_CASE(2, 1) XF_CASE(2, 2) XF_CASE(2, 3) XF_CASE(2, 4) XF_CASE(3, 1) XF_CASE(3, 2) XF_CASE(3, 3) XF_CASE(3, 4)
#undef XF_CASE
  return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_facet_elmat: unsupported (dim, order)");
}
