#pragma once
// facetpatch.cuh — (templates shared by facetpatch.cu and facetpatch_dmma_*.cu; split for build time)
// facetpatch.cu — ghost-penalty facet-patch matrices
// (SymbolicFacetPatchBilinearFormIntegrator::T_CalcFacetMatrix, xfem/symboliccutbfi.cpp:1982-2377,
//  point mapping MapPatchIntegrationPoint :1747-1846) for the integrand
//      coef * h^hpow * (u - u.Other()) * (v - v.Other()).
// One WARP per facet.  The patch rule is [rule on el1 ; rule on el2]; every lane maps one patch
// point into the other element's reference coordinates (affine elements: the affine solve the
// reference uses as Newton start is already exact, Newton exits after its first step),
// evaluates both bases and stages the jump vector J = [phi^1(x1) ; -phi^2(x2)] in shared memory;
// lanes own 2x2 blocks of the (2nd)^2 matrix and accumulate c_q J J^T over the staged points.
#include "common.cuh"
#include "cutgeom.cuh"
#include "deform.cuh"

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

struct ValSink {
  double* J; double sgn;
  template <int D> __device__ __forceinline__ void operator()(int i, const AD<D>& a) { J[i] = sgn * a.v; }
};

template <int D>
__device__ __forceinline__ void load_geom(const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz, int64_t el,
                                          int* v, double (*vx)[3], double* F, double* Fi, double* meas)
{
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
    v[k] = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
    for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[k] * D + c) : 0.0;
  }
  xg_affine_jac<D>(vx, F);
  *meas = fabs(xg_inv<D>(F, Fi));
}

template <int D>
__device__ __forceinline__ void map_fwd(const double (*vx)[3], const double* xh, double* x)
{
  double s = 1.0;
#pragma unroll
  for (int k = 0; k < D; k++) s -= xh[k];
#pragma unroll
  for (int c = 0; c < D; c++) {
    double t = s * vx[D][c];
#pragma unroll
    for (int k = 0; k < D; k++) t += xh[k] * vx[k][c];
    x[c] = t;
  }
}

template <int D>
__device__ __forceinline__ void map_inv(const double (*vx)[3], const double* Fi, const double* x, double* xh)
{
#pragma unroll
  for (int i = 0; i < D; i++) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) t += Fi[i * D + j] * (x[j] - vx[D][j]);
    xh[i] = t;
  }
  for (int i = D; i < 3; i++) xh[i] = 0.0;
}

// dFacetPatch(downscale = s), ScaleAndMapIntegrationRule (symboliccutbfi.cpp:1849-1885): the volume rule of an element is
// scaled by s towards the centre of the common facet, xhat <- s xhat + (1 - s) c_facet, and the weight by s^2 (the
// reference's factor in every dimension).  c_facet = mean of the reference vertices of the facet = of the vertices of `v`
// that `vo` has too; reference vertex k < D is e_k, vertex D the origin.  s = 1: shift 0, bit-identical points.
template <int D>
__device__ __forceinline__ void fp_shift(const int* v, const int* vo, double s, double* sh)
{
#pragma unroll
  for (int k = 0; k < 3; k++) sh[k] = 0.0;
  if (s == 1.0) return;
#pragma unroll
  for (int k = 0; k < D; k++) {
    bool shared = false;
#pragma unroll
    for (int j = 0; j < D + 1; j++) shared |= (v[k] == vo[j]);
    if (shared) sh[k] = (1.0 - s) / D;
  }
}

template <int D>
__device__ __forceinline__ void fp_point(const double* __restrict__ rp, const double* __restrict__ rw, int q, double s,
                                         const double* sh, double* x, double* w)
{
#pragma unroll
  for (int c = 0; c < 3; c++) x[c] = (c < D) ? s * rp[3 * q + c] + sh[c] : rp[3 * q + c];
  *w = rw[q] * s * s;
}

// x(xhat) and F(xhat) of a (deformed) element: affine part from the vertices + deformation
// DO > 0: the deformation order at compile time (one set of shape functions per inlined copy instead of all four)
template <int D, int DO = 0>
__device__ __forceinline__ void trafo_eval(const double (*vx)[3], const double* Faff, const int* v, int def_order,
                                           const double* defs, const double* xh, double* F, double* x)
{
  map_fwd<D>(vx, xh, x);
#pragma unroll
  for (int k = 0; k < D * D; k++) F[k] = Faff[k];
  if constexpr (DO > 0) {
    DefJacSink<D> sink{defs, F, x};
    xg_h1_shapes<D, DO>(v, xh, sink);
  } else {
    xg_def_apply<D>(def_order, v, xh, defs, F, x);
  }
}

// Geometry of the two elements of a facet for the deformed path, per warp in shared memory: the data are the same for all
// lanes, and a `from` / `to` element chosen by POINTER keeps one copy of the mapping code (each inlined evaluation of a
// deformed transformation carries the shape functions of all deformation orders: with the elements in registers and the
// forward / backward cases as two branches the kernels were 20 000 - 25 000 SASS instructions and instruction-fetch bound,
// profiles/r02_notes.md).  Layout in doubles: vertices [4][3], affine Jacobian [9], vertex numbers (4 ints).
constexpr int FP_GEO = 24;

template <int D>
struct FpEl {
  const double (*vx)[3];
  const double* F;
  const int* v;
  const double* defs;
};

template <int D>
__device__ __forceinline__ FpEl<D> fp_el(const double* geo, const double* defs)
{
  return FpEl<D>{reinterpret_cast<const double (*)[3]>(geo), geo + 12, reinterpret_cast<const int*>(geo + 21), defs};
}

template <int D>
__device__ __forceinline__ void fp_store_geo(double* geo, const double (*vx)[3], const double* F, const int* v)
{
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
#pragma unroll
    for (int c = 0; c < 3; c++) geo[3 * k + c] = vx[k][c];
    reinterpret_cast<int*>(geo + 21)[k] = v[k];
  }
#pragma unroll
  for (int k = 0; k < D * D; k++) geo[12 + k] = F[k];
}

// MapPatchIntegrationPoint<D> (symboliccutbfi.cpp:1747-1846, non space-time) on deformed elements:
// (xf, wf) on the `from` element -> xt, wt on the `to` element; affine start value, Newton with the
// reference's constants (ngsxstd.cpp:66-82: eps 1e-12, 15 iterations, max distance 10), low-order
// guess kept on failure.
// the point independent part of the mapping INTO element t: the affine start value uses the images of the
// reference vertices (out: pc[0..2] = a0, pc[3..3+D*D) = A^{-1}, pc[12] = |det A|)
template <int D, int DO>
__device__ __forceinline__ void patch_consts(const FpEl<D>& t, int def_order, double* pc)
{
  double F[D * D], a0[3] = {0, 0, 0}, A[D * D], Ainv[D * D];
#pragma unroll
  for (int k = 0; k < D * D; k++) A[k] = 0.0;
#pragma unroll 1
  for (int d = -1; d < D; d++) {  // the origin, then the unit points: ONE inlined evaluation
    double e[3], ai[3] = {0, 0, 0};
    e[0] = (d == 0) ? 1.0 : 0.0; e[1] = (d == 1) ? 1.0 : 0.0; e[2] = (d == 2) ? 1.0 : 0.0;
    trafo_eval<D, DO>(t.vx, t.F, t.v, def_order, t.defs, e, F, ai);
    if (d < 0) {
      a0[0] = ai[0]; a0[1] = ai[1]; a0[2] = ai[2];
    } else {
#pragma unroll
      for (int c = 0; c < D; c++)
        if (d == c) {
#pragma unroll
          for (int i = 0; i < D; i++) A[i * D + c] = ai[i] - a0[i];
        }
    }
  }
  const double w00 = fabs(xg_inv<D>(A, Ainv));
  pc[0] = a0[0]; pc[1] = a0[1]; pc[2] = a0[2];
#pragma unroll
  for (int k = 0; k < D * D; k++) pc[3 + k] = Ainv[k];
  pc[12] = w00;
}

// One patch point on deformed elements as a state machine around ONE inlined evaluation of the deformed transformation:
// stage 0 the point on the `from` element (physical point, measure), stage 1 Newton on the `to` element
// (MapPatchIntegrationPoint), stage 2 the measure at the point's coordinates on element 1.  The arithmetic and its order are
// those of the three separate evaluations; the code of the shape functions exists once.
// out: x1 / x2 = the point in the reference coordinates of element 1 / 2, w1 = its weight on element 1, measq = |det F_1(x1)|
template <int D, int DO>
__device__ __forceinline__ void fp_map_one(const FpEl<D>& ef, const FpEl<D>& et, const FpEl<D>& e1, bool fwd, int def_order,
                                           const double* pc, const double* xf, double wf, double* x1, double* x2, double* w1,
                                           double* measq)
{
  double F[D * D], Fi[D * D], vec[3] = {0, 0, 0};
  double a0[3] = {pc[0], pc[1], pc[2]}, Ainv[D * D], x00[3] = {0, 0, 0}, x0[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < D * D; k++) Ainv[k] = pc[3 + k];
  const double w00 = pc[12];
  double h = 0.0, mipw = 0.0, w = 0.0, diffn = 0.0;
  int its = 0, stage = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) { x1[c] = 0.0; x2[c] = 0.0; }
#pragma unroll 1
  for (;;) {
    FpEl<D> el;
    el.vx = (stage == 0) ? ef.vx : ((stage == 1) ? et.vx : e1.vx);
    el.F = (stage == 0) ? ef.F : ((stage == 1) ? et.F : e1.F);
    el.v = (stage == 0) ? ef.v : ((stage == 1) ? et.v : e1.v);
    el.defs = (stage == 0) ? ef.defs : ((stage == 1) ? et.defs : e1.defs);
    double xe[3], xm[3] = {0, 0, 0};
#pragma unroll
    for (int c = 0; c < 3; c++) xe[c] = (stage == 0) ? xf[c] : ((stage == 1) ? x0[c] : x1[c]);
    trafo_eval<D, DO>(el.vx, el.F, el.v, def_order, el.defs, xe, F, xm);
    const double det = xg_inv<D>(F, Fi);
    if (stage == 0) {
      vec[0] = xm[0]; vec[1] = xm[1]; vec[2] = xm[2];
      h = sqrt(det);
      mipw = wf * fabs(det);
#pragma unroll
      for (int i = 0; i < D; i++) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) t += Ainv[i * D + j] * (vec[j] - a0[j]);
        x00[i] = t; x0[i] = t;
      }
      stage = 1;
    } else if (stage == 1) {
      double diff[3];
      diffn = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) { diff[i] = vec[i] - xm[i]; diffn += diff[i] * diff[i]; }
      diffn = sqrt(diffn);
#pragma unroll
      for (int i = 0; i < D; i++) {
        double u = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) u += Fi[i * D + j] * diff[j];
        x0[i] += u;
      }
      its++;
      w = fabs(det);
      if (!(diffn > 1e-12 * h && its < 15)) {  // the reference's loop ends here (ngsxstd.cpp:66-82: eps 1e-12, 15 iterations)
        double dist = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) dist += (x0[i] - x00[i]) * (x0[i] - x00[i]);
        dist = sqrt(dist);
        const bool failed = its >= 15 || diffn > 1e-12 * h || dist > 10.0;  // low-order guess kept on failure (max distance 10)
        const double wt = failed ? mipw / w00 : mipw / w;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const double xtc = failed ? x00[c] : x0[c];
          x1[c] = fwd ? xf[c] : xtc;
          x2[c] = fwd ? xtc : xf[c];
        }
        *w1 = fwd ? wf : wt;
        stage = 2;
      }
    } else {
      *measq = fabs(det);
      break;
    }
  }
}

// all patch points of one facet on deformed elements: (x1, x2, weight) -> mp[7 l ..], once per facet, all lanes busy.
// geo: the two geometry blocks (element 1, element 2); a point of element 1's rule is mapped into element 2 and vice versa
// through the same code, the elements chosen by pointer
template <int D, int DO>
__device__ __forceinline__ void fp_map_points(int lane, int np, const double* __restrict__ rp, const double* __restrict__ rw,
                                              const double* geo, const double* defs1, const double* defs2,
                                              int def_order, double coef, int hpow, double scale, const double* sh1,
                                              const double* sh2, double* pcs, double* mp)
{
  if (lane < 2)  // lane 0: constants of the map INTO element 2, lane 1: into element 1
    patch_consts<D, DO>(fp_el<D>(geo + (lane == 0 ? FP_GEO : 0), lane == 0 ? defs2 : defs1), def_order, pcs + 16 * lane);
  __syncwarp();
  const FpEl<D> e1 = fp_el<D>(geo, defs1);
#pragma unroll 1
  for (int l = lane; l < 2 * np; l += 32) {
    const bool fwd = l < np;
    const FpEl<D> ef = fp_el<D>(geo + (fwd ? 0 : FP_GEO), fwd ? defs1 : defs2);
    const FpEl<D> et = fp_el<D>(geo + (fwd ? FP_GEO : 0), fwd ? defs2 : defs1);
    double shf[3], xf[3], wf;
#pragma unroll
    for (int c = 0; c < 3; c++) shf[c] = fwd ? sh1[c] : sh2[c];
    fp_point<D>(rp, rw, fwd ? l : l - np, scale, shf, xf, &wf);
    double x1[3], x2[3], w1, measq;
    fp_map_one<D, DO>(ef, et, e1, fwd, def_order, pcs + (fwd ? 0 : 16), xf, wf, x1, x2, &w1, &measq);
    const double hq = xg_mesh_size<D>(measq);
    double hpq = 1.0;
    { const int ap = hpow < 0 ? -hpow : hpow; for (int i = 0; i < ap; i++) hpq *= hq; if (hpow < 0) hpq = 1.0 / hpq; }
    double* m = mp + 7 * l;
    m[0] = x1[0]; m[1] = x1[1]; m[2] = x1[2]; m[3] = x2[0]; m[4] = x2[1]; m[5] = x2[2];
    m[6] = coef * hpq * (w1 * measq);
  }
  __syncwarp();
}

constexpr int FP_MAXB = 13;

template <int D, bool DEF>
__global__ void __launch_bounds__(128) k_facetpatch(
    int64_t n, const int32_t* __restrict__ facets, const int32_t* __restrict__ facet_els,
    const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
    const double* __restrict__ rp, const double* __restrict__ rw, int np,
    int p, int nd, int CH, double coef, int hpow, double scale, double* __restrict__ mats, DefDev df)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int N = 2 * nd, Np = (N + 1) & ~1, nbk = Np >> 1, nblocks = nbk * nbk;
  const int PS = Np + 1;
  const int npatch_ = 2 * np;
  double* Bw = smem + (size_t)wib * (CH * PS + (DEF ? 2 * XG_DEFS + 32 + 2 * FP_GEO + 7 * npatch_ : 0));
  double* defs1 = Bw + (size_t)CH * PS;
  double* defs2 = defs1 + XG_DEFS;
  double* pcs = defs2 + XG_DEFS;   // mapping constants: [0..15] into element 2, [16..31] into element 1
  double* geo = pcs + 32;          // geometry blocks of the two elements (deformed path)
  double* mp = geo + 2 * FP_GEO;   // mapped patch points: x1[3], x2[3], weight  (DEF: computed ONCE per facet)
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int npatch = 2 * np;
  for (int64_t f = warp0; f < n; f += nwarps) {
    const int64_t fac = facets ? (int64_t)facets[f] : f;
    const int e1 = __ldg(facet_els + 2 * fac), e2 = __ldg(facet_els + 2 * fac + 1);
    double* M = mats + (size_t)f * N * N;
    if (e2 < 0) {  // the reference throws for LocalFacetNr2==-1 (:1992); boundary facets yield zero
      for (int k = lane; k < N * N; k += 32) M[k] = 0.0;
      continue;
    }
    int v1[D + 1], v2[D + 1];
    double vx1[D + 1][3], vx2[D + 1][3], F1[D * D], Fi1[D * D], F2[D * D], Fi2[D * D], meas1, meas2;
    load_geom<D>(el_vtx, vtx_xyz, e1, v1, vx1, F1, Fi1, &meas1);
    load_geom<D>(el_vtx, vtx_xyz, e2, v2, vx2, F2, Fi2, &meas2);
    double sh1[3], sh2[3];
    fp_shift<D>(v1, v2, scale, sh1);
    fp_shift<D>(v2, v1, scale, sh2);
    if (DEF) {
      __syncwarp();
      if (lane == 0) { fp_store_geo<D>(geo, vx1, F1, v1); fp_store_geo<D>(geo + FP_GEO, vx2, F2, v2); }
      xg_def_gather<D>(df, e1, lane, defs1);
      xg_def_gather<D>(df, e2, lane, defs2);
      __syncwarp();
      // the usual case (deformation order = 2 or 1 on these kernels' orders) with the order at compile time
      if (df.order == 2) fp_map_points<D, 2>(lane, np, rp, rw, geo, defs1, defs2, df.order, coef, hpow, scale, sh1, sh2, pcs, mp);
      else if (df.order == 1) fp_map_points<D, 1>(lane, np, rp, rw, geo, defs1, defs2, df.order, coef, hpow, scale, sh1, sh2, pcs, mp);
      else fp_map_points<D, 0>(lane, np, rp, rw, geo, defs1, defs2, df.order, coef, hpow, scale, sh1, sh2, pcs, mp);
    }
    const double h = (D == 1) ? meas1 : ((D == 2) ? sqrt(meas1) : cbrt(meas1));
    double hp = 1.0;
    { const int ap = hpow < 0 ? -hpow : hpow; for (int i = 0; i < ap; i++) hp *= h; if (hpow < 0) hp = 1.0 / hp; }
    for (int panel = 0; panel * FP_MAXB * 32 < nblocks; panel++) {
      double acc[FP_MAXB][4];
#pragma unroll
      for (int b = 0; b < FP_MAXB; b++) { acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.0; }
      for (int c0 = 0; c0 < npatch; c0 += CH) {
        const int cs = min(CH, npatch - c0);
        __syncwarp();
        if (lane < cs) {
          const int l = c0 + lane;
          double* B = Bw + (size_t)lane * PS;
          double x1[3], x2[3], x[3], w;
          double cwq = coef * hp;
          if (DEF) {
            const double* m = mp + 7 * l;
            x1[0] = m[0]; x1[1] = m[1]; x1[2] = m[2]; x2[0] = m[3]; x2[1] = m[4]; x2[2] = m[5];
            w = m[6]; cwq = 1.0;
          } else if (l < np) {
            fp_point<D>(rp, rw, l, scale, sh1, x1, &w);
            map_fwd<D>(vx1, x1, x);
            map_inv<D>(vx2, Fi2, x, x2);
            w = w * meas1;
          } else {
            fp_point<D>(rp, rw, l - np, scale, sh2, x2, &w);
            map_fwd<D>(vx2, x2, x);
            map_inv<D>(vx1, Fi1, x, x1);
            w = w * meas2;
          }
          for (int i = N; i < Np; i++) B[i] = 0.0;
          ValSink s1{B, 1.0}, s2{B + nd, -1.0};
          xg_h1_shapes_rt<D>(p, v1, x1, s1);
          xg_h1_shapes_rt<D>(p, v2, x2, s2);
          B[Np] = cwq * w;
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < FP_MAXB; b++) {
          const int blk = (panel * FP_MAXB + b) * 32 + lane;
          if (blk < nblocks) {
            const int i0 = 2 * (blk / nbk), j0 = 2 * (blk % nbk);
            for (int q = 0; q < cs; q++) {
              const double* B = Bw + (size_t)q * PS;
              const double cw = B[Np];
              const double a0 = cw * B[i0], a1 = cw * B[i0 + 1], b0 = B[j0], b1 = B[j0 + 1];
              acc[b][0] += a0 * b0; acc[b][1] += a0 * b1; acc[b][2] += a1 * b0; acc[b][3] += a1 * b1;
            }
          }
        }
      }
#pragma unroll
      for (int b = 0; b < FP_MAXB; b++) {
        const int blk = (panel * FP_MAXB + b) * 32 + lane;
        if (blk < nblocks) {
          const int i0 = 2 * (blk / nbk), j0 = 2 * (blk % nbk);
#pragma unroll
          for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
              const int i = i0 + a, j = j0 + c;
              if (i < N && j < N) M[i * N + j] = acc[b][2 * a + c];
            }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- FP64 tensor-core contraction (order >= 3)
// M = sum_q c_q J_q J_q^T with J_q = [phi^1(x1_q) ; -phi^2(x2_q)] is a GEMM over the patch points: FOUR points are one
// k-step of the FP64 tensor instruction (mma.sync m8n8k4: A = c J (8 dofs x 4 points), B = J^T), on the symmetric set of
// 8x8 tiles (north_star (d): the facet-patch matrices reuse the batched contraction of the element matrices).  The jump
// vectors are staged ONCE per chunk of points (no accumulator panels: the 70x70 matrix of two P4 tets is 45 tile pairs =
// 90 accumulator registers), row stride = 4 mod 16 doubles so that the four points of a fragment load hit disjoint banks.
__device__ __forceinline__ void fp_dmma884(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int D, int P>
struct FpGeom {
  static constexpr int ND = (D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2
                                     : 4 + 6 * (P - 1) + 4 * ((P - 1) * (P - 2) / 2) + (P - 1) * (P - 2) * (P - 3) / 6;
  static constexpr int N = 2 * ND, NT = (N + 7) / 8, NDP = 8 * NT;
  static constexpr int PS = NDP + ((NDP % 16 == 0) ? 4 : 12);  // = 4 mod 16, slot NDP holds the point's weight
  static constexpr int NPAIR = NT * (NT + 1) / 2;
};

template <int D, int P, bool DEF>
__global__ void __launch_bounds__(128) k_facetpatch_dmma(
    int64_t n, const int32_t* __restrict__ facets, const int32_t* __restrict__ facet_els,
    const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
    const double* __restrict__ rp, const double* __restrict__ rw, int np,
    int CH, double coef, int hpow, double scale, double* __restrict__ mats, DefDev df)
{
  using G = FpGeom<D, P>;
  constexpr int ND = G::ND, N = G::N, NT = G::NT, NDP = G::NDP, PS = G::PS, NPAIR = G::NPAIR;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const int npatch = 2 * np, npad = (npatch + 3) & ~3;
  double* Bw = smem + (size_t)wib * (CH * PS + (DEF ? 2 * XG_DEFS + 32 + 2 * FP_GEO + 7 * npatch : 0));
  double* defs1 = Bw + (size_t)CH * PS;
  double* defs2 = defs1 + XG_DEFS;
  double* pcs = defs2 + XG_DEFS;
  double* geo = pcs + 32;
  double* mp = geo + 2 * FP_GEO;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t f = warp0; f < n; f += nwarps) {
    const int64_t fac = facets ? (int64_t)facets[f] : f;
    const int e1 = __ldg(facet_els + 2 * fac), e2 = __ldg(facet_els + 2 * fac + 1);
    double* M = mats + (size_t)f * N * N;
    if (e2 < 0) {
      for (int k = lane; k < N * N; k += 32) M[k] = 0.0;
      continue;
    }
    int v1[D + 1], v2[D + 1];
    double vx1[D + 1][3], vx2[D + 1][3], F1[D * D], Fi1[D * D], F2[D * D], Fi2[D * D], meas1, meas2;
    load_geom<D>(el_vtx, vtx_xyz, e1, v1, vx1, F1, Fi1, &meas1);
    load_geom<D>(el_vtx, vtx_xyz, e2, v2, vx2, F2, Fi2, &meas2);
    double sh1[3], sh2[3];
    fp_shift<D>(v1, v2, scale, sh1);
    fp_shift<D>(v2, v1, scale, sh2);
    if (DEF) {
      __syncwarp();
      if (lane == 0) { fp_store_geo<D>(geo, vx1, F1, v1); fp_store_geo<D>(geo + FP_GEO, vx2, F2, v2); }
      xg_def_gather<D>(df, e1, lane, defs1);
      xg_def_gather<D>(df, e2, lane, defs2);
      __syncwarp();
      // the usual case (deformation order = FE order) with the order at compile time
      if (df.order == P) fp_map_points<D, P>(lane, np, rp, rw, geo, defs1, defs2, df.order, coef, hpow, scale, sh1, sh2, pcs, mp);
      else fp_map_points<D, 0>(lane, np, rp, rw, geo, defs1, defs2, df.order, coef, hpow, scale, sh1, sh2, pcs, mp);
    }
    const double h = xg_mesh_size<D>(meas1);
    double hp = 1.0;
    { const int ap = hpow < 0 ? -hpow : hpow; for (int i = 0; i < ap; i++) hp *= h; if (hpow < 0) hp = 1.0 / hp; }
    double acc[NPAIR][2];
#pragma unroll
    for (int b = 0; b < NPAIR; b++) acc[b][0] = acc[b][1] = 0.0;
    for (int c0 = 0; c0 < npad; c0 += CH) {
      const int cs = min(CH, npad - c0);  // multiple of 4
      __syncwarp();
      // staging of the jump vectors: with at most 16 points in the chunk the two bases of a point go to two lanes (lane q and
      // q + 16: all 32 lanes busy), else one lane evaluates them one after the other; ONE inlined copy of the shape functions
      const bool split = 2 * cs <= 32;
      const int q = split ? (lane & 15) : lane;
      const int half0 = split ? (lane >> 4) : 0, nrep = split ? 1 : 2;
      if (q < cs) {
        const int l = c0 + q;
        double* B = Bw + (size_t)q * PS;
        if (l >= npatch) {
          if (half0 == 0)
            for (int i = 0; i <= NDP; i++) B[i] = 0.0;  // padding point: weight 0
        } else {
          double x1[3], x2[3], x[3], w;
          if (DEF) {
            const double* m = mp + 7 * l;
            x1[0] = m[0]; x1[1] = m[1]; x1[2] = m[2]; x2[0] = m[3]; x2[1] = m[4]; x2[2] = m[5];
            w = m[6];
          } else if (l < np) {
            fp_point<D>(rp, rw, l, scale, sh1, x1, &w);
            map_fwd<D>(vx1, x1, x);
            map_inv<D>(vx2, Fi2, x, x2);
            w = coef * hp * (w * meas1);
          } else {
            fp_point<D>(rp, rw, l - np, scale, sh2, x2, &w);
            map_fwd<D>(vx2, x2, x);
            map_inv<D>(vx1, Fi1, x, x1);
            w = coef * hp * (w * meas2);
          }
          if (half0 == 0) {
#pragma unroll
            for (int i = N; i < NDP; i++) B[i] = 0.0;
            B[NDP] = w;
          }
#pragma unroll 1
          for (int rep = 0; rep < nrep; rep++) {
            const int half = half0 + rep;
            int vs[D + 1];
            double xs[3];
#pragma unroll
            for (int k = 0; k < D + 1; k++) vs[k] = half ? v2[k] : v1[k];
#pragma unroll
            for (int c = 0; c < 3; c++) xs[c] = half ? x2[c] : x1[c];
            ValSink sk{B + half * ND, half ? -1.0 : 1.0};
            xg_h1_shapes<D, P>(vs, xs, sk);
          }
        }
      }
      __syncwarp();
      for (int q = 0; q < cs; q += 4) {
        const double* Sq = Bw + (size_t)(q + fc) * PS;
        const double wq = Sq[NDP];
        double sv[NT], av[NT];
#pragma unroll
        for (int t = 0; t < NT; t++) { sv[t] = Sq[8 * t + fr]; av[t] = wq * sv[t]; }
        int idx = 0;
#pragma unroll
        for (int ti = 0; ti < NT; ti++)
#pragma unroll
          for (int tj = ti; tj < NT; tj++, idx++) fp_dmma884(acc[idx][0], acc[idx][1], av[ti], sv[tj]);
      }
    }
    {  // tiles -> global (upper tiles and their mirror images)
      int idx = 0;
#pragma unroll
      for (int ti = 0; ti < NT; ti++)
#pragma unroll
        for (int tj = ti; tj < NT; tj++, idx++) {
          const int i = 8 * ti + fr, j = 8 * tj + 2 * fc;
          if (i < N) {
            if (j < N) M[i * N + j] = acc[idx][0];
            if (j + 1 < N) M[i * N + j + 1] = acc[idx][1];
            if (ti != tj) {
              if (j < N) M[j * N + i] = acc[idx][0];
              if (j + 1 < N) M[(j + 1) * N + i] = acc[idx][1];
            }
          }
        }
    }
  }
}

template <int D, int P>
int launch_fp_dmma(xgpu_ctx* ctx, int64_t n, const int32_t* facets, const int32_t* facet_els, const int32_t* el_vtx,
                   const double* vtx_xyz, const xgpu_rule_host* r, double coef, int hpow, double* mats)
{
  using G = FpGeom<D, P>;
  const DefDev df = xg_defdev(ctx, D);
  const size_t extra = df.order > 0 ? (size_t)(2 * XG_DEFS + 32 + 2 * FP_GEO + 7 * 2 * r->n) : 0;  // doubles per warp
  int CH = 32;
  while (CH > 8 && (size_t)4 * (CH * G::PS + extra) * sizeof(double) > 110 * 1024) CH >>= 1;
  const size_t sm = (size_t)4 * (CH * G::PS + extra) * sizeof(double);
  if (sm > 220 * 1024) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "xgpu_facetpatch: patch rule too large for the shared-memory staging");
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 8) grid = XGPU_NSM * 8;
  if (df.order > 0) {
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_facetpatch_dmma<D, P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_facetpatch_dmma<D, P, true><<<grid, 128, sm, ctx->stream>>>(n, facets, facet_els, el_vtx, vtx_xyz, r->d_pts, r->d_wts, r->n,
                                                                  CH, coef, hpow, ctx->fp_downscale, mats, df);
  } else {
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_facetpatch_dmma<D, P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_facetpatch_dmma<D, P, false><<<grid, 128, sm, ctx->stream>>>(n, facets, facet_els, el_vtx, vtx_xyz, r->d_pts, r->d_wts, r->n,
                                                                   CH, coef, hpow, ctx->fp_downscale, mats, df);
  }
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace
