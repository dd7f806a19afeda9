// elmat_st_impl.cuh — generic space-time element matrix / vector kernels (any space order 1..4, time order 0..3) and their
// launchers; instantiated per space dimension in elmat_st_d1.cu / _d2.cu / _d3.cu (separate translation units: build time).
#pragma once
#include "common.cuh"
#include "cutgeom.cuh"
#include "deform.cuh"
#include "h1.cuh"
#include "elmat_st.cuh"

#define XG_ST_MAXT 4  // time orders 0..3

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ double ipow_st(double h, int p)
{
  if (p == 0) return 1.0;
  double r = 1.0;
  const int ap = p < 0 ? -p : p;
  for (int i = 0; i < ap; i++) r *= h;
  return p < 0 ? 1.0 / r : r;
}

// NodalTimeFE::CalcInterpolationPoints (SpaceTimeFE.cpp:258-282)
__device__ __forceinline__ void st_time_nodes(int order, double* nd)
{
  if (order == 0) { nd[0] = 0.0; return; }
  if (order == 1) { nd[0] = 0.0; nd[1] = 1.0; return; }
  if (order == 2) { nd[0] = 0.0; nd[1] = 0.5; nd[2] = 1.0; return; }
  nd[0] = 0.0; nd[1] = 0.5 * (1.0 - 1.0 / sqrt(5.0)); nd[2] = 0.5 * (1.0 + 1.0 / sqrt(5.0)); nd[3] = 1.0;
}

// Lagrange values and first derivatives on the nodes
__device__ __forceinline__ void st_time_shapes(int order, const double* nd, double t, double* L, double* dL)
{
  for (int i = 0; i <= order; i++) {
    double s = 1.0, ds = 0.0;
    for (int j = 0; j <= order; j++) if (j != i) s *= (t - nd[j]) / (nd[i] - nd[j]);
    for (int k = 0; k <= order; k++) if (k != i) {
      double pr = 1.0 / (nd[i] - nd[k]);
      for (int j = 0; j <= order; j++) if (j != i && j != k) pr *= (t - nd[j]) / (nd[i] - nd[j]);
      ds += pr;
    }
    L[i] = s; dL[i] = ds;
  }
}

// staging sink: tensor product of one space function with all time functions
template <int D>
struct StSink {
  double* B;   // this point's slab: [val | g0..g(D-1) | dt], each ndp long
  int ndp, nds, nt;
  const double* Fi;
  const double* L;
  const double* dL;
  __device__ __forceinline__ void operator()(int i, const AD<D>& a)
  {
    double g[D];
#pragma unroll
    for (int c = 0; c < D; c++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) s += Fi[j * D + c] * a.d[j];
      g[c] = s;
    }
    for (int j = 0; j < nt; j++) {
      const int k = j * nds + i;
      B[k] = L[j] * a.v;
#pragma unroll
      for (int c = 0; c < D; c++) B[(1 + c) * ndp + k] = L[j] * g[c];
      B[(1 + D) * ndp + k] = dL[j] * a.v;
    }
  }
};

template <int D>
struct ElGeom {
  int v[D + 1];
  double vx[D + 1][3], F[D * D], Fi[D * D], meas, h;
};

template <int D>
__device__ __forceinline__ void st_load_geom(int64_t el, const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
                                             ElGeom<D>& G)
{
#pragma unroll
  for (int k = 0; k < D + 1; k++) {
    G.v[k] = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
    for (int c = 0; c < 3; c++) G.vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)G.v[k] * D + c) : 0.0;
  }
  xg_affine_jac<D>(G.vx, G.F);
  G.meas = fabs(xg_inv<D>(G.F, G.Fi));
  G.h = xg_mesh_size<D>(G.meas);
}

__device__ __forceinline__ int st_slot(int op, int D) { return op == 0 ? 0 : (1 + D); }  // ID -> val, DT -> dt

template <int D, int MAXB>
__global__ void __launch_bounds__(128) k_elmat_st(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
    const int64_t* __restrict__ offsets, const double* __restrict__ pts, const double* __restrict__ tvals,
    const double* __restrict__ wts, int p, int nds, int kt, int CH, StTermsDev T, int accumulate,
    const int32_t* __restrict__ slots, double* __restrict__ mats)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nt = kt + 1, nd = nds * nt;
  const int ndp = (nd + 1) & ~1, nbk = ndp >> 1;
  const int PS = ndp * (2 + D) + XG_ST_MAXTERMS;
  double* Bw = smem + (size_t)wib * CH * PS;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nblocks = nbk * nbk;
  double tn[XG_ST_MAXT];
  st_time_nodes(kt, tn);
  for (int64_t e = warp0; e < n; e += nwarps) {
    const int64_t el = elems ? (int64_t)elems[e] : e;
    ElGeom<D> G;
    st_load_geom<D>(el, el_vtx, vtx_xyz, G);
    const int64_t q0 = offsets[e];
    const int npe = (int)(offsets[e + 1] - q0);
    double acc[MAXB][4];
#pragma unroll
    for (int b = 0; b < MAXB; b++) { acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.0; }
    for (int c0 = 0; c0 < npe; c0 += CH) {
      const int cs = min(CH, npe - c0);
      __syncwarp();
      if (lane < cs) {
        double* B = Bw + (size_t)lane * PS;
        const double* xq = pts + 3 * (q0 + c0 + lane);
        double xh[3] = {xq[0], xq[1], xq[2]};
        for (int i = nd; i < ndp; i++) for (int k = 0; k < 2 + D; k++) B[k * ndp + i] = 0.0;
        double L[XG_ST_MAXT], dL[XG_ST_MAXT];
        st_time_shapes(kt, tn, tvals[q0 + c0 + lane], L, dL);
        StSink<D> sink{B, ndp, nds, nt, G.Fi, L, dL};
        xg_h1_shapes_rt<D>(p, G.v, xh, sink);
        const double fac = wts[q0 + c0 + lane] * G.meas;
        for (int t = 0; t < T.n; t++) B[(2 + D) * ndp + t] = T.coef[t] * ipow_st(G.h, T.hpow[t]) * fac;
      }
      __syncwarp();
#pragma unroll
      for (int b = 0; b < MAXB; b++) {
        const int blk = lane + 32 * b;
        if (blk < nblocks) {
          const int i0 = 2 * (blk / nbk), j0 = 2 * (blk % nbk);
          for (int q = 0; q < cs; q++) {
            const double* B = Bw + (size_t)q * PS;
            for (int t = 0; t < T.n; t++) {
              const double cw = B[(2 + D) * ndp + t];
              if (T.op_test[t] == 1) {  // GRAD . GRAD
                double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
                for (int c = 0; c < D; c++) {
                  const double a0 = B[(1 + c) * ndp + i0], a1 = B[(1 + c) * ndp + i0 + 1];
                  const double b0 = B[(1 + c) * ndp + j0], b1 = B[(1 + c) * ndp + j0 + 1];
                  s00 += a0 * b0; s01 += a0 * b1; s10 += a1 * b0; s11 += a1 * b1;
                }
                acc[b][0] += cw * s00; acc[b][1] += cw * s01; acc[b][2] += cw * s10; acc[b][3] += cw * s11;
              } else {
                const int ot = st_slot(T.op_test[t], D), ou = st_slot(T.op_trial[t], D);
                const double a0 = cw * B[ot * ndp + i0], a1 = cw * B[ot * ndp + i0 + 1];
                const double b0 = B[ou * ndp + j0], b1 = B[ou * ndp + j0 + 1];
                acc[b][0] += a0 * b0; acc[b][1] += a0 * b1; acc[b][2] += a1 * b0; acc[b][3] += a1 * b1;
              }
            }
          }
        }
      }
    }
    double* M = mats + (size_t)(slots ? (int64_t)slots[e] : e) * nd * nd;
#pragma unroll
    for (int b = 0; b < MAXB; b++) {
      const int blk = lane + 32 * b;
      if (blk < nblocks) {
        const int i0 = 2 * (blk / nbk), j0 = 2 * (blk % nbk);
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int c = 0; c < 2; c++) {
            const int i = i0 + a, j = j0 + c;
            if (i < nd && j < nd) {
              if (accumulate) M[i * nd + j] += acc[b][2 * a + c];
              else M[i * nd + j] = acc[b][2 * a + c];
            }
          }
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(128) k_elvec_st(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
    const int64_t* __restrict__ offsets, const double* __restrict__ pts, const double* __restrict__ tvals,
    const double* __restrict__ wts, int p, int nds, int kt, int CH, StTermsDev T, int accumulate,
    const int32_t* __restrict__ slots, double* __restrict__ vecs)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nt = kt + 1, nd = nds * nt;
  const int ndp = (nd + 1) & ~1;
  const int PS = ndp * (2 + D) + XG_ST_MAXTERMS;
  double* Bw = smem + (size_t)wib * CH * PS;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  double tn[XG_ST_MAXT];
  st_time_nodes(kt, tn);
  const int nrounds = (nd + 31) / 32;
  for (int64_t e = warp0; e < n; e += nwarps) {
    const int64_t el = elems ? (int64_t)elems[e] : e;
    ElGeom<D> G;
    st_load_geom<D>(el, el_vtx, vtx_xyz, G);
    const int64_t q0 = offsets[e];
    const int npe = (int)(offsets[e + 1] - q0);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // nd <= 128
    for (int c0 = 0; c0 < npe; c0 += CH) {
      const int cs = min(CH, npe - c0);
      __syncwarp();
      if (lane < cs) {
        double* B = Bw + (size_t)lane * PS;
        const double* xq = pts + 3 * (q0 + c0 + lane);
        double xh[3] = {xq[0], xq[1], xq[2]};
        const double tq = tvals[q0 + c0 + lane];
        double L[XG_ST_MAXT], dL[XG_ST_MAXT];
        st_time_shapes(kt, tn, tq, L, dL);
        StSink<D> sink{B, ndp, nds, nt, G.Fi, L, dL};
        xg_h1_shapes_rt<D>(p, G.v, xh, sink);
        // physical point of the affine element: x = sum lam_k v_k, lam = (xh, 1 - sum xh)
        double x[3] = {0.0, 0.0, 0.0}, lamD = 1.0;
#pragma unroll
        for (int k = 0; k < D; k++) lamD -= xh[k];
#pragma unroll
        for (int c = 0; c < D; c++) {
          double s = lamD * G.vx[D][c];
#pragma unroll
          for (int k = 0; k < D; k++) s += xh[k] * G.vx[k][c];
          x[c] = s;
        }
        const double fac = wts[q0 + c0 + lane] * G.meas;
        for (int t = 0; t < T.n; t++) {
          double f = T.coef[t] * ipow_st(G.h, T.hpow[t]) * fac;
          for (int c = 0; c < 3; c++) f *= ipow_st(x[c], T.expo[t][c]);
          f *= ipow_st(tq, T.expo[t][3]);
          B[(2 + D) * ndp + t] = f;
        }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int i = lane + 32 * r;
        if (r < nrounds && i < nd)
          for (int q = 0; q < cs; q++) {
            const double* B = Bw + (size_t)q * PS;
            for (int t = 0; t < T.n; t++) acc[r] += B[(2 + D) * ndp + t] * B[st_slot(T.op_test[t], D) * ndp + i];
          }
      }
    }
    double* V = vecs + (size_t)(slots ? (int64_t)slots[e] : e) * nd;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int i = lane + 32 * r;
      if (r < nrounds && i < nd) { if (accumulate) V[i] += acc[r]; else V[i] = acc[r]; }
    }
  }
}

int st_chunk(int ndp, int D, size_t* sm)
{
  const int PS = ndp * (2 + D) + XG_ST_MAXTERMS;
  int CH = 32;
  while (CH > 2 && (size_t)4 * CH * PS * sizeof(double) > 96 * 1024) CH >>= 1;
  *sm = (size_t)4 * CH * PS * sizeof(double);
  return CH;
}

template <int D>
int launch_elmat_st(xgpu_ctx* ctx, int p, int kt, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                    const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,
                    const double* wts, const StTermsDev& T, int accumulate, const int32_t* slots, double* mats)
{
  const int nds = xg_h1_ndof(D, p), nd = nds * (kt + 1);
  const int ndp = (nd + 1) & ~1;
  size_t sm;
  const int CH = st_chunk(ndp, D, &sm);
  const int nblocks = (ndp / 2) * (ndp / 2);
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 8) grid = XGPU_NSM * 8;
#define XG_LST(MB)                                                                                                  \
  {                                                                                                                 \
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_st<D, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));  \
    k_elmat_st<D, MB><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, p, nds, kt, CH, T, \
                                                      accumulate, slots, mats);                                     \
  }
  if (nblocks <= 32) XG_LST(1) else if (nblocks <= 128) XG_LST(4) else if (nblocks <= 512) XG_LST(16)
  else return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elmat_st: space-time element too large (nd_s*(k_t+1) <= 44)");
#undef XG_LST
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

template <int D>
int launch_elvec_st(xgpu_ctx* ctx, int p, int kt, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                    const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,
                    const double* wts, const StTermsDev& T, int accumulate, const int32_t* slots, double* vecs)
{
  const int nds = xg_h1_ndof(D, p), nd = nds * (kt + 1);
  if (nd > 128) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elvec_st: nd_s*(k_t+1) <= 128");
  const int ndp = (nd + 1) & ~1;
  size_t sm;
  const int CH = st_chunk(ndp, D, &sm);
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 8) grid = XGPU_NSM * 8;
  XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elvec_st<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_elvec_st<D><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, p, nds, kt, CH, T,
                                                accumulate, slots, vecs);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

#define XG_ST_INSTANTIATE(DD)                                                                                          \
  int xg_launch_elmat_st_d##DD(xgpu_ctx* ctx, int p, int kt, int64_t n, const int32_t* elems, const int32_t* el_vtx,    \
                               const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,   \
                               const double* wts, const StTermsDev& T, int accumulate, const int32_t* slots, double* mats) \
  {                                                                                                                    \
    return launch_elmat_st<DD>(ctx, p, kt, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, mats); \
  }                                                                                                                    \
  int xg_launch_elvec_st_d##DD(xgpu_ctx* ctx, int p, int kt, int64_t n, const int32_t* elems, const int32_t* el_vtx,    \
                               const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,   \
                               const double* wts, const StTermsDev& T, int accumulate, const int32_t* slots, double* vecs) \
  {                                                                                                                    \
    return launch_elvec_st<DD>(ctx, p, kt, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, vecs); \
  }
