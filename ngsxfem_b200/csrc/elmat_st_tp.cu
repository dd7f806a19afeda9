// elmat_st_tp.cu — space-time element matrices, tensor-structured kernel for the common orders (space P1 / P2, time 1 / 2,
// dim 2 / 3; BASELINE config 5 is P1 x P2 in 3D).  Same results as k_elmat_st (elmat_st.cu) up to summation order.
//
// Every menu term of xgpu_elmat_st factorises at a point into (time matrix) x (space pair):
//     M[(j,i),(l,k)] += w ( TS[j][l] N_i N_k + TG[j][l] grad N_i . grad N_k ),
//     TS[j][l] = c_mm L_j L_l + c_dtrial L_j L_l' + c_dtest L_j' L_l + c_dd L_j' L_l',   TG[j][l] = c_ss L_j L_l
// (j = test time node, l = trial time node).  N_i N_k and grad N_i . grad N_k are symmetric in (i,k), so a lane owns ONE
// unordered space pair {i,k} and all (k_t+1)^2 time pairs: per point it loads 2 + 2D staged space values and the 2 (k_t+1)^2
// broadcast time coefficients and updates (k_t+1)^2 accumulators — a third of the shared-memory loads per entry of the
// generic 2x2-block kernel.  Small elements (P1: 10 pairs in 3D) split the points over 32 / #pairs lane groups; the groups
// are summed by warp shuffles in a fixed order, and the matrix goes through a shared-memory tile to coalesced stores.
#include "common.cuh"
#include "cutgeom.cuh"
#include "deform.cuh"
#include "h1.cuh"
#include "elmat_st.cuh"


namespace {

__device__ __forceinline__ double ipow_tp(double h, int p)
{
  if (p == 0) return 1.0;
  double r = 1.0;
  const int ap = p < 0 ? -p : p;
  for (int i = 0; i < ap; i++) r *= h;
  return p < 0 ? 1.0 / r : r;
}

// Lagrange values / derivatives on NodalTimeFE's nodes (SpaceTimeFE.cpp:258-282), closed forms for orders 1 and 2
template <int KT>
__device__ __forceinline__ void tp_time_shapes(double t, double* L, double* dL)
{
  if (KT == 1) {
    L[0] = 1.0 - t; L[1] = t; dL[0] = -1.0; dL[1] = 1.0;
  } else {  // nodes 0, 1/2, 1
    const double a = t, b = t - 0.5, c = t - 1.0;
    L[0] = 2.0 * (b * c); L[1] = -4.0 * (a * c); L[2] = 2.0 * (a * b);
    dL[0] = 2.0 * (b + c); dL[1] = -4.0 * (a + c); dL[2] = 2.0 * (a + b);
  }
}

template <int D, int NDS>
struct TpSink {
  double* B;  // [N: NDS][G: D x NDS]
  const double* Fi;
  __device__ __forceinline__ void operator()(int i, const AD<D>& a)
  {
    B[i] = a.v;
#pragma unroll
    for (int c = 0; c < D; c++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) s += Fi[j * D + c] * a.d[j];
      B[NDS * (1 + c) + i] = s;
    }
  }
};

template <int D, int P, int KT>
struct TpCfg {
  static constexpr int NDS = (D == 2) ? (P == 1 ? 3 : 6) : (P == 1 ? 4 : 10);
  static constexpr int NT = KT + 1, NTT = NT * NT, ND = NDS * NT;
  static constexpr int NPAIR = NDS * (NDS + 1) / 2;
  static constexpr int GROUPS = (NPAIR >= 32) ? 1 : 32 / NPAIR;
  static constexpr int ROUNDS = (NPAIR + 31) / 32;
  static constexpr int PS = NDS * (1 + D) + 2 * NTT;
  static constexpr int CH = (P == 1) ? 32 : 16;
  static constexpr int WARPS = 4;
  static constexpr size_t SMEM = (size_t)WARPS * (CH * PS + ND * ND) * sizeof(double);
};

template <int D, int P, int KT>
__global__ void __launch_bounds__(128, (P == 1) ? 5 : 3) k_elmat_st_tp(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx, const double* __restrict__ vtx_xyz,
    const int64_t* __restrict__ offsets, const double* __restrict__ pts, const double* __restrict__ tvals,
    const double* __restrict__ wts, StTermsDev T, int accumulate, const int32_t* __restrict__ slots, double* __restrict__ mats)
{
  using C = TpCfg<D, P, KT>;
  constexpr int NDS = C::NDS, NT = C::NT, NTT = C::NTT, ND = C::ND, NPAIR = C::NPAIR, GROUPS = C::GROUPS, ROUNDS = C::ROUNDS,
                PS = C::PS, CH = C::CH;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* Bw = smem + (size_t)wib * (CH * PS + ND * ND);
  double* tile = Bw + CH * PS;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  // this lane's space pairs {i <= k} and point group
  const int grp = (GROUPS > 1) ? lane / NPAIR : 0;
  int pi_[ROUNDS], pk_[ROUNDS];
  bool pv_[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int pr = (GROUPS > 1) ? lane % NPAIR : lane + 32 * r;
    pv_[r] = (GROUPS > 1) ? (lane < NPAIR * GROUPS) : (pr < NPAIR);
    int i = 0, rem = pv_[r] ? pr : 0;
    while (rem >= NDS - i) { rem -= NDS - i; i++; }
    pi_[r] = i; pk_[r] = i + rem;
  }
  for (int64_t e = warp0; e < n; e += nwarps) {
    const int64_t el = elems ? (int64_t)elems[e] : e;
    int v[D + 1];
    double vx[D + 1][3], F[D * D], Fi[D * D];
#pragma unroll
    for (int k = 0; k < D + 1; k++) {
      v[k] = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
      for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[k] * D + c) : 0.0;
    }
    xg_affine_jac<D>(vx, F);
    const double meas = fabs(xg_inv<D>(F, Fi));
    const double h = xg_mesh_size<D>(meas);
    // folded coefficients of the five term classes
    double c_mm = 0.0, c_dtr = 0.0, c_dte = 0.0, c_dd = 0.0, c_ss = 0.0;
    for (int t = 0; t < T.n; t++) {
      const double c = T.coef[t] * ipow_tp(h, T.hpow[t]) * meas;
      const int ou = T.op_trial[t], ot = T.op_test[t];
      if (ot == 1) c_ss += c;
      else if (ou == 0 && ot == 0) c_mm += c;
      else if (ou != 0 && ot == 0) c_dtr += c;
      else if (ou == 0) c_dte += c;
      else c_dd += c;
    }
    const int64_t q0 = offsets[e];
    const int npe = (int)(offsets[e + 1] - q0);
    double acc[ROUNDS][NTT];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++)
#pragma unroll
      for (int m = 0; m < NTT; m++) acc[r][m] = 0.0;
    for (int c0 = 0; c0 < npe; c0 += CH) {
      const int cs = min(CH, npe - c0);
      __syncwarp();
      if (lane < cs) {
        double* B = Bw + lane * PS;
        const double* xq = pts + 3 * (q0 + c0 + lane);
        double xh[3] = {xq[0], xq[1], xq[2]};
        TpSink<D, NDS> sink{B, Fi};
        xg_h1_shapes<D, P>(v, xh, sink);
        double L[NT], dL[NT];
        tp_time_shapes<KT>(tvals[q0 + c0 + lane], L, dL);
        const double w = wts[q0 + c0 + lane];
        double* TS = B + NDS * (1 + D);
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
          for (int l = 0; l < NT; l++) {
            TS[j * NT + l] = w * (c_mm * (L[j] * L[l]) + c_dtr * (L[j] * dL[l]) + c_dte * (dL[j] * L[l]) + c_dd * (dL[j] * dL[l]));
            TS[NTT + j * NT + l] = w * (c_ss * (L[j] * L[l]));
          }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < ROUNDS; r++) {
        if (pv_[r]) {
          const int i = pi_[r], k = pk_[r];
          for (int q = grp; q < cs; q += GROUPS) {
            const double* B = Bw + q * PS;
            const double s = B[i] * B[k];
            double g = 0.0;
#pragma unroll
            for (int c = 0; c < D; c++) g += B[NDS * (1 + c) + i] * B[NDS * (1 + c) + k];
            const double* TS = B + NDS * (1 + D);
#pragma unroll
            for (int m = 0; m < NTT; m++) acc[r][m] += TS[m] * s + TS[NTT + m] * g;
          }
        }
      }
    }
    // point groups -> group 0 by warp shuffles in a fixed order ((g0 + g1) + g2 ...), group 0 fills the tile, coalesced stores
    __syncwarp();
    if (GROUPS > 1) {
#pragma unroll
      for (int m = 0; m < NTT; m++) {
        double a = acc[0][m];
#pragma unroll
        for (int g = 1; g < GROUPS; g++) a += __shfl_down_sync(0xffffffffu, acc[0][m], g * NPAIR);
        acc[0][m] = a;
      }
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      if (pv_[r] && grp == 0) {
        const int i = pi_[r], k = pk_[r];
        double* t1 = tile + i * ND + k;
        double* t2 = tile + k * ND + i;
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
          for (int l = 0; l < NT; l++) {
            const double a = acc[r][j * NT + l];
            t1[j * NDS * ND + l * NDS] = a;
            t2[j * NDS * ND + l * NDS] = a;
          }
      }
    }
    __syncwarp();
    double* M = mats + (size_t)(slots ? (int64_t)slots[e] : e) * ND * ND;
    if (accumulate) {
      for (int idx = lane; idx < ND * ND; idx += 32) M[idx] += tile[idx];
    } else {
      for (int idx = lane; idx < ND * ND; idx += 32) M[idx] = tile[idx];
    }
  }
}

template <int D, int P, int KT>
int launch_tp(xgpu_ctx* ctx, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
              const int64_t* offsets, const double* pts, const double* tvals, const double* wts, const StTermsDev& T,
              int accumulate, const int32_t* slots, double* mats)
{
  using C = TpCfg<D, P, KT>;
  unsigned grid = (unsigned)((n + 3) / 4);
  if (grid > XGPU_NSM * 16) grid = XGPU_NSM * 16;
  XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_st_tp<D, P, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  k_elmat_st_tp<D, P, KT><<<grid, 128, C::SMEM, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T,
                                                               accumulate, slots, mats);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

// returns 1 if this (dim, space order, time order) has no tensor-structured instantiation (caller uses k_elmat_st)
int xg_launch_elmat_st_tp(xgpu_ctx* ctx, int dim, int p, int kt, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                          const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,
                          const double* wts, const StTermsDev& T, int accumulate, const int32_t* slots, double* mats)
{
#define XG_TP(DD, PP, KK)                                                                                             \
  if (dim == DD && p == PP && kt == KK)                                                                               \
    return launch_tp<DD, PP, KK>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, tvals, wts, T, accumulate, slots, mats);
  XG_TP(3, 1, 1) XG_TP(3, 1, 2) XG_TP(3, 2, 1) XG_TP(3, 2, 2)
  XG_TP(2, 1, 1) XG_TP(2, 1, 2) XG_TP(2, 2, 1) XG_TP(2, 2, 2)
#undef XG_TP
  return 1;
}
