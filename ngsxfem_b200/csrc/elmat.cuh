#pragma once
// elmat.cuh — kernels + launch templates shared by elmat.cu, elmat_ms_*.cu, elmat_dmma_*.cu (split for build time)
// elmat.cu — batched cut / uncut element matrices  B^T diag(w D) B
// (SymbolicCutBilinearFormIntegrator::T_CalcElementMatrixAdd, xfem/symboliccutbfi.cpp:76-486).
//
//   k_elmat_points<D,MAXB>  general path, one WARP per element: (1) every lane evaluates the H1
//        basis, physical gradients (F^{-T}) and normal derivatives at one quadrature point of the
//        current chunk and stages them in shared memory, (2) the lanes own 2x2 register blocks of
//        the element matrix and contract over the staged points.  Used for cut elements
//        (per-element rules from cutquad) and for uncut elements with a shared standard rule
//        when the fast path does not apply.
//   k_elmat_affine<D>       fast path for uncut AFFINE elements and ID*ID / GRAD.GRAD terms:
//        sum_q w_q (F^{-T} dphi_i).(F^{-T} dphi_j) = sum_ab G_ab sum_q w_q d_a phi_i d_b phi_j is the
//        same sum re-associated, so the matrix is  |det F| (c_m M_ref + c_s sum_ab G_ab K_ref^ab)
//        with reference tensors contracted once on the host from the same rule.  Each lane
//        prepares the 1 + D(D+1)/2 geometry coefficients of one element, then the warp streams
//        the 32 matrices out with fully coalesced 256 B stores: HBM-write bound
//        (nd^2 * 8 B out per element, 16 B + vertex gathers in).
#include "common.cuh"
#include "cutgeom.cuh"
#include "deform.cuh"
#include <vector>

#define XG_MAXTERMS 8
struct TermsDev {
  int n;
  int op_trial[XG_MAXTERMS], op_test[XG_MAXTERMS], hpow[XG_MAXTERMS];
  double coef[XG_MAXTERMS];
};

// optional second integrator of the same pass (xgpu_elmat_rule_pair): its own per-element rule and term weights
struct Rule2 {
  const int64_t* offsets;  // nullptr: absent
  const double* pts;
  const double* wts;
  TermsDev T;
};


namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__device__ __forceinline__ double ipow(double h, int p)
{
  if (p == 0) return 1.0;
  double r = 1.0;
  const int ap = p < 0 ? -p : p;
  for (int i = 0; i < ap; i++) r *= h;
  return p < 0 ? 1.0 / r : r;
}

// folded weights of the ID*ID (mass) and GRAD.GRAD (stiffness) terms at measure `meas`, mesh size h
__device__ __forceinline__ void ms_coefs(const TermsDev& T, double h, double meas, double& cm, double& cs)
{
  cm = 0.0; cs = 0.0;
  for (int t = 0; t < T.n; t++) {
    const double c = T.coef[t] * ipow(h, T.hpow[t]) * meas;
    if (T.op_test[t] == 0) cm += c; else cs += c;
  }
}

// staging sink: values, physical gradients and normal derivative of shape i at one point
template <int D>
struct StageSink {
  double* B;         // this point's slab: [val | g0 | g1 | g2 | dn], each ndp long
  int ndp;
  const double* Fi;  // D x D inverse Jacobian
  const double* nrm; // unit normal (physical) or zeros
  __device__ __forceinline__ void operator()(int i, const AD<D>& a)
  {
    B[i] = a.v;
    double dn = 0.0;
#pragma unroll
    for (int c = 0; c < D; c++) {
      double g = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) g += Fi[j * D + c] * a.d[j];
      B[(1 + c) * ndp + i] = g;
      dn += g * nrm[c];
    }
    B[(1 + D) * ndp + i] = dn;
  }
};

template <int D, int MAXB, bool DEF>
__global__ void __launch_bounds__(128) k_elmat_points(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
    const double* __restrict__ vtx_xyz, const double* __restrict__ phi,
    const int64_t* __restrict__ offsets, const double* __restrict__ pts, const double* __restrict__ wts, int np_shared,
    int p, int nd, int CH, TermsDev T, int accumulate, const int32_t* __restrict__ slots, double* __restrict__ mats,
    DefDev df)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int ndp = (nd + 1) & ~1, nbk = ndp >> 1;
  const int PS = ndp * (2 + D) + XG_MAXTERMS;  // per point: val, D grads, dn, term weights
  double* Bw = smem + (size_t)wib * (CH * PS + (DEF ? XG_DEFS : 0));
  double* defs = Bw + (size_t)CH * PS;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nblocks = nbk * nbk;
  for (int64_t e = warp0; e < n; e += nwarps) {
    const int64_t el = elems ? (int64_t)elems[e] : e;
    int v[D + 1];
    double vx[D + 1][3], l[D + 1], F[D * D], Fi[D * D], nrm[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < D + 1; k++) {
      v[k] = __ldg(el_vtx + el * (D + 1) + k);
      l[k] = phi ? __ldg(phi + v[k]) : 0.0;
#pragma unroll
      for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[k] * D + c) : 0.0;
    }
    xg_affine_jac<D>(vx, F);
    const double meas = fabs(xg_inv<D>(F, Fi));
    const double h = xg_mesh_size<D>(meas);
    if (phi) xg_phys_normal<D>(Fi, l, nrm);  // n = F^{-T} grad_ref(lset) normalised (Normalize(grad(lsetp1)))
    if (DEF) { __syncwarp(); xg_def_gather<D>(df, el, lane, defs); __syncwarp(); }
    const int64_t q0 = offsets ? offsets[e] : 0;
    const int npe = offsets ? (int)(offsets[e + 1] - q0) : np_shared;
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
        double Fiq[D * D], nq[3] = {0, 0, 0}, measq = meas, hq = h;
        if (DEF) {  // Jacobian of the deformed element at this point
          double Fq[D * D], xd[3] = {0, 0, 0};
#pragma unroll
          for (int k = 0; k < D * D; k++) Fq[k] = F[k];
          xg_def_apply<D>(df.order, v, xh, defs, Fq, xd);
          measq = fabs(xg_inv<D>(Fq, Fiq));
          hq = xg_mesh_size<D>(measq);
          if (phi) xg_phys_normal<D>(Fiq, l, nq);
        }
        StageSink<D> sink{B, ndp, DEF ? Fiq : Fi, DEF ? nq : nrm};
        xg_h1_shapes_rt<D>(p, v, xh, sink);
        const double fac = wts[q0 + c0 + lane] * measq;
        for (int t = 0; t < T.n; t++) B[(2 + D) * ndp + t] = T.coef[t] * ipow(hq, T.hpow[t]) * fac;
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
              if (T.op_test[t] == 1 && T.op_trial[t] == 1) {
                double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
                for (int c = 0; c < D; c++) {
                  const double a0 = B[(1 + c) * ndp + i0], a1 = B[(1 + c) * ndp + i0 + 1];
                  const double b0 = B[(1 + c) * ndp + j0], b1 = B[(1 + c) * ndp + j0 + 1];
                  s00 += a0 * b0; s01 += a0 * b1; s10 += a1 * b0; s11 += a1 * b1;
                }
                acc[b][0] += cw * s00; acc[b][1] += cw * s01; acc[b][2] += cw * s10; acc[b][3] += cw * s11;
              } else {
                const int ot = (T.op_test[t] == 0) ? 0 : (1 + D), ou = (T.op_trial[t] == 0) ? 0 : (1 + D);
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

// ---------------------------------------------------------------- mass + stiffness on point rules
// Specialisation of k_elmat_points for sums of ID*ID and GRAD.GRAD terms (the volume and interface
// integrators of CutFEM Poisson-type forms): compile-time order (the basis evaluation unrolls, no
// operator dispatch in the contraction), the coefficients of all terms are folded per element into
// one mass and one stiffness weight per point.  Per point and lane: 9 shared loads (128 bit) and
// 22 FP64 operations on a 2x2 register block.
template <int D>
struct StageSinkMS {
  double* B;         // this point's slab: [val | g0 | .. | g(D-1)], each NDP long
  int ndp;
  const double* Fi;
  __device__ __forceinline__ void operator()(int i, const AD<D>& a)
  {
    B[i] = a.v;
#pragma unroll
    for (int c = 0; c < D; c++) {
      double g = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) g += Fi[j * D + c] * a.d[j];
      B[(1 + c) * ndp + i] = g;
    }
  }
};

// Work split: the matrices of these forms are symmetric, so only the 2x2 blocks (bi <= bj) are computed — NSYM =
// NBK(NBK+1)/2 of them (15 for P2 tets) — by the 16 lanes of a half warp; the two half warps take the even / odd
// staged points of every chunk and are summed (fixed order) at the end, half 0 then stores the upper blocks, half 1
// the mirrored ones.  Two integrators (NEG volume terms + IF interface term of CutFEM forms) run in one pass.
template <int D, int P, bool DEF>
__global__ void __launch_bounds__(128, (P <= 2 ? 5 : 1)) k_elmat_ms(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
    const double* __restrict__ vtx_xyz, const int64_t* __restrict__ offsets, const double* __restrict__ pts,
    const double* __restrict__ wts, int np_shared, int CH, TermsDev T, int accumulate,
    const int32_t* __restrict__ slots, double* __restrict__ mats, DefDev df, Rule2 R2)
{
  constexpr int ND = (D == 1) ? P + 1 : ((D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2
                                                   : 4 + 6 * (P - 1) + 4 * ((P - 1) * (P - 2) / 2) + (P - 1) * (P - 2) * (P - 3) / 6);
  constexpr int NDP = (ND + 1) & ~1, NBK = NDP / 2, NSYM = NBK * (NBK + 1) / 2, MAXS = (NSYM + 15) / 16;
  constexpr int PS = NDP * (1 + D) + 2;  // per point: val, D physical gradients, (mass, stiffness) weights
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int half = lane >> 4, hl = lane & 15;
  double* Bw = smem + (size_t)wib * (CH * PS + (DEF ? XG_DEFS : 0));
  double* defs = Bw + (size_t)CH * PS;
  int bi_[MAXS], bj_[MAXS];  // the blocks of this lane (bi = -1: none)
#pragma unroll
  for (int b = 0; b < MAXS; b++) {
    const int blk = hl + 16 * b;
    bi_[b] = -1; bj_[b] = 0;
    if (blk < NSYM) {
      int r = 0, rem = blk;
      while (rem >= NBK - r) { rem -= NBK - r; r++; }
      bi_[b] = r; bj_[b] = r + rem;
    }
  }
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nseg = R2.offsets ? 2 : 1;
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
    if (DEF) { __syncwarp(); xg_def_gather<D>(df, el, lane, defs); __syncwarp(); }
    double acc[MAXS][4];
#pragma unroll
    for (int b = 0; b < MAXS; b++) { acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.0; }
    for (int seg = 0; seg < nseg; seg++) {
      const TermsDev* Ts = seg ? &R2.T : &T;
      const int64_t* off_s = seg ? R2.offsets : offsets;
      const double* pts_s = seg ? R2.pts : pts;
      const double* wts_s = seg ? R2.wts : wts;
      double cm, cs;
      ms_coefs(*Ts, xg_mesh_size<D>(meas), meas, cm, cs);
      const int64_t q0 = off_s ? off_s[e] : 0;
      const int npe = off_s ? (int)(off_s[e + 1] - q0) : np_shared;
      for (int c0 = 0; c0 < npe; c0 += CH) {
        const int cnt = min(CH, npe - c0);
        __syncwarp();
        for (int ql = lane; ql < cnt; ql += 32) {
          double* B = Bw + (size_t)ql * PS;
          const double* xq = pts_s + 3 * (q0 + c0 + ql);
          double xh[3] = {xq[0], xq[1], xq[2]};
          if (ND < NDP) {
#pragma unroll
            for (int k = 0; k < 1 + D; k++) B[k * NDP + ND] = 0.0;
          }
          double Fiq[D * D], cmq = cm, csq = cs;
          if (DEF) {
            double Fq[D * D], xd[3] = {0, 0, 0};
#pragma unroll
            for (int k = 0; k < D * D; k++) Fq[k] = F[k];
            xg_def_apply<D>(df.order, v, xh, defs, Fq, xd);
            const double measq = fabs(xg_inv<D>(Fq, Fiq));
            ms_coefs(*Ts, xg_mesh_size<D>(measq), measq, cmq, csq);
          }
          StageSinkMS<D> sink{B, NDP, DEF ? Fiq : Fi};
          xg_h1_shapes<D, P>(v, xh, sink);
          const double w = wts_s[q0 + c0 + ql];
          B[NDP * (1 + D)] = cmq * w;
          B[NDP * (1 + D) + 1] = csq * w;
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < MAXS; b++) {
          if (bi_[b] >= 0) {
            const int i0 = 2 * bi_[b], j0 = 2 * bj_[b];
            for (int q = half; q < cnt; q += 2) {
              const double* B = Bw + (size_t)q * PS;
              const double2 wq = *reinterpret_cast<const double2*>(B + NDP * (1 + D));
              double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
              for (int c = 0; c < D; c++) {
                const double2 a = *reinterpret_cast<const double2*>(B + (1 + c) * NDP + i0);
                const double2 bb = *reinterpret_cast<const double2*>(B + (1 + c) * NDP + j0);
                s00 += a.x * bb.x; s01 += a.x * bb.y; s10 += a.y * bb.x; s11 += a.y * bb.y;
              }
              const double2 a = *reinterpret_cast<const double2*>(B + i0);
              const double2 bb = *reinterpret_cast<const double2*>(B + j0);
              const double m0 = wq.x * a.x, m1 = wq.x * a.y;
              acc[b][0] += wq.y * s00 + m0 * bb.x; acc[b][1] += wq.y * s01 + m0 * bb.y;
              acc[b][2] += wq.y * s10 + m1 * bb.x; acc[b][3] += wq.y * s11 + m1 * bb.y;
            }
          }
        }
      }
    }
    double* M = mats + (size_t)(slots ? (int64_t)slots[e] : e) * ND * ND;
#pragma unroll
    for (int b = 0; b < MAXS; b++) {
#pragma unroll
      for (int k = 0; k < 4; k++) acc[b][k] += __shfl_xor_sync(0xffffffffu, acc[b][k], 16);  // even + odd points
      if (bi_[b] >= 0) {
        const int i0 = 2 * bi_[b], j0 = 2 * bj_[b];
        if (half == 0 || bi_[b] != bj_[b]) {
#pragma unroll
          for (int a = 0; a < 2; a++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
              const int i = i0 + a, j = j0 + c;
              if (i < ND && j < ND) {
                const int k = half ? j * ND + i : i * ND + j;  // half 1: the mirrored block
                if (accumulate) M[k] += acc[b][2 * a + c];
                else M[k] = acc[b][2 * a + c];
              }
            }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- FP64 tensor-core contraction (order >= 3)
// M = sum_q  S_q^T diag(w_q) S_q  with S_q the 4 x nd slab [val ; g0 ; g1 ; g2] of one point: a point
// IS one k-step of the FP64 tensor instruction (sm_100a lowers every mma.sync f64 shape to
// DMMA.8x8x4, checked with cuobjdump), so the contraction over the quadrature points is a chain of
// m8n8k4 MMAs on 8x8 tiles of the element matrix:
//   A fragment (8 x 4, row i, col c) = w_c S_q[c][8 ti + i]   (lane = 4 i + c)
//   B fragment (4 x 8, row c, col j) =     S_q[c][8 tj + j]   (lane = 4 j + c)
// — the SAME shared-memory word, weighted for A, plain for B: per point a lane loads NT values and
// one weight and issues NT(NT+1)/2 DMMAs (the forms on this path have op_test == op_trial, so only
// the tiles ti <= tj are computed and mirrored in the epilogue; the reference computes both halves,
// they agree to rounding).  One warp per element; lanes stage one point each (basis + physical
// gradients), row stride = 8 mod 16 doubles so that the four component rows of a fragment load fall
// into disjoint bank groups; the finished tiles go through shared memory for coalesced stores.
__device__ __forceinline__ void xg_dmma884(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int D>
struct StageSinkT {
  double* B;         // this point's slab: 4 rows of RS doubles
  int rs;
  const double* Fi;
  __device__ __forceinline__ void operator()(int i, const AD<D>& a)
  {
    B[i] = a.v;
#pragma unroll
    for (int c = 0; c < D; c++) {
      double g = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) g += Fi[j * D + c] * a.d[j];
      B[(1 + c) * rs + i] = g;
    }
  }
};

template <int D, int P>
struct DmmaGeom {
  static constexpr int ND = (D == 1) ? P + 1 : ((D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2
                                                         : 4 + 6 * (P - 1) + 4 * ((P - 1) * (P - 2) / 2) + (P - 1) * (P - 2) * (P - 3) / 6);
  static constexpr int NT = (ND + 7) / 8, NDP = NT * 8;
  static constexpr int RS = NDP + ((NDP % 16 == 0) ? 8 : 0);  // row stride: 8 mod 16 doubles
  static constexpr int PS = 4 * RS + 4;                       // 4 component rows + 4 weights (PS = 4 mod 8: 16 B aligned slabs)
  static constexpr int NPAIR = NT * (NT + 1) / 2;
};

template <int D, int P, bool DEF>
__global__ void __launch_bounds__(160) k_elmat_dmma(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
    const double* __restrict__ vtx_xyz, const int64_t* __restrict__ offsets, const double* __restrict__ pts,
    const double* __restrict__ wts, int np_shared, int CH, TermsDev T, int accumulate,
    const int32_t* __restrict__ slots, double* __restrict__ mats, DefDev df)
{
  using G = DmmaGeom<D, P>;
  constexpr int ND = G::ND, NT = G::NT, NDP = G::NDP, RS = G::RS, PS = G::PS, NPAIR = G::NPAIR;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int fr = lane >> 2, fc = lane & 3;  // fragment row / component
  double* Bw = smem + (size_t)wib * (CH * PS + (DEF ? XG_DEFS : 0));
  double* defs = Bw + (size_t)CH * PS;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
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
    double cm, cs;
    ms_coefs(T, xg_mesh_size<D>(meas), meas, cm, cs);
    if (DEF) { __syncwarp(); xg_def_gather<D>(df, el, lane, defs); __syncwarp(); }
    const int64_t q0 = offsets ? offsets[e] : 0;
    const int npe = offsets ? (int)(offsets[e + 1] - q0) : np_shared;
    double acc[NPAIR][2];
#pragma unroll
    for (int b = 0; b < NPAIR; b++) acc[b][0] = acc[b][1] = 0.0;
    for (int c0 = 0; c0 < npe; c0 += CH) {
      const int cnt = min(CH, npe - c0);
      __syncwarp();
      for (int ql = lane; ql < cnt; ql += 32) {
        double* B = Bw + (size_t)ql * PS;
        const double* xq = pts + 3 * (q0 + c0 + ql);
        double xh[3] = {xq[0], xq[1], xq[2]};
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
          for (int i = ND; i < NDP; i++) B[k * RS + i] = 0.0;
        if (D < 3) {
#pragma unroll
          for (int k = 1 + D; k < 4; k++)
            for (int i = 0; i < ND; i++) B[k * RS + i] = 0.0;
        }
        double Fiq[D * D], cmq = cm, csq = cs;
        if (DEF) {
          double Fq[D * D], xd[3] = {0, 0, 0};
#pragma unroll
          for (int k = 0; k < D * D; k++) Fq[k] = F[k];
          xg_def_apply<D>(df.order, v, xh, defs, Fq, xd);
          const double measq = fabs(xg_inv<D>(Fq, Fiq));
          ms_coefs(T, xg_mesh_size<D>(measq), measq, cmq, csq);
        }
        StageSinkT<D> sink{B, RS, DEF ? Fiq : Fi};
        xg_h1_shapes<D, P>(v, xh, sink);
        const double w = wts[q0 + c0 + ql];
        B[4 * RS] = cmq * w;
        B[4 * RS + 1] = B[4 * RS + 2] = B[4 * RS + 3] = csq * w;
      }
      __syncwarp();
      for (int q = 0; q < cnt; q++) {
        const double* Sq = Bw + (size_t)q * PS;
        const double wq = Sq[4 * RS + fc];
        double sv[NT], av[NT];
#pragma unroll
        for (int t = 0; t < NT; t++) { sv[t] = Sq[fc * RS + 8 * t + fr]; av[t] = wq * sv[t]; }
        int idx = 0;
#pragma unroll
        for (int ti = 0; ti < NT; ti++)
#pragma unroll
          for (int tj = ti; tj < NT; tj++, idx++) xg_dmma884(acc[idx][0], acc[idx][1], av[ti], sv[tj]);
      }
    }
    // epilogue: tiles -> shared (full NDP x NDP matrix, both halves) -> coalesced global stores
    __syncwarp();
    double* Ms = Bw;  // NDP * NDP doubles <= CH * PS (checked on the host)
    {
      int idx = 0;
#pragma unroll
      for (int ti = 0; ti < NT; ti++)
#pragma unroll
        for (int tj = ti; tj < NT; tj++, idx++) {
          const int i = 8 * ti + fr, j = 8 * tj + 2 * fc;
          Ms[i * NDP + j] = acc[idx][0];
          Ms[i * NDP + j + 1] = acc[idx][1];
          if (ti != tj) {
            Ms[j * NDP + i] = acc[idx][0];
            Ms[(j + 1) * NDP + i] = acc[idx][1];
          }
        }
    }
    __syncwarp();
    double* M = mats + (size_t)(slots ? (int64_t)slots[e] : e) * ND * ND;
    for (int k = lane; k < ND * ND; k += 32) {
      const int i = k / ND, j = k - i * ND;
      const double val = Ms[i * NDP + j];
      if (accumulate) M[k] += val; else M[k] = val;
    }
  }
}

// ---------------------------------------------------------------- affine fast path
template <int D>
__global__ void __launch_bounds__(256) k_elmat_affine(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
    const double* __restrict__ vtx_xyz, int nd, const double* __restrict__ Tref, TermsDev T, int accumulate,
    const int32_t* __restrict__ slots, double* __restrict__ mats)
{
  constexpr int NG = D * (D + 1) / 2;
  constexpr int NC = 1 + NG;
  const int lane = threadIdx.x & 31;
  const int nn = nd * nd;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // reference tensor entries owned by this lane (idx = lane + 32 k), k < 4 kept in registers
  double tr[4][NC];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int idx = lane + 32 * k;
#pragma unroll
    for (int m = 0; m < NC; m++) tr[k][m] = (idx < nn) ? __ldg(Tref + m * nn + idx) : 0.0;
  }
  const int64_t ngroups = (n + 31) >> 5;
  for (int64_t g = warp0; g < ngroups; g += nwarps) {
    const int64_t e = (g << 5) + lane;
    double coef[NC];
#pragma unroll
    for (int m = 0; m < NC; m++) coef[m] = 0.0;
    if (e < n) {
      const int64_t el = elems ? (int64_t)elems[e] : e;
      double vx[D + 1][3], F[D * D], Fi[D * D];
#pragma unroll
      for (int k = 0; k < D + 1; k++) {
        const int v = __ldg(el_vtx + el * (D + 1) + k);
#pragma unroll
        for (int c = 0; c < 3; c++) vx[k][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v * D + c) : 0.0;
      }
      xg_affine_jac<D>(vx, F);
      const double meas = fabs(xg_inv<D>(F, Fi));
      const double h = (D == 1) ? meas : ((D == 2) ? sqrt(meas) : cbrt(meas));
      double cm = 0.0, cs = 0.0;
      for (int t = 0; t < T.n; t++) {
        const double c = T.coef[t] * ipow(h, T.hpow[t]) * meas;
        if (T.op_test[t] == 0) cm += c; else cs += c;
      }
      coef[0] = cm;
      int m = 1;
#pragma unroll
      for (int a = 0; a < D; a++)
#pragma unroll
        for (int b = a; b < D; b++) {
          double gab = 0.0;
#pragma unroll
          for (int c = 0; c < D; c++) gab += Fi[a * D + c] * Fi[b * D + c];
          coef[m++] = cs * gab;
        }
    }
    const int cnt = (int)min((int64_t)32, n - (g << 5));
    for (int s = 0; s < cnt; s++) {
      double cc[NC];
#pragma unroll
      for (int m = 0; m < NC; m++) cc[m] = __shfl_sync(0xffffffffu, coef[m], s);
      const int64_t me = (g << 5) + s;
      double* M = mats + (size_t)(slots ? (int64_t)slots[me] : me) * nn;
      if (nn <= 128) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int idx = lane + 32 * k;
          if (idx < nn) {
            double r = 0.0;
#pragma unroll
            for (int m = 0; m < NC; m++) r += cc[m] * tr[k][m];
            if (accumulate) M[idx] += r; else M[idx] = r;
          }
        }
      } else {
        for (int idx = lane; idx < nn; idx += 32) {
          double r = 0.0;
#pragma unroll
          for (int m = 0; m < NC; m++) r += cc[m] * __ldg(Tref + m * nn + idx);
          if (accumulate) M[idx] += r; else M[idx] = r;
        }
      }
    }
  }
}

int make_terms(xgpu_ctx* ctx, const xgpu_term* th, int nt, TermsDev& T, bool* only_mass_stiff)
{
  if (nt < 0 || nt > XG_MAXTERMS) return xgpu_fail(ctx, XGPU_ERR_ARG, "elmat: too many terms (max 8)");
  T.n = nt;
  *only_mass_stiff = true;
  for (int t = 0; t < nt; t++) {
    const int ou = th[t].op_trial, ot = th[t].op_test;
    if (ou < 0 || ou > 2 || ot < 0 || ot > 2 || ((ou == 1) != (ot == 1)))
      return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elmat: unsupported operator pair (menu: ID/DN x ID/DN, GRAD.GRAD)");
    T.op_trial[t] = ou; T.op_test[t] = ot; T.hpow[t] = th[t].hpow; T.coef[t] = th[t].coef;
    if (!((ou == 0 && ot == 0) || (ou == 1 && ot == 1))) *only_mass_stiff = false;
  }
  return XGPU_OK;
}

template <int D, int P>
int launch_ms(xgpu_ctx* ctx, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
              const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T,
              int accumulate, const int32_t* slots, double* mats, const Rule2* r2 = nullptr)
{
  const int nd = xg_h1_ndof(D, P);
  const int ndp = (nd + 1) & ~1;
  const int PS = ndp * (1 + D) + 2;
  int CH = 32;
  while (CH > 4 && (size_t)4 * CH * PS * sizeof(double) > 56 * 1024) CH >>= 1;
  const DefDev df = xg_defdev(ctx, D);
  Rule2 R2;
  if (r2) R2 = *r2; else { R2.offsets = nullptr; R2.pts = nullptr; R2.wts = nullptr; R2.T.n = 0; }
  const size_t sm = (size_t)4 * (CH * PS + (df.order > 0 ? XG_DEFS : 0)) * sizeof(double);
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 16 && !ctx->opt_short_blocks) grid = XGPU_NSM * 16;
  if (df.order > 0) {
    if constexpr (D >= 2) {
      XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_ms<D, P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_elmat_ms<D, P, true><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, CH, T,
                                                             accumulate, slots, mats, df, R2);
    } else return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elmat: deformation needs dim >= 2");
  } else {
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_ms<D, P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_elmat_ms<D, P, false><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, CH, T,
                                                            accumulate, slots, mats, df, R2);
  }
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

template <int D, int P>
int launch_dmma(xgpu_ctx* ctx, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
                const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T,
                int accumulate, const int32_t* slots, double* mats)
{
  using G = DmmaGeom<D, P>;
  // points per chunk: 32 (one per lane) when 4-5 warps of slabs fit the 227 KB of an SM
  const DefDev df = xg_defdev(ctx, D);
  const size_t dfs = df.order > 0 ? XG_DEFS : 0;
  int CH = 32, nw = 4;
  if ((size_t)5 * (CH * G::PS + dfs) * sizeof(double) <= 220 * 1024 && (size_t)4 * CH * G::PS * sizeof(double) > 110 * 1024) nw = 5;
  while (CH > 8 && (size_t)nw * (CH * G::PS + dfs) * sizeof(double) > 220 * 1024) CH >>= 1;
  if (CH * G::PS < G::NDP * G::NDP) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "elmat: staging buffer smaller than the element matrix");
  const size_t sm = (size_t)nw * (CH * G::PS + dfs) * sizeof(double);
  unsigned grid = nblk(n, nw);
  if (grid > XGPU_NSM * 8) grid = XGPU_NSM * 8;
  if (df.order > 0) {
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_dmma<D, P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_elmat_dmma<D, P, true><<<grid, nw * 32, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, CH, T,
                                                                 accumulate, slots, mats, df);
  } else {
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_dmma<D, P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_elmat_dmma<D, P, false><<<grid, nw * 32, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, CH, T,
                                                                  accumulate, slots, mats, df);
  }
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace
