// elmat_lp.cu — cut-element matrices of mass + stiffness forms at order <= 2: ONE LANE PER QUADRATURE POINT
// (SymbolicCutBilinearFormIntegrator::T_CalcElementMatrixAdd, xfem/symboliccutbfi.cpp:76-486, for the ID*ID and
// GRAD.GRAD terms; same entry points and semantics as k_elmat_ms in elmat.cuh, which stays the path for order >= 3).
//
// Why: k_elmat_ms stages the basis of every point in shared memory and lets half warps own 2x2 blocks of the matrix.
// On B200 an LDS.128 costs 4 wavefronts whatever the broadcast (measured: 40.7 M shared loads = 155 M wavefronts,
// profiles/r02_notes.md), so the 9 loads per 22 DFMA of that scheme keep the LSU 70 % busy and the FP64 pipe at 46 %,
// half of it basis evaluation in rounds with 6-14 active lanes (NEG and IF rules staged separately).
// Here a lane takes one point of the element (NEG and IF points in ONE sequence: 20 / 48 / 54 points on config 2),
// evaluates the basis with PHYSICAL gradients carried by the AD type (lam_i.d = F^{-T} e_i: no per-function
// F^{-T} product), and accumulates the whole symmetric matrix (ND(ND+1)/2 = 55 entries for P2 tets) in registers:
// 4 DFMA per entry and point, no shared-memory traffic in the contraction.  The lanes' partial matrices are summed
// once per element through a [NSYM][34] shared-memory tile (conflict-free column writes, LDS.128 row reads), in a
// fixed order: bit-reproducible run to run.
#include "elmat.cuh"

namespace {

template <int D, int ND>
struct RegSink {   // (kept for orders whose shapes need the AD recurrences)
  double* val;   // [ND]
  double* g;     // [D][ND]
  __device__ __forceinline__ void operator()(int i, const AD<D>& a)
  {
    val[i] = a.v;
#pragma unroll
    for (int c = 0; c < D; c++) g[c * ND + i] = a.d[c];
  }
};

// the shapes of order P <= 2 from barycentric coordinates with gradients, in closed form: vertex functions lam_i, edge
// functions L_2(lam_e - lam_s, lam_e + lam_s) = 0.5 ((lam_e - lam_s)^2 - (lam_e + lam_s)^2) = -2 lam_s lam_e (h1.cuh; the
// orientation does not matter at order 2) — 7 FP64 instructions per edge function instead of ~30 through the AD recurrences
template <int D, int P, int ND>
__device__ __forceinline__ void lp_shapes(const AD<D>* lam, double* val, double* g)
{
  static_assert(P >= 1 && P <= 2, "closed forms for orders 1, 2");
#pragma unroll
  for (int i = 0; i < D + 1; i++) {
    val[i] = lam[i].v;
#pragma unroll
    for (int c = 0; c < D; c++) g[c * ND + i] = lam[i].d[c];
  }
  if (P == 2) {
    constexpr int NED = (D == 1) ? 1 : ((D == 2) ? 3 : 6);
#pragma unroll
    for (int e = 0; e < NED; e++) {
      const int ea = (D == 1) ? 0 : ((D == 2) ? XG_TRIG_EDGE_A(e) : XG_TET_EDGE_A(e));
      const int eb = (D == 1) ? 1 : ((D == 2) ? XG_TRIG_EDGE_B(e) : XG_TET_EDGE_B(e));
      const double m2a = -2.0 * lam[ea].v, m2b = -2.0 * lam[eb].v;
      val[D + 1 + e] = m2a * lam[eb].v;
#pragma unroll
      for (int c = 0; c < D; c++) g[c * ND + D + 1 + e] = m2a * lam[eb].d[c] + m2b * lam[ea].d[c];
    }
  }
}

template <int D, int P, bool DEF>
__global__ void __launch_bounds__(128, 2) k_elmat_lp(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
    const double* __restrict__ vtx_xyz, const int64_t* __restrict__ offsets, const double* __restrict__ pts,
    const double* __restrict__ wts, TermsDev T, int accumulate,
    const int32_t* __restrict__ slots, double* __restrict__ mats, DefDev df, Rule2 R2)
{
  constexpr int ND = (D == 1) ? P + 1 : ((D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2
                                                   : 4 + 6 * (P - 1) + 4 * ((P - 1) * (P - 2) / 2) + (P - 1) * (P - 2) * (P - 3) / 6);
  constexpr int NSYM = ND * (ND + 1) / 2, RS = 34, MAXK = (NSYM + 31) / 32;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* S = smem + (size_t)wib * (NSYM * RS + (DEF ? XG_DEFS : 0));
  double* defs = S + NSYM * RS;
  // the entries this lane sums and stores: k = lane + 32 m  ->  (i <= j)
  int ki_[MAXK], kj_[MAXK];
#pragma unroll
  for (int m = 0; m < MAXK; m++) {
    const int k = lane + 32 * m;
    ki_[m] = -1; kj_[m] = 0;
    if (k < NSYM) {
      int r = 0, rem = k;
      while (rem >= ND - r) { rem -= ND - r; r++; }
      ki_[m] = r; kj_[m] = r + rem;
    }
  }
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
    if (DEF) { __syncwarp(); xg_def_gather<D>(df, el, lane, defs); __syncwarp(); }
    double cm0, cs0, cm1 = 0.0, cs1 = 0.0;
    ms_coefs(T, xg_mesh_size<D>(meas), meas, cm0, cs0);
    const int64_t q0 = offsets[e];
    const int n0 = (int)(offsets[e + 1] - q0);
    int64_t q1 = 0;
    int n1 = 0;
    if (R2.offsets) {
      ms_coefs(R2.T, xg_mesh_size<D>(meas), meas, cm1, cs1);
      q1 = R2.offsets[e];
      n1 = (int)(R2.offsets[e + 1] - q1);
    }
    const int ntot = n0 + n1;
    double acc[NSYM];
#pragma unroll
    for (int k = 0; k < NSYM; k++) acc[k] = 0.0;
#pragma unroll 1
    for (int base = 0; base < ntot; base += 32) {
      const int q = base + lane;
      if (q < ntot) {
        const bool second = q >= n0;
        const int64_t idx = second ? q1 + (q - n0) : q0 + q;
        const double* xq = (second ? R2.pts : pts) + 3 * idx;
        const double xh[3] = {xq[0], xq[1], xq[2]};
        const double w = (second ? R2.wts : wts)[idx];
        double wm = (second ? cm1 : cm0) * w, ws = (second ? cs1 : cs0) * w;
        double Fiq[D * D];
#pragma unroll
        for (int k = 0; k < D * D; k++) Fiq[k] = Fi[k];
        if (DEF) {
          double Fq[D * D], xd[3] = {0, 0, 0};
#pragma unroll
          for (int k = 0; k < D * D; k++) Fq[k] = F[k];
          xg_def_apply<D>(df.order, v, xh, defs, Fq, xd);
          const double measq = fabs(xg_inv<D>(Fq, Fiq));
          double cmq, csq;
          ms_coefs(second ? R2.T : T, xg_mesh_size<D>(measq), measq, cmq, csq);
          wm = cmq * w; ws = csq * w;
        }
        // barycentric coordinates with physical gradients: grad lam_i = F^{-T} e_i (row i of Fi), grad lam_D = -sum
        AD<D> lam[D + 1];
        double s = 1.0;
#pragma unroll
        for (int c = 0; c < D; c++) lam[D].d[c] = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          lam[i].v = xh[i]; s -= xh[i];
#pragma unroll
          for (int c = 0; c < D; c++) { lam[i].d[c] = Fiq[i * D + c]; lam[D].d[c] -= Fiq[i * D + c]; }
        }
        lam[D].v = s;
        double val[ND], g[D * ND];
        lp_shapes<D, P, ND>(lam, val, g);
        {
          int k = 0;
#pragma unroll
          for (int i = 0; i < ND; i++) {
            const double ai = wm * val[i];
#pragma unroll
            for (int j = i; j < ND; j++) acc[k++] += ai * val[j];
          }
        }
#pragma unroll
        for (int c = 0; c < D; c++) {
          int k = 0;
#pragma unroll
          for (int i = 0; i < ND; i++) {
            const double ai = ws * g[c * ND + i];
#pragma unroll
            for (int j = i; j < ND; j++) acc[k++] += ai * g[c * ND + j];
          }
        }
      }
    }
    // sum the 32 partial matrices: column `lane` of every entry row, then lane k adds row k in index order
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NSYM; k++) S[k * RS + lane] = acc[k];
    __syncwarp();
    double* M = mats + (size_t)(slots ? (int64_t)slots[e] : e) * ND * ND;
#pragma unroll
    for (int m = 0; m < MAXK; m++) {
      if (ki_[m] >= 0) {
        const double2* row = reinterpret_cast<const double2*>(S + (lane + 32 * m) * RS);
        double r = 0.0;
#pragma unroll
        for (int t = 0; t < 16; t++) { const double2 x = row[t]; r += x.x; r += x.y; }
        const int i = ki_[m], j = kj_[m];
        if (accumulate) { M[i * ND + j] += r; if (i != j) M[j * ND + i] += r; }
        else { M[i * ND + j] = r; if (i != j) M[j * ND + i] = r; }
      }
    }
  }
}

template <int D, int P>
int launch_lp(xgpu_ctx* ctx, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
              const int64_t* offsets, const double* pts, const double* wts, const TermsDev& T,
              int accumulate, const int32_t* slots, double* mats, const Rule2* r2)
{
  constexpr int ND = (D == 1) ? P + 1 : ((D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2 : 4 + 6 * (P - 1));
  constexpr int NSYM = ND * (ND + 1) / 2;
  const DefDev df = xg_defdev(ctx, D);
  Rule2 R2;
  if (r2) R2 = *r2; else { R2.offsets = nullptr; R2.pts = nullptr; R2.wts = nullptr; R2.T.n = 0; }
  const size_t sm = (size_t)4 * (NSYM * 34 + (df.order > 0 ? XG_DEFS : 0)) * sizeof(double);
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 8 && !ctx->opt_short_blocks) grid = XGPU_NSM * 8;
  if (df.order > 0) {
    if constexpr (D >= 2) {
      XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_lp<D, P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_elmat_lp<D, P, true><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, T, accumulate, slots, mats, df, R2);
    } else return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "elmat: deformation needs dim >= 2");
  } else {
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_lp<D, P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_elmat_lp<D, P, false><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, offsets, pts, wts, T, accumulate, slots, mats, df, R2);
  }
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}


// ---------------------------------------------------------------- cut rules made on the fly
// The reference never stores a cut rule: T_CalcElementMatrixAdd calls CreateCutIntegrationRule for the element at hand
// (xfem/symboliccutbfi.cpp:105-121) and contracts at once.  k_elmat_cut does the same for up to two integrators (the NEG
// volume terms and the IF interface term of a CutFEM form): a block takes 32 listed elements; phase 1, one THREAD per
// (element, integrator): vertex / level-set gathers, F^{-1}, |det F|, folded term weights, Decompose (<= 3 sub-simplices,
// straightcutrule.cpp:131-196) and the interface factor go to shared memory — all dependent global loads of the batch are
// in flight together; phase 2, one WARP per element and one LANE per quadrature point as in k_elmat_lp, the points mapped
// from the staged sub-simplices (GetPlainIntegrationRule :80-100): no global loads, no rule arrays in HBM (32 B per point
// written by xgpu_cutquad_fill and read back by the element-matrix kernel before).
struct CutSeg {
  int dt;                                              // XGPU_NEG / POS / IF, -1: absent
  const double* rs_p; const double* rs_w; int np_sub;  // rule on the sub-simplices (dimension D, D-1 for IF)
  const double* rv_p; const double* rv_w; int np_vol;  // rule on the whole element (uncut element of the matching type)
  TermsDev T;
};

struct CutDesc {
  double p[3][4][3];  // sub-simplex vertices (reference coordinates)
  double vol[3];      // their measures
  double iffac;       // interface weight factor ||F^{-T} n_ref|| (1 for NEG / POS)
  double cm, cs;      // folded mass / stiffness term weights of this integrator on this element
  int ns;             // > 0 sub-simplices, 0 = standard rule of an uncut matching element, < 0 = nothing / error
  int sd;             // dimension of the sub-simplices
  int cnt;            // points
};

// rows [R0, R1) of the symmetric matrix of one element: one lane per quadrature point, the entries (a, b >= a) of these rows
// in registers, partial matrices summed through the shared tile S in fixed order.  The full P2-tet matrix (55 entries) needs
// ~245 registers = 8 warps per SM; two passes over the points with 27 / 28 entries each need ~150 (12+ warps per SM) and
// recompute only the point map and the closed-form shapes (~100 of ~500 instructions per round).
template <int ND>
__host__ __device__ constexpr int lp_entries(int r0, int r1)
{
  int n = 0;
  for (int a = r0; a < r1; a++) n += ND - a;
  return n;
}
template <int ND>
__host__ __device__ constexpr int lp_split_row()
{
  int r = 0;
  while (2 * lp_entries<ND>(0, r) < ND * (ND + 1) / 2) r++;
  return r;
}

struct CutSeg;
struct CutDesc;
template <int D, int P, int R0, int R1>
__device__ __forceinline__ void cut_pass(const CutDesc& A, const CutDesc& B, const CutSeg& SA, const CutSeg& SB,
                                         const double* ruleA, const double* ruleB, int npa, int npb, const double* Fi,
                                         int lane, double* S, double* M, int accumulate);

#define XG_CUT_EB 32  // elements per block and pass

template <int D, int P, bool SPLIT>
__global__ void __launch_bounds__(128, SPLIT ? 3 : 2) k_elmat_cut(
    int64_t n, const int32_t* __restrict__ elems, const int32_t* __restrict__ el_vtx,
    const double* __restrict__ vtx_xyz, const double* __restrict__ phi, CutSeg SA, CutSeg SB, int accumulate,
    const int32_t* __restrict__ slots, double* __restrict__ mats, int* status, unsigned long long* status_elem)
{
  constexpr int ND = (D == 1) ? P + 1 : ((D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2
                                                   : 4 + 6 * (P - 1) + 4 * ((P - 1) * (P - 2) / 2) + (P - 1) * (P - 2) * (P - 3) / 6);
  constexpr int NSYM = ND * (ND + 1) / 2, RS = 34;
  constexpr int NTILE = SPLIT ? lp_entries<ND>(lp_split_row<ND>(), ND) > lp_entries<ND>(0, lp_split_row<ND>()) ? lp_entries<ND>(lp_split_row<ND>(), ND) : lp_entries<ND>(0, lp_split_row<ND>()) : NSYM;   // rows of the shared tile
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int npa = SA.np_sub, npb = SB.dt >= 0 ? SB.np_sub : 0;
  double* ruleA = smem;                              // npa x 4 (xi0, xi1, xi2, w)
  double* ruleB = ruleA + 4 * ((npa + 1) & ~1);
  double* geo = ruleB + 4 * ((npb + 1) & ~1);        // XG_CUT_EB x D*D inverse Jacobians
  CutDesc* desc = reinterpret_cast<CutDesc*>(geo + XG_CUT_EB * D * D);   // [2][XG_CUT_EB]
  int* next = reinterpret_cast<int*>(desc + 2 * XG_CUT_EB);   // phase-2 element counter (16 bytes reserved)
  double* S = reinterpret_cast<double*>(desc + 2 * XG_CUT_EB) + 2 + (size_t)wib * NTILE * RS;
  {
    const int da = (SA.dt == XGPU_IF) ? D - 1 : D, db = (SB.dt == XGPU_IF) ? D - 1 : D;   // dimension of the sub-simplices
    for (int t = threadIdx.x; t < npa; t += blockDim.x) {
#pragma unroll
      for (int m = 0; m < 3; m++) ruleA[4 * t + m] = (m < da) ? SA.rs_p[3 * t + m] : 0.0;
      ruleA[4 * t + 3] = SA.rs_w[t];
    }
    for (int t = threadIdx.x; t < npb; t += blockDim.x) {
#pragma unroll
      for (int m = 0; m < 3; m++) ruleB[4 * t + m] = (m < db) ? SB.rs_p[3 * t + m] : 0.0;
      ruleB[4 * t + 3] = SB.rs_w[t];
    }
  }
#pragma unroll 1
  for (int64_t base = (int64_t)blockIdx.x * XG_CUT_EB; base < n; base += (int64_t)gridDim.x * XG_CUT_EB) {
    __syncthreads();
    {  // phase 1: thread = (element, integrator)
      const int k = threadIdx.x & 31, role = threadIdx.x >> 5;   // warp 0: integrator A, warp 1: integrator B (uniform domain type per warp)
      const int64_t i = base + k;
      if (role < 2) {
        CutDesc& Dd = desc[role * XG_CUT_EB + k];
        const CutSeg& Sg = role ? SB : SA;
        Dd.ns = -1; Dd.cnt = 0; Dd.sd = D;
        if (i < n && Sg.dt >= 0) {
          const int64_t el = elems ? (int64_t)elems[i] : i;
          int v[D + 1];
          double l[D + 1], vx[D + 1][3], F[D * D], Fi[D * D];
#pragma unroll
          for (int a = 0; a < D + 1; a++) { v[a] = __ldg(el_vtx + el * (D + 1) + a); l[a] = __ldg(phi + v[a]); }
#pragma unroll
          for (int a = 0; a < D + 1; a++)
#pragma unroll
            for (int c = 0; c < 3; c++) vx[a][c] = (c < D) ? __ldg(vtx_xyz + (int64_t)v[a] * D + c) : 0.0;
          xg_affine_jac<D>(vx, F);
          const double meas = fabs(xg_inv<D>(F, Fi));
          if (role == 0) {
#pragma unroll
            for (int a = 0; a < D * D; a++) geo[k * D * D + a] = Fi[a];
          }
          ms_coefs(Sg.T, xg_mesh_size<D>(meas), meas, Dd.cm, Dd.cs);
          const int dom = xg_check_cut(l, D + 1);
          if (dom != XGPU_IF) {
            if (dom == Sg.dt) { Dd.ns = 0; Dd.cnt = Sg.np_vol; }  // uncut, matching: the standard rule (:669-674)
          } else {
            XgSimplex ref, sub[3];
            xg_ref_simplex(D, ref);
            const int ns = xg_decompose(ref, l, Sg.dt, sub);
            if (ns < 0) xgpu_record_error(status, status_elem, ns, el);
            else {
              for (int s = 0; s < ns; s++) {
                Dd.vol[s] = xg_volume(sub[s]);
                for (int a = 0; a < 4; a++) for (int c = 0; c < 3; c++) Dd.p[s][a][c] = sub[s].p[a][c];
              }
              Dd.sd = sub[0].D;
              Dd.iffac = (Sg.dt == XGPU_IF) ? xg_if_weight_factor<D>(Fi, l) : 1.0;   // (:558-578)
              Dd.cnt = ns * Sg.np_sub;
            }
            Dd.ns = ns;
          }
        }
      }
    }
    if (threadIdx.x == 0) *next = 0;
    __syncthreads();
    // phase 2: warp = element, lane = quadrature point
    // (not unrolled: the body is ~1400 straight-line instructions, eight copies of it thrash the instruction cache)
#pragma unroll 1
    for (;;) {
      // the warps take the next element of the batch as they finish one (20 to 54 points per element: static shares wait)
      int k = 0;
      if (lane == 0) k = atomicAdd(next, 1);
      k = __shfl_sync(0xffffffffu, k, 0);
      const int64_t i = base + k;
      if (k >= XG_CUT_EB || i >= n) break;
      const CutDesc& A = desc[k];
      const CutDesc& B = desc[XG_CUT_EB + k];
      const int n0 = A.cnt, ntot = n0 + B.cnt;
      double Fi[D * D];
#pragma unroll
      for (int a = 0; a < D * D; a++) Fi[a] = geo[k * D * D + a];
      double* M = mats + (size_t)(slots ? (int64_t)slots[i] : i) * ND * ND;
      if constexpr (SPLIT) {
        constexpr int RM = lp_split_row<ND>();
        cut_pass<D, P, 0, RM>(A, B, SA, SB, ruleA, ruleB, npa, npb, Fi, lane, S, M, accumulate);
        cut_pass<D, P, RM, ND>(A, B, SA, SB, ruleA, ruleB, npa, npb, Fi, lane, S, M, accumulate);
      } else {
        cut_pass<D, P, 0, ND>(A, B, SA, SB, ruleA, ruleB, npa, npb, Fi, lane, S, M, accumulate);
      }
    }
  }
}

template <int D, int P, int R0, int R1>
__device__ __forceinline__ void cut_pass(const CutDesc& A, const CutDesc& B, const CutSeg& SA, const CutSeg& SB,
                                         const double* ruleA, const double* ruleB, int npa, int npb, const double* Fi,
                                         int lane, double* S, double* M, int accumulate)
{
  constexpr int ND = (D == 1) ? P + 1 : ((D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2
                                                   : 4 + 6 * (P - 1) + 4 * ((P - 1) * (P - 2) / 2) + (P - 1) * (P - 2) * (P - 3) / 6);
  constexpr int NH = lp_entries<ND>(R0, R1), RS = 34, MAXK = (NH + 31) / 32;
  const int n0 = A.cnt, ntot = n0 + B.cnt;
  double acc[NH];
#pragma unroll
  for (int a = 0; a < NH; a++) acc[a] = 0.0;
#pragma unroll 1
  for (int qb = 0; qb < ntot; qb += 32) {
    const int q = qb + lane;
    if (q < ntot) {
      const bool second = q >= n0;
      const CutDesc& Dd = second ? B : A;
      const int qq = second ? q - n0 : q;
      double xh[3], w;
      if (Dd.ns == 0) {
        const double* rp = second ? SB.rv_p : SA.rv_p;
        xh[0] = rp[3 * qq]; xh[1] = rp[3 * qq + 1]; xh[2] = rp[3 * qq + 2];
        w = (second ? SB.rv_w : SA.rv_w)[qq];
      } else {
        const int nps = second ? npb : npa;
        const int s = (qq >= nps) + (qq >= 2 * nps), qi = qq - s * nps;   // <= 3 sub-simplices
        // branch-free point map: the staged rule has zeros in the coordinates beyond the sub-simplex dimension and
        // xg_set fills the unused vertices with finite values, so the sum over all three is the sum over sd of them
        const double* xi = (second ? ruleB : ruleA) + 4 * qi;
        const double x0 = xi[0], x1 = xi[1], x2 = xi[2];
        const double ow = ((1.0 - x0) - x1) - x2;
        const double* pp = &Dd.p[s][0][0];
#pragma unroll
        for (int c = 0; c < 3; c++) xh[c] = ((ow * pp[c] + x0 * pp[3 + c]) + x1 * pp[6 + c]) + x2 * pp[9 + c];
        w = xi[3] * Dd.vol[s] * Dd.iffac;
      }
      const double wm = Dd.cm * w, ws = Dd.cs * w;
      AD<D> lam[D + 1];
      double sl = 1.0;
#pragma unroll
      for (int c = 0; c < D; c++) lam[D].d[c] = 0.0;
#pragma unroll
      for (int a = 0; a < D; a++) {
        lam[a].v = xh[a]; sl -= xh[a];
#pragma unroll
        for (int c = 0; c < D; c++) { lam[a].d[c] = Fi[a * D + c]; lam[D].d[c] -= Fi[a * D + c]; }
      }
      lam[D].v = sl;
      double val[ND], g[D * ND];
      lp_shapes<D, P, ND>(lam, val, g);
      {
        int kk = 0;
#pragma unroll
        for (int a = R0; a < R1; a++) {
          const double ai = wm * val[a];
#pragma unroll
          for (int b = a; b < ND; b++) acc[kk++] += ai * val[b];
        }
      }
#pragma unroll
      for (int c = 0; c < D; c++) {
        int kk = 0;
#pragma unroll
        for (int a = R0; a < R1; a++) {
          const double ai = ws * g[c * ND + a];
#pragma unroll
          for (int b = a; b < ND; b++) acc[kk++] += ai * g[c * ND + b];
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int a = 0; a < NH; a++) S[a * RS + lane] = acc[a];
  __syncwarp();
#pragma unroll
  for (int m = 0; m < MAXK; m++) {
    const int kk = lane + 32 * m;
    if (kk < NH) {
      const double2* row = reinterpret_cast<const double2*>(S + kk * RS);
      double r0 = 0.0, r1 = 0.0;
#pragma unroll
      for (int t = 0; t < 16; t++) { const double2 x = row[t]; r0 += x.x; r1 += x.y; }
      const double r = r0 + r1;
      int a = R0, rem = kk;
      while (rem >= ND - a) { rem -= ND - a; a++; }
      const int b = a + rem;
      if (accumulate) { M[a * ND + b] += r; if (a != b) M[b * ND + a] += r; }
      else { M[a * ND + b] = r; if (a != b) M[b * ND + a] = r; }
    }
  }
  __syncwarp();
}

template <int D, int P>
int launch_cut(xgpu_ctx* ctx, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
               const CutSeg& SA, const CutSeg& SB, int accumulate, const int32_t* slots, double* mats)
{
  constexpr int ND = (D == 1) ? P + 1 : ((D == 2) ? 3 + 3 * (P - 1) + (P - 1) * (P - 2) / 2 : 4 + 6 * (P - 1));
  constexpr int NSYM = ND * (ND + 1) / 2;
  // SPLIT = two passes over the points with half of the matrix rows each (167 registers, 12 warps per SM instead of 244 / 8):
  // measured 0.74 ms against 0.705 ms for the single pass on config 2 — more warps do not pay for the repeated point map and
  // shapes, so it stays off (profiles/r02_notes.md)
  constexpr bool SPLIT = false;
  constexpr int RM = lp_split_row<ND>();
  constexpr int NTILE = SPLIT ? (lp_entries<ND>(RM, ND) > lp_entries<ND>(0, RM) ? lp_entries<ND>(RM, ND) : lp_entries<ND>(0, RM)) : NSYM;
  const int npb = SB.dt >= 0 ? SB.np_sub : 0;
  const size_t sm_need = sizeof(double) * (4 * ((SA.np_sub + 1) & ~1) + 4 * ((npb + 1) & ~1) + XG_CUT_EB * D * D + 2 + (size_t)4 * NTILE * 34)
                         + 2 * XG_CUT_EB * sizeof(CutDesc);
  size_t sm = sm_need;
  if (ctx->opt_cut_bpsm == 1 && sm < 116 * 1024) sm = 116 * 1024;   // more than half of an SM's shared memory: one block per SM
  const int bpsm = ctx->opt_cut_bpsm == 1 ? 1 : (SPLIT ? 3 : 2);
  unsigned grid = nblk(n, XG_CUT_EB);
  if (grid > (unsigned)(XGPU_NSM * 2 * bpsm) && !ctx->opt_short_blocks) grid = XGPU_NSM * 2 * bpsm;
  XGPU_CUDA(ctx, cudaFuncSetAttribute(k_elmat_cut<D, P, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  k_elmat_cut<D, P, SPLIT><<<grid, 128, sm, ctx->stream>>>(n, elems, el_vtx, vtx_xyz, phi, SA, SB, accumulate, slots, mats,
                                                           ctx->d_status, ctx->d_status_elem);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

int make_seg(xgpu_ctx* ctx, int dim, int dt, int order, const xgpu_term* th, int nt, CutSeg& S)
{
  S.dt = dt;
  if (dt < 0) { S.rs_p = S.rs_w = S.rv_p = S.rv_w = nullptr; S.np_sub = S.np_vol = 0; S.T.n = 0; return XGPU_OK; }
  if (dt > 2) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_cut_pair: domain type");
  const xgpu_rule_host* rv = xgpu_rule_lookup(ctx, dim, order);
  const xgpu_rule_host* rs = (dt == XGPU_IF) ? xgpu_rule_lookup(ctx, dim - 1, order) : rv;
  if (!rv || !rs) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_cut_pair: no rule of that order");
  S.rs_p = rs->d_pts; S.rs_w = rs->d_wts; S.np_sub = rs->n;
  S.rv_p = rv->d_pts; S.rv_w = rv->d_wts; S.np_vol = rv->n;
  bool ms;
  int rc = make_terms(ctx, th, nt, S.T, &ms);
  if (rc) return rc;
  if (!ms) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_elmat_cut_pair: ID*ID and GRAD.GRAD terms only (use xgpu_cutquad_* + xgpu_elmat_rule)");
  return XGPU_OK;
}

}  // namespace

// per-element rules (offsets != nullptr), order 1 or 2
int xg_launch_lp(xgpu_ctx* ctx, int dim, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz,
                 const int64_t* offsets, const double* pts, const double* wts, const TermsDev& T, int accumulate,
                 const int32_t* slots, double* mats, const Rule2* r2)
{
#define XG_LP_CASE(D_, P_) if (dim == D_ && p == P_) return launch_lp<D_, P_>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, T, accumulate, slots, mats, r2)
  XG_LP_CASE(1, 1); XG_LP_CASE(1, 2); XG_LP_CASE(2, 1); XG_LP_CASE(2, 2); XG_LP_CASE(3, 1); XG_LP_CASE(3, 2);
#undef XG_LP_CASE
  return xgpu_fail(ctx, XGPU_ERR_ARG, "elmat_lp: orders 1..2");
}

extern "C" int xgpu_elmat_cut_pair(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* elems,
                                   const int32_t* el_vtx, const double* vtx_xyz, const double* phi,
                                   int dt_a, int order_a, const xgpu_term* terms_a_host, int nterms_a,
                                   int dt_b, int order_b, const xgpu_term* terms_b_host, int nterms_b,
                                   int accumulate, const int32_t* slots, double* mats)
{
  XGPU_NVTX(ctx, "xgpu_elmat_cut_pair");
  if (!ctx || dim < 1 || dim > 3 || dt_a < 0 || !el_vtx || !vtx_xyz || !phi || (!mats && n > 0))
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_cut_pair: bad argument");
  if (fe_order < 1 || fe_order > 2)
    return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_elmat_cut_pair: orders 1..2 (higher orders: xgpu_cutquad_* + xgpu_elmat_rule)");
  if (ctx->def_order > 0)
    return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_elmat_cut_pair: affine meshes only (deformed: xgpu_cutquad_* + xgpu_elmat_rule)");
  if (n == 0) return XGPU_OK;
  CutSeg SA, SB;
  int rc = make_seg(ctx, dim, dt_a, order_a, terms_a_host, nterms_a, SA);
  if (rc) return rc;
  rc = make_seg(ctx, dim, dt_b < 0 ? -1 : dt_b, order_b, terms_b_host, nterms_b, SB);
  if (rc) return rc;
#define XG_CUT_CASE(D_, P_) if (dim == D_ && fe_order == P_) return launch_cut<D_, P_>(ctx, n, elems, el_vtx, vtx_xyz, phi, SA, SB, accumulate, slots, mats)
  XG_CUT_CASE(1, 1); XG_CUT_CASE(1, 2); XG_CUT_CASE(2, 1); XG_CUT_CASE(2, 2); XG_CUT_CASE(3, 1); XG_CUT_CASE(3, 2);
#undef XG_CUT_CASE
  return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elmat_cut_pair: bad argument");
}
