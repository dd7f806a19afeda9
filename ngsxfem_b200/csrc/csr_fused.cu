// csr_fused.cu — sparsity pattern AND values in ONE pass over the matrix rows.
// (RestrictedBilinearForm::GetGraph utils/restrictedblf.cpp:142-163 -> NGSolve MatrixGraph, followed by the
//  SparseMatrix::AddElementMatrix loop of BilinearForm::Assemble.)
//
// The two-pass path of csr.cu walks every row's adjacency and dof rows twice (hash-set pattern pass, then a
// gather that looks every contribution's column up again) and reads every element matrix from HBM.  Here a
// row is finished in one visit:
//   * a GROUP of GS lanes (GS = widest table row, e.g. 3 groups of 10 lanes for P2 tets) owns one matrix row.
//     Its open-addressing table in shared memory holds (column, FP64 accumulator) per slot; the adjacency is
//     walked in ascending (batch, table row, local index) order and the lanes of the group take the columns of
//     one adjacent table row per step: find-or-insert, then add — columns inside a table row are distinct, so
//     there is no conflict inside a step and every nonzero is summed in exactly the order of a sequential
//     AddElementMatrix loop (no floating-point atomics, bit-reproducible).
//   * PROCEDURAL batches: an uncut affine element's matrix is |det F| (c_m M_ref + c_s sum_ab G_ab K_ref^ab);
//     its rows are generated from the element's 1 + D(D+1)/2 coefficients (xgpu_affine_coefs, 64 B per element)
//     and the reference tensors in shared memory with the FMA chain of k_elmat_affine — the 800 B matrices of
//     93 % of config 2's elements are never written to or read from HBM.  Other batches (cut elements) bring
//     their matrices in memory as before.
//   * the slots are compacted in place, rank-sorted (all distinct: rank = #smaller) and written straight to
//     their final CSR position — no candidate-offset scratch, no compaction copy.
//   * the row offsets come from a lean COUNT pass (k_rows_count: the same walk with keys only — no values, no
//     compaction, no sort; ~1/4 of the instructions of the full pass) + a scan, run by xgpu_csr_plan_rows, so nnz
//     is known before the caller allocates colidx / vals, and the full pass sizes every row's table from its
//     exact length.  (A single-pass variant with a decoupled look-back over per-warp tiles was measured first:
//     1.6 M tiles of 3 rows publish faster than a 32-wide look-back window can follow — the prefix propagation,
//     not the row work, set the time: 11.6 ms.  profiles/r02_notes.md)
#include "csr.cuh"
#include <algorithm>

namespace {

#define XF_WARPS 4
#define XF_GOLD 2654435761u
#define XF_CHUNK 16

struct FusedPar {
  int ndof, NG, GS, HG, nn;
  int skip_na;  // rows with 1..skip_na adjacent table rows (full pass: and at most 32 nonzeros) belong to csr_rows_small.cu
  int skip_cand;  // the other rows with at most skip_cand candidates belong to csr_rows_warp.cu (-1: none)
  unsigned ntiles;
  const long long* adj_ptr;
  const long long* adj;
  const int32_t* cand;
  BatchDev bd;
  const double* tref;
  unsigned int* ticket;
  const long long* rowptr;  // full pass: exclusive scan of the row lengths (count pass)
  int32_t* rowcnt;          // count pass: distinct columns per row
  int32_t* colidx;
  double* vals;
  int* stats;               // count pass: [0] longest row
};

// The tables live at 32-bit shared-window addresses and are touched through PTX accessors: with generic pointers
// the compiler rebuilt the window base (S2UR / ULEA) at every probe inside the loops.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_i32(unsigned a) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_i32(unsigned a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ double lds_f64_ro(unsigned a) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ int4 lds_i4(unsigned a)
{
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_i4(unsigned a, int x) { asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(x) : "memory"); }
__device__ __forceinline__ void sts_zero16(unsigned a) { asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0) : "memory"); }
__device__ __forceinline__ int cas_i32(unsigned a, int cmp, int val)
{
  int o;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(o) : "r"(a), "r"(cmp), "r"(val) : "memory");
  return o;
}

// find-or-insert: ONE compare-and-swap per probe (no separate load, no insert / found branch); returns the slot
__device__ __forceinline__ unsigned probe(unsigned keyA, unsigned hmask, int shift, int c, int& fresh)
{
  unsigned h = ((unsigned)c * XF_GOLD) >> shift;
  while (true) {
    const int old = cas_i32(keyA + (h << 2), XG_EMPTY, c);
    if (old == XG_EMPTY) { fresh++; break; }
    if (old == c) break;
    h = (h + 1) & hmask;
  }
  return h;
}

// ranks of the group's nu distinct keys (K per lane; rank = #smaller) and the final stores
template <int K>
__device__ __forceinline__ void rank_write(unsigned keyA, unsigned accA, int nu, int ls, int GS, int k0,
                                           int32_t* __restrict__ col, double* __restrict__ val)
{
  int x[K], r[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int i = k0 + ls + GS * k;
    x[k] = (i < nu) ? lds_i32(keyA + (i << 2)) : XG_EMPTY;
    r[k] = 0;
  }
  for (int j = 0; j < nu; j += 4) {
    const int4 v = lds_i4(keyA + (j << 2));  // padded with XG_EMPTY (never smaller)
#pragma unroll
    for (int k = 0; k < K; k++)  // keys are in [0, 2^31): the sign bit of the difference is the comparison
      r[k] += (int)(((unsigned)(v.x - x[k]) >> 31) + ((unsigned)(v.y - x[k]) >> 31) + ((unsigned)(v.z - x[k]) >> 31) +
                    ((unsigned)(v.w - x[k]) >> 31));
  }
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int i = k0 + ls + GS * k;
    if (i < nu) { col[r[k]] = x[k]; val[r[k]] = lds_f64(accA + (i << 3)); }
  }
}

#define XF_DECODE(e, bi, w, ro, loc)      \
  const int bi = XG_ADJ_BATCH(e);         \
  const int w = P.bd.width[bi];           \
  const int ro = XG_ADJ_ROW(e) * w; /* nrows*width < 2^31 (checked on the host) */ \
  const int loc = XG_ADJ_LOC(e);

template <int NC, int TRIPS>
__global__ void __launch_bounds__(XF_WARPS * 32) k_rows_fused(const FusedPar P)
{
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int GS = P.GS, NG = P.NG, HG = P.HG, nn = P.nn;
  const int ntref = (NC * nn + 1) & ~1;
  if (NC > 0) {
    for (int k = threadIdx.x; k < NC * nn; k += blockDim.x) sm[k] = __ldg(P.tref + k);
    __syncthreads();
  }
  const int g = lane / GS, ls = lane - g * GS;
  const bool live = g < NG;
  const int gi = live ? g : 0;
  const unsigned smA = smem_u32(sm);
  const unsigned trefA = smA + (ls << 3);
  const unsigned accA = smA + ((ntref + (wib * NG + gi) * HG) << 3);
  const unsigned keyA = smA + ((ntref + XF_WARPS * NG * HG) << 3) + (((wib * NG + gi) * (HG + 4)) << 2);
  const unsigned gmask = (GS >= 32) ? 0xffffffffu : ((1u << GS) - 1u);
  const unsigned FULL = 0xffffffffu;

  // tiles (NG consecutive rows) are handed out in chunks of XF_CHUNK by an atomic ticket: one ticket per tile made the
  // single counter the bottleneck (1.6 M same-address atomics ~ 1.4 ms, profiles/r02_notes.md)
  unsigned tk = 0, tk_end = 0;
  for (;;) {
    if (tk == tk_end) {
      if (lane == 0) tk = atomicAdd(P.ticket, (unsigned)XF_CHUNK);
      tk = __shfl_sync(FULL, tk, 0);
      if (tk >= P.ntiles) break;
      tk_end = min(tk + (unsigned)XF_CHUNK, P.ntiles);
    }
    const long long d = (long long)tk * NG + g;
    tk++;
    const bool have = live && d < P.ndof;
    long long a = 0, rowbase = 0;
    int na = 0, nr = 0;
    if (have) {
      a = __ldg(P.adj_ptr + d);
      na = (int)(__ldg(P.adj_ptr + d + 1) - a);
      rowbase = __ldg(P.rowptr + d);
      nr = (int)(__ldg(P.rowptr + d + 1) - rowbase);
      const bool srow = na >= 1 && na <= P.skip_na;
      if (srow ? (nr <= 32) : (__ldg(P.cand + d) <= P.skip_cand)) { na = 0; nr = 0; }  // a row of the other kernels
    }
    if (!__any_sync(FULL, nr > 0)) continue;
    int H = 32, shift = 27;
    while (H < nr + (nr >> 1) + 1) { H <<= 1; shift--; }  // load factor <= 2/3 with the exact row length of the count pass
    if (!have || nr == 0) H = 0;
    const unsigned hmask = (unsigned)(H - 1);
    for (int k = ls << 2; k < H; k += GS << 2) sts_i4(keyA + (k << 2), XG_EMPTY);
    for (int k = ls << 1; k < H; k += GS << 1) sts_zero16(accA + (k << 3));
    __syncwarp();

    // ---- walk the adjacency: one table row per step and group, operands of step t+1 in flight during step t
    const long long* __restrict__ adjd = P.adj + a;
    const int maxna = __reduce_max_sync(FULL, na);
    long long e_nx = (0 < na) ? __ldg(adjd) : -1;      // entry of the step whose loads are issued next
    long long e_n2 = (1 < na) ? __ldg(adjd + 1) : -1;  // the one after
    int c_ld[TRIPS], ti_ld = -1;
    double m_ld[TRIPS], cf_ld[NC > 0 ? NC : 1];
#define XF_ISSUE(e)                                                                                    \
  {                                                                                                    \
    ti_ld = -1;                                                                                        \
    _Pragma("unroll") for (int tr = 0; tr < TRIPS; tr++) { c_ld[tr] = -1; m_ld[tr] = 0.0; }            \
    if ((e) >= 0) {                                                                                    \
      XF_DECODE(e, bi, w, ro, loc)                                                                     \
      const int32_t* __restrict__ trow = P.bd.table[bi] + ro;                                          \
      _Pragma("unroll") for (int tr = 0; tr < TRIPS; tr++) {                                           \
        const int lj = tr * GS + ls;                                                                   \
        if (lj < w) c_ld[tr] = __ldg(trow + lj);                                                       \
      }                                                                                                \
      const double* __restrict__ cfp = (NC > 0) ? P.bd.coefs[bi] : nullptr;                            \
      if (NC > 0 && cfp) {                                                                             \
        const double2* __restrict__ c2 = reinterpret_cast<const double2*>(cfp + (size_t)XG_ADJ_ROW(e) * 8); \
        _Pragma("unroll") for (int m = 0; m < (NC + 1) / 2; m++) {                                     \
          const double2 v = __ldg(c2 + m);                                                             \
          cf_ld[2 * m] = v.x;                                                                          \
          if (2 * m + 1 < NC) cf_ld[2 * m + 1] = v.y;                                                  \
        }                                                                                              \
        ti_ld = loc * w;                                                                               \
      } else {                                                                                         \
        const double* __restrict__ mb = P.bd.mats[bi];                                                 \
        if (mb) {                                                                                      \
          const double* __restrict__ mrow = mb + (size_t)(ro + loc) * w;                               \
          _Pragma("unroll") for (int tr = 0; tr < TRIPS; tr++) {                                       \
            const int lj = tr * GS + ls;                                                               \
            if (lj < w) m_ld[tr] = __ldg(mrow + lj);                                                   \
          }                                                                                            \
        } /* else: the batch only contributes sparsity (explicit zeros) */                             \
      }                                                                                                \
    }                                                                                                  \
  }
    XF_ISSUE(e_nx);
    for (int t = 0; t < maxna; t++) {
      int c_cur[TRIPS];
      double m_cur[TRIPS];
      const int ti = ti_ld;
#pragma unroll
      for (int tr = 0; tr < TRIPS; tr++) { c_cur[tr] = c_ld[tr]; m_cur[tr] = m_ld[tr]; }
      if (NC > 0 && ti >= 0) {
#pragma unroll
        for (int tr = 0; tr < TRIPS; tr++) {
          if (c_cur[tr] >= 0) {  // lj < w
            const unsigned ta = trefA + ((ti + tr * GS) << 3);
            double r = 0.0;
#pragma unroll
            for (int m = 0; m < NC; m++) r += cf_ld[m] * lds_f64_ro(ta + ((m * nn) << 3));
            m_cur[tr] = r;
          }
        }
      }
      e_nx = e_n2;
      e_n2 = (t + 2 < na) ? __ldg(adjd + t + 2) : -1;
      XF_ISSUE(e_nx);
#pragma unroll
      for (int tr = 0; tr < TRIPS; tr++) {
        const int c = c_cur[tr];
        if (c >= 0) {
          int fresh = 0;
          const unsigned h = probe(keyA, hmask, shift, c, fresh);
          const unsigned aa = accA + (h << 3);
          sts_f64(aa, lds_f64(aa) + m_cur[tr]);
        }
      }
      __syncwarp();
    }
#undef XF_ISSUE

    // ---- in-place compaction of the occupied slots, GS slots of every group per round (the slots of a round are
    // read into registers before any lane writes, and a round writes below the end of the slots it has read)
    int nu = 0;
    const int Hw = __reduce_max_sync(FULL, H);
    for (int base = 0; base < Hw; base += GS) {
      const int idx = base + ls;
      int key = XG_EMPTY;
      double v = 0.0;
      if (idx < H) {
        key = lds_i32(keyA + (idx << 2));
        v = lds_f64(accA + (idx << 3));
      }
      __syncwarp();
      const unsigned bal = __ballot_sync(FULL, key != XG_EMPTY);
      const unsigned gb = live ? ((bal >> (g * GS)) & gmask) : 0u;
      if (key != XG_EMPTY) {
        const int p = nu + __popc(gb & ((1u << ls) - 1u));
        sts_i32(keyA + (p << 2), key);
        sts_f64(accA + (p << 3), v);
      }
      nu += __popc(gb);
      __syncwarp();
    }
    if (have) for (int k = ls; k < 4; k += GS) sts_i32(keyA + ((nu + k) << 2), XG_EMPTY);  // pad for the int4 reads (HG + 4 slots exist)
    __syncwarp();

    // ---- rank sort of the distinct columns, values follow their column to the final CSR position (nu == nr)
    const bool wr = have && nu > 0;  // skipped rows: na = 0 -> nu = 0
    const int numax = __reduce_max_sync(FULL, wr ? nu : 0);
    if (wr) {
      int32_t* __restrict__ col = P.colidx + rowbase;
      double* __restrict__ val = P.vals + rowbase;
      if (numax <= GS) rank_write<1>(keyA, accA, nu, ls, GS, 0, col, val);
      else if (numax <= 2 * GS) rank_write<2>(keyA, accA, nu, ls, GS, 0, col, val);
      else if (numax <= 3 * GS) rank_write<3>(keyA, accA, nu, ls, GS, 0, col, val);
      else if (numax <= 4 * GS) rank_write<4>(keyA, accA, nu, ls, GS, 0, col, val);
      else if (numax <= 6 * GS) rank_write<6>(keyA, accA, nu, ls, GS, 0, col, val);
      else
        for (int k0 = 0; k0 < nu; k0 += 8 * GS) rank_write<8>(keyA, accA, nu, ls, GS, k0, col, val);
    }
    __syncwarp();
  }
}

// ---- count pass: the same walk with keys only.  No ordering is needed (a set), so there is no synchronisation between
// the steps; every lane counts the columns it inserted and the group total is the row length.  H > cand (the number of
// candidates bounds the number of distinct columns), so a probe always ends.
template <int TRIPS>
__global__ void __launch_bounds__(XF_WARPS * 32) k_rows_count(const FusedPar P)
{
  extern __shared__ int smk[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int GS = P.GS, NG = P.NG, HG = P.HG;
  const int g = lane / GS, ls = lane - g * GS;
  const bool live = g < NG;
  const int gi = live ? g : 0;
  const unsigned keyA = smem_u32(smk) + (((wib * NG + gi) * HG) << 2);
  const unsigned FULL = 0xffffffffu;
  int longest = 0;
  // tiles (NG consecutive rows) are handed out in chunks of XF_CHUNK by an atomic ticket: one ticket per tile made the
  // single counter the bottleneck (1.6 M same-address atomics ~ 1.4 ms, profiles/r02_notes.md)
  unsigned tk = 0, tk_end = 0;
  for (;;) {
    if (tk == tk_end) {
      if (lane == 0) tk = atomicAdd(P.ticket, (unsigned)XF_CHUNK);
      tk = __shfl_sync(FULL, tk, 0);
      if (tk >= P.ntiles) break;
      tk_end = min(tk + (unsigned)XF_CHUNK, P.ntiles);
    }
    const long long d = (long long)tk * NG + g;
    tk++;
    const bool have = live && d < P.ndof;
    long long a = 0;
    int na = 0, cand = 0;
    if (have) {
      a = __ldg(P.adj_ptr + d);
      na = (int)(__ldg(P.adj_ptr + d + 1) - a);
      cand = __ldg(P.cand + d);
    }
    const bool mine = have && !(na >= 1 && na <= P.skip_na) && cand > P.skip_cand;  // else: a row of the other kernels
    if (!mine) { na = 0; cand = 0; }
    if (!__any_sync(FULL, mine)) continue;
    int H = 32, shift = 27;
    while (H <= cand) { H <<= 1; shift--; }
    if (!mine) H = 0;
    const unsigned hmask = (unsigned)(H - 1);
    for (int k = ls << 2; k < H; k += GS << 2) sts_i4(keyA + (k << 2), XG_EMPTY);
    __syncwarp();
    const long long* __restrict__ adjd = P.adj + a;
    const int maxna = __reduce_max_sync(FULL, na);
    long long e_nx = (0 < na) ? __ldg(adjd) : -1;
    long long e_n2 = (1 < na) ? __ldg(adjd + 1) : -1;
    int c_ld[TRIPS];
#define XF_ISSUE_C(e)                                                                      \
  {                                                                                        \
    _Pragma("unroll") for (int tr = 0; tr < TRIPS; tr++) c_ld[tr] = -1;                    \
    if ((e) >= 0) {                                                                        \
      XF_DECODE(e, bi, w, ro, loc)                                                         \
      (void)loc;                                                                           \
      const int32_t* __restrict__ trow = P.bd.table[bi] + ro;                              \
      _Pragma("unroll") for (int tr = 0; tr < TRIPS; tr++) {                               \
        const int lj = tr * GS + ls;                                                       \
        if (lj < w) c_ld[tr] = __ldg(trow + lj);                                           \
      }                                                                                    \
    }                                                                                      \
  }
    XF_ISSUE_C(e_nx);
    int nins = 0;
    for (int t = 0; t < maxna; t++) {
      int c_cur[TRIPS];
#pragma unroll
      for (int tr = 0; tr < TRIPS; tr++) c_cur[tr] = c_ld[tr];
      e_nx = e_n2;
      e_n2 = (t + 2 < na) ? __ldg(adjd + t + 2) : -1;
      XF_ISSUE_C(e_nx);
#pragma unroll
      for (int tr = 0; tr < TRIPS; tr++)
        if (c_cur[tr] >= 0) probe(keyA, hmask, shift, c_cur[tr], nins);
    }
#undef XF_ISSUE_C
    // group totals: inclusive scan over the warp, difference at the group boundaries
    int incl = nins;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int y = __shfl_up_sync(FULL, incl, s);
      if (lane >= s) incl += y;
    }
    const int hi = __shfl_sync(FULL, incl, min(31, gi * GS + GS - 1));
    const int lo = __shfl_sync(FULL, incl - nins, gi * GS);
    if (mine && ls == 0) { P.rowcnt[d] = hi - lo; longest = max(longest, hi - lo); }
    __syncwarp();  // the next sweep re-initialises the tables
  }
  longest = __reduce_max_sync(FULL, longest);
  if (lane == 0 && longest > 0) atomicMax(P.stats, longest);
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

static int fused_shape(xgpu_ctx* ctx, xgpu_plan* P, FusedPar& F, int* trips)
{
  if (P->dups) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "one-pass CSR: a table row holds a dof twice (facet patches): use xgpu_csr_plan + xgpu_csr_assemble");
  if (P->wmax > 64) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "one-pass CSR: table rows wider than 64");
  F.ndof = (int)P->ndof;
  // lanes per row: the widest table row, or half of it with two columns per lane (ctx option "csr_trips": 0 = policy)
  *trips = P->wmax > 32 ? 2 : (ctx->opt_csr_trips == 2 && P->wmax >= 4 ? 2 : 1);
  F.GS = *trips == 2 ? std::min(32, (P->wmax + 1) / 2) : std::max(1, P->wmax);
  F.NG = 32 / F.GS;
  const int64_t ntiles = (P->ndof + F.NG - 1) / F.NG;
  if (ntiles >= (int64_t)0x7fffffff) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "one-pass CSR: too many rows");
  F.ntiles = (unsigned)ntiles;
  F.adj_ptr = P->adj_ptr; F.adj = P->adj; F.cand = P->tmp32; F.bd = P->bd;
  // thread-per-row kernels for the short rows (ctx option "csr_small", default on): all table rows of one width
  bool same_w = true;
  for (int i = 0; i < P->bd.nb; i++) same_w = same_w && P->bd.width[i] == P->wmax;
  F.skip_na = (ctx->opt_csr_small && same_w) ? xg_rows_small_na(P->wmax, P->ndof) : 0;
  F.skip_cand = (ctx->opt_csr_warp && same_w && xg_rows_warp_ok(P->wmax, P->ndof)) ? 255 : -1;
  return XGPU_OK;
}

// xgpu_csr_count_rows: count pass + scan on a transposed table.  The plan then holds rowptr / nnz / the longest row; its
// pattern scratch is not filled (xgpu_csr_pattern / xgpu_csr_assemble need a plan of xgpu_csr_plan).
static int count_rows(xgpu_ctx* ctx, xgpu_plan* P, int64_t* nnz_host);

extern "C" int xgpu_csr_count_rows(xgpu_ctx* ctx, xgpu_plan* P, int64_t* nnz_host)
{
  XGPU_NVTX(ctx, "xgpu_csr_count_rows");
  if (!ctx || !P) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_count_rows: bad argument");
  if (P->counted) { if (nnz_host) *nnz_host = P->nnz; return XGPU_OK; }
  return count_rows(ctx, P, nnz_host);
}

extern "C" int xgpu_csr_plan_rows(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                                  xgpu_plan** plan_out, int64_t* nnz_host)
{
  XGPU_NVTX(ctx, "xgpu_csr_plan_rows");
  if (!plan_out) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_plan_rows: bad argument");
  int rc = xgpu_adjacency_plan(ctx, ndof, batches_host, nbatches, plan_out);
  if (rc) return rc;
  rc = count_rows(ctx, *plan_out, nnz_host);
  if (rc) { xgpu_plan_destroy(*plan_out); *plan_out = nullptr; }
  return rc;
}

static int count_rows(xgpu_ctx* ctx, xgpu_plan* P, int64_t* nnz_host)
{
  const int64_t ndof = P->ndof;
  int rc;
  cudaStream_t st = ctx->stream;
  FusedPar F;
  int trips;
  rc = fused_shape(ctx, P, F, &trips);
  int HG = 32;
  while (HG <= P->maxcand) HG <<= 1;
  const size_t smem = sizeof(int) * (size_t)XF_WARPS * F.NG * HG;
  if (!rc && P->maxcand > F.skip_cand && smem > 200 * 1024) rc = xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "one-pass CSR: a row has too many candidate columns");
  if (rc) return rc;
  if (ndof == 0) { P->nnz = 0; P->counted = true; if (nnz_host) *nnz_host = 0; return XGPU_OK; }
  F.HG = HG;
  void* sp;
  rc = xgpu_scratch(ctx, 9, 64, &sp);
  if (rc) return rc;
#define XF_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return xgpu_fail(ctx, XGPU_ERR_CUDA, cudaGetErrorString(e_)); } while (0)
  XF_TRY(cudaMemsetAsync(sp, 0, 64, st));
  F.ticket = (unsigned int*)sp;
  F.stats = (int*)sp + 2;
  XF_TRY(cudaMallocAsync(&P->rowcnt, sizeof(int32_t) * (size_t)(ndof + 2), st));
  F.rowcnt = P->rowcnt;
  unsigned int* slotcount = (unsigned int*)sp + 4;
  if (F.skip_na > 0) XF_TRY(cudaMallocAsync(&P->pat_small, xg_rows_small_record_bytes(ndof), st));
  if (F.skip_cand >= 0) {
    P->nbig_cap = F.skip_na > 0 ? P->nbig : ndof;  // without the thread-per-row kernels every row may come here
    XF_TRY(cudaMallocAsync(&P->bigslot, sizeof(int32_t) * (size_t)ndof, st));
    XF_TRY(cudaMallocAsync(&P->pat_big, std::max<size_t>(16, xg_rows_warp_record_bytes(P->nbig_cap)), st));
  }
  if (P->maxcand > F.skip_cand) {  // rows left for the lane-group hash kernel
    int per_sm = (int)std::min<size_t>(12, (220 * 1024) / (smem + 1024));
    unsigned grid = nblk(F.ntiles, XF_WARPS);
    if (grid > (unsigned)(XGPU_NSM * per_sm)) grid = XGPU_NSM * per_sm;
    if (trips == 1) {
      XF_TRY(cudaFuncSetAttribute(k_rows_count<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_rows_count<1><<<grid, XF_WARPS * 32, smem, st>>>(F);
    } else {
      XF_TRY(cudaFuncSetAttribute(k_rows_count<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_rows_count<2><<<grid, XF_WARPS * 32, smem, st>>>(F);
    }
    ctx->launches++;
    XF_TRY(cudaGetLastError());
  }
  P->warp_cand = F.skip_cand;
  // the two kernels take disjoint rows and are bound by different units (warp per row: latency / shuffles, thread per row:
  // issue): with both present the warp-per-row one runs on the context's internal second stream
  const bool fork = ctx->opt_csr_fork && F.skip_cand >= 0 && F.skip_na > 0;
  if (F.skip_cand >= 0) {
    if (fork) XF_TRY(xgpu_aux_begin(ctx));
    rc = xg_rows_warp_run(ctx, P, P->bd, nullptr, 0, true, F.skip_na, (unsigned int*)sp + 1, slotcount, P->rowcnt, F.stats, nullptr, nullptr);
    if (fork) { cudaError_t e_ = xgpu_aux_end(ctx, st); if (!rc && e_ != cudaSuccess) rc = xgpu_fail(ctx, XGPU_ERR_CUDA, cudaGetErrorString(e_)); }
    if (rc) return rc;
  }
  P->small_na = F.skip_na;
  if (F.skip_na > 0) {
    rc = xg_rows_small_run(ctx, P, P->bd, nullptr, 0, true, P->rowcnt, F.stats, nullptr, nullptr);
    if (rc) return rc;
  }
  rc = xgpu_exclusive_scan_i32(ctx, P->rowcnt, ndof, (int64_t*)P->rowptr);
  if (rc) return rc;
  XF_TRY(cudaEventRecord(P->ev[2], st));
  XF_TRY(xgpu_d2h_small(ctx->h_pinned, P->rowptr + ndof, sizeof(long long), st));
  XF_TRY(xgpu_d2h_small(ctx->h_pinned + 1, F.stats, 2 * sizeof(int), st));
  XF_TRY(cudaStreamSynchronize(st));
#undef XF_TRY
  P->nnz = ctx->h_pinned[0];
  P->maxnr = ((int*)(ctx->h_pinned + 1))[0];
  P->small_overflow = ((int*)(ctx->h_pinned + 1))[1];
  P->counted = true;
  if (nnz_host) *nnz_host = P->nnz;
  return XGPU_OK;
}

static int assemble_fused_rows(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                               const double* tref, int ncoef, int64_t* rowptr, int32_t* colidx, double* vals,
                               int64_t row_lo, int64_t row_hi, bool copy_rowptr);

extern "C" int xgpu_csr_assemble_fused(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                                       const double* tref, int ncoef, int64_t* rowptr, int32_t* colidx, double* vals)
{
  XGPU_NVTX(ctx, "xgpu_csr_assemble_fused");
  return assemble_fused_rows(ctx, P, batches_host, nbatches, tref, ncoef, rowptr, colidx, vals, 0, (int64_t)1 << 40, true);
}

extern "C" int xgpu_csr_assemble_fused_rows(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                                            const double* tref, int ncoef, int64_t* rowptr, int32_t* colidx, double* vals,
                                            int64_t row_lo, int64_t row_hi, int copy_rowptr)
{
  XGPU_NVTX(ctx, "xgpu_csr_assemble_fused_rows");
  if (row_lo < 0 || row_hi < row_lo) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_assemble_fused_rows: bad row range");
  return assemble_fused_rows(ctx, P, batches_host, nbatches, tref, ncoef, rowptr, colidx, vals, row_lo, row_hi, copy_rowptr != 0);
}

static int assemble_fused_rows(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                               const double* tref, int ncoef, int64_t* rowptr, int32_t* colidx, double* vals,
                               int64_t row_lo, int64_t row_hi, bool copy_rowptr)
{
  if (!ctx || !P || !batches_host || nbatches != P->bd.nb || !rowptr)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_assemble_fused: bad argument");
  if (!P->counted) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_assemble_fused: the plan has no row counts (build it with xgpu_csr_plan_rows)");
  FusedPar F;
  int trips;
  int rc = fused_shape(ctx, P, F, &trips);
  if (rc) return rc;
  int nproc = 0, wproc = 0;
  for (int i = 0; i < nbatches; i++) {
    if (batches_host[i].nrows != F.bd.rowoff[i + 1] - F.bd.rowoff[i] || batches_host[i].width != F.bd.width[i])
      return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_assemble_fused: batches do not match the plan");
    F.bd.table[i] = batches_host[i].table;
    F.bd.mats[i] = batches_host[i].mats;
    F.bd.coefs[i] = batches_host[i].coefs;
    if (F.bd.coefs[i]) {
      if (F.bd.mats[i]) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_assemble_fused: a batch has both matrices and coefficients");
      if (nproc && wproc != F.bd.width[i]) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_csr_assemble_fused: procedural batches of different widths");
      nproc++; wproc = F.bd.width[i];
    }
  }
  if (nproc && (!tref || (ncoef != 2 && ncoef != 4 && ncoef != 7)))
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_assemble_fused: procedural batches need the reference tensors (ncoef 2, 4 or 7)");
  cudaStream_t st = ctx->stream;
  if (copy_rowptr) XGPU_CUDA(ctx, cudaMemcpyAsync(rowptr, P->rowptr, sizeof(long long) * (P->ndof + 1), cudaMemcpyDeviceToDevice, st));
  if (P->ndof == 0 || P->nnz == 0) return XGPU_OK;
  const bool partial = row_lo > 0 || row_hi < P->ndof;
  const bool fast_ok = xg_rows_small_supports(P->wmax, nproc ? ncoef : 0) != 0;
  F.skip_na = (P->small_na > 0 && fast_ok) ? P->small_na : 0;
  F.skip_cand = (P->warp_cand >= 0 && fast_ok) ? P->warp_cand : -1;
  void* sp;
  rc = xgpu_scratch(ctx, 9, 64, &sp);
  if (rc) return rc;
  XGPU_CUDA(ctx, cudaMemsetAsync(sp, 0, 64, st));
  const bool fork = ctx->opt_csr_fork && F.skip_cand >= 0 && F.skip_na > 0;
  if (F.skip_cand >= 0) {
    if (fork) XGPU_CUDA(ctx, xgpu_aux_begin(ctx));
    rc = xg_rows_warp_run(ctx, P, F.bd, tref, nproc ? ncoef : 0, false, F.skip_na, (unsigned int*)sp + 1, nullptr, nullptr, nullptr, colidx, vals,
                          row_lo, row_hi);
    if (fork) { cudaError_t e_ = xgpu_aux_end(ctx, st); if (!rc && e_ != cudaSuccess) rc = xgpu_fail(ctx, XGPU_ERR_CUDA, cudaGetErrorString(e_)); }
    if (rc) return rc;
  }
  if (F.skip_na > 0) {
    rc = xg_rows_small_run(ctx, P, F.bd, tref, nproc ? ncoef : 0, false, nullptr, nullptr, colidx, vals, row_lo, row_hi);
    if (rc) return rc;
  }
  if (P->maxcand <= F.skip_cand && !(F.skip_na > 0 && P->small_overflow)) return XGPU_OK;  // no row left for the lane-group hash kernel
  if (partial) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_csr_assemble_fused_rows: rows of the lane-group kernel cannot be split by range (assemble all rows at once)");
  int HG = 32;
  while (HG < P->maxnr + (P->maxnr >> 1) + 1) HG <<= 1;
  F.HG = HG;
  F.nn = nproc ? wproc * wproc : 0;
  F.tref = tref;
  const int NC = nproc ? ncoef : 0;
  const size_t ntref = ((size_t)NC * F.nn + 1) & ~(size_t)1;
  const size_t smem = sizeof(double) * (ntref + (size_t)XF_WARPS * F.NG * HG) + sizeof(int) * (size_t)XF_WARPS * F.NG * (HG + 4);
  if (smem > 200 * 1024) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_csr_assemble_fused: a row is too long for the one-pass kernel");
  F.ticket = (unsigned int*)sp;
  F.stats = nullptr;
  F.rowcnt = nullptr;
  F.rowptr = P->rowptr; F.colidx = colidx; F.vals = vals;
  int per_sm = (int)std::min<size_t>(12, (220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  unsigned grid = nblk(F.ntiles, XF_WARPS);
  if (grid > (unsigned)(XGPU_NSM * per_sm)) grid = XGPU_NSM * per_sm;
#define XF_LAUNCH(NCV, TR)                                                                                              \
  {                                                                                                                     \
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_rows_fused<NCV, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_rows_fused<NCV, TR><<<grid, XF_WARPS * 32, smem, st>>>(F);                                                        \
  }
  if (trips == 1) {
    if (NC == 0) XF_LAUNCH(0, 1) else if (NC == 2) XF_LAUNCH(2, 1) else if (NC == 4) XF_LAUNCH(4, 1) else XF_LAUNCH(7, 1)
  } else {
    if (NC == 0) XF_LAUNCH(0, 2) else if (NC == 2) XF_LAUNCH(2, 2) else if (NC == 4) XF_LAUNCH(4, 2) else XF_LAUNCH(7, 2)
  }
#undef XF_LAUNCH
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

