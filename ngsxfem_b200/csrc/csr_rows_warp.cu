// csr_rows_warp.cu — matrix rows with many adjacent elements (e.g. the vertex dofs of a P2 tet mesh: ~24 tets, 240
// candidate columns, 65 distinct), ONE WARP PER ROW, no hash table:
//   * the lanes pack the row's candidates (column << 8 | candidate index), 8 per lane, and the warp sorts the 256 keys
//     with an all-ascending bitonic network: the three innermost strides of every merge are compare-exchanges between
//     registers of one lane, the others one shuffle + one min/max per key (~450 instructions per row);
//   * duplicates are runs of the sorted sequence: run heads = the row's columns (ascending), number of heads = row
//     length (count pass); inside a run the candidates are in (batch, table row, local index) order, the summation
//     order of a sequential AddElementMatrix loop;
//   * the pattern pass stores, per row, the candidate indices in SORTED order (256 bytes), the start of every run (one byte
//     per column) and the columns: the row's record;
//   * values pass: every lane fetches / generates the values of its 8 candidates in candidate order (procedural affine
//     batches: the element's coefficients x the reference tensors in shared memory) into shared memory; then ONE LANE PER
//     COLUMN sums its run front to back from a register (the loads do not depend on the sum: no load-add-store chain
//     through shared memory) — the result has the bits of the sequential order (= the two-pass path).
// The lane-group hash kernel of csr_fused.cu needs ~2200 instructions per such row (24 sequential steps with 10 of 32
// lanes, rank sort); it remains the fallback for rows with more than 256 candidates.
#include "csr.cuh"
#include <algorithm>

namespace {

#define XW_WARPS 4
#define XW_KPL 8  // keys per lane: 256 candidates per row

#define XW_RECB 1536  // bytes of a row record: 256 sorted candidate indices + 256 run starts + 256 columns

struct WarpPar {
  int ndof, small_na;
  int half_keys;            // pattern pass: rows with at most 128 candidates on the 4-keys-per-lane network (option "csr_warp_half")
  int row_lo, row_hi, tick0;   // values pass: rows [row_lo, row_hi) only; the ticket counter starts at row tick0 (multiple of 32)
  const long long* adj_ptr;
  const long long* adj;
  BatchDev bd;
  const double* tref;       // procedural batches: ncoef x W x W
  const long long* rowptr;  // values pass
  int32_t* rowcnt;          // pattern pass
  int* stats;               // pattern pass: [0] longest row
  unsigned int* ticket;
  unsigned int* slotcount;  // pattern pass: records handed out so far
  unsigned int slotcap;
  int32_t* bigslot;         // record number of a row (pattern pass: out, values pass: in)
  unsigned char* records;
  int accstride;            // values pass: doubles per row accumulator (>= longest row of this kernel)
  int namax;                // values pass: adjacent table rows staged per row (>= maxcand / W, even)
  int32_t* colidx;
  double* vals;
};

__device__ __forceinline__ void ce(unsigned& a, unsigned& b)
{
  const unsigned lo = min(a, b), hi = max(a, b);
  a = lo; b = hi;
}

// bitonic sorting network over the warp: 32 * KPL keys, position g = lane * KPL + k, ascending.  Mirror stage of a merge of
// `size` keys: g <-> g ^ (size - 1); half cleaners: g <-> g ^ stride.  Partners inside a lane are register pairs, partners in
// another lane come by shuffle (mirror: lane ^ (size / KPL - 1), register KPL - 1 - k; cleaner: lane ^ (stride / KPL)).
template <int KPL>
__device__ __forceinline__ void sort_warp(unsigned (&K)[KPL], int lane)
{
  const unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int size = 2; size <= 32 * KPL; size <<= 1) {
    if (size <= KPL) {
#pragma unroll
      for (int k = 0; k < KPL; k++) {
        const int o = k ^ (size - 1);
        if (o > k) ce(K[k], K[o]);
      }
    } else {
      const int lm = size / KPL - 1;
      const bool lower = (lane & (size / (2 * KPL))) == 0;
      unsigned o[KPL];
#pragma unroll
      for (int k = 0; k < KPL; k++) o[k] = __shfl_xor_sync(FULL, K[KPL - 1 - k], lm);
#pragma unroll
      for (int k = 0; k < KPL; k++) K[k] = lower ? min(K[k], o[k]) : max(K[k], o[k]);
    }
#pragma unroll
    for (int stride = size >> 2; stride >= 1; stride >>= 1) {
      if (stride >= KPL) {  // half cleaner across lanes
        const int lm = stride / KPL;
        const bool lower = (lane & lm) == 0;
#pragma unroll
        for (int k = 0; k < KPL; k++) {
          const unsigned o = __shfl_xor_sync(FULL, K[k], lm);
          K[k] = lower ? min(K[k], o) : max(K[k], o);
        }
      } else {
#pragma unroll
        for (int k = 0; k < KPL; k++) {
          const int o = k ^ stride;
          if (o > k) ce(K[k], K[o]);
        }
      }
    }
  }
}

// one row of the pattern pass with 32 * KPL key slots (KPL = 8: up to 255 candidates; KPL = 4: up to 128 — the rows of the
// axis-parallel edges of a Kuhn mesh have 80 — at less than half the shuffles and compare-exchanges).  The record is the same
// for both: byte g of the index area = candidate of sorted position g, run starts are sorted positions.  returns the row length
template <int W, int KPL>
__device__ __forceinline__ int warp_row_pattern(const WarpPar& P, long long d, long long a, int cand, int lane,
                                                unsigned char* rs_b, const unsigned int* rs_u, int* __restrict__ cols_s)
{
  const unsigned FULL = 0xffffffffu;
  // candidates: index k * 32 + lane (neighbouring lanes read neighbouring table entries)
  unsigned K[KPL];
#pragma unroll
  for (int k = 0; k < KPL; k++) {
    const int idx = k * 32 + lane;
    K[k] = 0xffffffffu;
    if (idx < cand) {
      const int t = idx / W, j = idx - t * W;
      const long long e = __ldg(P.adj + a + t);
      const int c = __ldg(P.bd.table[XG_ADJ_BATCH(e)] + (size_t)XG_ADJ_ROW(e) * W + j);
      if (c >= 0) K[k] = ((unsigned)c << 8) | (unsigned)idx;
    }
  }
  sort_warp<KPL>(K, lane);
  // runs: heads (first key of an equal-column run)
  const unsigned prevlast = __shfl_up_sync(FULL, K[KPL - 1], 1);
  unsigned fresh = 0u;
  int nval = 0;
#pragma unroll
  for (int k = 0; k < KPL; k++) {
    if (K[k] != 0xffffffffu) {
      nval++;
      const unsigned pk = (k == 0) ? ((lane == 0) ? 0xffffffffu : prevlast) : K[k - 1];
      if ((pk >> 8) != (K[k] >> 8) || (k == 0 && lane == 0)) fresh |= 1u << k;
    }
  }
  const int nf = __popc(fresh);
  int incl = nf;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const int y = __shfl_up_sync(FULL, incl, s);
    if (lane >= s) incl += y;
  }
  const int nr = __shfl_sync(FULL, incl, 31);
  const int nvalid = __reduce_add_sync(FULL, nval);   // the valid keys are the first nvalid of the sorted sequence
  unsigned slot = 0;
  if (lane == 0) {
    slot = atomicAdd(P.slotcount, 1u);
    P.rowcnt[d] = nr;
    P.bigslot[d] = (int)slot;
  }
  slot = __shfl_sync(FULL, slot, 0);
  __syncwarp();
  {
    int q = incl - nf;  // index of the lane's first head
#pragma unroll
    for (int k = 0; k < KPL; k++) {
      if ((fresh >> k) & 1u) { cols_s[q] = (int)(K[k] >> 8); rs_b[q] = (unsigned char)(lane * KPL + k); q++; }
    }
    if (lane == 31) rs_b[nr] = (unsigned char)nvalid;   // nr <= nvalid <= 255
  }
  __syncwarp();
  if (slot < P.slotcap) {
    unsigned char* __restrict__ rec = P.records + (size_t)slot * XW_RECB;
    // sorted candidate indices: KPL bytes per lane, straight from the registers (KPL = 4: the upper half of the index area is
    // not written and never read: the values pass stops at the row's number of valid keys)
    if constexpr (KPL == 8) {
      reinterpret_cast<uint2*>(rec)[lane] =
          make_uint2((K[0] & 255u) | ((K[1] & 255u) << 8) | ((K[2] & 255u) << 16) | ((K[3] & 255u) << 24),
                     (K[4] & 255u) | ((K[5] & 255u) << 8) | ((K[6] & 255u) << 16) | ((K[7] & 255u) << 24));
    } else {
      static_assert(KPL == 4, "record layout: 4 or 8 keys per lane");
      reinterpret_cast<unsigned*>(rec)[lane] = (K[0] & 255u) | ((K[1] & 255u) << 8) | ((K[2] & 255u) << 16) | ((K[3] & 255u) << 24);
    }
    reinterpret_cast<uint2*>(rec + 256)[lane] = make_uint2(rs_u[2 * lane], rs_u[2 * lane + 1]);
    int* __restrict__ rc = reinterpret_cast<int*>(rec + 512);
    for (int i = lane; i < nr; i += 32) rc[i] = cols_s[i];
  }
  __syncwarp();
  return nr;
}

// ---- pattern pass: sort, runs -> row length, columns and the position of every candidate, stored as the row's record
template <int W>
__global__ void __launch_bounds__(XW_WARPS * 32) k_rows_warp_pattern(const WarpPar P)
{
  __shared__ unsigned int rs_sm[XW_WARPS][64];    // 256 run-start bytes per warp
  __shared__ int cols_sm[XW_WARPS][32 * XW_KPL];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  unsigned char* rs_b = reinterpret_cast<unsigned char*>(rs_sm[wib]);  // aliases rs_sm: no __restrict__
  int* __restrict__ cols_s = cols_sm[wib];
  int longest = 0;
  for (;;) {
    // 32 consecutive rows per ticket; every lane looks at one of them, the warp then works through the eligible ones
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(P.ticket, 32u);
    base = __shfl_sync(FULL, base, 0);
    if (base >= (unsigned)P.ndof) break;
    const long long dl = (long long)base + lane;
    long long al = 0;
    int nal = 0;
    if (dl < P.ndof) {
      al = __ldg(P.adj_ptr + dl);
      nal = (int)(__ldg(P.adj_ptr + dl + 1) - al);
    }
    const bool small = nal >= 1 && nal <= P.small_na;  // rows of the thread-per-row kernel
    const bool mine = dl < P.ndof && !small && nal * W < 32 * XW_KPL;  // < 256 candidates: run starts and indices fit a byte
    if (mine && nal == 0) P.rowcnt[dl] = 0;
    unsigned todo = __ballot_sync(FULL, mine && nal > 0);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const long long d = (long long)base + src;
      const long long a = __shfl_sync(FULL, al, src);
      const int cand = __shfl_sync(FULL, nal, src) * W;
      const int nr = (cand <= 128 && P.half_keys) ? warp_row_pattern<W, 4>(P, d, a, cand, lane, rs_b, rs_sm[wib], cols_s)
                                                  : warp_row_pattern<W, XW_KPL>(P, d, a, cand, lane, rs_b, rs_sm[wib], cols_s);
      longest = max(longest, nr);
    }
  }
  longest = __reduce_max_sync(FULL, longest);
  if (lane == 0 && longest > 0) atomicMax(P.stats, longest);
}

// ---- values pass: one row per warp at a time, small footprint (high occupancy hides the dependent global loads).
// Phase 1: the values of the row's candidates, fetched / generated in candidate order (neighbouring lanes share the
// adjacency entry and the element's coefficients: few distinct cache lines per load) into shared memory; the sorted
// candidate indices and the run starts of the row's record go there too.  Phase 2: lane q sums run q (the contributions to
// column q, in candidate order = the summation order of a sequential AddElementMatrix loop) in a register and stores
// column + value.
template <int W, int NC>
__global__ void __launch_bounds__(XW_WARPS * 32) k_rows_warp_values(const WarpPar P)
{
  constexpr int NCAND = 32 * XW_KPL;
  constexpr int NTREF = (NC > 0) ? NC * W * W : 0;
  extern __shared__ double smw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  if (NC > 0) {
    for (int k = threadIdx.x; k < NC * W * W; k += blockDim.x) smw[k] = __ldg(P.tref + k);
    __syncthreads();
  }
  double* vals_w = smw + NTREF + (size_t)wib * (NCAND + NCAND / 4);
  unsigned char* sidx_w = reinterpret_cast<unsigned char*>(vals_w + NCAND);   // 256 sorted candidate indices
  unsigned char* rs_w = sidx_w + NCAND;                                       // 256 run starts
  for (;;) {
    unsigned base = 0;
    if (lane == 0) base = (unsigned)P.tick0 + atomicAdd(P.ticket, 32u);
    base = __shfl_sync(FULL, base, 0);
    if (base >= (unsigned)P.row_hi) break;
    const long long dl = (long long)base + lane;
    long long al = 0, rbl = 0;
    int nal = 0, nrl = 0, sll = -1;
    if (dl < P.ndof) {
      al = __ldg(P.adj_ptr + dl);
      nal = (int)(__ldg(P.adj_ptr + dl + 1) - al);
      rbl = __ldg(P.rowptr + dl);
      nrl = (int)(__ldg(P.rowptr + dl + 1) - rbl);
    }
    const bool small = nal >= 1 && nal <= P.small_na;
    const bool mine = dl >= P.row_lo && dl < P.row_hi && !small && nal * W < NCAND && nal > 0;
    if (mine) sll = __ldg(P.bigslot + dl);
    unsigned todo = __ballot_sync(FULL, mine);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const long long a = __shfl_sync(FULL, al, src), rowbase = __shfl_sync(FULL, rbl, src);
      const int na = __shfl_sync(FULL, nal, src), nr = __shfl_sync(FULL, nrl, src);
      const unsigned char* __restrict__ rec = P.records + (size_t)__shfl_sync(FULL, sll, src) * XW_RECB;
      // ---- phase 1: lane t takes adjacent element t — ONE coalesced load of the row's adjacency entries, then every lane's
      // coefficient / matrix-row loads are independent of each other (two memory latencies per row, not two per candidate)
      reinterpret_cast<uint2*>(sidx_w)[lane] = __ldg(reinterpret_cast<const uint2*>(rec) + lane);
      reinterpret_cast<uint2*>(rs_w)[lane] = __ldg(reinterpret_cast<const uint2*>(rec + 256) + lane);
      for (int t = lane; t < na; t += 32) {
        const long long e = __ldg(P.adj + a + t);
        const int bi = XG_ADJ_BATCH(e);
        const size_t row = (size_t)XG_ADJ_ROW(e);
        const int loc = XG_ADJ_LOC(e);
        double* __restrict__ vout = vals_w + t * W;
        const double* __restrict__ cfp = (NC > 0) ? P.bd.coefs[bi] : nullptr;
        if (NC > 0 && cfp) {
          const double2* __restrict__ c2 = reinterpret_cast<const double2*>(cfp + row * 8);
          double cf[8];
#pragma unroll
          for (int m = 0; m < (NC + 1) / 2; m++) {
            const double2 x = __ldg(c2 + m);
            cf[2 * m] = x.x; cf[2 * m + 1] = x.y;
          }
          const double* __restrict__ tr = smw + loc * W;
#pragma unroll
          for (int j = 0; j < W; j++) {
            double val = 0.0;
#pragma unroll
            for (int m = 0; m < NC; m++) val += cf[m] * tr[m * W * W + j];
            vout[j] = val;
          }
        } else {
          const double* __restrict__ mb = P.bd.mats[bi];
          if (mb) {
            const double* __restrict__ mrow = mb + (row * W + loc) * W;
            if ((W & 1) == 0) {
              double2 x[W / 2];
#pragma unroll
              for (int j = 0; j < W / 2; j++) x[j] = __ldg(reinterpret_cast<const double2*>(mrow) + j);
#pragma unroll
              for (int j = 0; j < W / 2; j++) { vout[2 * j] = x[j].x; vout[2 * j + 1] = x[j].y; }
            } else {
#pragma unroll
              for (int j = 0; j < W; j++) vout[j] = __ldg(mrow + j);
            }
          } else {
#pragma unroll
            for (int j = 0; j < W; j++) vout[j] = 0.0;  // sparsity only (explicit zeros)
          }
        }
      }
      __syncwarp();
      // ---- phase 2: lane q sums run q
      {
        const int* __restrict__ rc = reinterpret_cast<const int*>(rec + 512);
        for (int q = lane; q < nr; q += 32) {
          const int c = __ldg(rc + q);
          const int g0 = rs_w[q], g1 = rs_w[q + 1];
          double sum = 0.0;
          for (int g = g0; g < g1; g++) sum += vals_w[sidx_w[g]];
          P.colidx[rowbase + q] = c;
          P.vals[rowbase + q] = sum;
        }
      }
      __syncwarp();
    }
  }
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <int W>
int launch_warp(xgpu_ctx* ctx, const WarpPar& S, int NC, bool pattern)
{
  cudaStream_t st = ctx->stream;
  unsigned grid = nblk(((int64_t)S.ndof + 31) / 32, XW_WARPS);
  const unsigned bpsm = (unsigned)std::max(1, std::min(16, ctx->opt_csr_warp_bpsm));
  if (grid > XGPU_NSM * bpsm) grid = XGPU_NSM * bpsm;  // persistent grid (row tickets); fewer blocks leave room for the forked thread-per-row kernel
  if (pattern) {
    k_rows_warp_pattern<W><<<grid, XW_WARPS * 32, 0, st>>>(S);
  } else {
    const size_t smem = sizeof(double) * ((size_t)NC * W * W + (size_t)XW_WARPS * (32 * XW_KPL + 32 * XW_KPL / 4));
#define XW_LAUNCH(NCV)                                                                                                  \
  {                                                                                                                     \
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_rows_warp_values<W, NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_rows_warp_values<W, NCV><<<grid, XW_WARPS * 32, smem, st>>>(S);                                                   \
  }
    bool done = false;
    if (NC == 0) { XW_LAUNCH(0) done = true; }
    if constexpr (W == 4 || W == 10) { if (NC == 7) { XW_LAUNCH(7) done = true; } }
    if constexpr (W == 6) { if (NC == 4) { XW_LAUNCH(4) done = true; } }
    if (!done) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "csr warp rows: coefficient count");
#undef XW_LAUNCH
  }
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

// the warp-per-row kernels take rows with fewer than 256 candidates when (column << 8 | candidate) fits 32 bits
int xg_rows_warp_ok(int width, int64_t ndof)
{
  if (ndof >= ((int64_t)1 << 24)) return 0;
  return width == 4 || width == 6 || width == 10 || width == 12;
}

size_t xg_rows_warp_record_bytes(int64_t nrows) { return (size_t)nrows * XW_RECB; }

int xg_rows_warp_run(xgpu_ctx* ctx, xgpu_plan* P, const BatchDev& bd, const double* tref, int ncoef, bool pattern, int small_na,
                     unsigned int* ticket, unsigned int* slotcount, int32_t* rowcnt, int* stats, int32_t* colidx, double* vals,
                     int64_t row_lo, int64_t row_hi)
{
  WarpPar S;
  S.ndof = (int)P->ndof; S.small_na = small_na;
  S.half_keys = ctx->opt_csr_warp_half;
  S.row_lo = (int)row_lo; S.row_hi = (int)std::min<int64_t>(row_hi, P->ndof); S.tick0 = (int)(row_lo & ~(int64_t)31);
  if (!pattern && S.row_hi <= S.row_lo) return XGPU_OK;
  S.adj_ptr = P->adj_ptr; S.adj = P->adj; S.bd = bd; S.tref = tref;
  S.rowptr = P->rowptr; S.rowcnt = rowcnt; S.stats = stats; S.ticket = ticket; S.colidx = colidx; S.vals = vals;
  S.slotcount = slotcount; S.slotcap = (unsigned)P->nbig_cap; S.bigslot = P->bigslot; S.records = P->pat_big;
  S.accstride = (std::max(P->maxnr, 8) + 7) & ~7;
  if (S.accstride > 256) S.accstride = 256;
  S.namax = (std::min(P->maxcand, 255) / std::max(1, P->wmax) + 2) & ~1;
  switch (P->wmax) {
    case 4: return launch_warp<4>(ctx, S, ncoef, pattern);
    case 6: return launch_warp<6>(ctx, S, ncoef, pattern);
    case 10: return launch_warp<10>(ctx, S, ncoef, pattern);
    case 12: return launch_warp<12>(ctx, S, ncoef, pattern);
  }
  return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "csr warp rows: width");
}
