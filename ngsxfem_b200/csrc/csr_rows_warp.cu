// csr_rows_warp.cu — matrix rows with many adjacent elements (e.g. the vertex dofs of a P2 tet mesh: ~24 tets, 240
// candidate columns, 65 distinct), ONE WARP PER ROW, no hash table:
//   * the lanes pack the row's candidates (column << 8 | candidate index), 8 per lane, and the warp sorts the 256 keys
//     with an all-ascending bitonic network: the three innermost strides of every merge are compare-exchanges between
//     registers of one lane, the others one shuffle + one min/max per key (~450 instructions per row);
//   * duplicates are runs of the sorted sequence: run heads = the row's columns (ascending), number of heads = row
//     length (count pass); inside a run the candidates are in (batch, table row, local index) order, the summation
//     order of a sequential AddElementMatrix loop;
//   * full pass: every lane fetches / generates the values of its 8 candidates (procedural affine batches: the element's
//     coefficients x the reference tensors in shared memory), walks them in order and stores a run's sum at its end.
//     Runs that continue over a lane boundary get the running sum of the previous lane as carry — resolved
//     by a short shuffle iteration, so the result has the bits of the sequential order (= the two-pass path).
// The lane-group hash kernel of csr_fused.cu needs ~2200 instructions per such row (24 sequential steps with 10 of 32
// lanes, rank sort); it remains the fallback for rows with more than 256 candidates.
#include "csr.cuh"

namespace {

#define XW_WARPS 4
#define XW_KPL 8  // keys per lane: 256 candidates per row

struct WarpPar {
  int ndof, small_na;
  const long long* adj_ptr;
  const long long* adj;
  BatchDev bd;
  const double* tref;       // procedural batches: ncoef x W x W
  const long long* rowptr;  // full pass
  int32_t* rowcnt;          // count pass
  int* stats;               // count pass: [0] longest row
  unsigned int* ticket;
  int32_t* colidx;
  double* vals;
};

__device__ __forceinline__ void ce(unsigned& a, unsigned& b)
{
  const unsigned lo = min(a, b), hi = max(a, b);
  a = lo; b = hi;
}

// 256 keys, position g = lane * 8 + k, ascending
__device__ __forceinline__ void sort256(unsigned (&K)[XW_KPL], int lane)
{
  const unsigned FULL = 0xffffffffu;
  // blocks of 2, 4, 8: inside a lane (mirror stage, then half cleaners)
  ce(K[0], K[1]); ce(K[2], K[3]); ce(K[4], K[5]); ce(K[6], K[7]);
  ce(K[0], K[3]); ce(K[1], K[2]); ce(K[4], K[7]); ce(K[5], K[6]);
  ce(K[0], K[1]); ce(K[2], K[3]); ce(K[4], K[5]); ce(K[6], K[7]);
  ce(K[0], K[7]); ce(K[1], K[6]); ce(K[2], K[5]); ce(K[3], K[4]);
  ce(K[0], K[2]); ce(K[1], K[3]); ce(K[4], K[6]); ce(K[5], K[7]);
  ce(K[0], K[1]); ce(K[2], K[3]); ce(K[4], K[5]); ce(K[6], K[7]);
#pragma unroll
  for (int size = 16; size <= 256; size <<= 1) {
    {  // mirror stage: g <-> g ^ (size - 1): lane ^ (size/8 - 1), register 7 - k
      const int lm = (size >> 3) - 1;
      const bool lower = (lane & (size >> 4)) == 0;
      unsigned o[XW_KPL];
#pragma unroll
      for (int k = 0; k < XW_KPL; k++) o[k] = __shfl_xor_sync(FULL, K[XW_KPL - 1 - k], lm);
#pragma unroll
      for (int k = 0; k < XW_KPL; k++) K[k] = lower ? min(K[k], o[k]) : max(K[k], o[k]);
    }
#pragma unroll
    for (int stride = size >> 2; stride >= 8; stride >>= 1) {  // half cleaners across lanes
      const int lm = stride >> 3;
      const bool lower = (lane & lm) == 0;
#pragma unroll
      for (int k = 0; k < XW_KPL; k++) {
        const unsigned o = __shfl_xor_sync(FULL, K[k], lm);
        K[k] = lower ? min(K[k], o) : max(K[k], o);
      }
    }
    ce(K[0], K[4]); ce(K[1], K[5]); ce(K[2], K[6]); ce(K[3], K[7]);
    ce(K[0], K[2]); ce(K[1], K[3]); ce(K[4], K[6]); ce(K[5], K[7]);
    ce(K[0], K[1]); ce(K[2], K[3]); ce(K[4], K[5]); ce(K[6], K[7]);
  }
}

template <int W, int NC, bool COUNT>
__global__ void __launch_bounds__(XW_WARPS * 32) k_rows_warp(const WarpPar P)
{
  constexpr int NA_SMALL_W = 64 / W;
  extern __shared__ double smw[];
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  if (!COUNT && NC > 0) {
    for (int k = threadIdx.x; k < NC * W * W; k += blockDim.x) smw[k] = __ldg(P.tref + k);
    __syncthreads();
  }
  (void)NA_SMALL_W;
  constexpr int NTREF = (!COUNT && NC > 0) ? NC * W * W : 0;
  double* __restrict__ vals_s = smw + NTREF + (threadIdx.x >> 5) * (32 * XW_KPL);  // this warp's candidate values
  int longest = 0;
  for (;;) {
    // 32 consecutive rows per ticket; every lane looks at one of them, the warp then works through the eligible ones
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(P.ticket, 32u);
    base = __shfl_sync(FULL, base, 0);
    if (base >= (unsigned)P.ndof) break;
    const long long dl = (long long)base + lane;
    long long al = 0, rbl = 0;
    int nal = 0, nrl = 0;
    if (dl < P.ndof) {
      al = __ldg(P.adj_ptr + dl);
      nal = (int)(__ldg(P.adj_ptr + dl + 1) - al);
      if (!COUNT) {
        rbl = __ldg(P.rowptr + dl);
        nrl = (int)(__ldg(P.rowptr + dl + 1) - rbl);
      }
    }
    bool small = nal >= 1 && nal <= P.small_na;  // rows of the thread-per-row kernel
    if (!COUNT) small = small && nrl <= 32;
    const bool mine = dl < P.ndof && !small && nal * W <= 32 * XW_KPL;
    if (COUNT && mine && nal == 0) P.rowcnt[dl] = 0;
    unsigned todo = __ballot_sync(FULL, mine && nal > 0);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const long long d = (long long)base + src;
      const long long a = __shfl_sync(FULL, al, src);
      const int na = __shfl_sync(FULL, nal, src);
      const long long rowbase = __shfl_sync(FULL, rbl, src);
      const int cand = na * W;
      // ---- candidates: index k * 32 + lane (neighbouring lanes read neighbouring table entries and share the adjacency
      // entry / the element's coefficients: few distinct cache lines per load).  Full pass: the candidate's value is
      // fetched / generated here, in candidate order, and parked in shared memory — after the sort every lane would
      // touch 32 different elements per load and the L1 pipe sets the time (profiles/r02_notes.md)
      unsigned K[XW_KPL];
#pragma unroll
      for (int k = 0; k < XW_KPL; k++) {
        const int idx = k * 32 + lane;
        K[k] = 0xffffffffu;
        if (idx < cand) {
          const int t = idx / W, j = idx - t * W;
          const long long e = __ldg(P.adj + a + t);
          const int bi = XG_ADJ_BATCH(e);
          const size_t row = (size_t)XG_ADJ_ROW(e);
          const int c = __ldg(P.bd.table[bi] + row * W + j);
          if (c >= 0) K[k] = ((unsigned)c << 8) | (unsigned)idx;
          if (!COUNT) {
            const int loc = XG_ADJ_LOC(e);
            double val = 0.0;
            const double* __restrict__ cfp = (NC > 0) ? P.bd.coefs[bi] : nullptr;
            if (NC > 0 && cfp) {
              const double2* __restrict__ c2 = reinterpret_cast<const double2*>(cfp + row * 8);
              double cf[8];
#pragma unroll
              for (int m = 0; m < (NC + 1) / 2; m++) {
                const double2 x = __ldg(c2 + m);
                cf[2 * m] = x.x; cf[2 * m + 1] = x.y;
              }
              const double* __restrict__ tr = smw + loc * W + j;
#pragma unroll
              for (int m = 0; m < NC; m++) val += cf[m] * tr[m * W * W];
            } else {
              const double* __restrict__ mb = P.bd.mats[bi];
              if (mb) val = __ldg(mb + (row * W + loc) * W + j);  // else: sparsity only (explicit zero)
            }
            vals_s[idx] = val;
          }
        }
      }
      if (!COUNT) __syncwarp();
      sort256(K, lane);
      // ---- runs: heads (first of an equal-column run) and ends
      const unsigned prevlast = __shfl_up_sync(FULL, K[XW_KPL - 1], 1);
      const unsigned nextfirst = __shfl_down_sync(FULL, K[0], 1);
      unsigned fresh = 0u, last = 0u, valid = 0u;
#pragma unroll
      for (int k = 0; k < XW_KPL; k++) {
        if (K[k] != 0xffffffffu) {
          valid |= 1u << k;
          const unsigned pk = (k == 0) ? ((lane == 0) ? 0xffffffffu : prevlast) : K[k - 1];
          const unsigned nk = (k == XW_KPL - 1) ? ((lane == 31) ? 0xffffffffu : nextfirst) : K[k + 1];
          if ((pk >> 8) != (K[k] >> 8) || (k == 0 && lane == 0)) fresh |= 1u << k;
          if ((nk >> 8) != (K[k] >> 8) || (k == XW_KPL - 1 && lane == 31)) last |= 1u << k;
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
      if (COUNT) {
        if (lane == 0) P.rowcnt[d] = nr;
        longest = max(longest, nr);
        continue;
      }
      // ---- full pass: columns at the run heads
      int slot = incl - nf - 1;  // slot of the run the lane's first key continues (head in an earlier lane)
      {
        int q = slot;
#pragma unroll
        for (int k = 0; k < XW_KPL; k++)
          if ((fresh >> k) & 1u) { q++; P.colidx[rowbase + q] = (int)(K[k] >> 8); }
      }
      // values of the lane's (sorted) candidates
      double v[XW_KPL];
#pragma unroll
      for (int k = 0; k < XW_KPL; k++) v[k] = ((valid >> k) & 1u) ? vals_s[K[k] & 255u] : 0.0;
      // carry over the lane boundaries: a lane whose 8 keys all continue the previous lane's run passes the
      // running sum on (sequential adds), every other lane's carry-out is the sum of its own last run
      const bool cont_in = (valid & 1u) && !(fresh & 1u);
      const bool through = valid == 0xffu && fresh == 0u;
      double cout = 0.0;
      {
        double r = 0.0;
#pragma unroll
        for (int k = 0; k < XW_KPL; k++) {
          if ((fresh >> k) & 1u) r = 0.0;
          if ((valid >> k) & 1u) r += v[k];
        }
        cout = r;
      }
      double cin = 0.0;
      const int iters = (na >> 3) + 1;  // a run has at most na keys: it covers at most na/8 whole lanes
      for (int it = 0; it < iters; it++) {
        cin = __shfl_up_sync(FULL, cout, 1);
        if (lane == 0) cin = 0.0;
        if (through) {
          double r = cin;
#pragma unroll
          for (int k = 0; k < XW_KPL; k++) r += v[k];
          cout = r;
        }
      }
      {
        double r = cont_in ? cin : 0.0;
        int q = slot;
#pragma unroll
        for (int k = 0; k < XW_KPL; k++) {
          if ((fresh >> k) & 1u) { r = 0.0; q++; }
          if ((valid >> k) & 1u) r += v[k];
          if ((last >> k) & 1u) P.vals[rowbase + q] = r;
        }
      }
      __syncwarp();  // the next row overwrites vals_s
    }
  }
  if (COUNT) {
    longest = __reduce_max_sync(FULL, longest);
    if (lane == 0 && longest > 0) atomicMax(P.stats, longest);
  }
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <int W>
int launch_warp(xgpu_ctx* ctx, const WarpPar& S, int NC, bool count)
{
  cudaStream_t st = ctx->stream;
  unsigned grid = nblk(((int64_t)S.ndof + 31) / 32, XW_WARPS);
  if (grid > XGPU_NSM * 12) grid = XGPU_NSM * 12;
  if (count) {
    k_rows_warp<W, 0, true><<<grid, XW_WARPS * 32, 0, st>>>(S);
  } else {
    const size_t smem = sizeof(double) * ((size_t)NC * W * W + (size_t)XW_WARPS * 32 * XW_KPL);
    bool done = false;
    if (NC == 0) { k_rows_warp<W, 0, false><<<grid, XW_WARPS * 32, smem, st>>>(S); done = true; }
    if constexpr (W == 4 || W == 10) { if (NC == 7) { k_rows_warp<W, 7, false><<<grid, XW_WARPS * 32, smem, st>>>(S); done = true; } }
    if constexpr (W == 6) { if (NC == 4) { k_rows_warp<W, 4, false><<<grid, XW_WARPS * 32, smem, st>>>(S); done = true; } }
    if (!done) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "csr warp rows: coefficient count");
  }
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

// the warp-per-row kernels take rows with at most 256 candidates when (column << 8 | candidate) fits 32 bits
int xg_rows_warp_ok(int width, int64_t ndof)
{
  if (ndof >= ((int64_t)1 << 24)) return 0;
  return width == 4 || width == 6 || width == 10 || width == 12;
}

int xg_rows_warp_run(xgpu_ctx* ctx, xgpu_plan* P, const BatchDev& bd, const double* tref, int ncoef, bool count, int small_na,
                     unsigned int* ticket, int32_t* rowcnt, int* stats, int32_t* colidx, double* vals)
{
  WarpPar S;
  S.ndof = (int)P->ndof; S.small_na = small_na;
  S.adj_ptr = P->adj_ptr; S.adj = P->adj; S.bd = bd; S.tref = tref;
  S.rowptr = P->rowptr; S.rowcnt = rowcnt; S.stats = stats; S.ticket = ticket; S.colidx = colidx; S.vals = vals;
  switch (P->wmax) {
    case 4: return launch_warp<4>(ctx, S, ncoef, count);
    case 6: return launch_warp<6>(ctx, S, ncoef, count);
    case 10: return launch_warp<10>(ctx, S, ncoef, count);
    case 12: return launch_warp<12>(ctx, S, ncoef, count);
  }
  return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "csr warp rows: width");
}
