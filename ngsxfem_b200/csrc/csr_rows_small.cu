// csr_rows_small.cu — matrix rows with few adjacent elements (e.g. the edge dofs of a P2 tet mesh: 4-6 tets, 40-60
// candidate columns, 19-27 distinct), ONE THREAD PER ROW, no hash table:
//   * the thread packs its candidates (column << 6 | candidate index) into 64 REGISTERS and sorts them with a fully
//     unrolled bitonic network (static register indices, 2 instructions per compare-exchange, all 32 lanes busy on 32
//     different rows) — duplicates become runs, the columns come out ascending, equal columns stay in candidate
//     (= batch, table row, local index) order, the summation order of a sequential AddElementMatrix loop;
//   * count pass: the number of runs is the row length (xgpu_csr_plan_rows);
//   * full pass: a walk over the sorted registers gives every candidate its position in the row (a byte per candidate in
//     shared memory), then the adjacent elements are visited in order with STATIC local column indices: the matrix row
//     (or, for a procedural affine batch, the element's coefficients) arrives with vector loads, procedural values read
//     the reference tensors at bank-conflict-free addresses (all lanes share j, the row index i differs: padded rows),
//     and the values are added into the warp's staging buffer, which leaves with coalesced stores.
// The lane-group kernels of csr_fused.cu spend ~200 instructions per 30 contributions on per-step bookkeeping
// (profiles/r02_notes.md); here the bookkeeping is per 32 rows.
#include "csr.cuh"

namespace {

#define XS_BLOCK 128
#define XS_STAGE 1056  // doubles of warp-private shared memory: 32 slots x 33 (padded lane stride: also the transposed reads are conflict free)
#define XS_LS 33

struct SmallPar {
  int ndof;
  const long long* adj_ptr;
  const long long* adj;
  BatchDev bd;
  const double* tref;       // procedural batches: ncoef x W x W
  const long long* rowptr;  // full pass
  int32_t* rowcnt;          // count pass
  int* stats;               // count pass: [0] longest row
  int32_t* colidx;
  double* vals;
};

template <int N>
__device__ __forceinline__ void bitonic_regs(unsigned (&r)[N])
{
#pragma unroll
  for (int size = 2; size <= N; size <<= 1)
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll
      for (int i = 0; i < N; i++) {
        const int l = i ^ stride;
        if (l > i) {
          const unsigned lo = min(r[i], r[l]), hi = max(r[i], r[l]);
          if ((i & size) == 0) { r[i] = lo; r[l] = hi; } else { r[i] = hi; r[l] = lo; }
        }
      }
}

// W: width of every table row; NC: coefficients of the procedural batches (0: none); COUNT: count pass.
// Only the sorting network is unrolled (static register indices); every other phase is a loop over shared memory, so the
// program stays ~2000 instructions: the first version (everything unrolled, 5800 instructions, each executed once per
// warp) stalled on instruction fetch (profiles/r02_notes.md).
template <int W, int NC, bool COUNT>
__global__ void __launch_bounds__(XS_BLOCK, 4) k_rows_small(const SmallPar P)
{
  constexpr int NA = 64 / W;          // adjacent table rows a thread can take
  constexpr int WP = W + 1;           // padded row stride of the reference tensors in shared memory
  constexpr int NTREF = (!COUNT && NC > 0) ? ((NC * W * WP + 1) & ~1) : 0;
  extern __shared__ double sms[];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  if (!COUNT && NC > 0) {
    for (int k = tid; k < NC * W * W; k += XS_BLOCK) {
      const int m = k / (W * W), ij = k - m * W * W, i = ij / W, j = ij - i * W;
      sms[(m * W + i) * WP + j] = __ldg(P.tref + k);
    }
    __syncthreads();
  }
  const long long d = (long long)blockIdx.x * XS_BLOCK + tid;
  long long a = 0;
  int na = 0;
  if (d < P.ndof) {
    a = __ldg(P.adj_ptr + d);
    na = (int)(__ldg(P.adj_ptr + d + 1) - a);
  }
  long long rowbase = 0;
  int nr = 0;
  bool small = na >= 1 && na <= NA;
  if (!COUNT && small) {
    rowbase = __ldg(P.rowptr + d);
    nr = (int)(__ldg(P.rowptr + d + 1) - rowbase);
    small = nr <= 32;  // longer rows go to the lane-group kernel (same test there)
  }
  const unsigned FULL = 0xffffffffu;
  if (!__any_sync(FULL, small)) return;  // the whole warp leaves together: warp-level primitives below
  if (!small) na = 0;

  // ---- candidates -> registers (unrolled: static register indices)
  unsigned K[64];
  unsigned long long okmask = 0ull;  // candidate (t, j) holds a dof (>= 0)
#pragma unroll
  for (int t = 0; t < NA; t++) {
    int c[W];
#pragma unroll
    for (int j = 0; j < W; j++) c[j] = -1;
    if (t < na) {
      const long long e = __ldg(P.adj + a + t);
      const int32_t* __restrict__ trow = P.bd.table[XG_ADJ_BATCH(e)] + (size_t)XG_ADJ_ROW(e) * W;
      if ((W & 1) == 0) {
#pragma unroll
        for (int j = 0; j < W; j += 2) {
          const int2 v = __ldg(reinterpret_cast<const int2*>(trow + j));
          c[j] = v.x; c[j + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < W; j++) c[j] = __ldg(trow + j);
      }
    }
#pragma unroll
    for (int j = 0; j < W; j++) {
      const int idx = t * W + j;
      K[idx] = c[j] >= 0 ? (((unsigned)c[j] << 6) | (unsigned)idx) : 0xffffffffu;
      if (c[j] >= 0) okmask |= 1ull << idx;
    }
  }
#pragma unroll
  for (int idx = NA * W; idx < 64; idx++) K[idx] = 0xffffffffu;
  bitonic_regs<64>(K);

  if (COUNT) {
    int nu = 0;
#pragma unroll
    for (int p = 0; p < 64; p++) {
      const bool fresh = K[p] != 0xffffffffu && (p == 0 || (K[p] >> 6) != (K[p - 1] >> 6));
      nu += fresh ? 1 : 0;
    }
    if (small) P.rowcnt[d] = nu;
    const int longest = __reduce_max_sync(FULL, small ? nu : 0);
    if (lane == 0 && longest > 0) atomicMax(P.stats, longest);
    return;
  }

  // ---- full pass.  Warp-private shared memory, always indexed [slot][lane] (bank = lane: conflict free whatever the
  // slot): 64 x 32 words = the sorted keys, compacted IN PLACE to the row's columns by the rank loop, then (the same
  // 8 KB) 32 x 32 doubles = the row's values.  Block: position bytes [candidate][thread].
  unsigned* __restrict__ Ks = reinterpret_cast<unsigned*>(sms + NTREF + (size_t)wib * XS_STAGE) + lane;
  double* __restrict__ Vs = sms + NTREF + (size_t)wib * XS_STAGE + lane;
  unsigned char* __restrict__ pos_s = reinterpret_cast<unsigned char*>(sms + NTREF + (size_t)(XS_BLOCK / 32) * XS_STAGE) + tid;
#pragma unroll
  for (int p = 0; p < 64; p++) Ks[p * XS_LS] = K[p];
  {
    // position of every candidate in its row = index of its run; the run heads are the row's columns (q <= p: the
    // in-place store never overtakes the reads of this lane's column)
    int q = -1;
    unsigned prevcol = 0xffffffffu;
#pragma unroll 4
    for (int p = 0; p < 64; p++) {
      const unsigned key = Ks[p * XS_LS];
      if (key != 0xffffffffu) {
        const unsigned col = key >> 6;
        if (col != prevcol) { q++; Ks[q * XS_LS] = col; prevcol = col; }
        pos_s[(key & 63u) * XS_BLOCK] = (unsigned char)q;
      }
    }
  }
  __syncwarp();
  // columns out: one row per iteration, lanes over its entries (contiguous stores)
  const int mynr = small ? nr : 0;
  {
    const unsigned* __restrict__ Kw = reinterpret_cast<const unsigned*>(sms + NTREF + (size_t)wib * XS_STAGE);
    for (int L = 0; L < 32; L++) {
      const int nL = __shfl_sync(FULL, mynr, L);
      const long long bL = __shfl_sync(FULL, rowbase, L);
      if (lane < nL) P.colidx[bL + lane] = (int)Kw[lane * XS_LS + L];
    }
  }
  __syncwarp();
#pragma unroll 8
  for (int q = 0; q < 32; q++) Vs[q * XS_LS] = 0.0;  // own column only: no synchronisation needed before the adds
#pragma unroll 1
  for (int t = 0; t < na; t++) {
    const long long e = __ldg(P.adj + a + t);
    const int bi = XG_ADJ_BATCH(e);
    const size_t row = (size_t)XG_ADJ_ROW(e);
    const int loc = XG_ADJ_LOC(e);
    double v[W];
    bool have = true;
    const double* __restrict__ cfp = (NC > 0) ? P.bd.coefs[bi] : nullptr;
    if (NC > 0 && cfp) {
      double cf[8];
      const double2* __restrict__ c2 = reinterpret_cast<const double2*>(cfp + row * 8);
#pragma unroll
      for (int m = 0; m < (NC + 1) / 2; m++) {
        const double2 x = __ldg(c2 + m);
        cf[2 * m] = x.x; cf[2 * m + 1] = x.y;
      }
      const double* __restrict__ tr = sms + loc * WP;
#pragma unroll
      for (int j = 0; j < W; j++) {
        double r = 0.0;
#pragma unroll
        for (int m = 0; m < NC; m++) r += cf[m] * tr[m * W * WP + j];
        v[j] = r;
      }
    } else {
      const double* __restrict__ mb = P.bd.mats[bi];
      have = mb != nullptr;  // a batch without matrices only contributes sparsity (explicit zeros)
      if (have) {
        const double* __restrict__ mrow = mb + (row * W + loc) * W;
        if ((W & 1) == 0) {
#pragma unroll
          for (int j = 0; j < W; j += 2) {
            const double2 x = __ldg(reinterpret_cast<const double2*>(mrow + j));
            v[j] = x.x; v[j + 1] = x.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < W; j++) v[j] = __ldg(mrow + j);
        }
      }
    }
    if (have) {
      // the columns of one table row are distinct: the W read-modify-writes are independent — all loads first
      const unsigned ok = (unsigned)(okmask >> (t * W)) & ((1u << W) - 1u);
      int pp[W];
      double cur[W];
#pragma unroll
      for (int j = 0; j < W; j++) pp[j] = pos_s[(t * W + j) * XS_BLOCK];
#pragma unroll
      for (int j = 0; j < W; j++) cur[j] = ((ok >> j) & 1u) ? Vs[pp[j] * XS_LS] : 0.0;
#pragma unroll
      for (int j = 0; j < W; j++)
        if ((ok >> j) & 1u) Vs[pp[j] * XS_LS] = cur[j] + v[j];
    }
  }
  __syncwarp();
  {
    const double* __restrict__ Vw = sms + NTREF + (size_t)wib * XS_STAGE;
    for (int L = 0; L < 32; L++) {
      const int nL = __shfl_sync(FULL, mynr, L);
      const long long bL = __shfl_sync(FULL, rowbase, L);
      if (lane < nL) P.vals[bL + lane] = Vw[lane * XS_LS + L];
    }
  }
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <int W>
int launch_small(xgpu_ctx* ctx, const SmallPar& S, int NC, bool count)
{
  const unsigned grid = nblk(S.ndof, XS_BLOCK);
  cudaStream_t st = ctx->stream;
  if (count) {
    k_rows_small<W, 0, true><<<grid, XS_BLOCK, 0, st>>>(S);
  } else {
    const size_t ntref = ((size_t)NC * W * (W + 1) + 1) & ~(size_t)1;
    const size_t smem = sizeof(double) * (ntref + (size_t)(XS_BLOCK / 32) * XS_STAGE) + (size_t)64 * XS_BLOCK;
#define XS_LAUNCH(NCV)                                                                                                   \
  {                                                                                                                      \
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_rows_small<W, NCV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_rows_small<W, NCV, false><<<grid, XS_BLOCK, smem, st>>>(S);                                                        \
  }
    bool done = false;
    if (NC == 0) { XS_LAUNCH(0) done = true; }
    if constexpr (W == 4 || W == 10) { if (NC == 7) { XS_LAUNCH(7) done = true; } }   // P1 / P2 tets
    if constexpr (W == 6) { if (NC == 4) { XS_LAUNCH(4) done = true; } }              // P2 triangles
    if (!done) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "csr small rows: coefficient count");
#undef XS_LAUNCH
  }
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

}  // namespace

// widths the thread-per-row kernels are instantiated for; 0: not available (the lane-group kernels take every row)
int xg_rows_small_supports(int width, int ncoef)
{
  if (ncoef == 0) return 1;
  return ((width == 4 || width == 10) && ncoef == 7) || (width == 6 && ncoef == 4);
}

int xg_rows_small_na(int width, int64_t ndof)
{
  if (ndof >= ((int64_t)1 << 26)) return 0;  // (column << 6 | candidate) must fit 32 bits
  switch (width) {
    case 4: case 6: case 10: case 12: return 64 / width;
    default: return 0;
  }
}

int xg_rows_small_run(xgpu_ctx* ctx, xgpu_plan* P, const BatchDev& bd, const double* tref, int ncoef, bool count,
                      int32_t* rowcnt, int* stats, int32_t* colidx, double* vals)
{
  SmallPar S;
  S.ndof = (int)P->ndof;
  S.adj_ptr = P->adj_ptr; S.adj = P->adj; S.bd = bd; S.tref = tref;
  S.rowptr = P->rowptr; S.rowcnt = rowcnt; S.stats = stats; S.colidx = colidx; S.vals = vals;
  switch (P->wmax) {
    case 4: return launch_small<4>(ctx, S, ncoef, count);
    case 6: return launch_small<6>(ctx, S, ncoef, count);
    case 10: return launch_small<10>(ctx, S, ncoef, count);
    case 12: return launch_small<12>(ctx, S, ncoef, count);
  }
  return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "csr small rows: width");
}
