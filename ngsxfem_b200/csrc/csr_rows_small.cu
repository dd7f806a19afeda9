// csr_rows_small.cu — matrix rows with few adjacent elements (e.g. the edge dofs of a P2 tet mesh: 4-6 tets, 40-60
// candidate columns, 19-27 distinct), ONE THREAD PER ROW, no hash table.
//
// PATTERN pass (xgpu_csr_count_rows): the thread packs its candidates (column << 6 | candidate index) into 64 REGISTERS
// and sorts them with a fully unrolled bitonic network (static register indices, 2 instructions per compare-exchange,
// all 32 lanes busy on 32 different rows).  Duplicates become runs: the run heads are the row's columns (ascending),
// their number is the row length, and every candidate's run index is its POSITION in the row.  Columns (<= 32 x 4 B) and
// positions (64 x 1 B) are stored as a 192 B record per row — interleaved over the 32 rows of a warp so that every store
// / load of the records is a fully coalesced 512 B access.
// VALUES pass (xgpu_csr_assemble_fused): no table rows, no sort — the thread reads its record, visits its adjacent
// elements in order with STATIC local column indices (matrix row / procedural coefficients with vector loads, the
// reference tensors at bank-conflict-free addresses: all lanes share j, the row index differs -> padded rows) and adds
// value j at position pos[t*W + j] of its own [slot][lane] column of shared memory (bank = lane).  Entries are visited in
// (batch, table row) order = the summation order of a sequential AddElementMatrix loop.  Columns and values leave
// row by row with contiguous stores.  The same plan can be assembled again with other values (values pass only).
// The lane-group hash kernels of csr_fused.cu spend ~200 instructions per 30 contributions on per-step bookkeeping
// (profiles/r02_notes.md); here the bookkeeping is per 32 rows.
#include <algorithm>
#include "csr.cuh"
#include "h1.cuh"

namespace {

#define XS_BLOCK 128
#define XS_LS 33          // padded lane stride of the warp-private [slot][lane] arrays (transposed reads conflict free too)
#define XS_STAGE 1056     // doubles of warp-private shared memory: 32 slots x 33
#define XS_CHUNKS 12      // 16-byte chunks of a row record: 64 position bytes + 32 columns = 192 B

struct SmallPar {
  int ndof;
  int row_lo, row_hi, d0;   // values pass: rows [row_lo, row_hi) only, thread 0 of block 0 is row d0 (multiple of 32)
  const long long* adj_ptr;
  const long long* adj;
  BatchDev bd;
  const double* tref;       // procedural batches: ncoef x W x W
  const long long* rowptr;  // values pass
  int32_t* rowcnt;          // pattern pass
  int* stats;               // pattern pass: [0] longest row, [1] a row of this kernel has more than 32 nonzeros
  uint4* records;           // row records, chunk c of row d at [((d >> 5) * XS_CHUNKS + c) * 32 + (d & 31)]
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

// ---- row `loc` of the mass + stiffness matrix of an affine P2 tetrahedron in CLOSED FORM (h1.cuh basis: vertex hats
// lam_v, edge functions -2 lam_a lam_b; reference tet lam = (x, y, z, 1-x-y-z), edges XG_TET_EDGE_A/B).
// cf[0] = c_m |det F|, cf[1..6] = c_s |det F| G_ab (a <= b) as xgpu_affine_coefs writes them.  With the barycentric Gram
// matrix  Gam_ij = grad lam_i . (c_s |det| G) grad lam_j  and  w(x,y) = 1 + delta_xy:
//   vertex a , vertex c   :  cm w(a,c)/120 + Gam_ac/6
//   vertex a , edge (c,d) : -cm (w(a,c)+w(a,d)-1)/360 - (Gam_ac+Gam_ad)/12
//   edge (a,b), vertex c  : -cm (w(a,c)+w(b,c)-1)/360 - (Gam_ac+Gam_bc)/12
//   edge (a,b), edge (c,d):  cm w(a,c)w(a,d)w(b,c)w(b,d)/1260 + (Gam_bd w(a,c) + Gam_bc w(a,d) + Gam_ad w(b,c) + Gam_ac w(b,d))/30
// (integrals of barycentric monomials: alpha! / (|alpha|+3)! on the reference tet).  Equal to sum_m cf[m] tref[m][loc][:]
// up to rounding (tools check: 3e-16); no table loads — the thread-per-row kernel is bound by shared-memory wavefronts
// when every lane reads its own row of the 7 reference tensors (profiles/r02_notes.md).
__device__ __forceinline__ void p2tet_row(const double (&cf)[8], int loc, double (&v)[10])
{
  const double cm = cf[0];
  const double g03 = -(cf[1] + cf[2] + cf[3]), g13 = -(cf[2] + cf[4] + cf[5]), g23 = -(cf[3] + cf[5] + cf[6]);
  const double g33 = -(g03 + g13 + g23);
  const double R[4][4] = {{cf[1], cf[2], cf[3], g03}, {cf[2], cf[4], cf[5], g13}, {cf[3], cf[5], cf[6], g23}, {g03, g13, g23, g33}};
  const int e = loc - 4;
  const int a = (loc < 4) ? loc : (e < 3 ? 3 : (e < 5 ? 0 : 1));
  const int b = (e < 3) ? e : (e == 3 ? 1 : 2);   // edge rows only
  double ra[4], wa[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    ra[c] = (a == 0) ? R[0][c] : ((a == 1) ? R[1][c] : ((a == 2) ? R[2][c] : R[3][c]));
    wa[c] = (a == c) ? 2.0 : 1.0;
  }
  if (loc < 4) {
#pragma unroll
    for (int c = 0; c < 4; c++) v[c] = cm * (wa[c] * (1.0 / 120.0)) + ra[c] * (1.0 / 6.0);
#pragma unroll
    for (int f = 0; f < 6; f++) {
      const int c = XG_TET_EDGE_A(f), d = XG_TET_EDGE_B(f);
      v[4 + f] = -(cm * ((wa[c] + wa[d] - 1.0) * (1.0 / 360.0)) + (ra[c] + ra[d]) * (1.0 / 12.0));
    }
    return;
  }
  double rb[4], wb[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    rb[c] = (b == 0) ? R[0][c] : ((b == 1) ? R[1][c] : R[2][c]);
    wb[c] = (b == c) ? 2.0 : 1.0;
  }
#pragma unroll
  for (int c = 0; c < 4; c++) v[c] = -(cm * ((wa[c] + wb[c] - 1.0) * (1.0 / 360.0)) + (ra[c] + rb[c]) * (1.0 / 12.0));
#pragma unroll
  for (int f = 0; f < 6; f++) {
    const int c = XG_TET_EDGE_A(f), d = XG_TET_EDGE_B(f);
    const double k = rb[d] * wa[c] + rb[c] * wa[d] + ra[d] * wb[c] + ra[c] * wb[d];
    v[4 + f] = cm * ((wa[c] * wa[d]) * (wb[c] * wb[d]) * (1.0 / 1260.0)) + k * (1.0 / 30.0);
  }
}

// ---- pattern pass.  W: width of every table row
template <int W>
__global__ void __launch_bounds__(XS_BLOCK, 4) k_rows_small_pattern(const SmallPar P)
{
  constexpr int NA = 64 / W;  // adjacent table rows a thread can take
  extern __shared__ double sms[];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const long long d = (long long)blockIdx.x * XS_BLOCK + tid;
  long long a = 0;
  int na = 0;
  if (d < P.ndof) {
    a = __ldg(P.adj_ptr + d);
    na = (int)(__ldg(P.adj_ptr + d + 1) - a);
  }
  const bool small = na >= 1 && na <= NA;
  const unsigned FULL = 0xffffffffu;
  if (!__any_sync(FULL, small)) return;  // the whole warp leaves together
  if (!small) na = 0;
  // ---- candidates -> registers (unrolled: static register indices)
  unsigned K[64];
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
    }
  }
#pragma unroll
  for (int idx = NA * W; idx < 64; idx++) K[idx] = 0xffffffffu;
  bitonic_regs<64>(K);

  // ---- runs -> columns [slot][lane] and position bytes [lane][candidate] (row stride 17 words: conflict free both ways)
  // (64 column slots: a row of this kernel may have up to 64 distinct columns; only the first 32 are stored)
  unsigned* Cs = reinterpret_cast<unsigned*>(sms) + (size_t)wib * (64 * XS_LS + 32 * 17) + lane;
  unsigned* Pw = reinterpret_cast<unsigned*>(sms) + (size_t)wib * (64 * XS_LS + 32 * 17) + 64 * XS_LS + lane * 17;
  unsigned char* Pb = reinterpret_cast<unsigned char*>(Pw);  // the same bytes: no __restrict__ on either
#pragma unroll
  for (int i = 0; i < 16; i++) Pw[i] = 0xffffffffu;  // 255: the candidate holds no dof
  int q = -1;
  unsigned prevcol = 0xffffffffu;
#pragma unroll
  for (int p = 0; p < 64; p++) {
    const unsigned key = K[p];
    if (key != 0xffffffffu) {
      const unsigned col = key >> 6;
      if (col != prevcol) {
        q++;
        prevcol = col;
        Cs[q * XS_LS] = col;
      }
      Pb[key & 63u] = (unsigned char)q;
    }
  }
  const int nu = q + 1;
  if (small) P.rowcnt[d] = nu;
  const int longest = __reduce_max_sync(FULL, small ? nu : 0);
  if (lane == 0 && longest > 0) {
    atomicMax(P.stats, longest);
    if (longest > 32) atomicOr(P.stats + 1, 1);  // such rows are assembled by the lane-group kernel
  }
  // ---- record out: chunk c of the warp's 32 rows is one contiguous 512 B store
  uint4* __restrict__ rec = P.records + (size_t)(d >> 5) * (XS_CHUNKS * 32) + lane;
#pragma unroll
  for (int c = 0; c < 4; c++) rec[c * 32] = make_uint4(Pw[4 * c], Pw[4 * c + 1], Pw[4 * c + 2], Pw[4 * c + 3]);
#pragma unroll
  for (int c = 0; c < 8; c++)
    rec[(4 + c) * 32] = make_uint4(Cs[(4 * c) * XS_LS], Cs[(4 * c + 1) * XS_LS], Cs[(4 * c + 2) * XS_LS], Cs[(4 * c + 3) * XS_LS]);
}

// ---- values pass.  NC: coefficients of the procedural batches (0: none); CF: their rows in closed form (P2 tets)
template <int W, int NC, bool CF>
__global__ void __launch_bounds__(XS_BLOCK, 4) k_rows_small_values(const SmallPar P)
{
  constexpr int NA = 64 / W;
  constexpr int WP = W + 1;  // padded row stride of the reference tensors in shared memory
  constexpr int NTREF = (NC > 0 && !CF) ? ((NC * W * WP + 1) & ~1) : 0;
  extern __shared__ double sms[];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  if (NC > 0 && !CF) {
    for (int k = tid; k < NC * W * W; k += XS_BLOCK) {
      const int m = k / (W * W), ij = k - m * W * W, i = ij / W, j = ij - i * W;
      sms[(m * W + i) * WP + j] = __ldg(P.tref + k);
    }
    __syncthreads();
  }
  const long long d = (long long)P.d0 + (long long)blockIdx.x * XS_BLOCK + tid;
  long long a = 0, rowbase = 0;
  int na = 0, nr = 0;
  if (d >= P.row_lo && d < P.row_hi) {
    a = __ldg(P.adj_ptr + d);
    na = (int)(__ldg(P.adj_ptr + d + 1) - a);
  }
  bool small = na >= 1 && na <= NA;
  if (small) {
    rowbase = __ldg(P.rowptr + d);
    nr = (int)(__ldg(P.rowptr + d + 1) - rowbase);
    small = nr <= 32;  // longer rows: the lane-group kernel (same test there)
  }
  const unsigned FULL = 0xffffffffu;
  if (!__any_sync(FULL, small)) return;
  if (!small) { na = 0; nr = 0; }
  // warp-private shared memory [slot][lane] (bank = lane whatever the slot): first the row's columns, then (the same
  // 8 KB) the row's FP64 values
  unsigned* __restrict__ Cs = reinterpret_cast<unsigned*>(sms + NTREF + (size_t)wib * XS_STAGE) + lane;
  double* __restrict__ Vs = sms + NTREF + (size_t)wib * XS_STAGE + lane;
  const uint4* __restrict__ rec = P.records + (size_t)(d >> 5) * (XS_CHUNKS * 32) + lane;
  unsigned pw[16];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const uint4 x = small ? __ldg(rec + c * 32) : make_uint4(~0u, ~0u, ~0u, ~0u);
    pw[4 * c] = x.x; pw[4 * c + 1] = x.y; pw[4 * c + 2] = x.z; pw[4 * c + 3] = x.w;
  }
#pragma unroll
  for (int c = 0; c < 8; c++) {
    if (4 * c < nr) {  // chunks beyond the row hold nothing
      const uint4 x = __ldg(rec + (4 + c) * 32);
      Cs[(4 * c) * XS_LS] = x.x; Cs[(4 * c + 1) * XS_LS] = x.y; Cs[(4 * c + 2) * XS_LS] = x.z; Cs[(4 * c + 3) * XS_LS] = x.w;
    }
  }
  __syncwarp();
  {  // columns out: one row per iteration, lanes over its entries (contiguous stores)
    const unsigned* __restrict__ Cw = reinterpret_cast<const unsigned*>(sms + NTREF + (size_t)wib * XS_STAGE);
    for (int L = 0; L < 32; L++) {
      const int nL = __shfl_sync(FULL, nr, L);
      const long long bL = __shfl_sync(FULL, rowbase, L);
      if (lane < nL) P.colidx[bL + lane] = (int)Cw[lane * XS_LS + L];
    }
  }
  __syncwarp();
#pragma unroll 8
  for (int q = 0; q < 32; q++) Vs[q * XS_LS] = 0.0;  // own column only: no synchronisation needed before the adds
#pragma unroll
  for (int t = 0; t < NA; t++) {
    if (t < na) {
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
        if constexpr (CF && W == 10 && NC == 7) {
          p2tet_row(cf, loc, v);
        } else {
          const double* __restrict__ tr = sms + loc * WP;
#pragma unroll
          for (int j = 0; j < W; j++) {
            double r = 0.0;
#pragma unroll
            for (int m = 0; m < NC; m++) r += cf[m] * tr[m * W * WP + j];
            v[j] = r;
          }
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
        int pp[W];
        double cur[W];
#pragma unroll
        for (int j = 0; j < W; j++) pp[j] = (int)((pw[(t * W + j) >> 2] >> (8 * ((t * W + j) & 3))) & 255u);
#pragma unroll
        for (int j = 0; j < W; j++) cur[j] = (pp[j] != 255) ? Vs[pp[j] * XS_LS] : 0.0;
#pragma unroll
        for (int j = 0; j < W; j++)
          if (pp[j] != 255) Vs[pp[j] * XS_LS] = cur[j] + v[j];
      }
    }
  }
  __syncwarp();
  {
    const double* __restrict__ Vw = sms + NTREF + (size_t)wib * XS_STAGE;
    for (int L = 0; L < 32; L++) {
      const int nL = __shfl_sync(FULL, nr, L);
      const long long bL = __shfl_sync(FULL, rowbase, L);
      if (lane < nL) P.vals[bL + lane] = Vw[lane * XS_LS + L];
    }
  }
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <int W>
int launch_small(xgpu_ctx* ctx, const SmallPar& S, int NC, bool pattern, bool closed_form)
{
  const unsigned grid = pattern ? nblk(S.ndof, XS_BLOCK) : nblk((int64_t)S.row_hi - S.d0, XS_BLOCK);
  cudaStream_t st = ctx->stream;
  if (pattern) {
    const size_t smem = sizeof(unsigned) * (size_t)(XS_BLOCK / 32) * (64 * XS_LS + 32 * 17);
    k_rows_small_pattern<W><<<grid, XS_BLOCK, smem, st>>>(S);
  } else {
    const bool cf = closed_form && W == 10 && NC == 7;
    const size_t ntref = cf ? 0 : (((size_t)NC * W * (W + 1) + 1) & ~(size_t)1);
    const size_t smem = sizeof(double) * (ntref + (size_t)(XS_BLOCK / 32) * XS_STAGE);
#define XS_LAUNCH(NCV, CFV)                                                                                              \
  {                                                                                                                      \
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_rows_small_values<W, NCV, CFV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_rows_small_values<W, NCV, CFV><<<grid, XS_BLOCK, smem, st>>>(S);                                                   \
  }
    bool done = false;
    if (NC == 0) { XS_LAUNCH(0, false) done = true; }
    if constexpr (W == 10) { if (NC == 7 && cf) { XS_LAUNCH(7, true) done = true; } }   // P2 tets, closed-form rows
    if constexpr (W == 4 || W == 10) { if (NC == 7 && !done) { XS_LAUNCH(7, false) done = true; } }   // P1 / P2 tets
    if constexpr (W == 6) { if (NC == 4) { XS_LAUNCH(4, false) done = true; } }              // P2 triangles
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

size_t xg_rows_small_record_bytes(int64_t ndof) { return (size_t)((ndof + 31) / 32) * XS_CHUNKS * 32 * sizeof(uint4); }

int xg_rows_small_run(xgpu_ctx* ctx, xgpu_plan* P, const BatchDev& bd, const double* tref, int ncoef, bool pattern,
                      int32_t* rowcnt, int* stats, int32_t* colidx, double* vals, int64_t row_lo, int64_t row_hi)
{
  // closed-form rows only when the reference tensors are the library's own P2-tet tensors (the formula is THAT basis)
  const bool cfm = !pattern && ctx->opt_csr_closed_form && tref && xg_reftensor_is(ctx, tref, 3, 2);
  SmallPar S;
  S.ndof = (int)P->ndof;
  S.row_lo = (int)row_lo; S.row_hi = (int)std::min<int64_t>(row_hi, P->ndof); S.d0 = (int)(row_lo & ~(int64_t)31);
  if (!pattern && S.row_hi <= S.row_lo) return XGPU_OK;
  S.adj_ptr = P->adj_ptr; S.adj = P->adj; S.bd = bd; S.tref = tref;
  S.rowptr = P->rowptr; S.rowcnt = rowcnt; S.stats = stats; S.records = (uint4*)P->pat_small; S.colidx = colidx; S.vals = vals;
  switch (P->wmax) {
    case 4: return launch_small<4>(ctx, S, ncoef, pattern, cfm);
    case 6: return launch_small<6>(ctx, S, ncoef, pattern, cfm);
    case 10: return launch_small<10>(ctx, S, ncoef, pattern, cfm);
    case 12: return launch_small<12>(ctx, S, ncoef, pattern, cfm);
  }
  return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "csr small rows: width");
}
