// csr.cu — sparsity graph + deterministic element-matrix assembly.
// (RestrictedBilinearForm::GetGraph utils/restrictedblf.cpp:142-163 -> NGSolve MatrixGraph;
//  SparseMatrix::AddElementMatrix.)
//
// The reference scatters element by element with a binary search per entry (and relies on
// element colouring for thread safety).  Here assembly is a ROW-WISE GATHER, one warp per
// matrix row, everything staged in shared memory:
//   plan    : transposed table  dof -> [(table row, local index)], built by histogram + scan +
//             atomic-cursor fill, then each segment is sorted so the order is deterministic.
//             Then ONE pattern pass: the warp inserts the dofs of all adjacent table rows into a
//             shared-memory hash set (duplicates vanish, ~60 % of the candidates), compacts it and
//             rank-sorts the survivors (all distinct -> rank = #smaller) straight into a candidate-
//             offset scratch; a scan of the unique counts gives rowptr.
//             "union of cliques, ascending columns, explicit zeros kept" — bit-exact pattern.
//   pattern : compaction of the scratch rows into the caller's colidx (pure copy).
//   scatter : the warp stages the row's columns and FP64 accumulators in shared memory, walks the
//             sorted adjacency, reads each element-matrix row exactly once (contiguous w doubles),
//             finds the column slot by binary search in shared memory and adds in (batch, table
//             row, local index) order: no atomics, bit-reproducible, every matrix entry is read
//             once and every value written once (write-only: no memset of vals needed).
// All kernels are HBM/L2-gather bound; algorithmic bytes are stated in DESIGN.md.
#include "csr.cuh"
#include <algorithm>
#include <vector>

namespace {

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

__global__ void k_adj_count(const int32_t* __restrict__ table, int64_t nent, int32_t* __restrict__ cnt)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nent) return;
  const int32_t d = __ldg(table + i);
  if (d >= 0) atomicAdd(cnt + d, 1);
}

__global__ void k_adj_fill(const int32_t* __restrict__ table, int64_t nent, int width, int bi,
                           const long long* __restrict__ adj_ptr, int32_t* __restrict__ cursor, long long* __restrict__ adj)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nent) return;
  const int32_t d = __ldg(table + i);
  if (d < 0) return;
  const int64_t r = i / width;
  const int loc = (int)(i - r * width);
  const int pos = atomicAdd(cursor + d, 1);
  adj[adj_ptr[d] + pos] = XG_ADJ(bi, r, loc);
}

// bitonic network on K registers (compile-time indices: the array never leaves the register file)
template <int K>
__device__ __forceinline__ void sort_regs(long long (&r)[K])
{
#pragma unroll
  for (int size = 2; size <= K; size <<= 1)
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll
      for (int i = 0; i < K; i++) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = ((i & size) == 0);
          const long long a = r[i], b = r[j];
          const bool sw = (a > b) == up;
          r[i] = sw ? b : a;
          r[j] = sw ? a : b;
        }
      }
}

template <int K>
__device__ __forceinline__ void sort_segment_regs(long long* __restrict__ seg, int na)
{
  long long r[K];
#pragma unroll
  for (int k = 0; k < K; k++) r[k] = (k < na) ? seg[k] : 0x7fffffffffffffffLL;  // all loads in flight together
  sort_regs<K>(r);
#pragma unroll
  for (int k = 0; k < K; k++) if (k < na) seg[k] = r[k];
}

// sort every adjacency segment ascending (the atomic-cursor fill leaves them in arbitrary order): segments of up to
// 32 entries go through a register sorting network (loads / stores of the whole segment are independent, so they
// pipeline; the former in-place insertion sort was a chain of dependent global accesses), longer ones fall back to
// insertion sort.  Also computes the candidate count of each row = sum of the widths of the adjacent table rows, and
// notes whether any dof occurs twice inside one table row (facet-patch rows [dofs(el1) ; dofs(el2)] do)
__global__ void k_adj_sort(int64_t ndof, const long long* __restrict__ adj_ptr, long long* __restrict__ adj,
                           BatchDev bd, int32_t* __restrict__ cand, int* __restrict__ stats)
{
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0, dup = 0;
  if (d < ndof) {
    const long long a = adj_ptr[d], b = adj_ptr[d + 1];
    const int na = (int)(b - a);
    if (na <= 1) {
    } else if (na <= 8) sort_segment_regs<8>(adj + a, na);
    else if (na <= 16) sort_segment_regs<16>(adj + a, na);
    else if (na <= 32) sort_segment_regs<32>(adj + a, na);
    else {
      for (long long i = a + 1; i < b; i++) {
        const long long key = adj[i];
        long long j = i - 1;
        while (j >= a && adj[j] > key) { adj[j + 1] = adj[j]; j--; }
        adj[j + 1] = key;
      }
    }
    long long prev = -1;
    for (long long i = a; i < b; i++) {
      const long long e = adj[i];
      dup |= (XG_ADJ_ROWKEY(e) == prev);
      prev = XG_ADJ_ROWKEY(e);
      c += bd.width[XG_ADJ_BATCH(e)];
    }
    cand[d] = c;
  }
  const int nbig = __popc(__ballot_sync(0xffffffffu, c > 64));  // rows of the warp-per-row kernels (csr_rows_warp.cu)
  for (int s = 16; s > 0; s >>= 1) c = max(c, __shfl_down_sync(0xffffffffu, c, s));
  dup = __any_sync(0xffffffffu, dup);
  if ((threadIdx.x & 31) == 0) {
    if (c > 0) atomicMax(stats, c);
    if (dup) atomicOr(stats + 1, 1);
    if (nbig) atomicAdd(stats + 3, nbig);
  }
}

// rank sort of nu distinct values in tab (shared): K values per lane, rank = #smaller values
template <int K>
__device__ __forceinline__ void rank_sort_store(const int* tab, int nu, int lane, int32_t* __restrict__ out)
{
  int x[K], r[K];
#pragma unroll
  for (int k = 0; k < K; k++) { const int i = lane + 32 * k; x[k] = (i < nu) ? tab[i] : XG_EMPTY; r[k] = 0; }
  for (int j = 0; j < nu; j += 4) {
    const int4 v = *reinterpret_cast<const int4*>(tab + j);  // padded with XG_EMPTY (never smaller)
#pragma unroll
    for (int k = 0; k < K; k++) r[k] += (v.x < x[k]) + (v.y < x[k]) + (v.z < x[k]) + (v.w < x[k]);
  }
#pragma unroll
  for (int k = 0; k < K; k++) if (lane + 32 * k < nu) out[r[k]] = x[k];
}

// ---- one pattern pass: hash-set dedupe + rank sort, one warp per matrix row
template <bool DUPS>
__global__ void __launch_bounds__(256) k_pattern(int ndof, const long long* __restrict__ adj_ptr,
                                                 const long long* __restrict__ adj, BatchDev bd, int HCAP, int SL, int G,
                                                 const long long* __restrict__ candoff,
                                                 int32_t* __restrict__ tmpcol, int32_t* __restrict__ rowcnt,
                                                 int* __restrict__ stats, int small_max, int big_max)
{
  extern __shared__ int sbuf[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int* tab = sbuf + wib * (HCAP + 4);
  const int warp0 = blockIdx.x * (blockDim.x >> 5) + wib;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int g = lane / SL, ls = lane - g * SL;
  int mymax = 0;
  for (int d = warp0; d < ndof; d += nwarps) {
    const long long o = candoff[d];
    const int cand = (int)(candoff[d + 1] - o);
    if (cand <= small_max || cand > big_max) continue;  // rows of k_pattern_small (it also writes the empty rows) / of another size class
    if (cand == 0) { if (lane == 0) rowcnt[d] = 0; continue; }
    // table size: power of two >= 1.25 * candidates (duplicates keep the load far lower)
    int H = 32, shift = 27;
    while (H < cand + (cand >> 2)) { H <<= 1; shift--; }
    for (int k = lane; k < H; k += 32) tab[k] = XG_EMPTY;
    __syncwarp();
    const long long a = adj_ptr[d];
    const int na = (int)(adj_ptr[d + 1] - a);
    const long long* __restrict__ adjd = adj + a;
    for (int t0 = 0; t0 < na; t0 += G) {
      const int t = t0 + g;
      if (g < G && t < na) {
        const long long e = adjd[t];
        bool fresh = true;
        if (DUPS) fresh = (t == 0) || (XG_ADJ_ROWKEY(adjd[t - 1]) != XG_ADJ_ROWKEY(e));  // same table row twice
        if (fresh) {
          const int bi = XG_ADJ_BATCH(e);
          const int w = bd.width[bi];
          const int32_t* __restrict__ row = bd.table[bi] + XG_ADJ_ROW(e) * w;  // nrows*width < 2^31 (checked on the host)
          for (int lj = ls; lj < w; lj += SL) {
            const int c = __ldg(row + lj);
            if (c >= 0) {
              unsigned h = ((unsigned)c * 2654435761u) >> shift;
              while (true) {
                const int old = atomicCAS(&tab[h], XG_EMPTY, c);
                if (old == XG_EMPTY || old == c) break;
                h = (h + 1) & (H - 1);
              }
            }
          }
        }
      }
    }
    __syncwarp();
    // in-place compaction of the occupied slots (write index <= read index)
    int nu = 0;
    for (int base = 0; base < H; base += 32) {
      const int x = tab[base + lane];
      const bool keep = (x != XG_EMPTY);
      __syncwarp();
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) tab[nu + __popc(bal & ((1u << lane) - 1u))] = x;
      nu += __popc(bal);
      __syncwarp();
    }
    if (lane == 0) rowcnt[d] = nu;
    mymax = max(mymax, nu);
    if (nu <= 128) {
      if (lane < 4) tab[nu + lane] = XG_EMPTY;  // pad for the int4 reads (slots exist: HCAP + 4)
      __syncwarp();
      if (nu <= 32) rank_sort_store<1>(tab, nu, lane, tmpcol + o);
      else if (nu <= 64) rank_sort_store<2>(tab, nu, lane, tmpcol + o);
      else rank_sort_store<4>(tab, nu, lane, tmpcol + o);
    } else {
      int P = 256;
      while (P < nu) P <<= 1;
      for (int k = nu + lane; k < P; k += 32) tab[k] = XG_EMPTY;  // P <= HCAP (host sizes HCAP >= 2 * maxcand)
      __syncwarp();
      for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int t = lane; t < (P >> 1); t += 32) {
            const int lo = 2 * t - (t & (stride - 1));
            const int hi = lo + stride;
            const bool up = ((lo & size) == 0);
            const int xa = tab[lo], xb = tab[hi];
            if ((xa > xb) == up) { tab[lo] = xb; tab[hi] = xa; }
          }
          __syncwarp();
        }
      for (int k = lane; k < nu; k += 32) tmpcol[o + k] = tab[k];
    }
    __syncwarp();
  }
  for (int s = 16; s > 0; s >>= 1) mymax = max(mymax, __shfl_down_sync(0xffffffffu, mymax, s));
  if (lane == 0 && mymax > 0) atomicMax(stats + 2, mymax);
}

// ---- short rows (at most XG_SMALL_CAND candidates, e.g. the edge dofs of a P2 tet mesh: 4-6
// adjacent elements): one THREAD per row.  The warp-per-row pass above spends most of its
// instructions on per-row fixed costs when a row only has 40-60 candidates; here every thread owns
// a private 64-slot hash set in shared memory, laid out slot-major (slot s of thread t at
// [s * blockDim + t]) so that the 32 lanes of a warp always hit 32 different banks, inserts its
// candidates serially (no atomics, no warp synchronisation), and compacts + insertion-sorts the
// survivors in place.
#define XG_SMALL_CAND 63
#define XG_SMALL_H 64
#define XG_SMALL_BLOCK 128
#define XG_SMALL_CH 10  // table entries loaded per batch (even)
template <bool DUPS>
__global__ void __launch_bounds__(XG_SMALL_BLOCK, 6) k_pattern_small(int ndof, const long long* __restrict__ adj_ptr,
                                                                  const long long* __restrict__ adj, BatchDev bd,
                                                                  const long long* __restrict__ candoff,
                                                                  int32_t* __restrict__ tmpcol, int32_t* __restrict__ rowcnt,
                                                                  int* __restrict__ stats)
{
  __shared__ int stab[XG_SMALL_H * XG_SMALL_BLOCK];
  int* tab = stab + threadIdx.x;
#define TAB(s) tab[(s) * XG_SMALL_BLOCK]
  int mymax = 0;
  for (int d = blockIdx.x * XG_SMALL_BLOCK + threadIdx.x; d < ndof; d += gridDim.x * XG_SMALL_BLOCK) {
    const long long o = candoff[d];
    const int cand = (int)(candoff[d + 1] - o);
    if (cand > XG_SMALL_CAND) continue;
    if (cand == 0) { rowcnt[d] = 0; continue; }
#pragma unroll
    for (int k = 0; k < XG_SMALL_H; k++) TAB(k) = XG_EMPTY;
    const long long a = adj_ptr[d];
    const int na = (int)(adj_ptr[d + 1] - a);
    long long prev = -1;
    long long e_next = adj[a];
    for (int t = 0; t < na; t++) {
      const long long e = e_next;
      if (t + 1 < na) e_next = adj[a + t + 1];
      if (DUPS) {
        const bool same = (XG_ADJ_ROWKEY(e) == prev);
        prev = XG_ADJ_ROWKEY(e);
        if (same) continue;
      }
      const int bi = XG_ADJ_BATCH(e);
      const int w = bd.width[bi];
      const int32_t* __restrict__ row = bd.table[bi] + XG_ADJ_ROW(e) * w;
      for (int j0 = 0; j0 < w; j0 += XG_SMALL_CH) {
        int cc[XG_SMALL_CH];  // the whole chunk is loaded before the first probe (loads in flight)
        if ((w & 1) == 0) {   // even width: rows are 8-byte aligned
#pragma unroll
          for (int u = 0; u < XG_SMALL_CH; u += 2) {
            int2 v = make_int2(-1, -1);
            if (j0 + u < w) v = __ldg(reinterpret_cast<const int2*>(row + j0 + u));
            cc[u] = v.x; cc[u + 1] = v.y;
          }
        } else {
#pragma unroll
          for (int u = 0; u < XG_SMALL_CH; u++) cc[u] = (j0 + u < w) ? __ldg(row + j0 + u) : -1;
        }
#pragma unroll
        for (int u = 0; u < XG_SMALL_CH; u++) {
          const int c = cc[u];
          if (c >= 0) {
            unsigned h = ((unsigned)c * 2654435761u) >> 26;
            while (true) {  // at most 63 keys in 64 slots: an empty slot always ends the probe
              const int old = TAB(h);
              if (old == c) break;
              if (old == XG_EMPTY) { TAB(h) = c; break; }
              h = (h + 1) & (XG_SMALL_H - 1);
            }
          }
        }
      }
    }
    // compaction of the occupied slots (write index <= read index) ...
    int nu = 0;
#pragma unroll
    for (int k = 0; k < XG_SMALL_H; k++) {
      const int x = TAB(k);
      if (x != XG_EMPTY) { TAB(nu) = x; nu++; }
    }
    rowcnt[d] = nu;
    mymax = max(mymax, nu);
    if (nu <= 32) {
      // ... then a bitonic network on 32 REGISTERS (fully unrolled, static indices: no memory traffic,
      // no divergence between the rows of a warp), padded with XG_EMPTY which sorts to the end
      int r[32];
#pragma unroll
      for (int k = 0; k < 32; k++) r[k] = (k < nu) ? TAB(k) : XG_EMPTY;
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const int l = i ^ stride;
            if (l > i) {
              const int lo = min(r[i], r[l]), hi = max(r[i], r[l]);
              if ((i & size) == 0) { r[i] = lo; r[l] = hi; } else { r[i] = hi; r[l] = lo; }
            }
          }
#pragma unroll
      for (int k = 0; k < 32; k++) if (k < nu) tmpcol[o + k] = r[k];
    } else {
      for (int i = 1; i < nu; i++) {  // rare: insertion sort in shared memory
        const int x = TAB(i);
        int j = i - 1;
        while (j >= 0) {
          const int y = TAB(j);
          if (y <= x) break;
          TAB(j + 1) = y;
          j--;
        }
        TAB(j + 1) = x;
      }
      for (int k = 0; k < nu; k++) tmpcol[o + k] = TAB(k);
    }
  }
#undef TAB
  mymax = __reduce_max_sync(0xffffffffu, mymax);
  if ((threadIdx.x & 31) == 0 && mymax > 0) atomicMax(stats + 2, mymax);
}

// copy the unique columns of every row from the candidate-offset scratch to the final CSR
__global__ void __launch_bounds__(256) k_compact(int64_t ndof, const long long* __restrict__ candoff,
                                                 const long long* __restrict__ rowptr, const int32_t* __restrict__ tmpcol,
                                                 int32_t* __restrict__ colidx)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t d = warp0; d < ndof; d += nwarps) {
    const long long r0 = rowptr[d];
    const int nr = (int)(rowptr[d + 1] - r0);
    const long long o = candoff[d];
    for (int k = lane; k < nr; k += 32) colidx[r0 + k] = tmpcol[o + k];
  }
}

// row-wise gather of the element matrices (see header comment).  A warp is split into NG lane
// groups of GS lanes (GS = widest table row, e.g. 3 groups of 10 lanes for P2 tets); every group
// owns ONE matrix row: FP64 accumulators and a small open-addressing hash  column id -> slot  live
// in shared memory (1-2 probes per lookup; a binary search over the sorted columns cost 6
// dependent shared loads per contribution and made the kernel shared-memory bound,
// profiles/r01_notes.md), its lanes take the local columns of one adjacency entry per pass.
// Entries are walked in ascending (batch, table row, local index) order, so every nonzero is summed
// in exactly the order of a sequential AddElementMatrix loop.  The per-row fixed cost is shared by
// NG rows and no pass ever waits for another group.
template <bool DUPS>
__global__ void __launch_bounds__(256) k_scatter(int ndof, const long long* __restrict__ adj_ptr,
                                                 const long long* __restrict__ adj, BatchDev bd, int RCAP, int HS, int GS,
                                                 int NG, int trips, const long long* __restrict__ rowptr,
                                                 const int32_t* __restrict__ colidx, double* __restrict__ vals, int small_nr,
                                                 int big_nr, const long long* __restrict__ src_off, int32_t* __restrict__ col_out)
{
  extern __shared__ double sacc[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
  const int g = lane / GS, ls = lane - g * GS;
  const bool live = g < NG;
  const int slot_id = wib * NG + (live ? g : 0);
  double* acc = sacc + slot_id * RCAP;
  int* hkey = reinterpret_cast<int*>(sacc + nwb * NG * RCAP) + slot_id * HS;
  unsigned short* hpos = reinterpret_cast<unsigned short*>(reinterpret_cast<int*>(sacc + nwb * NG * RCAP) + nwb * NG * HS) + slot_id * HS;
  const int rows_per_sweep = gridDim.x * nwb * NG;
  for (int d0 = (blockIdx.x * nwb + wib) * NG; d0 < ndof; d0 += rows_per_sweep) {
    const int d = d0 + g;
    const bool have = live && d < ndof;
    long long r0 = 0, a = 0;
    int nr = 0, na = 0;
    if (have) {
      r0 = rowptr[d]; nr = (int)(rowptr[d + 1] - r0);
      a = adj_ptr[d]; na = (int)(adj_ptr[d + 1] - a);
      if (nr <= small_nr || nr > big_nr) { nr = 0; na = 0; }  // empty rows, the short rows of k_scatter_small, other size classes
    }
    // hash size for this row: power of two >= 1.5 * nr (<= HS by the host's sizing)
    int H = 16, shift = 28;
    while (H < nr + (nr >> 1)) { H <<= 1; shift--; }
    for (int k = ls; k < (have ? H : 0); k += GS) hkey[k] = XG_EMPTY;
    __syncwarp();
    {
      // fused compaction (xgpu_csr_assemble): the row's columns come from the pattern scratch at its candidate
      // offset and are copied to their final place here
      const int32_t* __restrict__ cr = src_off ? colidx + (have ? src_off[d] : 0) : colidx + r0;
      for (int k = ls; k < nr; k += GS) {
        const int c = __ldg(cr + k);
        if (col_out) col_out[r0 + k] = c;
        acc[k] = 0.0;
        unsigned h = ((unsigned)c * 2654435761u) >> shift;
        while (atomicCAS(&hkey[h], XG_EMPTY, c) != XG_EMPTY) h = (h + 1) & (H - 1);  // columns of a row are distinct
        hpos[h] = (unsigned short)k;
      }
    }
    __syncwarp();
    const long long* __restrict__ adjd = adj + a;
    const int maxna = __reduce_max_sync(0xffffffffu, na);
    // The passes of a row are independent loads followed by ordered shared-memory adds.  To keep
    // the memory latency off the critical path the entries are taken in batches of XG_UB: the
    // adjacency entries of the NEXT batch are prefetched by the first lanes of the group, the table
    // and matrix rows of the whole batch are loaded before the first add.
    const int UB = GS < XG_UB ? GS : XG_UB;
    const int src0 = live ? g * GS : lane;
    long long e_pref = (ls < UB && ls < na) ? adjd[ls] : 0;
    for (int t0 = 0; t0 < maxna; t0 += UB) {
      const long long e_cur = e_pref;
      e_pref = (ls < UB && t0 + UB + ls < na) ? adjd[t0 + UB + ls] : 0;
      for (int it = 0; it < trips; it++) {
        const int lj = it * GS + ls;
        int cc[XG_UB];
        double mm[XG_UB];
#pragma unroll
        for (int u = 0; u < XG_UB; u++) {
          const long long e = __shfl_sync(0xffffffffu, e_cur, live ? src0 + (u < UB ? u : 0) : lane);
          cc[u] = -1; mm[u] = 0.0;
          if (u < UB && t0 + u < na) {
            const int bi = XG_ADJ_BATCH(e);
            if (bd.mats[bi]) {  // a batch without matrices only contributes sparsity (explicit zeros)
              const int w = bd.width[bi];
              if (lj < w) {
                const int ro = XG_ADJ_ROW(e) * w;  // nrows*width < 2^31 (checked on the host)
                cc[u] = __ldg(bd.table[bi] + ro + lj);
                mm[u] = __ldg(bd.mats[bi] + (size_t)(ro + XG_ADJ_LOC(e)) * w + lj);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < XG_UB; u++) {
          if (u < UB) {
            const int c = cc[u];
            const double m = mm[u];
            int pos = -1;
            if (c >= 0) {
              unsigned h = ((unsigned)c * 2654435761u) >> shift;
              int key = hkey[h];
              while (key != c && key != XG_EMPTY) { h = (h + 1) & (H - 1); key = hkey[h]; }
              pos = (key == c) ? (int)hpos[h] : -1;
            }
            if (DUPS) {
              // a dof may occur twice in one table row: lanes of a group hitting the same slot are
              // summed in lane (= local index) order and the first lane adds
              const unsigned peers = __match_any_sync(0xffffffffu, pos >= 0 ? (pos | (g << 20)) : (-1 - lane));
              unsigned rem = (pos >= 0) ? peers : 0u;
              double sum = 0.0;
              while (__any_sync(0xffffffffu, rem != 0u)) {
                const int src = rem ? (__ffs(rem) - 1) : lane;
                const double v = __shfl_sync(0xffffffffu, m, src);
                if (rem) { sum += v; rem &= rem - 1; }
              }
              if (pos >= 0 && lane == __ffs(peers) - 1) acc[pos] += sum;
            } else {
              if (pos >= 0) acc[pos] += m;  // distinct dofs inside one table row: no conflict in a group
            }
            __syncwarp();
          }
        }
      }
    }
    {
      double* __restrict__ vr = vals + r0;
      for (int k = ls; k < nr; k += GS) vr[k] = acc[k];
    }
    __syncwarp();
  }
}

// ---- short rows (at most 32 nonzeros): one THREAD per row.  The row's sorted columns and FP64
// accumulators sit in the thread's private, bank-conflict-free column of shared memory
// (slot k of thread t at [k * blockDim + t]); every contribution finds its slot with a branch-free
// 5-step binary search and is added in (batch, table row, local index) order — the same order as
// the lane-group kernel above, without any per-row warp bookkeeping.
#define XG_SMALL_NR 32
__global__ void __launch_bounds__(XG_SMALL_BLOCK) k_scatter_small(int ndof, const long long* __restrict__ adj_ptr,
                                                                  const long long* __restrict__ adj, BatchDev bd,
                                                                  const long long* __restrict__ rowptr,
                                                                  const int32_t* __restrict__ colidx, double* __restrict__ vals,
                                                                  const long long* __restrict__ src_off,
                                                                  int32_t* __restrict__ col_out)
{
  extern __shared__ double ssm[];
  double* acc = ssm + threadIdx.x;
  int* cols = reinterpret_cast<int*>(ssm + XG_SMALL_NR * XG_SMALL_BLOCK) + threadIdx.x;
#define ACC(k) acc[(k) * XG_SMALL_BLOCK]
#define COL(k) cols[(k) * XG_SMALL_BLOCK]
  for (int d = blockIdx.x * XG_SMALL_BLOCK + threadIdx.x; d < ndof; d += gridDim.x * XG_SMALL_BLOCK) {
    const long long r0 = rowptr[d];
    const int nr = (int)(rowptr[d + 1] - r0);
    if (nr == 0 || nr > XG_SMALL_NR) continue;
    const int32_t* __restrict__ cr = src_off ? colidx + src_off[d] : colidx + r0;  // fused compaction, see k_scatter
#pragma unroll
    for (int k = 0; k < XG_SMALL_NR; k++) {
      COL(k) = (k < nr) ? __ldg(cr + k) : XG_EMPTY;
      ACC(k) = 0.0;
    }
    const long long a = adj_ptr[d];
    const int na = (int)(adj_ptr[d + 1] - a);
    for (int t = 0; t < na; t++) {
      const long long e = adj[a + t];
      const int bi = XG_ADJ_BATCH(e);
      const double* __restrict__ mb = bd.mats[bi];
      if (!mb) continue;  // sparsity-only batch
      const int w = bd.width[bi];
      const int ro = XG_ADJ_ROW(e) * w;
      const int32_t* __restrict__ trow = bd.table[bi] + ro;
      const double* __restrict__ mrow = mb + (size_t)(ro + XG_ADJ_LOC(e)) * w;
      for (int j0 = 0; j0 < w; j0 += XG_SMALL_CH) {
        int cc[XG_SMALL_CH];
        double mm[XG_SMALL_CH];  // table and matrix entries of the chunk are in flight together
        if ((w & 1) == 0) {      // even width: 8-byte aligned table rows, 16-byte aligned matrix rows
#pragma unroll
          for (int u = 0; u < XG_SMALL_CH; u += 2) {
            int2 v = make_int2(-1, -1);
            double2 m2 = make_double2(0.0, 0.0);
            if (j0 + u < w) {
              v = __ldg(reinterpret_cast<const int2*>(trow + j0 + u));
              m2 = __ldg(reinterpret_cast<const double2*>(mrow + j0 + u));
            }
            cc[u] = v.x; cc[u + 1] = v.y; mm[u] = m2.x; mm[u + 1] = m2.y;
          }
        } else {
#pragma unroll
          for (int u = 0; u < XG_SMALL_CH; u++) {
            const bool in = j0 + u < w;
            cc[u] = in ? __ldg(trow + j0 + u) : -1;
            mm[u] = in ? __ldg(mrow + j0 + u) : 0.0;
          }
        }
#pragma unroll
        for (int u = 0; u < XG_SMALL_CH; u++) {
          const int c = cc[u];
          if (c >= 0) {
            int pos = (COL(16) <= c) ? 16 : 0;
            pos += (COL(pos + 8) <= c) ? 8 : 0;
            pos += (COL(pos + 4) <= c) ? 4 : 0;
            pos += (COL(pos + 2) <= c) ? 2 : 0;
            pos += (COL(pos + 1) <= c) ? 1 : 0;
            ACC(pos) += mm[u];
          }
        }
      }
    }
    for (int k = 0; k < nr; k++) vals[r0 + k] = ACC(k);
    if (col_out) for (int k = 0; k < nr; k++) col_out[r0 + k] = COL(k);
  }
#undef ACC
#undef COL
}

// LinearForm assembly: thread per dof walks its (sorted) adjacency segment
__global__ void k_vec_scatter(int64_t ndof, const long long* __restrict__ adj_ptr, const long long* __restrict__ adj,
                              BatchDev bd, double* __restrict__ out)
{
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= ndof) return;
  double s = 0.0;
  for (long long i = adj_ptr[d]; i < adj_ptr[d + 1]; i++) {
    const long long e = adj[i];
    const int bi = XG_ADJ_BATCH(e);
    if (bd.mats[bi]) s += __ldg(bd.mats[bi] + (size_t)XG_ADJ_ROW(e) * bd.width[bi] + XG_ADJ_LOC(e));
  }
  out[d] = s;
}

// the mean instead of the sum (GridFunction.Set / ProjectShift average the element-local projections over the elements
// sharing a dof, lsetcurving/projshift.cpp:165-197); out has `stride` doubles per dof, component `comp` is written
__global__ void k_vec_scatter_mean(int64_t ndof, const long long* __restrict__ adj_ptr, const long long* __restrict__ adj,
                                   BatchDev bd, double* __restrict__ out, int stride, int comp)
{
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= ndof) return;
  double s = 0.0;
  int cnt = 0;
  for (long long i = adj_ptr[d]; i < adj_ptr[d + 1]; i++) {
    const long long e = adj[i];
    const int bi = XG_ADJ_BATCH(e);
    if (bd.mats[bi]) { s += __ldg(bd.mats[bi] + (size_t)XG_ADJ_ROW(e) * bd.width[bi] + XG_ADJ_LOC(e)); cnt++; }
  }
  out[d * stride + comp] = cnt > 0 ? s * (1.0 / (double)cnt) : 0.0;
}

__global__ void k_spmv(int64_t nrows, const long long* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                       const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y)
{
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nrows) return;
  double s = 0.0;
  for (long long k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) s += vals[k] * __ldg(x + colidx[k]);
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if (lane == 0) y[r] = s;
}

int fill_bd(xgpu_ctx* ctx, const xgpu_batch* bh, int nb, BatchDev& bd)
{
  if (nb < 1 || nb > XG_MAXBATCH) return xgpu_fail(ctx, XGPU_ERR_ARG, "csr: 1..8 batches");
  bd.nb = nb;
  bd.rowoff[0] = 0;
  for (int i = 0; i < nb; i++) {
    if (bh[i].width < 1 || bh[i].width > 255 || bh[i].nrows < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "csr: bad batch");
    if (bh[i].nrows * bh[i].width >= (int64_t)0x7fffffff) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "csr: a batch has 2^31 or more table entries; split it");
    bd.table[i] = bh[i].table; bd.mats[i] = bh[i].mats; bd.coefs[i] = nullptr; bd.width[i] = bh[i].width;
    bd.rowoff[i + 1] = bd.rowoff[i] + bh[i].nrows;
  }
  return XGPU_OK;
}

}  // namespace

extern "C" void xgpu_plan_destroy(xgpu_plan* p)
{
  if (!p) return;
  if (p->adj_ptr) cudaFreeAsync(p->adj_ptr, p->stream);
  if (p->adj) cudaFreeAsync(p->adj, p->stream);
  if (p->rowptr) cudaFreeAsync(p->rowptr, p->stream);
  if (p->candoff) cudaFreeAsync(p->candoff, p->stream);
  if (p->tmp32) cudaFreeAsync(p->tmp32, p->stream);
  if (p->rowcnt) cudaFreeAsync(p->rowcnt, p->stream);
  if (p->pat_small) cudaFreeAsync(p->pat_small, p->stream);
  if (p->pat_big) cudaFreeAsync(p->pat_big, p->stream);
  if (p->bigslot) cudaFreeAsync(p->bigslot, p->stream);
  for (int i = 0; i < 3; i++) if (p->ev[i]) cudaEventDestroy(p->ev[i]);
  delete p;
}

namespace {

// lanes of a warp are split into G groups of SL slots: group g handles adjacency entry t0+g,
// slot ls the local column lj = ls, ls+SL, ... (trips iterations) of that entry's table row
void lane_split(int wmax, int* SL, int* G, int* trips)
{
  *SL = wmax <= 32 ? (wmax < 1 ? 1 : wmax) : 32;
  *G = 32 / *SL;
  *trips = (wmax + *SL - 1) / *SL;
}

// the pattern pass: fills the shared scratch (slot 8) with the sorted unique columns of every row
// at the row's candidate offset, the unique counts in tmp32, the longest row in stats[2]
int run_pattern(xgpu_ctx* ctx, xgpu_plan* P, int* stats)
{
  void* sp;
  int rc = xgpu_scratch(ctx, 8, sizeof(int32_t) * (size_t)(P->ncand > 0 ? P->ncand : 1), &sp);
  if (rc) return rc;
  int SL, G, trips;
  lane_split(P->wmax, &SL, &G, &trips);
  {  // short rows first: thread per row
    unsigned gs = nblk(P->ndof, XG_SMALL_BLOCK);
    if (gs > XGPU_NSM * 64) gs = XGPU_NSM * 64;
    if (P->dups) k_pattern_small<true><<<gs, XG_SMALL_BLOCK, 0, ctx->stream>>>((int)P->ndof, P->adj_ptr, P->adj, P->bd, P->candoff, (int32_t*)sp, P->tmp32, stats);
    else k_pattern_small<false><<<gs, XG_SMALL_BLOCK, 0, ctx->stream>>>((int)P->ndof, P->adj_ptr, P->adj, P->bd, P->candoff, (int32_t*)sp, P->tmp32, stats);
    XGPU_LAUNCH_CHECK(ctx);
  }
  // warp-per-row classes by candidate count.  The hash table of a warp is sized for the LONGEST row of its class, and
  // the shared memory per warp decides the occupancy: with facet-patch batches a few rows have thousands of candidates
  // (config 3: ~5000) — sizing every warp for them leaves 2 warps per SM.  Rows up to 1023 candidates get their own
  // launch with 2048-slot tables (8 warps per block, several blocks per SM), only the rest pays for the big tables.
  const int MID = 1023;
  int lo = XG_SMALL_CAND;
  while (P->maxcand > lo) {
    const int hi = (P->maxcand > MID && lo < MID) ? MID : P->maxcand;
    int HCAP = 64;
    while (HCAP < 2 * hi) HCAP <<= 1;
    int nw = 8;
    while (nw > 1 && (size_t)nw * (HCAP + 4) * sizeof(int) > 200 * 1024) nw >>= 1;
    const size_t sm = (size_t)nw * (HCAP + 4) * sizeof(int);
    if (sm > 220 * 1024) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "csr: a row has too many candidate columns");
    unsigned grid = nblk(P->ndof, nw);
    if (grid > XGPU_NSM * 32) grid = XGPU_NSM * 32;
    if (P->dups) {
      XGPU_CUDA(ctx, cudaFuncSetAttribute(k_pattern<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_pattern<true><<<grid, nw * 32, sm, ctx->stream>>>((int)P->ndof, P->adj_ptr, P->adj, P->bd, HCAP, SL, G, P->candoff,
                                                          (int32_t*)sp, P->tmp32, stats, lo, hi);
    } else {
      XGPU_CUDA(ctx, cudaFuncSetAttribute(k_pattern<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_pattern<false><<<grid, nw * 32, sm, ctx->stream>>>((int)P->ndof, P->adj_ptr, P->adj, P->bd, HCAP, SL, G, P->candoff,
                                                           (int32_t*)sp, P->tmp32, stats, lo, hi);
    }
    XGPU_LAUNCH_CHECK(ctx);
    lo = hi;
  }
  P->epoch = ++ctx->epoch;
  return XGPU_OK;
}

}  // namespace

static int build_plan(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                      xgpu_plan** plan_out, int64_t* nnz_host, bool with_pattern);

extern "C" int xgpu_csr_plan(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                             xgpu_plan** plan_out, int64_t* nnz_host)
{
  XGPU_NVTX(ctx, "xgpu_csr_plan");
  return build_plan(ctx, ndof, batches_host, nbatches, plan_out, nnz_host, true);
}

int xgpu_build_adjacency(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches, xgpu_plan** plan_out)
{
  return build_plan(ctx, ndof, batches_host, nbatches, plan_out, nullptr, false);
}

extern "C" int xgpu_adjacency_plan(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                                   xgpu_plan** plan_out)
{
  return build_plan(ctx, ndof, batches_host, nbatches, plan_out, nullptr, false);
}

// second half of the transposed table: sort every segment, candidate counts (tmp32), statistics (max candidates,
// duplicate dofs inside a table row), optionally the candidate offsets of the two-pass pattern scratch.  Synchronises.
int xg_plan_finish(xgpu_ctx* ctx, xgpu_plan* P, bool want_candoff)
{
  cudaStream_t st = ctx->stream;
  const int64_t ndof = P->ndof;
  void* sp;
  int rc = xgpu_scratch(ctx, 6, 64, &sp);
  if (rc) return rc;
  int* stats = (int*)sp;  // [0] max candidates of a row, [1] duplicate dofs inside a table row, [2] longest row
  XGPU_CUDA(ctx, cudaMemsetAsync(stats, 0, 4 * sizeof(int), st));
  if (ndof > 0) {
    k_adj_sort<<<nblk(ndof, 128), 128, 0, st>>>(ndof, P->adj_ptr, P->adj, P->bd, P->tmp32, stats);
    XGPU_LAUNCH_CHECK(ctx);
  }
  if (want_candoff) {
    rc = xgpu_exclusive_scan_i32(ctx, P->tmp32, ndof, (int64_t*)P->candoff);
    if (rc) return rc;
  }
  XGPU_CUDA(ctx, cudaEventRecord(P->ev[1], st));
  XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, stats, 4 * sizeof(int), st));
  if (want_candoff) XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned + 2, P->candoff + ndof, sizeof(long long), st));
  XGPU_CUDA(ctx, cudaStreamSynchronize(st));
  P->maxcand = ((int*)ctx->h_pinned)[0];
  P->dups = ((int*)ctx->h_pinned)[1];
  P->nbig = ((int*)ctx->h_pinned)[3];
  P->ncand = want_candoff ? ctx->h_pinned[2] : -1;
  return XGPU_OK;
}

static int build_plan(xgpu_ctx* ctx, int64_t ndof, const xgpu_batch* batches_host, int nbatches,
                      xgpu_plan** plan_out, int64_t* nnz_host, bool with_pattern)
{
  if (!ctx || !plan_out || ndof < 0 || ndof >= (int64_t)0x7fffffff)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_plan: bad argument (0 <= ndof < 2^31: column indices are int32)");
  *plan_out = nullptr;
  xgpu_plan* P = new xgpu_plan();
  ctx->plans_created++;
  int rc = fill_bd(ctx, batches_host, nbatches, P->bd);
  if (rc) { delete P; return rc; }
  P->ndof = ndof;
  for (int i = 0; i < nbatches; i++) P->wmax = std::max(P->wmax, batches_host[i].width);
  cudaStream_t st = ctx->stream;
  P->stream = st;
#define XG_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { xgpu_plan_destroy(P); return xgpu_fail(ctx, XGPU_ERR_CUDA, cudaGetErrorString(e_)); } } while (0)
#define XG_TRYRC(call) do { int r_ = (call); if (r_) { xgpu_plan_destroy(P); return r_; } } while (0)
  for (int i = 0; i < 3; i++) XG_TRY(cudaEventCreate(&P->ev[i]));
  XG_TRY(cudaEventRecord(P->ev[0], st));
  XG_TRY(cudaMallocAsync(&P->adj_ptr, sizeof(long long) * (ndof + 1), st));
  XG_TRY(cudaMallocAsync(&P->rowptr, sizeof(long long) * (ndof + 1), st));
  XG_TRY(cudaMallocAsync(&P->candoff, sizeof(long long) * (ndof + 1), st));
  XG_TRY(cudaMallocAsync(&P->tmp32, sizeof(int32_t) * (ndof + 2), st));
  XG_TRY(cudaMemsetAsync(P->tmp32, 0, sizeof(int32_t) * (ndof + 2), st));
  for (int i = 0; i < nbatches; i++) {
    const int64_t nent = batches_host[i].nrows * batches_host[i].width;
    if (nent == 0) continue;
    k_adj_count<<<nblk(nent, 256), 256, 0, st>>>(batches_host[i].table, nent, P->tmp32);
    ctx->launches++;
  }
  XG_TRYRC(xgpu_exclusive_scan_i32(ctx, P->tmp32, ndof, (int64_t*)P->adj_ptr));
  XG_TRY(xgpu_d2h_small(ctx->h_pinned, P->adj_ptr + ndof, sizeof(long long), st));
  XG_TRY(cudaStreamSynchronize(st));
  P->nadj = ctx->h_pinned[0];
  XG_TRY(cudaMallocAsync(&P->adj, sizeof(long long) * (P->nadj > 0 ? P->nadj : 1), st));
  XG_TRY(cudaMemsetAsync(P->tmp32, 0, sizeof(int32_t) * (ndof + 2), st));
  for (int i = 0; i < nbatches; i++) {
    const int64_t nent = batches_host[i].nrows * batches_host[i].width;
    if (nent == 0) continue;
    k_adj_fill<<<nblk(nent, 256), 256, 0, st>>>(batches_host[i].table, nent, batches_host[i].width, i,
                                                 P->adj_ptr, P->tmp32, P->adj);
    ctx->launches++;
  }
  XG_TRYRC(xg_plan_finish(ctx, P, true));
  void* sp;
  XG_TRYRC(xgpu_scratch(ctx, 6, 64, &sp));
  int* stats = (int*)sp;
  P->with_pattern = with_pattern;
  if (!with_pattern) {  // transposed table only (vector assembly)
    XG_TRY(cudaEventRecord(P->ev[2], st));
    *plan_out = P;
    return XGPU_OK;
  }
  if (ndof > 0) XG_TRYRC(run_pattern(ctx, P, stats));
  XG_TRYRC(xgpu_exclusive_scan_i32(ctx, P->tmp32, ndof, (int64_t*)P->rowptr));
  XG_TRY(cudaEventRecord(P->ev[2], st));
  XG_TRY(xgpu_d2h_small(ctx->h_pinned, P->rowptr + ndof, sizeof(long long), st));
  XG_TRY(xgpu_d2h_small(ctx->h_pinned + 1, stats + 2, sizeof(int), st));
  XG_TRY(cudaStreamSynchronize(st));
  P->nnz = ctx->h_pinned[0];
  P->maxnr = ((int*)(ctx->h_pinned + 1))[0];
  if (nnz_host) *nnz_host = P->nnz;
  *plan_out = P;
  return XGPU_OK;
#undef XG_TRY
#undef XG_TRYRC
}

extern "C" int xgpu_plan_phase_ms(xgpu_plan* P, double* ms_host)
{
  if (!P || !ms_host || !P->ev[2]) return XGPU_ERR_ARG;
  float a = 0.f, b = 0.f;
  if (cudaEventElapsedTime(&a, P->ev[0], P->ev[1]) != cudaSuccess) return XGPU_ERR_CUDA;
  if (cudaEventElapsedTime(&b, P->ev[1], P->ev[2]) != cudaSuccess) return XGPU_ERR_CUDA;
  ms_host[0] = a; ms_host[1] = b;
  return XGPU_OK;
}

extern "C" int xgpu_csr_pattern(xgpu_ctx* ctx, xgpu_plan* P, int64_t* rowptr, int32_t* colidx)
{
  if (!ctx || !P || !P->with_pattern) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_pattern: bad argument (adjacency-only plan?)");
  cudaStream_t st = ctx->stream;
  XGPU_CUDA(ctx, cudaMemcpyAsync(rowptr, P->rowptr, sizeof(long long) * (P->ndof + 1), cudaMemcpyDeviceToDevice, st));
  if (P->ndof == 0 || P->nnz == 0) return XGPU_OK;
  if (P->epoch != ctx->epoch) {  // another plan has used the shared scratch since: redo the pass
    void* sp;
    int rc = xgpu_scratch(ctx, 6, 64, &sp);
    if (rc) return rc;
    rc = run_pattern(ctx, P, (int*)sp);
    if (rc) return rc;
  }
  unsigned grid = nblk(P->ndof, 8);
  if (grid > XGPU_NSM * 32) grid = XGPU_NSM * 32;
  k_compact<<<grid, 256, 0, st>>>(P->ndof, P->candoff, P->rowptr, (const int32_t*)ctx->scratch[8], colidx);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

static int run_scatter(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                       const int64_t* rowptr, const int32_t* colidx, double* vals, const long long* src_off, int32_t* col_out);

extern "C" int xgpu_csr_scatter(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                                const int64_t* rowptr, const int32_t* colidx, double* vals)
{
  return run_scatter(ctx, P, batches_host, nbatches, rowptr, colidx, vals, nullptr, nullptr);
}

extern "C" int xgpu_csr_assemble(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                                 int64_t* rowptr, int32_t* colidx, double* vals)
{
  XGPU_NVTX(ctx, "xgpu_csr_assemble");
  if (!ctx || !P || !P->with_pattern) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_assemble: bad argument");
  XGPU_CUDA(ctx, cudaMemcpyAsync(rowptr, P->rowptr, sizeof(long long) * (P->ndof + 1), cudaMemcpyDeviceToDevice, ctx->stream));
  if (P->ndof == 0 || P->nnz == 0) return XGPU_OK;
  if (P->epoch != ctx->epoch) {  // another plan has used the shared scratch since: redo the pass
    void* sp;
    int rc = xgpu_scratch(ctx, 6, 64, &sp);
    if (rc) return rc;
    rc = run_pattern(ctx, P, (int*)sp);
    if (rc) return rc;
  }
  return run_scatter(ctx, P, batches_host, nbatches, (const int64_t*)P->rowptr, (const int32_t*)ctx->scratch[8], vals,
                     P->candoff, colidx);
}

static int run_scatter(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches,
                       const int64_t* rowptr, const int32_t* colidx, double* vals, const long long* src_off, int32_t* col_out)
{
  if (!ctx || !P || nbatches != P->bd.nb || !P->with_pattern) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_scatter: bad argument");
  BatchDev bd = P->bd;
  for (int i = 0; i < nbatches; i++) {
    if (batches_host[i].nrows != bd.rowoff[i + 1] - bd.rowoff[i] || batches_host[i].width != bd.width[i])
      return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_csr_scatter: batches do not match the plan");
    bd.table[i] = batches_host[i].table;
    bd.mats[i] = batches_host[i].mats;  // NULL: the batch only contributes sparsity
  }
  if (P->ndof == 0 || P->nnz == 0) return XGPU_OK;
  int SL, G, trips;
  lane_split(P->wmax, &SL, &G, &trips);
  {  // short rows: thread per row
    unsigned gs = nblk(P->ndof, XG_SMALL_BLOCK);
    if (gs > XGPU_NSM * 64) gs = XGPU_NSM * 64;
    const size_t sms = (size_t)XG_SMALL_NR * XG_SMALL_BLOCK * (sizeof(double) + sizeof(int));
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_scatter_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sms));
    k_scatter_small<<<gs, XG_SMALL_BLOCK, sms, ctx->stream>>>((int)P->ndof, P->adj_ptr, P->adj, bd, (const long long*)rowptr, colidx, vals,
                                                              src_off, col_out);
    XGPU_LAUNCH_CHECK(ctx);
  }
  // warp-group-per-row classes by row length (the accumulators + slot hash of a group are sized for the longest row of
  // the class; see run_pattern): rows up to 255 entries, then the rest
  const int MID = 255;
  int lo = XG_SMALL_NR;
  while (P->maxnr > lo) {
    const int hi = (P->maxnr > MID && lo < MID) ? MID : P->maxnr;
    const int RCAP = std::max(32, (hi + 31) & ~31);
    int HS = 16;
    while (HS < hi + (hi >> 1)) HS <<= 1;
    if (HS > 65536) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "csr: a row is too long for the shared-memory gather");
    const size_t per_warp = ((size_t)RCAP * 8 + (size_t)HS * 6) * G;
    int nw = 8;
    while (nw > 1 && nw * per_warp > 200 * 1024) nw >>= 1;
    const size_t sm = nw * per_warp;
    if (sm > 220 * 1024) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "csr: a row is too long for the shared-memory gather");
    unsigned grid = nblk(P->ndof, nw * G);
    if (grid > XGPU_NSM * 32) grid = XGPU_NSM * 32;
    if (P->dups) {
      XGPU_CUDA(ctx, cudaFuncSetAttribute(k_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_scatter<true><<<grid, nw * 32, sm, ctx->stream>>>((int)P->ndof, P->adj_ptr, P->adj, bd, RCAP, HS, SL, G, trips,
                                                          (const long long*)rowptr, colidx, vals, lo, hi, src_off, col_out);
    } else {
      XGPU_CUDA(ctx, cudaFuncSetAttribute(k_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_scatter<false><<<grid, nw * 32, sm, ctx->stream>>>((int)P->ndof, P->adj_ptr, P->adj, bd, RCAP, HS, SL, G, trips,
                                                           (const long long*)rowptr, colidx, vals, lo, hi, src_off, col_out);
    }
    XGPU_LAUNCH_CHECK(ctx);
    lo = hi;
  }
  return XGPU_OK;
}

extern "C" int xgpu_vec_scatter(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches, double* out)
{
  if (!ctx || !P || nbatches != P->bd.nb) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_vec_scatter: bad argument");
  BatchDev bd = P->bd;
  for (int i = 0; i < nbatches; i++) {
    if (batches_host[i].nrows != bd.rowoff[i + 1] - bd.rowoff[i] || batches_host[i].width != bd.width[i])
      return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_vec_scatter: batches do not match the plan");
    bd.mats[i] = batches_host[i].mats;
  }
  if (P->ndof == 0) return XGPU_OK;
  k_vec_scatter<<<nblk(P->ndof, 256), 256, 0, ctx->stream>>>(P->ndof, P->adj_ptr, P->adj, bd, out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_vec_scatter_mean(xgpu_ctx* ctx, xgpu_plan* P, const xgpu_batch* batches_host, int nbatches, double* out,
                                     int stride, int comp)
{
  if (!ctx || !P || nbatches != P->bd.nb || stride < 1 || comp < 0 || comp >= stride)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_vec_scatter_mean: bad argument");
  BatchDev bd = P->bd;
  for (int i = 0; i < nbatches; i++) {
    if (batches_host[i].nrows != bd.rowoff[i + 1] - bd.rowoff[i] || batches_host[i].width != bd.width[i])
      return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_vec_scatter_mean: batches do not match the plan");
    bd.mats[i] = batches_host[i].mats;
  }
  if (P->ndof == 0) return XGPU_OK;
  k_vec_scatter_mean<<<nblk(P->ndof, 256), 256, 0, ctx->stream>>>(P->ndof, P->adj_ptr, P->adj, bd, out, stride, comp);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_csr_spmv(xgpu_ctx* ctx, int64_t nrows, const int64_t* rowptr, const int32_t* colidx,
                             const double* vals, const double* x, double* y)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (nrows == 0) return XGPU_OK;
  k_spmv<<<nblk(nrows * 32, 256), 256, 0, ctx->stream>>>(nrows, (const long long*)rowptr, colidx, vals, x, y);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}
