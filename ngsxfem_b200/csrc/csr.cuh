#pragma once
// csr.cuh — plan / batch structures shared by csr.cu (transposed table, two-pass pattern + scatter) and
// csr_fused.cu (one row pass: pattern + values + procedural affine batches)
#include "common.cuh"

#define XG_MAXBATCH 8
#define XG_EMPTY 0x7fffffff
#define XG_UB 4  // adjacency entries per prefetch batch

struct BatchDev {
  int nb;
  const int32_t* table[XG_MAXBATCH];
  const double* mats[XG_MAXBATCH];
  const double* coefs[XG_MAXBATCH];  // procedural affine batch: nrows x 8 coefficients (mats NULL)
  long long rowoff[XG_MAXBATCH + 1];
  int width[XG_MAXBATCH];
};

struct xgpu_plan {
  int64_t ndof = 0, nadj = 0, nnz = 0, ncand = 0;
  int maxcand = 0, maxnr = 0, wmax = 0, dups = 0;
  int64_t epoch = -1;             // ctx->epoch at which the pattern scratch (slot 8) was filled
  bool with_pattern = false;      // false: transposed table only (xgpu_adjacency_plan)
  int small_na = 0;               // counted plans: rows with 1..small_na adjacent table rows went through the thread-per-row kernels
  int warp_cand = 0;              // counted plans: the other rows with at most warp_cand candidates went through the warp-per-row kernels
  bool counted = false;           // xgpu_csr_plan_rows: rowptr / nnz / maxnr from the count pass (no pattern scratch)
  int32_t* rowcnt = nullptr;      // counted plans: distinct columns per row
  // stored pattern of counted plans (csr_rows_small.cu / csr_rows_warp.cu): the values pass needs no sort
  unsigned char* pat_small = nullptr;  // 192 B records of the thread-per-row kernel, interleaved per 32 rows
  unsigned char* pat_big = nullptr;    // 1536 B records of the warp-per-row kernel
  int32_t* bigslot = nullptr;          // ndof: record number of a warp-per-row row
  int64_t nbig_cap = 0;                // records allocated in pat_big
  int nbig = 0;                        // rows with more than 64 candidates (k_adj_sort)
  int small_overflow = 0;              // a thread-per-row row has more than 32 nonzeros: the lane-group kernel assembles those
  BatchDev bd;
  long long* adj_ptr = nullptr;   // ndof+1
  long long* adj = nullptr;       // nadj entries: (global table row << 8) | local index
  long long* rowptr = nullptr;    // ndof+1
  long long* candoff = nullptr;   // ndof+1: offsets of the rows in the candidate scratch
  int32_t* tmp32 = nullptr;       // ndof+2 counters
  cudaStream_t stream = 0;        // the buffers are stream-ordered allocations on this stream
  cudaEvent_t ev[3] = {};         // start / adjacency built / pattern pass done (xgpu_plan_phase_ms)
};

// adjacency entry: (batch << 56) | (local table row << 8) | local index — ascending order is
// (batch, row, local index), the summation order of the scatter
#define XG_ADJ(bi, r, loc) (((long long)(bi) << 56) | ((long long)(r) << 8) | (long long)(loc))
#define XG_ADJ_BATCH(e) ((int)((unsigned long long)(e) >> 56))
#define XG_ADJ_ROW(e) ((int)(unsigned)((unsigned long long)(e) >> 8))  // rows < 2^31 per batch
#define XG_ADJ_LOC(e) ((int)((e) & 255))
#define XG_ADJ_ROWKEY(e) ((e) >> 8)  // (batch, row)


// csr_rows_small.cu: thread-per-row kernels (pattern pass / values pass) for rows with at most xg_rows_small_na(width)
// adjacent table rows
int xg_rows_small_na(int width, int64_t ndof);
int xg_rows_small_supports(int width, int ncoef);
size_t xg_rows_small_record_bytes(int64_t ndof);
int xg_rows_small_run(xgpu_ctx* ctx, xgpu_plan* P, const BatchDev& bd, const double* tref, int ncoef, bool pattern,
                      int32_t* rowcnt, int* stats, int32_t* colidx, double* vals, int64_t row_lo = 0,
                      int64_t row_hi = ((int64_t)1 << 40));   // values pass: rows [row_lo, row_hi) only

// csr_rows_warp.cu: warp-per-row kernels (sorting network over the warp) for the other rows with fewer than 256 candidates
int xg_rows_warp_ok(int width, int64_t ndof);
size_t xg_rows_warp_record_bytes(int64_t nrows);
int xg_rows_warp_run(xgpu_ctx* ctx, xgpu_plan* P, const BatchDev& bd, const double* tref, int ncoef, bool pattern, int small_na,
                     unsigned int* ticket, unsigned int* slotcount, int32_t* rowcnt, int* stats, int32_t* colidx, double* vals,
                     int64_t row_lo = 0, int64_t row_hi = ((int64_t)1 << 40));

// elmat.cu: is `tref` the library's own cached reference tensor of H1 order p on the dim-simplex (any integration order)?
bool xg_reftensor_is(xgpu_ctx* ctx, const double* tref, int dim, int p);

// csr.cu: sort the segments of a filled transposed table + statistics (synchronises)
int xg_plan_finish(xgpu_ctx* ctx, xgpu_plan* P, bool want_candoff);
