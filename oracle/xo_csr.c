/* xo_csr.c — ORACLE (test infrastructure): sparsity graph and element-matrix scatter.
 * Restates utils/restrictedblf.cpp:142-163 (GetGraph: table of dof rows of the active entities
 * handed to NGSolve's MatrixGraph = union of the row cliques, columns sorted ascending, explicit
 * zeros kept, full non-symmetric storage) and NGSolve's SparseMatrix::AddElementMatrix
 * (third party; binary search of the column in the sorted row).
 */
#include "xoracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int cmp_i32(const void* a, const void* b)
{
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* ascending sort of a short candidate list: binary insertion for short rows, qsort otherwise */
static void sort_i32(int32_t* a, int64_t n)
{
  if (n > 256) { qsort(a, n, sizeof(int32_t), cmp_i32); return; }
  for (int64_t i = 1; i < n; i++) {
    int32_t key = a[i];
    int64_t j = i - 1;
    while (j >= 0 && a[j] > key) { a[j + 1] = a[j]; j--; }
    a[j + 1] = key;
  }
}

int64_t xo_graph(int64_t ndof, int64_t nrows, int width, const int32_t* table,
                 int64_t* rowptr, int32_t* colidx)
{
  /* transposed table: dof -> table rows (ascending).  Histogram with atomic increments, then every thread fills the
   * segments of its own dof range from a scan over the table (segments stay in ascending row order: the result does not
   * depend on the thread count) */
  int64_t* cnt = (int64_t*)calloc(ndof + 1, sizeof(int64_t));
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nrows; r++)
    for (int k = 0; k < width; k++) {
      int32_t d = table[r * width + k];
      if (d >= 0) {
#pragma omp atomic
        cnt[d + 1]++;
      }
    }
  for (int64_t d = 0; d < ndof; d++) cnt[d + 1] += cnt[d];
  int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (ndof + 1));
  memcpy(pos, cnt, sizeof(int64_t) * (ndof + 1));
  int64_t* adj = (int64_t*)malloc(sizeof(int64_t) * (cnt[ndof] > 0 ? cnt[ndof] : 1));
#pragma omp parallel
  {
    int nt = 1, me = 0;
#ifdef _OPENMP
    nt = omp_get_num_threads(); me = omp_get_thread_num();
#endif
    const int64_t d0 = ndof * me / nt, d1 = ndof * (me + 1) / nt;
    for (int64_t r = 0; r < nrows; r++)
      for (int k = 0; k < width; k++) {
        const int32_t d = table[r * width + k];
        if (d >= d0 && d < d1) adj[pos[d]++] = r;
      }
  }
  /* rows are independent: pass 0 counts the unique columns of every row, pass 1 writes them
   * (the result does not depend on the thread count) */
  rowptr[0] = 0;
  for (int pass = 0; pass < (colidx ? 2 : 1); pass++) {
#pragma omp parallel
    {
      int32_t* buf = NULL;
      int64_t bufcap = 0;
#pragma omp for schedule(static)
      for (int64_t d = 0; d < ndof; d++) {
        int64_t need = (cnt[d + 1] - cnt[d]) * width;
        if (need > bufcap) { bufcap = need * 2 + 64; buf = (int32_t*)realloc(buf, sizeof(int32_t) * bufcap); }
        int64_t m = 0;
        for (int64_t a = cnt[d]; a < cnt[d + 1]; a++)
          for (int k = 0; k < width; k++) { int32_t c = table[adj[a] * width + k]; if (c >= 0) buf[m++] = c; }
        sort_i32(buf, m);
        int64_t u = 0;
        for (int64_t i = 0; i < m; i++)
          if (i == 0 || buf[i] != buf[i - 1]) { if (pass == 1) colidx[rowptr[d] + u] = buf[i]; u++; }
        if (pass == 0) rowptr[d + 1] = u;
      }
      free(buf);
    }
    if (pass == 0) for (int64_t d = 0; d < ndof; d++) rowptr[d + 1] += rowptr[d];
  }
  free(adj); free(pos); free(cnt);
  return rowptr[ndof];
}

/* the same graph with ONE sort per row (timing variant for the CPU baseline, as NGSolve's MatrixGraph builds its table once):
 * every thread takes a contiguous dof range with an equal share of the transposed table, appends the unique sorted columns of
 * its rows to its own buffer, and the buffers are copied behind each other after the scan of the row lengths.  Same result
 * as xo_graph for every thread count.  *colidx_out is malloc'ed (xo_free). */
int64_t xo_graph_once(int64_t ndof, int64_t nrows, int width, const int32_t* table, int64_t* rowptr, int32_t** colidx_out)
{
  int64_t* cnt = (int64_t*)calloc(ndof + 1, sizeof(int64_t));
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nrows; r++)
    for (int k = 0; k < width; k++) {
      int32_t d = table[r * width + k];
      if (d >= 0) {
#pragma omp atomic
        cnt[d + 1]++;
      }
    }
  for (int64_t d = 0; d < ndof; d++) cnt[d + 1] += cnt[d];
  int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (ndof + 1));
  memcpy(pos, cnt, sizeof(int64_t) * (ndof + 1));
  const int64_t nadj = cnt[ndof];
  int64_t* adj = (int64_t*)malloc(sizeof(int64_t) * (nadj > 0 ? nadj : 1));
  int ntmax = 1;
#ifdef _OPENMP
  ntmax = omp_get_max_threads();
#endif
  int32_t** tbuf = (int32_t**)calloc(ntmax, sizeof(int32_t*));
  int64_t* tlo = (int64_t*)calloc(ntmax + 1, sizeof(int64_t));
  int nt_used = 1;
  rowptr[0] = 0;
#pragma omp parallel num_threads(ntmax)
  {
    int nt = 1, me = 0;
#ifdef _OPENMP
    nt = omp_get_num_threads(); me = omp_get_thread_num();
#endif
    /* dof range of this thread: an equal share of the transposed table (binary search in its prefix sums) */
    int64_t bounds[2];
    for (int side = 0; side < 2; side++) {
      const int64_t target = nadj / nt * (me + side) + (me + side == nt ? nadj % nt : 0);
      int64_t lo = 0, hi = ndof;
      while (lo < hi) { int64_t mid = (lo + hi) / 2; if (cnt[mid] < target) lo = mid + 1; else hi = mid; }
      bounds[side] = (me + side == nt) ? ndof : lo;
    }
    const int64_t d0 = me == 0 ? 0 : bounds[0], d1 = bounds[1];
    if (me == 0) nt_used = nt;
    tlo[me] = d0;
    if (me == nt - 1) tlo[nt] = ndof;
    for (int64_t r = 0; r < nrows; r++)
      for (int k = 0; k < width; k++) {
        const int32_t d = table[r * width + k];
        if (d >= d0 && d < d1) adj[pos[d]++] = r;
      }
    int32_t *buf = NULL, *out = NULL;
    int64_t bufcap = 0, outcap = 0, nout = 0;
    for (int64_t d = d0; d < d1; d++) {
      const int64_t need = (cnt[d + 1] - cnt[d]) * width;
      if (need > bufcap) { bufcap = need * 2 + 64; buf = (int32_t*)realloc(buf, sizeof(int32_t) * bufcap); }
      if (nout + need > outcap) { outcap = (nout + need) * 2 + 1024; out = (int32_t*)realloc(out, sizeof(int32_t) * outcap); }
      int64_t m = 0;
      for (int64_t a = cnt[d]; a < cnt[d + 1]; a++)
        for (int k = 0; k < width; k++) { int32_t c = table[adj[a] * width + k]; if (c >= 0) buf[m++] = c; }
      sort_i32(buf, m);
      int64_t u = 0;
      for (int64_t i = 0; i < m; i++)
        if (i == 0 || buf[i] != buf[i - 1]) { out[nout + u] = buf[i]; u++; }
      nout += u;
      rowptr[d + 1] = u;
    }
    free(buf);
    tbuf[me] = out;
  }
  for (int64_t d = 0; d < ndof; d++) rowptr[d + 1] += rowptr[d];
  int32_t* col = (int32_t*)malloc(sizeof(int32_t) * (rowptr[ndof] > 0 ? rowptr[ndof] : 1));
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < nt_used; t++) {
    const int64_t a = rowptr[tlo[t]], b = rowptr[tlo[t + 1]];
    if (b > a) memcpy(col + a, tbuf[t], sizeof(int32_t) * (b - a));
    free(tbuf[t]);
  }
  free(tbuf); free(tlo); free(adj); free(pos); free(cnt);
  *colidx_out = col;
  return rowptr[ndof];
}

void xo_free(void* p) { free(p); }

void xo_scatter(int64_t ndof, int64_t nrows, int width, const int32_t* table, const double* mats,
                const int64_t* rowptr, const int32_t* colidx, double* vals)
{
  (void)ndof;
  for (int64_t r = 0; r < nrows; r++) {
    const int32_t* dn = table + r * width;
    const double* m = mats + r * (int64_t)width * width;
    for (int i = 0; i < width; i++) {
      if (dn[i] < 0) continue;
      int64_t lo0 = rowptr[dn[i]], hi0 = rowptr[dn[i] + 1];
      for (int j = 0; j < width; j++) {
        if (dn[j] < 0) continue;
        int64_t lo = lo0, hi = hi0;
        while (lo < hi) { int64_t mid = (lo + hi) / 2; if (colidx[mid] < dn[j]) lo = mid + 1; else hi = mid; }
        vals[lo] += m[i * width + j];
      }
    }
  }
}

/* the same scatter with the table rows distributed over the host threads (atomic adds): the
 * summation order is not fixed, so this variant is used for CPU baseline TIMING only — the
 * checker is the serial xo_scatter above.  (NGSolve itself assembles in parallel, by element
 * colouring.) */
void xo_scatter_threads(int64_t ndof, int64_t nrows, int width, const int32_t* table, const double* mats,
                        const int64_t* rowptr, const int32_t* colidx, double* vals)
{
  (void)ndof;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nrows; r++) {
    const int32_t* dn = table + r * width;
    const double* m = mats + r * (int64_t)width * width;
    for (int i = 0; i < width; i++) {
      if (dn[i] < 0) continue;
      int64_t lo0 = rowptr[dn[i]], hi0 = rowptr[dn[i] + 1];
      for (int j = 0; j < width; j++) {
        if (dn[j] < 0) continue;
        int64_t lo = lo0, hi = hi0;
        while (lo < hi) { int64_t mid = (lo + hi) / 2; if (colidx[mid] < dn[j]) lo = mid + 1; else hi = mid; }
#pragma omp atomic
        vals[lo] += m[i * width + j];
      }
    }
  }
}

/* threads the OpenMP loops of this library use (reported by bench.py next to the CPU numbers) */
int xo_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void xo_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
