// classify.cu — CutInfo on the GPU: element classification sweep, cut ratios, BitArray algebra,
// ordered compaction and the marker helpers of xfem/cutinfo.cpp.
//
// Kernel plan for CutInformation::Update (xfem/cutinfo.cpp:61-133):
//   k_classify<D>   one thread per element, coalesced connectivity read (int4 for tets), level-set
//                   gathers served by L2 (nodal vector << 126 MB), warp ballot -> one 32-bit word
//                   per warp and domain type.  HBM bound: 16 B + 8 B per element + gathers.
//   k_cut_ratio<D>  only for the ~2 % sign-cut elements (one thread per bitset word, iterating the
//                   set bits): the two order-0 cut rules of the reference -> part volumes,
//                   cut_ratio, and the reference's volume based NEG/POS/IF decision.
#include "common.cuh"
#include "cutgeom.cuh"

namespace {

template <int D>
__device__ __forceinline__ void load_vtx(const int32_t* __restrict__ el_vtx, int64_t el, int* v)
{
  if (D == 3) {
    int4 q = __ldg(reinterpret_cast<const int4*>(el_vtx) + el);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else if (D == 1) {
    int2 q = __ldg(reinterpret_cast<const int2*>(el_vtx) + el);
    v[0] = q.x; v[1] = q.y;
  } else {
    for (int k = 0; k < D + 1; k++) v[k] = __ldg(el_vtx + el * (D + 1) + k);
  }
}

template <int D>
__global__ void __launch_bounds__(256) k_classify(int64_t ne, const int32_t* __restrict__ el_vtx,
                                                  const double* __restrict__ phi,
                                                  uint32_t* __restrict__ bits_neg, uint32_t* __restrict__ bits_pos,
                                                  uint32_t* __restrict__ bits_if, double* __restrict__ cut_ratio,
                                                  int32_t* __restrict__ cutlist, unsigned int* __restrict__ cutcount)
{
  __shared__ unsigned int s_cnt, s_base;
  if (threadIdx.x == 0) s_cnt = 0u;
  __syncthreads();
  const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = el < ne;
  bool haspos = false, hasneg = false;
  if (valid) {
    int v[D + 1];
    load_vtx<D>(el_vtx, el, v);
#pragma unroll
    for (int k = 0; k < D + 1; k++) {
      double p = __ldg(phi + v[k]);
      haspos |= (p > 0.0);
      hasneg |= (p < 0.0);
    }
  }
  // CheckIfStraightCut (straightcutrule.cpp:11-24): both or neither -> IF
  const bool is_if = valid && (haspos == hasneg);
  const bool is_neg = valid && hasneg && !haspos;
  const bool is_pos = valid && haspos && !hasneg;
  const unsigned bn = __ballot_sync(0xffffffffu, is_neg);
  const unsigned bp = __ballot_sync(0xffffffffu, is_pos);
  const unsigned bi = __ballot_sync(0xffffffffu, is_if);
  if ((threadIdx.x & 31) == 0 && el < ne) {
    const int64_t w = el >> 5;
    bits_neg[w] = bn; bits_pos[w] = bp; bits_if[w] = bi;
  }
  // uncut: part volumes are (ref vol, 0) resp. (0, ref vol) -> ratio 1 resp. 0 (cutinfo.cpp:104)
  if (valid && cut_ratio && !is_if) cut_ratio[el] = is_neg ? 1.0 : 0.0;
  // the sign-cut elements (~2 %) are appended to a list (one global atomic per block; the order does not matter): the
  // geometry of k_cut_ratio then runs with full warps instead of one or two lanes per warp of the element sweep
  unsigned int mypos = 0u;
  if (is_if) mypos = atomicAdd(&s_cnt, 1u);
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt > 0u) s_base = atomicAdd(cutcount, s_cnt);
  __syncthreads();
  if (is_if) cutlist[s_base + mypos] = (int32_t)el;
}

// one thread per element, only the sign-cut ones (bit set in bits_if) do work: in a warp that holds
// a cut element the geometry runs once with the cut lanes active
template <int D>
__global__ void __launch_bounds__(256) k_cut_ratio(const int32_t* __restrict__ cutlist, const unsigned int* __restrict__ cutcount,
                                                   const int32_t* __restrict__ el_vtx,
                                                   const double* __restrict__ phi, const double* __restrict__ w0, int nw0,
                                                   uint32_t* __restrict__ bits_neg, uint32_t* __restrict__ bits_pos,
                                                   uint32_t* __restrict__ bits_if, double* __restrict__ cut_ratio,
                                                   int* status, unsigned long long* status_elem)
{
  const unsigned int ncut = *cutcount;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncut; i += gridDim.x * blockDim.x) {
  const int64_t el = cutlist[i];
  const int64_t w = el >> 5;
  const uint32_t bit = 1u << (el & 31);
  int v[D + 1];
  double l[D + 1];
  load_vtx<D>(el_vtx, el, v);
#pragma unroll
  for (int k = 0; k < D + 1; k++) l[k] = __ldg(phi + v[k]);
  XgSimplex ref, sub[3];
  xg_ref_simplex(D, ref);
  double part[2] = {0.0, 0.0};
  int err = 0;
  // POS first, then NEG, as cutinfo.cpp:93
  for (int a = 0; a < 2 && !err; a++) {
    const int np = (a == 0) ? XGPU_POS : XGPU_NEG;
    int ns = xg_decompose(ref, l, np, sub);
    if (ns < 0) { err = ns; break; }
    double s = 0.0;
    for (int i = 0; i < ns; i++) {
      const double vol = xg_volume(sub[i]);
      for (int q = 0; q < nw0; q++) s += w0[q] * vol;
    }
    part[np] = s;
  }
  if (err) {
    xgpu_record_error(status, status_elem, err, el);
    if (cut_ratio) cut_ratio[el] = nan("");
    atomicAnd(bits_if + w, ~bit);
    continue;
  }
  if (cut_ratio) cut_ratio[el] = part[XGPU_NEG] / (part[XGPU_NEG] + part[XGPU_POS]);
  // volume based decision of cutinfo.cpp:104-124 (differs from the sign test only on underflow)
  if (!(part[XGPU_NEG] > 0.0 && part[XGPU_POS] > 0.0)) {
    atomicAnd(bits_if + w, ~bit);
    if (part[XGPU_NEG] > 0.0) atomicOr(bits_neg + w, bit); else atomicOr(bits_pos + w, bit);
  }
  }
}

__global__ void k_bitset_op(int op, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                            uint32_t* __restrict__ out, int64_t nwords, uint32_t tailmask)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nwords) return;
  uint32_t r;
  switch (op) {
    case 0: r = a[i] | b[i]; break;
    case 1: r = a[i] & b[i]; break;
    case 2: r = a[i] & ~b[i]; break;
    case 3: r = ~a[i]; break;
    case 4: r = a[i]; break;
    case 5: r = 0u; break;
    default: r = 0xffffffffu; break;
  }
  if (i == nwords - 1) r &= tailmask;
  out[i] = r;
}

__global__ void k_popc_words(const uint32_t* __restrict__ bits, int64_t nwords, int32_t* __restrict__ cnt)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nwords) cnt[i] = __popc(bits[i]);
}

__global__ void k_popc_total(const uint32_t* __restrict__ bits, int64_t nwords, unsigned long long* total)
{
  unsigned long long s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x)
    s += __popc(bits[i]);
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

__global__ void k_bits_to_list(const uint32_t* __restrict__ bits, int64_t nwords, const int32_t* __restrict__ woff,
                               int32_t* __restrict__ list)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nwords) return;
  uint32_t word = bits[i];
  int32_t o = woff[i];
  const int32_t base = (int32_t)(i << 5);
  while (word) {
    const int b = __ffs(word) - 1;
    word &= word - 1;
    list[o++] = base + b;
  }
}

__global__ void k_bits_rank(const uint32_t* __restrict__ bits, int64_t nbits, const int32_t* __restrict__ woff,
                            int32_t* __restrict__ map)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbits) return;
  const uint32_t word = bits[i >> 5];
  const int b = (int)(i & 31);
  map[i] = ((word >> b) & 1u) ? woff[i >> 5] + __popc(word & ((1u << b) - 1u)) : -1;
}

__global__ void k_threshold(int64_t ne, const double* __restrict__ cr, int dt, double thr, uint32_t* __restrict__ out)
{
  const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool m = false;
  if (el < ne) {
    const double r = cr[el];
    m = (dt == XGPU_NEG) ? (r >= thr) : ((dt == XGPU_POS) ? (r <= thr) : false);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && el < ne) out[el >> 5] = bal;
}

#define XG_ML_MAXM 16
struct MlTuples { int M, nt; signed char dts[64 * XG_ML_MAXM]; };

template <int D>
__global__ void __launch_bounds__(256) k_classify_ml(int64_t ne, const int32_t* __restrict__ el_vtx,
                                                     const double* __restrict__ phis, int64_t nv, MlTuples tp,
                                                     uint32_t* __restrict__ out)
{
  const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool match = false;
  if (el < ne) {
    int v[D + 1];
    load_vtx<D>(el_vtx, el, v);
    int dts_el[XG_ML_MAXM];
    for (int j = 0; j < tp.M; j++) {
      bool hp = false, hn = false;
#pragma unroll
      for (int k = 0; k < D + 1; k++) {
        const double p = __ldg(phis + (int64_t)j * nv + v[k]);
        hp |= p > 0.0; hn |= p < 0.0;
      }
      dts_el[j] = (hp == hn) ? XGPU_IF : (hn ? XGPU_NEG : XGPU_POS);
    }
    for (int t = 0; t < tp.nt; t++) {
      bool m = true;
      for (int j = 0; j < tp.M; j++) m &= (tp.dts[t * tp.M + j] == dts_el[j]);
      match |= m;
    }
  }
  const unsigned bal = __ballot_sync(0xffffffffu, match);
  if ((threadIdx.x & 31) == 0 && el < ne) out[el >> 5] = bal;
}

__device__ __forceinline__ bool tbit(const uint32_t* b, int64_t i) { return (__ldg(b + (i >> 5)) >> (i & 31)) & 1u; }

__global__ void k_facets_nb(int64_t nf, const int32_t* __restrict__ facet_els, const uint32_t* __restrict__ a,
                            const uint32_t* __restrict__ b, int bva, int bvb, int use_and, uint32_t* __restrict__ out)
{
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool m = false;
  if (f < nf) {
    const int2 e = __ldg(reinterpret_cast<const int2*>(facet_els) + f);
    const bool al = tbit(a, e.x), bl = tbit(b, e.x);
    const bool ar = e.y >= 0 ? tbit(a, e.y) : (bva != 0);
    const bool br = e.y >= 0 ? tbit(b, e.y) : (bvb != 0);
    m = use_and ? ((al && br) || (ar && bl)) : ((al || br) || (ar || bl));
  }
  const unsigned bal = __ballot_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && f < nf) out[f >> 5] = bal;
}

__global__ void k_els_of_facets(int64_t nf, const int32_t* __restrict__ facet_els, const uint32_t* __restrict__ fb,
                                uint32_t* __restrict__ out)
{
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf || !tbit(fb, f)) return;
  const int2 e = __ldg(reinterpret_cast<const int2*>(facet_els) + f);
  atomicOr(out + (e.x >> 5), 1u << (e.x & 31));
  if (e.y >= 0) atomicOr(out + (e.y >> 5), 1u << (e.y & 31));
}

// one thread per element; only marked elements read their dof row.  A plain load tests the bit
// before the atomic: most dofs are shared by 5-25 elements, so only the first toucher pays for the
// L2 atomic (a stale read merely repeats an idempotent atomicOr).
__global__ void __launch_bounds__(256) k_dofs_of_els(int64_t ne, int nd, const int32_t* __restrict__ el_dofs,
                                                     const uint32_t* __restrict__ eb, uint32_t* out)
{
  const int64_t el = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= ne || !tbit(eb, el)) return;
  const int32_t* __restrict__ row = el_dofs + el * nd;
  for (int k = 0; k < nd; k++) {
    const int32_t d = __ldg(row + k);
    if (d < 0) continue;
    const uint32_t m = 1u << (d & 31);
    volatile uint32_t* wp = out + (d >> 5);
    if (!(*wp & m)) atomicOr(out + (d >> 5), m);
  }
}

__global__ void k_gather_rows(int64_t n, const int32_t* __restrict__ list, int width, const int32_t* __restrict__ table,
                              const int32_t* __restrict__ map, int32_t* __restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int64_t r = i / width;
  const int k = (int)(i - r * width);
  const int64_t src = list ? (int64_t)list[r] : r;
  int32_t d = __ldg(table + src * width + k);
  if (map && d >= 0) d = __ldg(map + d);
  out[i] = d;
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" int xgpu_classify(xgpu_ctx* ctx, int dim, int64_t ne, const int32_t* el_vtx, const double* phi,
                             uint32_t* bits_neg, uint32_t* bits_pos, uint32_t* bits_if, double* cut_ratio)
{
  XGPU_NVTX(ctx, "xgpu_classify");
  if (!ctx || dim < 1 || dim > 3 || ne < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_classify: bad argument");
  if (ne == 0) return XGPU_OK;
  const xgpu_rule_host* r0 = xgpu_rule_lookup(ctx, dim, 0);
  if (!r0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_classify: no order-0 rule");
  cudaStream_t st = ctx->stream;
  if (ne >= (int64_t)0x7fffffff) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "xgpu_classify: 2^31 or more elements");
  void* sp;
  int rc = xgpu_scratch(ctx, 11, sizeof(int32_t) * (size_t)(ne + 4), &sp);
  if (rc) return rc;
  unsigned int* cutcount = (unsigned int*)sp;
  int32_t* cutlist = (int32_t*)sp + 4;
  XGPU_CUDA(ctx, cudaMemsetAsync(cutcount, 0, sizeof(unsigned int), st));
#define XG_CLS(DD)                                                                                         \
  k_classify<DD><<<nblk(ne, 256), 256, 0, st>>>(ne, el_vtx, phi, bits_neg, bits_pos, bits_if, cut_ratio, cutlist, cutcount);  \
  XGPU_LAUNCH_CHECK(ctx);                                                                                  \
  k_cut_ratio<DD><<<XGPU_NSM * 4, 256, 0, st>>>(cutlist, cutcount, el_vtx, phi, r0->d_wts, r0->n, bits_neg, bits_pos, \
                                                bits_if, cut_ratio, ctx->d_status, ctx->d_status_elem); \
  XGPU_LAUNCH_CHECK(ctx);
  if (dim == 3) { XG_CLS(3) } else if (dim == 2) { XG_CLS(2) } else { XG_CLS(1) }
#undef XG_CLS
  return XGPU_OK;
}

extern "C" int xgpu_bitset_op(xgpu_ctx* ctx, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, int64_t nbits)
{
  if (!ctx || op < 0 || op > 6) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_bitset_op: bad argument");
  const int64_t nwords = (nbits + 31) >> 5;
  if (nwords == 0) return XGPU_OK;
  const uint32_t tail = (nbits & 31) ? ((1u << (nbits & 31)) - 1u) : 0xffffffffu;
  k_bitset_op<<<nblk(nwords, 256), 256, 0, ctx->stream>>>(op, a, b, out, nwords, tail);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_bitset_count(xgpu_ctx* ctx, const uint32_t* bits, int64_t nbits, int64_t* count_host)
{
  if (!ctx || !count_host) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_bitset_count: bad argument");
  const int64_t nwords = (nbits + 31) >> 5;
  *count_host = 0;
  if (nwords == 0) return XGPU_OK;
  void* p;
  int rc = xgpu_scratch(ctx, 6, 64, &p);
  if (rc) return rc;
  XGPU_CUDA(ctx, cudaMemsetAsync(p, 0, 8, ctx->stream));
  unsigned g = nblk(nwords, 256);
  if (g > XGPU_NSM * 8) g = XGPU_NSM * 8;
  k_popc_total<<<g, 256, 0, ctx->stream>>>(bits, nwords, (unsigned long long*)p);
  XGPU_LAUNCH_CHECK(ctx);
  XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, p, 8, ctx->stream));
  XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *count_host = ctx->h_pinned[0];
  return XGPU_OK;
}

namespace {
// set bits with index < nbits (the last word masked: a PREFIX of a longer bitset may be counted)
__global__ void k_popc_prefix(const uint32_t* __restrict__ bits, int64_t nbits, unsigned long long* total)
{
  const int64_t nwords = (nbits + 31) >> 5;
  unsigned long long s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t w = bits[i];
    if (i == nwords - 1 && (nbits & 31)) w &= (1u << (nbits & 31)) - 1u;
    s += __popc(w);
  }
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}
}  // namespace

extern "C" int xgpu_bitset_count_many(xgpu_ctx* ctx, int k, const uint32_t* const* bits_host, const int64_t* nbits_host,
                                      int64_t* counts_host)
{
  if (!ctx || k < 1 || k > 8 || !bits_host || !nbits_host || !counts_host)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_bitset_count_many: bad argument (1..8 bitsets)");
  for (int i = 0; i < k; i++) counts_host[i] = 0;
  void* p;
  int rc = xgpu_scratch(ctx, 6, 64, &p);
  if (rc) return rc;
  XGPU_CUDA(ctx, cudaMemsetAsync(p, 0, 64, ctx->stream));
  bool any = false;
  for (int i = 0; i < k; i++) {
    if (nbits_host[i] < 0) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_bitset_count_many: negative length");
    const int64_t nwords = (nbits_host[i] + 31) >> 5;
    if (nwords == 0) continue;
    unsigned g = nblk(nwords, 256);
    if (g > XGPU_NSM * 8) g = XGPU_NSM * 8;
    k_popc_prefix<<<g, 256, 0, ctx->stream>>>(bits_host[i], nbits_host[i], (unsigned long long*)p + i);
    XGPU_LAUNCH_CHECK(ctx);
    any = true;
  }
  if (!any) return XGPU_OK;
  XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, p, 8 * k, ctx->stream));   // ONE read-back and ONE host wait for all of them
  XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < k; i++) counts_host[i] = ctx->h_pinned[i];
  return XGPU_OK;
}

static int word_offsets(xgpu_ctx* ctx, const uint32_t* bits, int64_t nwords, int32_t** woff)
{
  void* p;
  int rc = xgpu_scratch(ctx, 5, sizeof(int32_t) * (2 * nwords + 2), &p);
  if (rc) return rc;
  int32_t* cnt = (int32_t*)p;
  *woff = cnt + nwords;
  k_popc_words<<<nblk(nwords, 256), 256, 0, ctx->stream>>>(bits, nwords, cnt);
  XGPU_LAUNCH_CHECK(ctx);
  return xgpu_exclusive_scan_i32_o32(ctx, cnt, nwords, *woff);
}

extern "C" int xgpu_bitset_to_list(xgpu_ctx* ctx, const uint32_t* bits, int64_t nbits, int32_t* list)
{
  XGPU_NVTX(ctx, "xgpu_bitset_to_list");
  if (!ctx) return XGPU_ERR_ARG;
  const int64_t nwords = (nbits + 31) >> 5;
  if (nwords == 0) return XGPU_OK;
  int32_t* woff;
  int rc = word_offsets(ctx, bits, nwords, &woff);
  if (rc) return rc;
  k_bits_to_list<<<nblk(nwords, 256), 256, 0, ctx->stream>>>(bits, nwords, woff, list);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_bitset_rank(xgpu_ctx* ctx, const uint32_t* bits, int64_t nbits, int32_t* map, int64_t* nactive_host)
{
  if (!ctx) return XGPU_ERR_ARG;
  const int64_t nwords = (nbits + 31) >> 5;
  if (nactive_host) *nactive_host = 0;
  if (nwords == 0) return XGPU_OK;
  int32_t* woff;
  int rc = word_offsets(ctx, bits, nwords, &woff);
  if (rc) return rc;
  k_bits_rank<<<nblk(nbits, 256), 256, 0, ctx->stream>>>(bits, nbits, woff, map);
  XGPU_LAUNCH_CHECK(ctx);
  if (nactive_host) {
    XGPU_CUDA(ctx, xgpu_d2h_small(ctx->h_pinned, woff + nwords, 4, ctx->stream));
    XGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *nactive_host = *(int32_t*)ctx->h_pinned;
  }
  return XGPU_OK;
}

// ---- classification from part volumes (space-time CutInfo: cutinfo.cpp:90-124 with time_order >= 0) ----
// sums[i] = sum of the weights of element i's rule, added in rule order (as `for (auto w : wei_arr) part_vol += w`)
__global__ void __launch_bounds__(256) k_rule_sums(int64_t n, const int64_t* __restrict__ offsets,
                                                   const double* __restrict__ wts, double* __restrict__ sums)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int64_t q = offsets[i]; q < offsets[i + 1]; q++) s += wts[q];
  sums[i] = s;
}

// NEG volume > 0 and POS volume > 0 -> IF; only NEG > 0 -> NEG; else POS; ratio = neg / (neg + pos) (cutinfo.cpp:104-124)
__global__ void __launch_bounds__(256) k_classify_volumes(int64_t ne, const double* __restrict__ vneg,
                                                          const double* __restrict__ vpos, uint32_t* __restrict__ bneg,
                                                          uint32_t* __restrict__ bpos, uint32_t* __restrict__ bif,
                                                          double* __restrict__ ratio)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < ne;
  const double a = in ? vneg[i] : 0.0, b = in ? vpos[i] : 0.0;
  const bool hn = in && a > 0.0, hp = b > 0.0;
  const unsigned mn = __ballot_sync(0xffffffffu, hn && !hp);
  const unsigned mi = __ballot_sync(0xffffffffu, hn && hp);
  const unsigned mp = __ballot_sync(0xffffffffu, in && !hn);
  if (in && ratio) ratio[i] = a / (a + b);
  if ((threadIdx.x & 31) == 0 && in) {
    bneg[i >> 5] = mn; bif[i >> 5] = mi; bpos[i >> 5] = mp;
  }
}

extern "C" int xgpu_rule_element_sums(xgpu_ctx* ctx, int64_t n, const int64_t* offsets, const double* wts, double* sums)
{
  if (!ctx || !offsets || !sums) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_rule_element_sums: bad argument");
  if (n == 0) return XGPU_OK;
  k_rule_sums<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, offsets, wts, sums);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_classify_volumes(xgpu_ctx* ctx, int64_t ne, const double* vol_neg, const double* vol_pos,
                                     uint32_t* bits_neg, uint32_t* bits_pos, uint32_t* bits_if, double* cut_ratio)
{
  if (!ctx || !vol_neg || !vol_pos || !bits_neg || !bits_pos || !bits_if)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_classify_volumes: bad argument");
  if (ne == 0) return XGPU_OK;
  k_classify_volumes<<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, vol_neg, vol_pos, bits_neg, bits_pos, bits_if, cut_ratio);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_threshold_elements(xgpu_ctx* ctx, int64_t ne, const double* cut_ratio, int dt, double threshold,
                                       uint32_t* bits_out)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (ne == 0) return XGPU_OK;
  if (dt == XGPU_POS) threshold = 1 - threshold;  // cutinfo.cpp:269-270
  k_threshold<<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, cut_ratio, dt, threshold, bits_out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_classify_ml(xgpu_ctx* ctx, int dim, int64_t ne, const int32_t* el_vtx, const double* phis,
                                int64_t nv, int M, const int* dts_host, int ntuples, uint32_t* bits_out)
{
  if (!ctx || dim < 1 || dim > 3 || M < 1 || M > XG_ML_MAXM || ntuples < 0 || ntuples > 64)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_classify_ml: bad argument (M<=16, ntuples<=64)");
  if (ne == 0) return XGPU_OK;
  MlTuples tp;
  tp.M = M; tp.nt = ntuples;
  for (int i = 0; i < M * ntuples; i++) tp.dts[i] = (signed char)dts_host[i];
  if (dim == 3) k_classify_ml<3><<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, el_vtx, phis, nv, tp, bits_out);
  else if (dim == 2) k_classify_ml<2><<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, el_vtx, phis, nv, tp, bits_out);
  else k_classify_ml<1><<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, el_vtx, phis, nv, tp, bits_out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_facets_with_neighbor_types(xgpu_ctx* ctx, int64_t nf, const int32_t* facet_els, const uint32_t* a,
                                               const uint32_t* b, int bnd_val_a, int bnd_val_b, int use_and,
                                               uint32_t* bits_out)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (nf == 0) return XGPU_OK;
  k_facets_nb<<<nblk(nf, 256), 256, 0, ctx->stream>>>(nf, facet_els, a, b, bnd_val_a, bnd_val_b, use_and, bits_out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_elements_with_neighbor_facets(xgpu_ctx* ctx, int64_t nf, const int32_t* facet_els,
                                                  const uint32_t* facet_bits, int64_t ne, uint32_t* bits_out)
{
  if (!ctx) return XGPU_ERR_ARG;
  XGPU_CUDA(ctx, cudaMemsetAsync(bits_out, 0, sizeof(uint32_t) * ((ne + 31) >> 5), ctx->stream));
  if (nf == 0) return XGPU_OK;
  k_els_of_facets<<<nblk(nf, 256), 256, 0, ctx->stream>>>(nf, facet_els, facet_bits, bits_out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_dofs_of_elements(xgpu_ctx* ctx, int64_t ne, int nd, const int32_t* el_dofs, const uint32_t* el_bits,
                                     int64_t ndof, uint32_t* bits_out)
{
  if (!ctx) return XGPU_ERR_ARG;
  XGPU_CUDA(ctx, cudaMemsetAsync(bits_out, 0, sizeof(uint32_t) * ((ndof + 31) >> 5), ctx->stream));
  if (ne == 0) return XGPU_OK;
  k_dofs_of_els<<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, nd, el_dofs, el_bits, bits_out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

namespace {
// elements with at least one marked entry in their table row (vertices / dofs): thread per element, ballot -> words
__global__ void __launch_bounds__(256) k_els_with_marked(int64_t ne, int nd, const int32_t* __restrict__ table,
                                                         const uint32_t* __restrict__ marks, uint32_t* __restrict__ bits_out)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false;
  if (e < ne) {
    for (int k = 0; k < nd; k++) {
      const int32_t d = __ldg(table + e * nd + k);
      if (d >= 0 && ((__ldg(marks + (d >> 5)) >> (d & 31)) & 1u)) hit = true;
    }
  }
  const unsigned w = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && e < ne) bits_out[e >> 5] = w;
}
}  // namespace

extern "C" int xgpu_elements_with_marked(xgpu_ctx* ctx, int64_t ne, int nd, const int32_t* table, const uint32_t* marks,
                                         uint32_t* bits_out)
{
  if (!ctx || ne < 0 || nd < 1 || (ne > 0 && (!table || !marks || !bits_out))) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_elements_with_marked: bad argument");
  if (ne == 0) return XGPU_OK;
  k_els_with_marked<<<nblk(ne, 256), 256, 0, ctx->stream>>>(ne, nd, table, marks, bits_out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

extern "C" int xgpu_gather_rows(xgpu_ctx* ctx, int64_t n, const int32_t* list, int width, const int32_t* table,
                                const int32_t* map, int32_t* out)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (n == 0) return XGPU_OK;
  k_gather_rows<<<nblk(n * width, 256), 256, 0, ctx->stream>>>(n, list, width, table, map, out);
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}
