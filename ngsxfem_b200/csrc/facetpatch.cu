// facetpatch.cu — C ABI of the ghost-penalty facet-patch matrices; kernels in facetpatch.cuh, the tensor-tile
// instantiations in facetpatch_dmma_*.cu (separate translation units: build time)
#include "facetpatch.cuh"

#define XG_FP_ARGS xgpu_ctx* ctx, int64_t n, const int32_t* facets, const int32_t* facet_els, const int32_t* el_vtx, \
                   const double* vtx_xyz, const xgpu_rule_host* r, double coef, int hpow, double* mats
int xg_fp_dmma_23(XG_FP_ARGS);
int xg_fp_dmma_24(XG_FP_ARGS);
int xg_fp_dmma_33(XG_FP_ARGS);
int xg_fp_dmma_34(XG_FP_ARGS);

extern "C" int xgpu_facetpatch(xgpu_ctx* ctx, int dim, int fe_order, int64_t n, const int32_t* facets,
                               const int32_t* facet_els, const int32_t* el_vtx, const double* vtx_xyz,
                               int order, double coef, int hpow, double* mats)
{
  XGPU_NVTX(ctx, "xgpu_facetpatch");
  if (!ctx || dim < 1 || dim > 3 || fe_order < 1 || fe_order > XG_MAX_ORDER)
    return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_facetpatch: bad argument (orders 1..4)");
  if (n == 0) return XGPU_OK;
  if (order < 0) order = 2 * fe_order;  // symboliccutbfi.cpp:2004-2012
  const xgpu_rule_host* r = xgpu_rule_lookup(ctx, dim, order);
  if (!r) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_facetpatch: no rule");
  if (fe_order >= 3 && dim >= 2 && ctx->opt_dmma >= 0) {
    // dense contraction: FP64 tensor tiles (xgpu_set_option("elmat_dmma", -1) forces the CUDA-core kernel)
    if (dim == 3 && fe_order == 3) return xg_fp_dmma_33(ctx, n, facets, facet_els, el_vtx, vtx_xyz, r, coef, hpow, mats);
    if (dim == 3 && fe_order == 4) return xg_fp_dmma_34(ctx, n, facets, facet_els, el_vtx, vtx_xyz, r, coef, hpow, mats);
    if (dim == 2 && fe_order == 3) return xg_fp_dmma_23(ctx, n, facets, facet_els, el_vtx, vtx_xyz, r, coef, hpow, mats);
    if (dim == 2 && fe_order == 4) return xg_fp_dmma_24(ctx, n, facets, facet_els, el_vtx, vtx_xyz, r, coef, hpow, mats);
  }
  const int nd = xg_h1_ndof(dim, fe_order);
  const int Np = (2 * nd + 1) & ~1, PS = Np + 1;
  const DefDev df = xg_defdev(ctx, dim);
  if (df.order > 0 && dim < 2) return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_facetpatch: deformation needs dim >= 2");
  const size_t extra = df.order > 0 ? (size_t)(2 * XG_DEFS + 32 + 2 * FP_GEO + 7 * 2 * r->n) : 0;  // doubles per warp
  int CH = 32;  // staged points per pass: the kernels are register limited to 2 blocks / SM, so up to ~100 KB per block is free
  while (CH > 4 && (size_t)4 * (CH * PS + extra) * sizeof(double) > 100 * 1024) CH >>= 1;
  const size_t sm = (size_t)4 * (CH * PS + extra) * sizeof(double);
  if (sm > 220 * 1024) return xgpu_fail(ctx, XGPU_ERR_CAPACITY, "xgpu_facetpatch: patch rule too large for the shared-memory staging");
  unsigned grid = nblk(n, 4);
  if (grid > XGPU_NSM * 8) grid = XGPU_NSM * 8;
#define XG_FP(DD, DF)                                                                                               \
  {                                                                                                                 \
    XGPU_CUDA(ctx, cudaFuncSetAttribute(k_facetpatch<DD, DF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_facetpatch<DD, DF><<<grid, 128, sm, ctx->stream>>>(n, facets, facet_els, el_vtx, vtx_xyz, r->d_pts, r->d_wts, r->n, \
                                                         fe_order, nd, CH, coef, hpow, ctx->fp_downscale, mats, df);                   \
  }
  if (df.order > 0) { if (dim == 3) XG_FP(3, true) else XG_FP(2, true) }
  else if (dim == 3) XG_FP(3, false) else if (dim == 2) XG_FP(2, false) else XG_FP(1, false)
#undef XG_FP
  XGPU_LAUNCH_CHECK(ctx);
  return XGPU_OK;
}

// SymbolicFacetPatchBilinearFormIntegrator::SetIRScaling (xfem/python_xfem.cpp:966-967, symboliccutbfi.cpp:2020-2033): state
// of the context, read by xgpu_facetpatch / xgpu_facetpatch_st; only 0 < s <= 1 is valid
extern "C" int xgpu_facetpatch_set_downscale(xgpu_ctx* ctx, double s)
{
  if (!ctx) return XGPU_ERR_ARG;
  if (!(s > 0.0 && s <= 1.0)) return xgpu_fail(ctx, XGPU_ERR_ARG, "xgpu_facetpatch_set_downscale: only 0 < downscale <= 1 valid");
  ctx->fp_downscale = s;
  return XGPU_OK;
}
