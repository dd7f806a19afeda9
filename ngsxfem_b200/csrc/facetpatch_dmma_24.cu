// facetpatch_dmma_24.cu — k_facetpatch_dmma<2, 4, *> instantiations (see facetpatch.cuh)
#include "facetpatch.cuh"

int xg_fp_dmma_24(xgpu_ctx* ctx, int64_t n, const int32_t* facets, const int32_t* facet_els, const int32_t* el_vtx,
                  const double* vtx_xyz, const xgpu_rule_host* r, double coef, int hpow, double* mats)
{
  return launch_fp_dmma<2, 4>(ctx, n, facets, facet_els, el_vtx, vtx_xyz, r, coef, hpow, mats);
}
