// elmat_dmma_2.cu — k_elmat_dmma<2, P, *> instantiations (see elmat.cuh)
#include "elmat.cuh"

int xg_launch_dmma_2(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats)
{
  switch (p) {
    case 2: return launch_dmma<2, 2>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
    case 3: return launch_dmma<2, 3>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
    case 4: return launch_dmma<2, 4>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
  }
  return xgpu_fail(ctx, XGPU_ERR_ARG, "elmat: orders 2..4 on tensor tiles");
}
