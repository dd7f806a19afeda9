// elmat_dmma_3.cu — k_elmat_dmma<3, P, *> instantiations (see elmat.cuh)
#include "elmat.cuh"

int xg_launch_dmma_3(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats)
{
  switch (p) {
    case 2: return launch_dmma<3, 2>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
    case 3: return launch_dmma<3, 3>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
    case 4: return launch_dmma<3, 4>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats);
  }
  return xgpu_fail(ctx, XGPU_ERR_ARG, "elmat: orders 2..4 on tensor tiles");
}
