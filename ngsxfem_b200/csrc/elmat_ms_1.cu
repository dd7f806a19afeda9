// elmat_ms_1.cu — k_elmat_ms<1, P, *> instantiations (see elmat.cuh)
#include "elmat.cuh"

int xg_launch_ms_1(xgpu_ctx* ctx, int p, int64_t n, const int32_t* elems, const int32_t* el_vtx, const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* wts, int np_shared, const TermsDev& T, int accumulate, const int32_t* slots, double* mats, const Rule2* r2)
{
  switch (p) {
    case 1: return launch_ms<1, 1>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats, r2);
    case 2: return launch_ms<1, 2>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats, r2);
    case 3: return launch_ms<1, 3>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats, r2);
    case 4: return launch_ms<1, 4>(ctx, n, elems, el_vtx, vtx_xyz, offsets, pts, wts, np_shared, T, accumulate, slots, mats, r2);
  }
  return xgpu_fail(ctx, XGPU_ERR_ARG, "elmat: orders 1..4");
}
