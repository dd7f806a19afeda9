// cutquad_ext.cu — multi-level-set and space-time cut rules (placeholders until implemented).
#include "common.cuh"

extern "C" int xgpu_cutquad_ml_count(xgpu_ctx* ctx, int, int64_t, const int32_t*, const int32_t*, const double*, int64_t, int,
                                     const int*, int, int, int64_t*, int64_t*)
{ return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_cutquad_ml_count: not implemented yet"); }
extern "C" int xgpu_cutquad_ml_fill(xgpu_ctx* ctx, int, int64_t, const int32_t*, const int32_t*, const double*, const double*,
                                    int64_t, int, const int*, int, int, const int64_t*, double*, double*)
{ return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_cutquad_ml_fill: not implemented yet"); }
extern "C" int xgpu_cutquad_st_count(xgpu_ctx* ctx, int, int64_t, const int32_t*, const int32_t*, const double*, int64_t, int,
                                     int, int, int, int64_t*, int64_t*)
{ return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_cutquad_st_count: not implemented yet"); }
extern "C" int xgpu_cutquad_st_fill(xgpu_ctx* ctx, int, int64_t, const int32_t*, const int32_t*, const double*, const double*,
                                    int64_t, int, int, int, int, const int64_t*, double*, double*, double*)
{ return xgpu_fail(ctx, XGPU_ERR_UNSUPPORTED, "xgpu_cutquad_st_fill: not implemented yet"); }
