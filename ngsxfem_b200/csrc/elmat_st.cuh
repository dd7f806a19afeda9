// elmat_st.cuh — shared between elmat_st.cu (generic kernel, C ABI) and elmat_st_tp.cu (tensor-structured kernel)
#pragma once
#include "common.cuh"

#define XG_ST_MAXTERMS 8
struct StTermsDev {
  int n;
  int op_trial[XG_ST_MAXTERMS], op_test[XG_ST_MAXTERMS], hpow[XG_ST_MAXTERMS];
  int expo[XG_ST_MAXTERMS][4];  // element vectors only: x^e0 y^e1 z^e2 tref^e3
  double coef[XG_ST_MAXTERMS];
};

// elmat_st_tp.cu; returns 1 if (dim, space order, time order) has no tensor-structured instantiation
int xg_launch_elmat_st_tp(xgpu_ctx* ctx, int dim, int p, int kt, int64_t n, const int32_t* elems, const int32_t* el_vtx,
                          const double* vtx_xyz, const int64_t* offsets, const double* pts, const double* tvals,
                          const double* wts, const StTermsDev& T, int accumulate, const int32_t* slots, double* mats);
