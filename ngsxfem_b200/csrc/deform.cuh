// deform.cuh — isoparametric mesh deformation on the device (xgpu_set_deformation).
// NGSolve's ElementTransformation with AddDeformation (third party; call sites
// cutint/straightcutrule.cpp:568-569, xfem/symboliccutbfi.cpp:135,316,1747-1846):
//   x(xhat) = x_affine(xhat) + sum_a d_a phi_a(xhat),   F = F_affine + sum_a d_a (x) grad_ref phi_a
// with phi_a the H1 basis of order `order` (h1.cuh) and d_a in R^D the element's coefficients.
#pragma once
#include "common.cuh"
#include "cutgeom.cuh"

#define XG_DEFS 112  // doubles of shared memory per warp for the coefficients (35 dofs x 3 at order 4, padded)

struct DefDev {
  int order, ndd;
  const int32_t* el_dofs;
  const double* vals;
};

inline DefDev xg_defdev(const xgpu_ctx* ctx, int dim)
{
  DefDev d;
  d.order = ctx->def_order;
  d.ndd = ctx->def_order > 0 ? xg_h1_ndof(dim, ctx->def_order) : 0;
  d.el_dofs = ctx->def_el_dofs;
  d.vals = ctx->def_vals;
  return d;
}

// warp-cooperative gather of one element's coefficients (ndd x D, dof-major) into shared memory
template <int D>
__device__ __forceinline__ void xg_def_gather(const DefDev& df, int64_t el, int lane, double* defs)
{
  for (int k = lane; k < df.ndd * D; k += 32) {
    const int a = k / D, c = k - a * D;
    const int d = __ldg(df.el_dofs + el * df.ndd + a);
    defs[k] = (d >= 0) ? __ldg(df.vals + (int64_t)d * D + c) : 0.0;
  }
}

template <int D>
struct DefJacSink {
  const double* defs;
  double* F;
  double* x;
  __device__ __forceinline__ void operator()(int a, const AD<D>& s)
  {
#pragma unroll
    for (int c = 0; c < D; c++) {
      const double dc = defs[a * D + c];
      x[c] += dc * s.v;
#pragma unroll
      for (int j = 0; j < D; j++) F[c * D + j] += dc * s.d[j];
    }
  }
};

// in: F = affine Jacobian, x = affine image of xhat (or anything the displacement is to be added to)
template <int D>
__device__ __forceinline__ void xg_def_apply(int order, const int* v, const double* xh, const double* defs, double* F, double* x)
{
  DefJacSink<D> sink{defs, F, x};
  xg_h1_shapes_rt<D>(order, v, xh, sink);
}

// unit normal n = F^{-T} grad_ref(lset) / |.|  (Normalize(grad(lsetp1)) on the mapped element)
template <int D>
__device__ __forceinline__ void xg_phys_normal(const double* Fi, const double* l, double* nrm)
{
  double nn = 0.0;
#pragma unroll
  for (int c = 0; c < D; c++) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) s += Fi[j * D + c] * (l[j] - l[D]);
    nrm[c] = s; nn += s * s;
  }
  nn = sqrt(nn);
  if (nn > 0.0) {
#pragma unroll
    for (int c = 0; c < D; c++) nrm[c] /= nn;
  }
}

template <int D>
__device__ __forceinline__ double xg_mesh_size(double meas)
{
  return (D == 1) ? meas : ((D == 2) ? sqrt(meas) : cbrt(meas));
}
