// generic space-time element matrix / vector kernels, dimension 3 (kernels in elmat_st_impl.cuh)
#include "elmat_st_impl.cuh"
XG_ST_INSTANTIATE(3)
