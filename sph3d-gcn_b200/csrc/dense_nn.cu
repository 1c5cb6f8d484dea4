// y (R x Cout) = x (R x Cin*r, row-major) * w (Cin*r x Cout, row-major): the forward pointwise product.
#include "dense_gemm.cuh"
#ifdef SPH3D_NO_CUTLASS
SPH3D_DEFINE_FP32_GEMM(sph3d_dense_nn, _, _, _, _, _, _)
#else
SPH3D_DEFINE_FP32_GEMM(sph3d_dense_nn, cutlass::layout::RowMajor, cutlass::layout::RowMajor, KernelTmaWarpSpecialized1SmFastFP32Sm100, _128, _1, TmaWarpSpecialized1Sm)
#endif
