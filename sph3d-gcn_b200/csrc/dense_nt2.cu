// gx = g * w^T with cta_group::2: a CTA pair owns one 256 x 128 output tile (see dense_gemm.cuh).
#include "dense_gemm.cuh"
#ifdef SPH3D_NO_CUTLASS
SPH3D_DEFINE_FP32_GEMM(sph3d_dense_nt2, _, _, _, _, _, _)
#else
SPH3D_DEFINE_FP32_GEMM(sph3d_dense_nt2, cutlass::layout::RowMajor, cutlass::layout::ColumnMajor, KernelTmaWarpSpecialized2SmFastFP32Sm100, _256, _2, TmaWarpSpecialized2Sm)
#endif
