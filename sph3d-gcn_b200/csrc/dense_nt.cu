// gx (R x Cin*r) = g (R x Cout, row-major) * w^T: w (Cin*r x Cout) row-major read as the K-major (column-major) B operand.
#include "dense_gemm.cuh"
#ifdef SPH3D_NO_CUTLASS
SPH3D_DEFINE_FP32_GEMM(sph3d_dense_nt, _, _, _, _, _, _)
#else
SPH3D_DEFINE_FP32_GEMM(sph3d_dense_nt, cutlass::layout::RowMajor, cutlass::layout::ColumnMajor, KernelTmaWarpSpecialized1SmFastFP32Sm100, _128, _1, TmaWarpSpecialized1Sm)
#endif
