// gw[l] (Cin*r x Cout) = x[l]^T * g[l] over L row slabs (explicit split-K as a batch; the caller sums the slabs in order):
// x (rows x Cin*r) row-major read as the M-major (column-major) A operand, g (rows x Cout) row-major as the N-major B operand.
#include "dense_gemm.cuh"
#ifdef SPH3D_NO_CUTLASS
SPH3D_DEFINE_FP32_GEMM(sph3d_dense_tn, _, _, _, _, _, _)
#else
SPH3D_DEFINE_FP32_GEMM(sph3d_dense_tn, cutlass::layout::ColumnMajor, cutlass::layout::RowMajor, KernelTmaWarpSpecialized1SmFastFP32SmemSm100, _128, _1, TmaWarpSpecialized1Sm)
#endif
