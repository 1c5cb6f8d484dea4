// dense_gemm.cuh -- the pointwise product of every SPH3D layer (tf.matmul over the B*M rows,
// /root/reference/utils/sph3gcn_util.py:144-146, :203-205, :254-256) on the 5th-generation tensor cores, at fp32 accuracy.
//
// The products are fp32 by contract (1e-5 parity) and compute-bound on the fp32 SIMT pipe in cuBLAS (S3DIS step: 111
// GFLOP = 2.2 ms of a 10 ms step at ~50 TFLOP/s).  tcgen05 has no fp32 MMA; a single TF32 or BF16 pass loses 13-16
// mantissa bits.  Here every fp32 operand tile is split ON CHIP into three BF16 terms (a = a0 + a1*2^-8 + a2*2^-16) by a
// transform warp group after the TMA load, and the product is accumulated from all 9 cross terms (5 "bands") with
// block-scaled tcgen05.mma into TMEM -- the operands cross HBM once, as fp32.  The mainloop (TMA load -> transform ->
// UMMA -> TMEM accumulator -> tcgen05.ld epilogue, warp-specialised, persistent CLC tile scheduler) is the CuTe/CUTLASS
// collective `MainloopSm100TmaUmmaWarpSpecializedFastF32` from the header tree vendored in this image
// (flashinfer/data/cutlass/include, CUTLASS 4.5), instantiated per operand layout in dense_*.cu.
//
// SCHEDULE: KernelTmaWarpSpecialized1SmFastFP32Sm100 keeps the BF16 terms of a K-major A operand in TENSOR MEMORY (the
// transform warps write them with tcgen05.st; only B's terms go back to shared memory), ...SmemSm100 keeps both in
// shared memory (required for an M-major A, i.e. the weight gradient).
//
// TILE_M / CLUSTER_M / EPILOGUE: _128 / _1 / TmaWarpSpecialized1Sm = one CTA per 128 x 128 tile; _256 / _2 /
// TmaWarpSpecialized2Sm with a ...2Sm... schedule = cta_group::2, a CTA pair shares one 256 x 128 tile (each CTA loads half
// of B; the pair's MMA reads both halves).
//
// Each instantiation lives at namespace scope (nvcc's host pass cannot size CollectiveEpilogue::SharedStorage from inside
// a class template) and in its own translation unit (two minutes of template expansion each, compiled in parallel).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#ifndef SPH3D_NO_CUTLASS
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"

// D[l] (M x N, row-major) = A[l] (M x K, LAYOUT_A) * B[l] (K x N, LAYOUT_B), l < L, packed batch strides
#define SPH3D_DEFINE_FP32_GEMM(NS, LAYOUT_A, LAYOUT_B, SCHEDULE, TILE_M, CLUSTER_M, EPILOGUE)                              \
namespace NS {                                                                                                              \
    using namespace cute;                                                                                                   \
    using LayoutC = cutlass::layout::RowMajor;                                                                              \
    using MmaTileShape = Shape<TILE_M, _128, _16>;                                                                          \
    using ClusterShape = Shape<CLUSTER_M, _1, _1>;                                                                              \
    using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<                                   \
        cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,                                   \
        cutlass::epilogue::collective::EpilogueTileAuto, float, float, float, LayoutC, 4, float, LayoutC, 4,                \
        cutlass::epilogue::EPILOGUE>::CollectiveOp;                                                                        \
    using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<                                       \
        cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, float, LAYOUT_A, 4, float, LAYOUT_B, 4, float,                \
        MmaTileShape, ClusterShape,                                                                                         \
        cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>, \
        cutlass::gemm::SCHEDULE>::CollectiveOp;                                                                            \
    using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>; \
    using DeviceGemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;                                             \
    using StrideA = typename GemmKernel::StrideA;                                                                           \
    using StrideB = typename GemmKernel::StrideB;                                                                           \
    using StrideC = typename GemmKernel::StrideC;                                                                           \
    using StrideD = typename GemmKernel::StrideD;                                                                           \
    static typename DeviceGemm::Arguments make(int M, int N, int K, int L, const float* A, const float* B, float* D)        \
    {                                                                                                                       \
        typename DeviceGemm::Arguments args;                                                                                \
        args.mode = cutlass::gemm::GemmUniversalMode::kGemm;                                                                \
        args.problem_shape = cute::make_shape(M, N, K, L);                                                                  \
        args.mainloop.ptr_A = A; args.mainloop.dA = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, L)); \
        args.mainloop.ptr_B = B; args.mainloop.dB = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, L)); \
        args.epilogue.thread.alpha = 1.0f; args.epilogue.thread.beta = 0.0f;                                                \
        args.epilogue.ptr_C = D; args.epilogue.dC = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, L)); \
        args.epilogue.ptr_D = D; args.epilogue.dD = cutlass::make_cute_packed_stride(StrideD{}, cute::make_shape(M, N, L)); \
        return args;                                                                                                        \
    }                                                                                                                       \
    size_t workspace(int M, int N, int K, int L)                                                                            \
    {                                                                                                                       \
        auto args = make(M, N, K, L, nullptr, nullptr, nullptr);                                                            \
        return DeviceGemm::get_workspace_size(args);                                                                        \
    }                                                                                                                       \
    int run(int M, int N, int K, int L, const float* A, const float* B, float* D, void* ws, size_t ws_bytes, cudaStream_t st) \
    {                                                                                                                       \
        DeviceGemm gemm;                                                                                                    \
        auto args = make(M, N, K, L, A, B, D);                                                                              \
        if (gemm.can_implement(args) != cutlass::Status::kSuccess) return (int)cudaErrorInvalidValue;                       \
        if (DeviceGemm::get_workspace_size(args) > ws_bytes) return (int)cudaErrorInvalidValue;                             \
        if (gemm.initialize(args, ws, st) != cutlass::Status::kSuccess) return (int)cudaErrorInvalidValue;                  \
        if (gemm.run(st) != cutlass::Status::kSuccess) return (int)cudaErrorLaunchFailure;                                  \
        return 0;                                                                                                           \
    }                                                                                                                       \
}
#else
// built without the CUTLASS header tree: the entry points exist and report "not supported"
#define SPH3D_DEFINE_FP32_GEMM(NS, LAYOUT_A, LAYOUT_B, SCHEDULE, TILE_M, CLUSTER_M, EPILOGUE)                              \
namespace NS {                                                                                                              \
    size_t workspace(int, int, int, int) { return 0; }                                                                      \
    int run(int, int, int, int, const float*, const float*, float*, void*, size_t, cudaStream_t) { return (int)cudaErrorNotSupported; } \
}
#endif

namespace sph3d_dense_nn { size_t workspace(int, int, int, int); int run(int, int, int, int, const float*, const float*, float*, void*, size_t, cudaStream_t); }
namespace sph3d_dense_nt { size_t workspace(int, int, int, int); int run(int, int, int, int, const float*, const float*, float*, void*, size_t, cudaStream_t); }
namespace sph3d_dense_tn { size_t workspace(int, int, int, int); int run(int, int, int, int, const float*, const float*, float*, void*, size_t, cudaStream_t); }
namespace sph3d_dense_nn2 { size_t workspace(int, int, int, int); int run(int, int, int, int, const float*, const float*, float*, void*, size_t, cudaStream_t); }
namespace sph3d_dense_nt2 { size_t workspace(int, int, int, int); int run(int, int, int, int, const float*, const float*, float*, void*, size_t, cudaStream_t); }
