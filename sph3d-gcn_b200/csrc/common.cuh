// common.cuh -- shared device helpers for libsph3d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SPH3D_SM_COUNT_FALLBACK 148
#define FULL_MASK 0xffffffffu

// Reference launch geometry: part of the ball query's semantics.
#define REF_GRID 32
#define REF_BLOCK 1024

namespace sph3d {

__host__ inline int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = SPH3D_SM_COUNT_FALLBACK;
    }
    return n;
}

// counts kernels enqueued by the most recent C-ABI call (bench.py's gpu_launches evidence)
extern thread_local int g_last_launch_count;   // per host thread: the entry points stay reentrant

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Squared distance in the reference's contraction order: the y product is rounded
// alone, the x and z products are fused.  Intrinsics are never re-contracted by nvcc.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz)
{
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// 16-byte vector reduction to global memory (REDG.E.ADD.F32x4 on sm_90+).
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// lowest set bit index of a non-zero mask, and clear it
__device__ __forceinline__ int pop_lowest(unsigned& m)
{
    int b = __ffs(m) - 1;
    m &= m - 1;
    return b;
}

}  // namespace sph3d

#define SPH3D_CHECK_LAUNCH()                                   \
    do {                                                       \
        cudaError_t e__ = cudaPeekAtLastError();               \
        if (e__ != cudaSuccess) return (int)cudaGetLastError(); \
    } while (0)
