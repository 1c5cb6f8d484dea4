// conv_common.cuh -- pieces shared by the forward and backward convolution kernels.
#pragma once
#include "rowwarp.cuh"

namespace sph3d {

// stage filter[f][cbase + lane*VEC + v][j] for all f into a [F][warp-wide strip] shared array
template <int VEC, int R>
__device__ __forceinline__ void stage_filter(float* Wsh, const float* __restrict__ filter, int F, int C, int cbase)
{
    constexpr int E = VEC * R;
    using S = SmemStrip<E>;
    for (int t = threadIdx.x; t < F * S::FLOATS; t += blockDim.x) {
        int f = t / S::FLOATS, rem = t % S::FLOATS;
        int ln = rem / E, e = rem % E;
        int c = cbase + ln * VEC + e / R;
        Wsh[f * S::FLOATS + S::flat(ln, e)] = (c < C) ? __ldg(filter + ((size_t)f * C + c) * R + (e % R)) : 0.f;
    }
}

// s += v, with Blackwell's packed fp32x2 adds (FADD2) where the strip is wide enough
template <int VEC>
__device__ __forceinline__ void strip_add(float (&s)[VEC], const float (&v)[VEC])
{
    if constexpr (VEC == 4) {
        float2 a = __fadd2_rn(make_float2(s[0], s[1]), make_float2(v[0], v[1]));
        float2 b = __fadd2_rn(make_float2(s[2], s[3]), make_float2(v[2], v[3]));
        s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y;
    } else if constexpr (VEC == 2) {
        float2 a = __fadd2_rn(make_float2(s[0], s[1]), make_float2(v[0], v[1]));
        s[0] = a.x; s[1] = a.y;
    } else {
        s[0] += v[0];
    }
}

// highest set bit of a non-zero, warp-uniform mask; clears it (FLO + SHF + LOP: cheaper than ffs)
__device__ __forceinline__ int pop_highest(unsigned& m)
{
    int b = 31 - __clz(m);
    m ^= 1u << b;
    return b;
}

// Sum the feature strips of the neighbours selected by the warp-uniform mask m (bit k <-> the id held
// by lane k in `myidx`).  Two independent gathers per trip, every branch is warp-uniform, no padding.
template <int VEC>
__device__ __forceinline__ void gather_sum_lean(float (&s)[VEC], unsigned m, int myidx,
                                                const float* __restrict__ inb, int C, bool active)
{
    while (m) {
        const int n0 = __shfl_sync(FULL_MASK, myidx, pop_highest(m));
        float v0[VEC];
        VecIO<VEC>::ld(v0, inb + (size_t)n0 * C, active);
        if (m) {
            const int n1 = __shfl_sync(FULL_MASK, myidx, pop_highest(m));
            float v1[VEC];
            VecIO<VEC>::ld(v1, inb + (size_t)n1 * C, active);
            strip_add<VEC>(s, v1);
        }
        strip_add<VEC>(s, v0);
    }
}

struct ConvPlan {
    int vec;          // 4 / 2 / 1, 0 = generic fallback
    int chunks;       // gridDim.y: channel chunks of 32*vec
    int grid_x;       // persistent CTAs along rows
    int threads;      // CTA size
    int slots;        // backward: filter bins owned per warp
    size_t smem;      // dynamic shared memory bytes
};

static inline int pick_vec(int C) { return (C % 4 == 0) ? 4 : ((C % 2 == 0) ? 2 : 1); }
static const size_t SMEM_CAP = 227 * 1024;

// rows are handed to CTAs in contiguous chunks so that the warps sharing an SM's L1 work on
// neighbouring rows (the "first K by index" rule makes neighbouring rows share most neighbours)
constexpr int ROWS_PER_CHUNK = 128;

template <typename Kern>
static cudaError_t set_smem(Kern k, size_t bytes)
{
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace sph3d
