// conv_common.cuh -- pieces shared by the forward and backward convolution kernels.
//
// Index arithmetic in the hot loops is 32-bit on purpose: a neighbour id is turned ONCE per row into a
// byte offset inside its cloud (id * C * 4, host-checked to fit 32 bits), shuffles then broadcast
// ready-made offsets and an address is one 64-bit add.  Row/cloud bookkeeping is incremental (no
// 64-bit divisions in device code).
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

// unconditional strip load at byte offset `off` from `base` (callers clamp idle lanes to a valid strip)
template <int VEC>
__device__ __forceinline__ void ld_strip(float (&v)[VEC], const char* __restrict__ base, unsigned off)
{
    const float* p = reinterpret_cast<const float*>(base + off);
    if constexpr (VEC == 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (VEC == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = __ldg(p);
    }
}

// s += v, with Blackwell's packed fp32x2 add (FADD2) where the strip is wide enough
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

// highest set bit of a non-zero, warp-uniform mask; clears it (FLO + SHF + LOP)
__device__ __forceinline__ int pop_highest(unsigned& m)
{
    int b = 31 - __clz(m);
    m ^= 1u << b;
    return b;
}

// Sum the feature strips of the edges selected by the warp-uniform mask m.  Bit k <-> the byte offset
// held by lane k in `myoff`.  Up to four independent gathers in flight; every branch is warp-uniform;
// no padding work.
template <int VEC>
__device__ __forceinline__ void gather_sum_lean(float (&s)[VEC], unsigned m, unsigned myoff,
                                                const char* __restrict__ inb)
{
    while (m) {
        float v0[VEC], v1[VEC], v2[VEC], v3[VEC];
        ld_strip<VEC>(v0, inb, __shfl_sync(FULL_MASK, myoff, pop_highest(m)));
        if (m) {
            ld_strip<VEC>(v1, inb, __shfl_sync(FULL_MASK, myoff, pop_highest(m)));
            if (m) {
                ld_strip<VEC>(v2, inb, __shfl_sync(FULL_MASK, myoff, pop_highest(m)));
                if (m) {
                    ld_strip<VEC>(v3, inb, __shfl_sync(FULL_MASK, myoff, pop_highest(m)));
                    strip_add<VEC>(s, v3);
                }
                strip_add<VEC>(s, v2);
            }
            strip_add<VEC>(s, v1);
        }
        strip_add<VEC>(s, v0);
    }
}

// Incremental (cloud, point) bookkeeping for a warp that visits rows row0, row0+step, ... of a
// contiguous chunk: avoids a division per row.
struct RowCursor {
    unsigned b, m;
    __device__ __forceinline__ void init(unsigned row, unsigned M) { b = row / M; m = row - b * M; }
    __device__ __forceinline__ void advance(unsigned step, unsigned M)
    {
        m += step;
        while (m >= M) { m -= M; b++; }
    }
};

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

// the lean kernels use 32-bit row ids and 32-bit byte offsets inside a cloud
static inline bool fits_32bit(int B, int N, int M, int C, int r)
{
    return (long long)B * M < (1LL << 31) && (long long)N * C * 4 < (1LL << 32) &&
           (long long)M * C * r * 4 < (1LL << 40);
}

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
