// rowwarp.cuh -- the shared skeleton of every gather / segment-sum kernel in this library.
//
// Work unit: one WARP owns one (row, channel-chunk) pair, where a row is one output point (b,m)
// with its <=K neighbour list, and a chunk is 32*VEC consecutive input channels (lane l owns
// channels c0 = chunk*32*VEC + l*VEC .. +VEC).  A feature-row gather is then ONE coalesced warp
// load of 128*VEC bytes (LDG.128 per lane when VEC=4) -- the access shape that keeps the L1/L2
// wavefront count at its minimum for this irregular gather (DESIGN.md section "conv").
// Lanes whose channels fall beyond C stay idle (their loads are predicated off), so any C works;
// VEC=4 needs C%4==0 (16-byte aligned rows), VEC=2 needs C%2==0, VEC=1 always works.
#pragma once
#include "common.cuh"

namespace sph3d {

template <int VEC> struct VecIO;

template <> struct VecIO<4> {
    static __device__ __forceinline__ void ld(float (&v)[4], const float* p, bool pred)
    {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pred) t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[4])
    {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void sti(int* p, const int (&v)[4])
    {
        *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void red(float* p, const float (&v)[4])
    {
        red_add_v4(p, v[0], v[1], v[2], v[3]);
    }
};

template <> struct VecIO<2> {
    static __device__ __forceinline__ void ld(float (&v)[2], const float* p, bool pred)
    {
        float2 t = make_float2(0.f, 0.f);
        if (pred) t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[2])
    {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
    static __device__ __forceinline__ void sti(int* p, const int (&v)[2])
    {
        *reinterpret_cast<int2*>(p) = make_int2(v[0], v[1]);
    }
    static __device__ __forceinline__ void red(float* p, const float (&v)[2])
    {
        asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    }
};

template <> struct VecIO<1> {
    static __device__ __forceinline__ void ld(float (&v)[1], const float* p, bool pred)
    {
        v[0] = pred ? __ldg(p) : 0.f;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[1]) { *p = v[0]; }
    static __device__ __forceinline__ void sti(int* p, const int (&v)[1]) { *p = v[0]; }
    static __device__ __forceinline__ void red(float* p, const float (&v)[1]) { atomicAdd(p, v[0]); }
};

// widest vector (in floats) usable for a per-lane strip of E contiguous floats
__host__ __device__ constexpr int strip_vw(int E) { return (E % 4 == 0) ? 4 : ((E % 2 == 0) ? 2 : 1); }

// Per-lane strip of E floats kept in shared memory as [plane][lane][VW] so that every access
// is a conflict-free LDS/STS of VW words (a plain [lane][E] layout is 2-way conflicted for E=8).
template <int E>
struct SmemStrip {
    static constexpr int VW = strip_vw(E);
    static constexpr int PLANES = E / VW;
    static constexpr int FLOATS = E * 32;          // per warp-wide strip
    static __device__ __forceinline__ int offset(int plane, int lane) { return (plane * 32 + lane) * VW; }

    static __device__ __forceinline__ void load(float (&v)[E], const float* base, int lane)
    {
#pragma unroll
        for (int pl = 0; pl < PLANES; pl++) {
            const float* p = base + offset(pl, lane);
            if constexpr (VW == 4) {
                float4 t = *reinterpret_cast<const float4*>(p);
                v[pl * 4] = t.x; v[pl * 4 + 1] = t.y; v[pl * 4 + 2] = t.z; v[pl * 4 + 3] = t.w;
            } else if constexpr (VW == 2) {
                float2 t = *reinterpret_cast<const float2*>(p);
                v[pl * 2] = t.x; v[pl * 2 + 1] = t.y;
            } else {
                v[pl] = *p;
            }
        }
    }
    static __device__ __forceinline__ void store(float* base, int lane, const float (&v)[E])
    {
#pragma unroll
        for (int pl = 0; pl < PLANES; pl++) {
            float* p = base + offset(pl, lane);
            if constexpr (VW == 4) {
                *reinterpret_cast<float4*>(p) = make_float4(v[pl * 4], v[pl * 4 + 1], v[pl * 4 + 2], v[pl * 4 + 3]);
            } else if constexpr (VW == 2) {
                *reinterpret_cast<float2*>(p) = make_float2(v[pl * 2], v[pl * 2 + 1]);
            } else {
                *p = v[pl];
            }
        }
    }
    // index of strip element e of lane `lane` inside a warp-wide strip
    static __host__ __device__ __forceinline__ int flat(int lane, int e)
    {
        return ((e / VW) * 32 + lane) * VW + (e % VW);
    }
};

}  // namespace sph3d
