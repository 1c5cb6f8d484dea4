// conv_common.cuh -- pieces shared by the forward and backward convolution kernels.
//
// Index arithmetic in the hot loops is 32-bit on purpose: a neighbour id is turned ONCE per row into a
// byte offset inside its cloud (id * C * 4, host-checked to fit 32 bits), shuffles then broadcast
// ready-made offsets and an address is one 64-bit add.  Row/cloud bookkeeping is incremental (no
// 64-bit divisions in device code).
#pragma once
#include <cstdlib>
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

// -------------------------------------------------------------------------------------------------
// Warp-level counting sort of one 64-edge tile by filter bin.
//
// Lane l holds edge l (o0, b0) and edge 32+l (o1, b1); b = -1 marks "no edge".  After the call the
// warp's shared arrays hold the tile grouped by bin, ascending bin, original k order inside a bin:
//     sOff [p]  byte offset of the p-th edge's neighbour row,        p < n_edges
//     sCode[p]  (bin << 1) | last     last = 1 on the final edge of a bin's segment
//     hA[f]     start position of bin f (exclusive prefix over ALL bins, hA[FP] = n_edges)
// Deterministic (ranks come from MATCH.ANY lane masks, not from atomics).  ~85 instructions per tile,
// which buys a flat, counted, branch-light gather loop instead of per-bin mask walking.
// FP = number of bins rounded up to a multiple of 32 (<= 128); hA has FP+1 ints, hB has FP ints.
__device__ __forceinline__ void sort_tile_by_bin(unsigned o0, int b0, unsigned o1, int b1, int FP, int lane,
                                                 int* __restrict__ hA, int* __restrict__ hB,
                                                 unsigned* __restrict__ sOff, int* __restrict__ sCode)
{
    for (int f = lane; f < FP; f += 32) { hA[f] = 0; hB[f] = 0; }
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    const unsigned p0 = __match_any_sync(FULL_MASK, b0), p1 = __match_any_sync(FULL_MASK, b1);
    const int r0 = __popc(p0 & lt), r1 = __popc(p1 & lt);
    const int n0 = __popc(p0), n1 = __popc(p1);
    if (b0 >= 0 && r0 == 0) hA[b0] = n0;             // one writer per bin and half
    if (b1 >= 0 && r1 == 0) hB[b1] = n1;
    __syncwarp();
    const int other1 = (b0 >= 0) ? hB[b0] : 0;      // edges of my bin in the second half
    __syncwarp();
    int base = 0;
    for (int f0 = 0; f0 < FP; f0 += 32) {
        const int c0 = hA[f0 + lane], c1 = hB[f0 + lane], t = c0 + c1;
        int incl = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += y;
        }
        const int excl = base + incl - t;
        hA[f0 + lane] = excl;                        // where half-0 edges of this bin start
        hB[f0 + lane] = excl + c0;                   // where half-1 edges of this bin start
        base += __shfl_sync(FULL_MASK, incl, 31);
    }
    if (lane == 0) hA[FP] = base;
    __syncwarp();
    if (b0 >= 0) {
        const int pos = hA[b0] + r0;
        sOff[pos] = o0;
        sCode[pos] = (b0 << 1) | ((r0 == n0 - 1 && other1 == 0) ? 1 : 0);
    }
    if (b1 >= 0) {
        const int pos = hB[b1] + r1;
        sOff[pos] = o1;
        sCode[pos] = (b1 << 1) | ((r1 == n1 - 1) ? 1 : 0);
    }
    __syncwarp();
}

// ints of shared memory one warp needs for sort_tile_by_bin
static __host__ __device__ inline int sort_smem_ints(int F) { int FP = ((F + 31) / 32) * 32; return (FP + 1 + 3) / 4 * 4 + FP + 64 + 64; }

struct ConvPlan {
    int vec;          // 4 / 2 / 1, 0 = generic fallback
    int chunks;       // gridDim.y: channel chunks of 32*vec
    int grid_x;       // persistent CTAs along rows
    int threads;      // CTA size
    int slots;        // backward: filter bins owned per warp
    size_t smem;      // dynamic shared memory bytes
    int cta_reduce;   // backward: groups of a CTA are summed in shared memory before the partial is written
    int rpc;          // rows per contiguous chunk handed to a CTA
};

static inline int pick_vec(int C) { return (C % 4 == 0) ? 4 : ((C % 2 == 0) ? 2 : 1); }

// Strip width that fills the warp: the narrowest legal VEC whose single chunk (32*VEC channels) still covers C.
// C = 64 runs 32 lanes x 2 channels instead of 16 lanes x 4 (measured: backward 0.56 -> 0.38 ms at B=8, N=8192,
// C=64, r=2); C >= 128 keeps 16-byte strips.
static inline int pick_vec_full_warp(int C)
{
    const int widest = pick_vec(C);
    for (int v = 1; v <= widest; v <<= 1)
        if (C % v == 0 && C <= 32 * v) return v;
    return widest;
}
static const size_t SMEM_CAP = 227 * 1024;

// the lean kernels use 32-bit row ids and 32-bit byte offsets inside a cloud
static inline bool fits_32bit(int B, int N, int M, int C, int r)
{
    return (long long)B * M < (1LL << 31) && (long long)N * C * 4 < (1LL << 32) &&
           (long long)M * C * r * 4 < (1LL << 40);
}

// rows are handed to CTAs in contiguous chunks so that the warps sharing an SM's L1 work on
// neighbouring rows (the "first K by index" rule makes neighbouring rows share most neighbours)
// Launch-shape tunables (defaults are what bench.py measures; the SPH3D_* environment variables exist for the sweeps
// documented in DESIGN.md).  The environment is read ONCE, when the library is loaded, into this table; a process that
// changes a variable afterwards (tests, sweep scripts) calls sph3d_reload_tunables().  0 = unset -> the default applies.
struct Tunables {
    int rows_per_chunk;                                  // SPH3D_ROWS_PER_CHUNK (pins the chunk size when set)
    int fwd_threads, fwd_vec;                            // SPH3D_FWD_THREADS, SPH3D_FWD_VEC
    int bwd_algo, bwd_vec, bwd_g, bwd_threads, bwd_cta_reduce;   // SPH3D_BWD_*  (row-owned backward)
    int bwdt_threads, bwdt_depth, bwdt_g, bwdt_sort, bwdt_fold;   // SPH3D_BWDT_* (transposed backward)
    int nnquery_grid;                                    // SPH3D_NNQUERY_GRID: -1 unset, 0 never, 2 whenever possible
    int pool_stream;                                     // SPH3D_POOL_STREAM: -1 unset = streaming gather form, 0 = warp-per-point gather form
    int fps_handshake, fps_cluster_min_n;                // SPH3D_FPS_HANDSHAKE (-1 unset = on, 0 = cluster barrier), SPH3D_FPS_CLUSTER_MIN_N
    int fwd_smem_pad_kb;                                 // SPH3D_FWD_SMEM_PAD_KB: unused shared memory added to the forward kernel (shrinks its L1: the sensitivity sweep of DESIGN 4.10)
    int sepconv_tile, sepconv_stages;                    // SPH3D_SEPCONV_TILE (64 / 32 rows), SPH3D_SEPCONV_STAGES (weight ring depth)
};
const Tunables& tunables();                              // conv_fwd.cu
static inline int tun(int v, int dflt) { return v > 0 ? v : dflt; }
static inline int rows_per_chunk() { return tun(tunables().rows_per_chunk, 128); }

// Chunk size for a problem of `rows` rows when `want` CTAs (per channel chunk) would fill the machine.  Large problems
// keep the default (contiguous rows share neighbours, so long chunks keep the gathers in L1).  Small ones -- the deep
// levels of the segmentation networks: 1 000-6 000 rows, up to 8 channel chunks -- would leave most SMs idle with
// 128-row chunks (a 1024-row level is 8 chunks) while each warp walks its rows one after the other; there the chunk
// shrinks until every SM has one, down to `min_rows` (one row per warp or per warp group).
static inline int pick_rows_per_chunk(long long rows, long long want, int min_rows)
{
    int rpc = rows_per_chunk();
    if (tunables().rows_per_chunk > 0) return rpc;                   // sweeps pin it
    const long long nchunks = (rows + rpc - 1) / rpc;
    if (nchunks >= want) return rpc;
    long long per = (rows + want - 1) / want;                        // rows per CTA if every SM gets one chunk
    per = (per + min_rows - 1) / min_rows * min_rows;
    if (per < min_rows) per = min_rows;
    return (int)(per < rpc ? per : rpc);
}

// conv_bwd.cu: out[t] = sum_p part[p][t], fixed order (bit-reproducible)
int launch_reduce_partials(int P, size_t n, const float* part, float* out, cudaStream_t st);

// conv_bwd_t.cu: the transposed backward (default where it applies; SPH3D_BWD_ALGO=1 forces the row-owned form)
bool bwd_transposed_supported(int B, int N, int M, int F, int C, int r, int K);
size_t bwd_transposed_workspace_bytes(int B, int N, int M, int F, int C, int r, int K);
int bwd_transposed_run(int B, int N, int M, int F, int C, int r, int K, const int* nn_index, const int* nn_count,
                       const int* bin_index, const float* input, const float* filter, const float* grad_output,
                       float* grad_input, float* grad_filter, void* workspace, size_t workspace_bytes, cudaStream_t st);

// conv_bwd_t.cu: gather form of the avg-pool / mean- and weighted-interpolate gradients (pool3d.cu routes to it when the
// caller provides scratch).  S = source points per cloud (grad_input rows), R = referencing rows per cloud.
size_t pool_scatter_workspace_bytes(int B, int S, int R, int C, int K);
int pool_scatter_run(int B, int S, int R, int C, int K, const int* nn_index, const int* nn_count, const float* weight,
                     const float* grad_output, float* grad_input, void* workspace, size_t workspace_bytes, cudaStream_t st);

template <typename Kern>
static cudaError_t set_smem(Kern k, size_t bytes)
{
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace sph3d
