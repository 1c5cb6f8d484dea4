// conv3d.cu -- depthwise spherical graph convolution, forward and backward, sm_100a.
//
// Replaces depthwiseConv3dLauncher / depthwiseConv3dGradLauncher
// (/root/reference/tf_ops/convolution/tf_conv3d_gpu.cu:107-140; kernels :7-101) and the
// cudaMemset zero fills of tf_conv3d.cpp:90,152-153.
//
//   out[b,m,c*r+j] = (1/cnt) * sum_{k<cnt} in[b, nn[b,m,k], c] * W[bin[b,m,k], c, j]          (Q9)
//
// Design (see rowwarp.cuh for the work unit): a warp owns one output point and 32*VEC input
// channels.  It reads the point's K neighbour ids and bin ids ONCE (coalesced, two per lane),
// then walks the bins that actually occur in the row (a 64-bit presence mask built with one
// warp-OR): for each bin a ballot selects its edges, their feature strips are gathered and SUMMED
// (one LDG.128 per lane and edge, four in flight), and the bin's filter strip -- staged once per
// persistent CTA in shared memory, conflict-free layout -- is applied once per (row, bin) instead
// of once per edge.  That is the segment-weighted-sum form of the op: FMA count drops from
// E*C*r to (#row-bins)*C*r, shared-memory filter traffic drops by the mean segment length (~4x at
// K=64, F=33), and what remains is the irreducible gather of E*C*4 bytes through L1/L2.
// Nothing is accumulated in global memory (the reference does a global read-modify-write per
// edge and channel) and the index rows are read once per 32*VEC channels, not once per channel.
//
// Backward is ONE fused pass over the same structure: per (row, bin) it forms d = sum_j g*W
// once and scatters it to grad_input with 16-byte vector reductions (REDG.ADD.F32x4), while the
// gathered-and-summed features give the filter gradient g*sum(in), accumulated without atomics in
// a per-warp private shared-memory copy of the filter, reduced per CTA, written as a per-CTA
// partial and summed in a fixed order by a second tiny kernel (deterministic grad_filter; the
// reference uses shared+global float atomics and ceil(F*C*r/12288) full re-passes, Q13/Q14).
#include "rowwarp.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

int g_last_launch_count = 0;

constexpr int CONV_WARPS = 8;

// stage filter[f][cbase + lane*VEC + v][j] for all f into a [F][strip] shared array
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

template <int VEC, int R>
__global__ void __launch_bounds__(CONV_WARPS * 32)
conv_fwd_kernel(int B, int N, int M, int F, int C, int K,
                const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                const int* __restrict__ bin_index, const float* __restrict__ input,
                const float* __restrict__ filter, float* __restrict__ output)
{
    constexpr int E = VEC * R;
    using S = SmemStrip<E>;
    extern __shared__ __align__(16) float smem[];
    float* Wsh = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cbase = blockIdx.y * 32 * VEC;
    stage_filter<VEC, R>(Wsh, filter, F, C, cbase);
    __syncthreads();

    const int c0 = cbase + lane * VEC;
    const bool active = c0 < C;
    const long long rows = (long long)B * M;
    for (long long row = (long long)blockIdx.x * CONV_WARPS + warp; row < rows;
         row += (long long)gridDim.x * CONV_WARPS) {
        const int b = (int)(row / M);
        const int cnt = __ldg(nn_count + row);
        const float* inb = input + (size_t)b * N * C + c0;
        const int* idxrow = nn_index + (size_t)row * K;
        const int* binrow = bin_index + (size_t)row * K;
        float acc[E];
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = 0.f;

        for (int kt = 0; kt < cnt; kt += 64) {
            const int k0 = kt + lane, k1 = kt + 32 + lane;
            int i0 = 0, b0 = -1, i1 = 0, b1 = -1;
            if (k0 < cnt) { i0 = __ldg(idxrow + k0); b0 = __ldg(binrow + k0); }
            if (k1 < cnt) { i1 = __ldg(idxrow + k1); b1 = __ldg(binrow + k1); }
            auto do_bin = [&](int f) {
                unsigned m0 = __ballot_sync(FULL_MASK, b0 == f);
                unsigned m1 = __ballot_sync(FULL_MASK, b1 == f);
                if (!(m0 | m1)) return;
                float s[VEC];
#pragma unroll
                for (int v = 0; v < VEC; v++) s[v] = 0.f;
                gather_sum<VEC>(s, m0, i0, inb, C, active);
                gather_sum<VEC>(s, m1, i1, inb, C, active);
                float w[E];
                S::load(w, Wsh + f * S::FLOATS, lane);
#pragma unroll
                for (int e = 0; e < E; e++) acc[e] = fmaf(s[e / R], w[e], acc[e]);
            };
            unsigned plo, phi;
            present_bins(b0, b1, plo, phi);
            while (plo) do_bin(pop_lowest(plo));
            while (phi) do_bin(32 + pop_lowest(phi));
            for (int f = 64; f < F; f++) do_bin(f);
        }
        if (active) {
            const float inv = cnt > 0 ? 1.0f / (float)cnt : 0.f;
#pragma unroll
            for (int e = 0; e < E; e++) acc[e] *= inv;
            float* out = output + (size_t)row * C * R + (size_t)c0 * R;
            constexpr int VW = strip_vw(E);
#pragma unroll
            for (int pl = 0; pl < E / VW; pl++) {
                float t[VW];
#pragma unroll
                for (int u = 0; u < VW; u++) t[u] = acc[pl * VW + u];
                VecIO<VW>::st(out + pl * VW, t);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fused backward: grad_input (vector reductions) + grad_filter (private smem accumulators)
// smem: Wsh[F][strip] | acc[WARPS][F][strip]
template <int VEC, int R>
__global__ void __launch_bounds__(512)
conv_bwd_kernel(int B, int N, int M, int F, int C, int K,
                const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                const int* __restrict__ bin_index, const float* __restrict__ input,
                const float* __restrict__ filter, const float* __restrict__ grad_output,
                float* __restrict__ grad_input, float* __restrict__ gw_partial)
{
    constexpr int E = VEC * R;
    using S = SmemStrip<E>;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cbase = blockIdx.y * 32 * VEC;
    float* Wsh = smem;
    float* accsh = smem + (size_t)F * S::FLOATS * (1 + warp);
    stage_filter<VEC, R>(Wsh, filter, F, C, cbase);
    for (int t = threadIdx.x; t < F * S::FLOATS * nwarps; t += blockDim.x) smem[F * S::FLOATS + t] = 0.f;
    __syncthreads();

    const int c0 = cbase + lane * VEC;
    const bool active = c0 < C;
    const long long rows = (long long)B * M;
    for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows;
         row += (long long)gridDim.x * nwarps) {
        const int b = (int)(row / M);
        const int cnt = __ldg(nn_count + row);
        if (cnt <= 0) continue;
        const float* inb = input + (size_t)b * N * C + c0;
        float* gib = grad_input + (size_t)b * N * C + c0;
        const int* idxrow = nn_index + (size_t)row * K;
        const int* binrow = bin_index + (size_t)row * K;
        float g[E];
        {
            const float inv = 1.0f / (float)cnt;
            const float* go = grad_output + (size_t)row * C * R + (size_t)c0 * R;
            constexpr int VW = strip_vw(E);
#pragma unroll
            for (int pl = 0; pl < E / VW; pl++) {
                float t[VW];
                VecIO<VW>::ld(t, go + pl * VW, active);
#pragma unroll
                for (int u = 0; u < VW; u++) g[pl * VW + u] = t[u] * inv;
            }
        }
        for (int kt = 0; kt < cnt; kt += 64) {
            const int k0 = kt + lane, k1 = kt + 32 + lane;
            int i0 = 0, b0 = -1, i1 = 0, b1 = -1;
            if (k0 < cnt) { i0 = __ldg(idxrow + k0); b0 = __ldg(binrow + k0); }
            if (k1 < cnt) { i1 = __ldg(idxrow + k1); b1 = __ldg(binrow + k1); }
            auto scatter = [&](unsigned m, int myidx, float (&s)[VEC], const float (&d)[VEC]) {
                while (m) {
                    int l0 = pop_lowest(m);
                    bool p1 = m != 0; int l1 = p1 ? pop_lowest(m) : l0;
                    int n0 = __shfl_sync(FULL_MASK, myidx, l0);
                    int n1 = __shfl_sync(FULL_MASK, myidx, l1);
                    float v0[VEC], v1[VEC];
                    VecIO<VEC>::ld(v0, inb + (size_t)n0 * C, active);
                    VecIO<VEC>::ld(v1, inb + (size_t)n1 * C, active && p1);
                    if (active) {
                        VecIO<VEC>::red(gib + (size_t)n0 * C, d);
                        if (p1) VecIO<VEC>::red(gib + (size_t)n1 * C, d);
                    }
#pragma unroll
                    for (int v = 0; v < VEC; v++) s[v] += v0[v] + v1[v];
                }
            };
            auto do_bin = [&](int f) {
                unsigned m0 = __ballot_sync(FULL_MASK, b0 == f);
                unsigned m1 = __ballot_sync(FULL_MASK, b1 == f);
                if (!(m0 | m1)) return;
                float w[E];
                S::load(w, Wsh + f * S::FLOATS, lane);
                float d[VEC], s[VEC];
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    float t = 0.f;
#pragma unroll
                    for (int j = 0; j < R; j++) t = fmaf(g[v * R + j], w[v * R + j], t);
                    d[v] = t; s[v] = 0.f;
                }
                scatter(m0, i0, s, d);
                scatter(m1, i1, s, d);
                float a[E];
                float* ap = accsh + f * S::FLOATS;
                S::load(a, ap, lane);
#pragma unroll
                for (int e = 0; e < E; e++) a[e] = fmaf(g[e], s[e / R], a[e]);
                S::store(ap, lane, a);
            };
            unsigned plo, phi;
            present_bins(b0, b1, plo, phi);
            while (plo) do_bin(pop_lowest(plo));
            while (phi) do_bin(32 + pop_lowest(phi));
            for (int f = 64; f < F; f++) do_bin(f);
        }
    }
    __syncthreads();
    // CTA partial: sum the per-warp copies in warp order, write [blockIdx.x][f][c][j]
    float* part = gw_partial + (size_t)blockIdx.x * F * C * R;
    for (int t = threadIdx.x; t < F * S::FLOATS; t += blockDim.x) {
        int f = t / S::FLOATS, rem = t % S::FLOATS;
        int ln = rem / E, e = rem % E;
        int c = cbase + ln * VEC + e / R;
        if (c < C) {
            float sum = 0.f;
            int off = f * S::FLOATS + S::flat(ln, e);
            for (int w = 0; w < nwarps; w++) sum += smem[(size_t)F * S::FLOATS * (1 + w) + off];
            part[((size_t)f * C + c) * R + (e % R)] = sum;
        }
    }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(int P, size_t n, const float* __restrict__ part, float* __restrict__ out)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < P; p++) s += part[(size_t)p * n + t];
        out[t] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// generic fallbacks (any r, any F): one thread per output element, parallel over the whole grid
__global__ void __launch_bounds__(256)
conv_fwd_generic(int B, int N, int M, int C, int r, int K, const int* __restrict__ nn_index,
                 const int* __restrict__ nn_count, const int* __restrict__ bin_index,
                 const float* __restrict__ input, const float* __restrict__ filter, float* __restrict__ output)
{
    const int Co = C * r;
    const size_t total = (size_t)B * M * Co;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t row = t / Co;
        int co = (int)(t - row * Co), ci = co / r, b = (int)(row / M);
        int cnt = __ldg(nn_count + row);
        float acc = 0.f;
        for (int k = 0; k < cnt; k++) {
            int n = __ldg(nn_index + row * K + k), f = __ldg(bin_index + row * K + k);
            acc = fmaf(__ldg(input + ((size_t)b * N + n) * C + ci), __ldg(filter + (size_t)f * Co + co), acc);
        }
        output[t] = cnt > 0 ? acc / (float)cnt : 0.f;
    }
}

__global__ void __launch_bounds__(256)
conv_bwd_generic(int B, int N, int M, int C, int r, int K, const int* __restrict__ nn_index,
                 const int* __restrict__ nn_count, const int* __restrict__ bin_index,
                 const float* __restrict__ input, const float* __restrict__ filter,
                 const float* __restrict__ grad_output, float* __restrict__ grad_input,
                 float* __restrict__ grad_filter)
{
    const int Co = C * r;
    const size_t total = (size_t)B * M * Co;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        size_t row = t / Co;
        int co = (int)(t - row * Co), ci = co / r, b = (int)(row / M);
        int cnt = __ldg(nn_count + row);
        if (cnt <= 0) continue;
        float g = __ldg(grad_output + t) / (float)cnt;
        for (int k = 0; k < cnt; k++) {
            int n = __ldg(nn_index + row * K + k), f = __ldg(bin_index + row * K + k);
            size_t ii = ((size_t)b * N + n) * C + ci;
            atomicAdd(grad_input + ii, g * __ldg(filter + (size_t)f * Co + co));
            atomicAdd(grad_filter + (size_t)f * Co + co, g * __ldg(input + ii));
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct ConvPlan {
    int vec;          // 4 / 2 / 1, 0 = generic
    int chunks;       // gridDim.y
    int grid_x;       // persistent CTAs along rows
    int warps;        // warps per CTA (backward)
    size_t smem;      // dynamic shared memory bytes
};

static inline int pick_vec(int C) { return (C % 4 == 0) ? 4 : ((C % 2 == 0) ? 2 : 1); }
static const size_t SMEM_CAP = 227 * 1024;

static ConvPlan plan_fwd(int B, int M, int F, int C, int r)
{
    ConvPlan p{0, 0, 0, CONV_WARPS, 0};
    if (r != 1 && r != 2) return p;
    int vec = pick_vec(C);
    size_t smem = (size_t)F * 32 * vec * r * sizeof(float);
    while (smem > SMEM_CAP && vec > 1) { vec >>= 1; smem >>= 1; }
    if (smem > SMEM_CAP) return p;
    p.vec = vec; p.smem = smem;
    p.chunks = (C + 32 * vec - 1) / (32 * vec);
    long long tiles = ((long long)B * M + CONV_WARPS - 1) / CONV_WARPS;
    long long want = (long long)sm_count() * 8 / p.chunks;          // ~8 resident CTAs of 8 warps per SM
    if (want < 1) want = 1;
    p.grid_x = (int)(tiles < want ? tiles : want);
    return p;
}

static ConvPlan plan_bwd(int B, int M, int F, int C, int r)
{
    ConvPlan p{0, 0, 0, 0, 0};
    if (r != 1 && r != 2) return p;
    int vec = pick_vec(C);
    size_t strip = (size_t)F * 32 * vec * r * sizeof(float);
    while (strip * 5 > SMEM_CAP && vec > 1) { vec >>= 1; strip >>= 1; }     // want >= 4 warps
    if (strip * 3 > SMEM_CAP) return p;
    int warps = (int)(SMEM_CAP / strip) - 1;
    if (warps > 16) warps = 16;
    p.vec = vec; p.warps = warps; p.smem = strip * (1 + warps);
    p.chunks = (C + 32 * vec - 1) / (32 * vec);
    long long tiles = ((long long)B * M + warps - 1) / warps;
    int per_sm = (int)(SMEM_CAP / p.smem); if (per_sm < 1) per_sm = 1; if (per_sm > 4) per_sm = 4;
    long long want = (long long)sm_count() * per_sm / p.chunks;
    if (want < 1) want = 1;
    p.grid_x = (int)(tiles < want ? tiles : want);
    return p;
}

template <typename Kern>
static cudaError_t set_smem(Kern k, size_t bytes)
{
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace sph3d

using namespace sph3d;

static int conv_args_bad(int B, int N, int M, int F, int C, int r, int K)
{
    return B <= 0 || N <= 0 || M <= 0 || F <= 0 || C <= 0 || r <= 0 || K <= 0;
}

extern "C" int sph3d_depthwise_conv3d(int B, int N, int M, int F, int C, int r, int K,
                                      const int* nn_index, const int* nn_count, const int* bin_index,
                                      const float* input, const float* filter, float* output, void* stream)
{
    g_last_launch_count = 0;
    if (conv_args_bad(B, N, M, F, C, r, K) || !nn_index || !nn_count || !bin_index || !input || !filter || !output)
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    ConvPlan p = plan_fwd(B, M, F, C, r);
    if (p.vec == 0) {
        size_t total = (size_t)B * M * C * r;
        size_t want = (total + 255) / 256, cap = (size_t)sm_count() * 16;
        conv_fwd_generic<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(B, N, M, C, r, K, nn_index, nn_count,
                                                                              bin_index, input, filter, output);
        SPH3D_CHECK_LAUNCH();
        g_last_launch_count = 1;
        return 0;
    }
    dim3 grid(p.grid_x, p.chunks);
    cudaError_t e = cudaSuccess;
#define LAUNCH_FWD(V, RR)                                                                            \
    do {                                                                                             \
        e = set_smem(conv_fwd_kernel<V, RR>, p.smem);                                                \
        if (e != cudaSuccess) return (int)e;                                                         \
        conv_fwd_kernel<V, RR><<<grid, CONV_WARPS * 32, p.smem, st>>>(B, N, M, F, C, K, nn_index,    \
                                                                      nn_count, bin_index, input,    \
                                                                      filter, output);               \
    } while (0)
    if (p.vec == 4 && r == 1) LAUNCH_FWD(4, 1);
    else if (p.vec == 4 && r == 2) LAUNCH_FWD(4, 2);
    else if (p.vec == 2 && r == 1) LAUNCH_FWD(2, 1);
    else if (p.vec == 2 && r == 2) LAUNCH_FWD(2, 2);
    else if (p.vec == 1 && r == 1) LAUNCH_FWD(1, 1);
    else LAUNCH_FWD(1, 2);
#undef LAUNCH_FWD
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}

extern "C" size_t sph3d_depthwise_conv3d_grad_workspace_bytes(int B, int N, int M, int F, int C, int r, int K)
{
    if (conv_args_bad(B, N, M, F, C, r, K)) return 0;
    ConvPlan p = plan_bwd(B, M, F, C, r);
    if (p.vec == 0) return 0;
    return (size_t)p.grid_x * F * C * r * sizeof(float);
}

extern "C" int sph3d_depthwise_conv3d_grad(int B, int N, int M, int F, int C, int r, int K,
                                           const int* nn_index, const int* nn_count, const int* bin_index,
                                           const float* input, const float* filter, const float* grad_output,
                                           float* grad_input, float* grad_filter,
                                           void* workspace, size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    if (conv_args_bad(B, N, M, F, C, r, K) || !nn_index || !nn_count || !bin_index || !input || !filter ||
        !grad_output || !grad_input || !grad_filter)
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_input, 0, sizeof(float) * (size_t)B * N * C, st);
    if (e != cudaSuccess) return (int)e;
    ConvPlan p = plan_bwd(B, M, F, C, r);
    if (p.vec == 0) {
        e = cudaMemsetAsync(grad_filter, 0, sizeof(float) * (size_t)F * C * r, st);
        if (e != cudaSuccess) return (int)e;
        size_t total = (size_t)B * M * C * r;
        size_t want = (total + 255) / 256, cap = (size_t)sm_count() * 16;
        conv_bwd_generic<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(B, N, M, C, r, K, nn_index, nn_count,
                                                                              bin_index, input, filter, grad_output,
                                                                              grad_input, grad_filter);
        SPH3D_CHECK_LAUNCH();
        g_last_launch_count = 1;
        return 0;
    }
    size_t need = (size_t)p.grid_x * F * C * r * sizeof(float);
    if (!workspace || workspace_bytes < need) return (int)cudaErrorInvalidValue;
    dim3 grid(p.grid_x, p.chunks);
    float* part = (float*)workspace;
#define LAUNCH_BWD(V, RR)                                                                            \
    do {                                                                                             \
        e = set_smem(conv_bwd_kernel<V, RR>, p.smem);                                                \
        if (e != cudaSuccess) return (int)e;                                                         \
        conv_bwd_kernel<V, RR><<<grid, p.warps * 32, p.smem, st>>>(B, N, M, F, C, K, nn_index,       \
                                                                   nn_count, bin_index, input,       \
                                                                   filter, grad_output, grad_input,  \
                                                                   part);                            \
    } while (0)
    if (p.vec == 4 && r == 1) LAUNCH_BWD(4, 1);
    else if (p.vec == 4 && r == 2) LAUNCH_BWD(4, 2);
    else if (p.vec == 2 && r == 1) LAUNCH_BWD(2, 1);
    else if (p.vec == 2 && r == 2) LAUNCH_BWD(2, 2);
    else if (p.vec == 1 && r == 1) LAUNCH_BWD(1, 1);
    else LAUNCH_BWD(1, 2);
#undef LAUNCH_BWD
    SPH3D_CHECK_LAUNCH();
    size_t n = (size_t)F * C * r;
    reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.grid_x, n, part, grad_filter);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 2;        // kernels only (the cudaMemsetAsync of grad_input is not counted)
    return 0;
}

extern "C" int sph3d_abi_version(void) { return SPH3D_B200_ABI_VERSION; }
extern "C" int sph3d_last_launch_count(void) { return g_last_launch_count; }
