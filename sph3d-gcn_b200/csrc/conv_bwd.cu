// conv_bwd.cu -- depthwise spherical graph convolution, backward (grad_input + grad_filter), sm_100a.
//
// Replaces depthwiseConv3dGradLauncher (/root/reference/tf_ops/convolution/tf_conv3d_gpu.cu:115-140,
// kernels depthwise_input_backward :32-55 and depthwise_filter_backward :58-101) and the cudaMemset
// zero fills of tf_conv3d.cpp:152-153.
//
//   grad_input [b, nn[b,m,k], c]  += sum_j gO[b,m,c*r+j] * W[bin[b,m,k], c, j] / cnt
//   grad_filter[bin[b,m,k], c, j] +=       gO[b,m,c*r+j] * in[b, nn[b,m,k], c] / cnt
//
// ONE fused pass.  A group of G warps (G = 1, 2, 4 or 8, the smallest that keeps the accumulators in
// registers) shares one output row; warp w of the group OWNS the filter bins f = w, w+G, w+2G, ...
// (SLOTS of them).  For each of its bins that occurs in the row (ballot) it
//   - forms d = sum_j g*W[f] once (filter strip from shared memory),
//   - walks the bin's edges: gathers the input strip (LDG.128), adds it to a running sum (FADD2) and
//     scatters d into grad_input with ONE 16-byte vector reduction per lane (REDG.ADD.F32x4),
//   - folds g * sum(in) into its REGISTER accumulator for that bin.
// Because bins are owned, the filter gradient needs no atomics and no shared-memory accumulation at
// all: every warp keeps SLOTS x (VEC*R) accumulators in registers for the whole kernel and writes
// them once, as a per-group partial that a second tiny kernel sums in a fixed order (deterministic
// grad_filter).  Measured alternatives that lost on B200 (profiles/): sorting the tile by bin first
// (as the forward kernel does) costs more than it saves when two warps share a row, and replacing the
// vector reductions by TMA bulk reductions (UBLKRED.G.S.ADD.F32, one 512-byte cp.reduce.async.bulk per
// edge) is 2.4x slower: the TMA unit sustains only about one small bulk op per ~80 cycles per SM.
// The reference instead issues E*C*r shared-memory float atomics per 48 KB filter
// window and re-runs the whole pass ceil(F*C*r/12288) times (Q13/Q14).
#include "conv_common.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

constexpr int BWD_THREADS = 512;

template <int VEC>
__device__ __forceinline__ void red_strip(char* __restrict__ base, unsigned off, const float (&d)[VEC])
{
    VecIO<VEC>::red(reinterpret_cast<float*>(base + off), d);
}

// threads per CTA the register budget allows: 9 bins/warp -> <= 85 registers -> 24 warps; 17 -> 16 warps
__host__ __device__ constexpr int bwd_max_threads(int vec, int r, int slots) { return (slots * vec * r <= 36) ? 768 : BWD_THREADS; }

template <int VEC, int R, int SLOTS>
__global__ void __launch_bounds__(bwd_max_threads(VEC, R, SLOTS), 1)
conv_bwd_kernel(unsigned rows, unsigned rpc, int N, unsigned M, int F, int C, int K, int G,
                const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                const int* __restrict__ bin_index, const float* __restrict__ input,
                const float* __restrict__ filter, const float* __restrict__ grad_output,
                float* __restrict__ grad_input, float* __restrict__ gw_partial, int cta_reduce)
{
    constexpr int E = VEC * R;
    using S = SmemStrip<E>;
    extern __shared__ __align__(16) float smem[];
    float* Wsh = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wig = warp % G, group = warp / G, ngroups = (blockDim.x >> 5) / G;
    const int cbase = blockIdx.y * 32 * VEC;
    stage_filter<VEC, R>(Wsh, filter, F, C, cbase);
    __syncthreads();

    const int c0 = cbase + lane * VEC;
    const bool active = c0 < C;
    const int c0ld = active ? c0 : 0;
    const unsigned strideB = (unsigned)C * 4u;
    const size_t cloudB = (size_t)N * C * 4;
    const float* wlane = Wsh + S::offset(0, lane) + wig * S::FLOATS;      // strip of my first bin
    const int wstep = G * S::FLOATS;
    float acc[SLOTS][E];
#pragma unroll
    for (int s = 0; s < SLOTS; s++)
#pragma unroll
        for (int e = 0; e < E; e++) acc[s][e] = 0.f;

    const unsigned nchunks = (rows + rpc - 1) / rpc;
    for (unsigned chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const unsigned rbeg = chunk * rpc;
        const unsigned rend = min(rbeg + rpc, rows);
        unsigned row = rbeg + group;
        if (row >= rend) continue;
        RowCursor cur;
        cur.init(row, M);
        for (; row < rend; row += ngroups, cur.advance(ngroups, M)) {
            const int cnt = min(__ldg(nn_count + row), K);
            if (cnt <= 0) continue;
            const char* inb = reinterpret_cast<const char*>(input) + cur.b * cloudB + (size_t)c0ld * 4;
            char* gib = reinterpret_cast<char*>(grad_input) + cur.b * cloudB + (size_t)c0ld * 4;
            const int* idxrow = nn_index + (size_t)row * K;
            const int* binrow = bin_index + (size_t)row * K;
            float g[E];
            {
                const float inv = 1.0f / (float)cnt;
                const float* go = grad_output + ((size_t)row * C + c0ld) * R;
                constexpr int VW = strip_vw(E);
#pragma unroll
                for (int pl = 0; pl < E / VW; pl++) {
                    float t[VW];
                    VecIO<VW>::ld(t, go + pl * VW, true);
#pragma unroll
                    for (int u = 0; u < VW; u++) g[pl * VW + u] = active ? t[u] * inv : 0.f;
                }
            }
            for (int kt = 0; kt < cnt; kt += 64) {
                const int k0 = kt + lane, k1 = kt + 32 + lane;
                unsigned o0 = 0, o1 = 0;
                int b0 = -1, b1 = -1;
                if (k0 < cnt) { o0 = (unsigned)__ldg(idxrow + k0) * strideB; b0 = __ldg(binrow + k0); }
                if (k1 < cnt) { o1 = (unsigned)__ldg(idxrow + k1) * strideB; b1 = __ldg(binrow + k1); }
                // bins of mine that occur in this tile: bit s <-> bin wig + s*G
                int f = wig;
#pragma unroll
                for (int s = 0; s < SLOTS; s++, f += G) {
                    const unsigned m0 = __ballot_sync(FULL_MASK, b0 == f);
                    const unsigned m1 = __ballot_sync(FULL_MASK, b1 == f);
                    if (m0 | m1) {                      // (f >= F never matches: bins are < F)
                        float w[E];
                        S::load(w, wlane + s * wstep, 0);
                        float d[VEC], sv[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; v++) {
                            float t = 0.f;
#pragma unroll
                            for (int j = 0; j < R; j++) t = fmaf(g[v * R + j], w[v * R + j], t);
                            d[v] = t; sv[v] = 0.f;
                        }
                        auto edges = [&](unsigned m, unsigned myoff) {
                            while (m) {
                                const unsigned q0 = __shfl_sync(FULL_MASK, myoff, pop_highest(m));
                                float v0[VEC];
                                ld_strip<VEC>(v0, inb, q0);
                                if (active) red_strip<VEC>(gib, q0, d);
                                if (m) {
                                    const unsigned q1 = __shfl_sync(FULL_MASK, myoff, pop_highest(m));
                                    float v1[VEC];
                                    ld_strip<VEC>(v1, inb, q1);
                                    if (active) red_strip<VEC>(gib, q1, d);
                                    strip_add<VEC>(sv, v1);
                                }
                                strip_add<VEC>(sv, v0);
                            }
                        };
                        edges(m0, o0);
                        edges(m1, o1);
#pragma unroll
                        for (int e = 0; e < E; e++) acc[s][e] = fmaf(g[e], sv[e / R], acc[s][e]);
                    }
                }
            }
        }
    }
    if (cta_reduce) {
        // sum the groups of this CTA in shared memory (fixed group order) and emit ONE partial per CTA:
        // 8x fewer partial bytes for the second-stage reduction than one partial per group
        float* gsm = smem + (size_t)F * S::FLOATS;                   // [ngroups][F][strip]
        __syncthreads();
        {
            int f = wig;
#pragma unroll
            for (int s = 0; s < SLOTS; s++, f += G)
                if (f < F) S::store(gsm + ((size_t)group * F + f) * S::FLOATS, lane, acc[s]);
        }
        __syncthreads();
        float* part = gw_partial + (size_t)blockIdx.x * F * C * R;
        for (int t = threadIdx.x; t < F * S::FLOATS; t += blockDim.x) {
            const int f = t / S::FLOATS, rem = t % S::FLOATS;
            const int ln = rem / E, e = rem % E;
            const int c = cbase + ln * VEC + e / R;
            if (c < C) {
                float sum = 0.f;
                const int off = f * S::FLOATS + S::flat(ln, e);
                for (int gq = 0; gq < ngroups; gq++) sum += gsm[(size_t)gq * F * S::FLOATS + off];
                part[((size_t)f * C + c) * R + (e % R)] = sum;
            }
        }
        return;
    }
    // partial [blockIdx.x][group][f][c][j]: each (f, c-chunk) is written by exactly one warp of the group
    float* part = gw_partial + ((size_t)blockIdx.x * ngroups + group) * F * C * R;
    if (active) {
        int f = wig;
#pragma unroll
        for (int s = 0; s < SLOTS; s++, f += G) {
            if (f < F) {
                float* dst = part + ((size_t)f * C + c0) * R;
                constexpr int VW = strip_vw(E);
#pragma unroll
                for (int pl = 0; pl < E / VW; pl++) {
                    float t[VW];
#pragma unroll
                    for (int u = 0; u < VW; u++) t[u] = acc[s][pl * VW + u];
                    VecIO<VW>::st(dst + pl * VW, t);
                }
            }
        }
    }
}

// second stage: out[t] = sum_p part[p][t] in a FIXED order (bit-reproducible).  A CTA owns 32 consecutive
// elements; its 8 warps split the partials (warp w takes p = w, w+8, ...), then warp 0 adds the 8 sums.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(int P, size_t n, const float* __restrict__ part, float* __restrict__ out)
{
    __shared__ float sm[8][32];
    const int tx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const size_t t = (size_t)blockIdx.x * 32 + tx;
    float s0 = 0.f, s1 = 0.f;
    if (t < n) {
        int p = py;
        for (; p + 8 < P; p += 16) { s0 += part[(size_t)p * n + t]; s1 += part[(size_t)(p + 8) * n + t]; }
        if (p < P) s0 += part[(size_t)p * n + t];
    }
    sm[py][tx] = s0 + s1;
    __syncthreads();
    if (py == 0 && t < n) {
        float s = sm[0][tx];
#pragma unroll
        for (int w = 1; w < 8; w++) s += sm[w][tx];
        out[t] = s;
    }
}

// generic fallback (any r, any F, any size): float atomics, like the reference but grid-parallel
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
        int cnt = min(__ldg(nn_count + row), K);
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

// plan: vec in {4,2,1}, SLOTS in {9,17}, G in {1,2,4,8} with G*SLOTS >= F and SLOTS*vec*r <= 72 registers
static ConvPlan plan_bwd(int B, int N, int M, int F, int C, int r, int* G_out)
{
    ConvPlan p{};
    *G_out = 0;
    if ((r != 1 && r != 2) || !fits_32bit(B, N, M, C, r) || F > 136) return p;
    int vec = pick_vec_full_warp(C), slots = 0, G = 0;
    {   // sweep knob: force a strip width (must divide C)
        int v_env = tun(tunables().bwd_vec, vec);
        if ((v_env == 1 || v_env == 2 || v_env == 4) && C % v_env == 0) vec = v_env;
    }
    // narrow strips (vec <= 2: C <= 64) keep few accumulators per bin: 9 bins per warp (4 warps per row) then fit in
    // <= 36 registers and 24 warps per SM, which measured faster than 17 bins per warp (0.38 vs 0.48 ms, C=64, r=2)
    if (vec <= 2 && 9 * vec * r <= 36 && 4 * 9 >= F && tun(tunables().bwd_g, 4) == 4) { G = 4; slots = 9; }
    for (; vec >= 1 && !G; vec = (vec > 1 ? vec >> 1 : 0)) {
        const int slot_opts[2] = {17, 9};
        for (int i = 0; i < 2 && !G; i++) {
            if (slot_opts[i] * vec * r > 72) continue;
            for (int g = 1; g <= 8; g <<= 1)
                if (g * slot_opts[i] >= F) { G = g; slots = slot_opts[i]; break; }
        }
        if (G) break;
        if (vec == 1) break;
    }
    if (!G) return p;
    {   // sweep knob: a larger group (fewer bins per warp) is always legal
        int g_env = tun(tunables().bwd_g, G);
        if ((g_env == 1 || g_env == 2 || g_env == 4 || g_env == 8) && g_env > G) G = g_env;
        if (G * 9 >= F) slots = 9;                      // fewer bins per warp -> fewer registers -> more warps
    }
    size_t smem = (size_t)F * 32 * vec * r * sizeof(float);
    if (smem > SMEM_CAP) return p;
    p.vec = vec; p.slots = slots; p.smem = smem;
    p.chunks = (C + 32 * vec - 1) / (32 * vec);
    const int max_threads = bwd_max_threads(vec, r, slots) / (32 * G) * (32 * G);
    p.threads = tun(tunables().bwd_threads, max_threads);
    if (p.threads > max_threads || p.threads % (32 * G)) p.threads = max_threads;
    const long long rows = (long long)B * M;
    long long want = sm_count();
    if (p.chunks > 1) want = (want + p.chunks - 1) / p.chunks;
    if (want < 1) want = 1;
    // a warp group walks its rows one after the other: small problems get chunks of >= 2 rows per group so that all SMs work
    const int rpc = pick_rows_per_chunk(rows, want, 2 * (p.threads / (32 * G)));
    p.rpc = rpc;
    const long long nchunks = (rows + rpc - 1) / rpc;
    p.grid_x = (int)(nchunks < want ? nchunks : want);
    // tiny problems: fewer warp GROUPS per CTA (a CTA must stay a whole number of G-warp groups: a partial group would
    // re-process a row with some of its bins and write its partial past the end of the workspace)
    {
        int groups = p.threads / (32 * G);
        while (groups > 1 && groups * 32 * G > 64 &&
               (long long)p.grid_x * p.chunks * groups > rows && p.grid_x * p.chunks < sm_count())
            groups = (groups + 1) / 2;
        p.threads = groups * 32 * G;
    }
    {   // Summing the CTA's groups in shared memory (one partial per CTA) is OFF by default: the 8 extra
        // filter-sized slabs move the L1/shared carve-out from ~200 KB of L1 to ~76 KB and the gathers
        // lose more (2.27 -> 2.42 ms at Cfg-T) than the smaller second-stage reduction saves.
        const size_t need = smem * (1 + (size_t)(p.threads / 32 / G));
        if (need <= SMEM_CAP && tunables().bwd_cta_reduce == 1) { p.cta_reduce = 1; p.smem = need; }
    }
    *G_out = G;
    return p;
}

static size_t bwd_partials(const ConvPlan& p, int G)
{
    return p.cta_reduce ? (size_t)p.grid_x : (size_t)p.grid_x * (p.threads / 32 / G);
}

int launch_reduce_partials(int P, size_t n, const float* part, float* out, cudaStream_t st)
{
    reduce_partials_kernel<<<(unsigned)((n + 31) / 32), 256, 0, st>>>(P, n, part, out);
    SPH3D_CHECK_LAUNCH();
    return 0;
}

}  // namespace sph3d

using namespace sph3d;

extern "C" size_t sph3d_depthwise_conv3d_grad_workspace_bytes(int B, int N, int M, int F, int C, int r, int K)
{
    if (B <= 0 || N <= 0 || M <= 0 || F <= 0 || C <= 0 || r <= 0 || K <= 0) return 0;
    if (bwd_transposed_supported(B, N, M, F, C, r, K)) return bwd_transposed_workspace_bytes(B, N, M, F, C, r, K);
    int G = 0;
    ConvPlan p = plan_bwd(B, N, M, F, C, r, &G);
    if (p.vec == 0) return 0;
    return bwd_partials(p, G) * F * C * r * sizeof(float);
}

extern "C" int sph3d_depthwise_conv3d_grad(int B, int N, int M, int F, int C, int r, int K,
                                           const int* nn_index, const int* nn_count, const int* bin_index,
                                           const float* input, const float* filter, const float* grad_output,
                                           float* grad_input, float* grad_filter,
                                           void* workspace, size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    if (B <= 0 || N <= 0 || M <= 0 || F <= 0 || C <= 0 || r <= 0 || K <= 0 || !nn_index || !nn_count ||
        !bin_index || !input || !filter || !grad_output || !grad_input || !grad_filter)
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    if (bwd_transposed_supported(B, N, M, F, C, r, K))
        return bwd_transposed_run(B, N, M, F, C, r, K, nn_index, nn_count, bin_index, input, filter, grad_output,
                                  grad_input, grad_filter, workspace, workspace_bytes, st);
    cudaError_t e = cudaMemsetAsync(grad_input, 0, sizeof(float) * (size_t)B * N * C, st);
    if (e != cudaSuccess) return (int)e;
    int G = 0;
    ConvPlan p = plan_bwd(B, N, M, F, C, r, &G);
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
    const size_t nW = (size_t)F * C * r;
    const size_t P = bwd_partials(p, G);
    if (!workspace || workspace_bytes < P * nW * sizeof(float)) return (int)cudaErrorInvalidValue;
    dim3 grid(p.grid_x, p.chunks);
    const unsigned rows = (unsigned)((long long)B * M);
    const unsigned rpc = (unsigned)p.rpc;
    float* part = (float*)workspace;
#define LAUNCH_BWD(V, RR, SL)                                                                        \
    do {                                                                                             \
        e = set_smem(conv_bwd_kernel<V, RR, SL>, p.smem);                                            \
        if (e != cudaSuccess) return (int)e;                                                         \
        conv_bwd_kernel<V, RR, SL><<<grid, p.threads, p.smem, st>>>(rows, rpc, N, (unsigned)M, F, C, K, \
                                                                    G, nn_index, nn_count,           \
                                                                    bin_index, input, filter,        \
                                                                    grad_output, grad_input, part,   \
                                                                    p.cta_reduce);                   \
    } while (0)
#define DISPATCH_SLOTS(V, RR)                                                                        \
    do {                                                                                             \
        if (p.slots == 9) LAUNCH_BWD(V, RR, 9);                                                      \
        else LAUNCH_BWD(V, RR, 17);                                                                  \
    } while (0)
    if (p.vec == 4 && r == 1) DISPATCH_SLOTS(4, 1);
    else if (p.vec == 4 && r == 2) LAUNCH_BWD(4, 2, 9);
    else if (p.vec == 2 && r == 1) DISPATCH_SLOTS(2, 1);
    else if (p.vec == 2 && r == 2) DISPATCH_SLOTS(2, 2);
    else if (p.vec == 1 && r == 1) DISPATCH_SLOTS(1, 1);
    else DISPATCH_SLOTS(1, 2);
#undef DISPATCH_SLOTS
#undef LAUNCH_BWD
    SPH3D_CHECK_LAUNCH();
    reduce_partials_kernel<<<(unsigned)((nW + 31) / 32), 256, 0, st>>>((int)P, nW, part, grad_filter);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 2;        // kernels only (the cudaMemsetAsync of grad_input is not counted)
    return 0;
}
