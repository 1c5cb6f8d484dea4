// conv_fwd.cu -- depthwise spherical graph convolution, forward, sm_100a.
//
// Replaces depthwiseConv3dLauncher (/root/reference/tf_ops/convolution/tf_conv3d_gpu.cu:107-113,
// kernel :7-29) and the cudaMemset zero fill of tf_conv3d.cpp:90.
//
//   out[b,m,c*r+j] = (1/cnt) * sum_{k<cnt} in[b, nn[b,m,k], c] * W[bin[b,m,k], c, j]          (Q9)
//
// Design (work unit: rowwarp.cuh).  A warp owns one output point and 32*VEC input channels.  It
// reads the point's neighbour ids and bin ids ONCE (coalesced, two per lane per 64-edge tile), then
// walks only the bins that occur in the row (64-bit presence mask from one REDUX.OR): a ballot
// selects the bin's edges, their feature strips are gathered (one LDG.128 per lane and edge, up to
// four in flight per warp) and SUMMED with packed FADD2, and the bin's filter strip -- staged once per
// persistent CTA in shared memory, conflict-free layout -- is applied once per (row, bin) with
// FFMA2 instead of once per edge.  This is the segment-weighted-sum form of the op: FMA count drops
// from E*C*r to (#row-bins)*C*r, shared-memory filter traffic drops by the mean segment length
// (~4x at K=64, F=33); what remains is the irreducible gather of E*C*4 bytes through L1/L2.
// Nothing is accumulated in global memory (the reference does a global read-modify-write per edge
// and channel) and index rows are read once per 32*VEC channels, not once per channel.
//
// One persistent CTA of up to 32 warps per SM takes CONTIGUOUS chunks of rows: the reference's
// "first K in-range points by ascending index" rule (Q4) makes neighbouring rows draw their
// neighbours from the same low-index set, so a chunk's gathers mostly hit the SM's L1.
#include "conv_common.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

thread_local int g_last_launch_count = 0;

static int env_int(const char* name, int unset)
{
    const char* v = getenv(name);
    if (!v || !*v) return unset;
    return atoi(v);
}
static Tunables read_tunables()
{
    Tunables t{};
    t.rows_per_chunk = env_int("SPH3D_ROWS_PER_CHUNK", 0);
    t.fwd_threads = env_int("SPH3D_FWD_THREADS", 0);
    t.fwd_vec = env_int("SPH3D_FWD_VEC", 0);
    t.bwd_algo = env_int("SPH3D_BWD_ALGO", 0);
    t.bwd_vec = env_int("SPH3D_BWD_VEC", 0);
    t.bwd_g = env_int("SPH3D_BWD_G", 0);
    t.bwd_threads = env_int("SPH3D_BWD_THREADS", 0);
    t.bwd_cta_reduce = env_int("SPH3D_BWD_CTA_REDUCE", 0);
    t.bwdt_threads = env_int("SPH3D_BWDT_THREADS", 0);
    t.bwdt_depth = env_int("SPH3D_BWDT_DEPTH", 0);
    t.bwdt_g = env_int("SPH3D_BWDT_G", 0);
    t.bwdt_sort = env_int("SPH3D_BWDT_SORT", 0);
    t.bwdt_fold = env_int("SPH3D_BWDT_FOLD", 0);
    t.nnquery_grid = env_int("SPH3D_NNQUERY_GRID", -1);
    t.pool_stream = env_int("SPH3D_POOL_STREAM", -1);
    t.fps_handshake = env_int("SPH3D_FPS_HANDSHAKE", -1);
    t.fps_cluster_min_n = env_int("SPH3D_FPS_CLUSTER_MIN_N", 0);
    t.fwd_smem_pad_kb = env_int("SPH3D_FWD_SMEM_PAD_KB", 0);
    t.sepconv_tile = env_int("SPH3D_SEPCONV_TILE", 0);
    t.sepconv_stages = env_int("SPH3D_SEPCONV_STAGES", 0);
    return t;
}
static Tunables g_tunables = read_tunables();             // once, at library load
const Tunables& tunables() { return g_tunables; }

// PLANNED = false: nn_index / bin_index are the graph tensors and every 64-edge tile is counting-sorted by bin in
// shared memory.  PLANNED = true: `nn_index` points at the sorted edge words written by conv_sort_kernel
// (word = neighbour id << 8 | bin << 1 | last-of-its-bin), `bin_index` is unused, and the next row's words are
// prefetched while the current row gathers.
template <int VEC, int R, bool PLANNED>
__global__ void __launch_bounds__(1024, 1)
conv_fwd_kernel(unsigned rows, unsigned rpc, int N, unsigned M, int F, int C, int K,
                const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                const int* __restrict__ bin_index, const float* __restrict__ input,
                const float* __restrict__ filter, float* __restrict__ output)
{
    constexpr int E = VEC * R;
    using S = SmemStrip<E>;
    extern __shared__ __align__(16) float smem[];
    float* Wsh = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cbase = blockIdx.y * 32 * VEC;
    stage_filter<VEC, R>(Wsh, filter, F, C, cbase);
    __syncthreads();

    const int c0 = cbase + lane * VEC;
    const bool active = c0 < C;
    const int c0ld = active ? c0 : 0;                      // idle lanes load a valid strip, never store
    const unsigned strideB = (unsigned)C * 4u;
    const size_t cloudB = (size_t)N * C * 4;
    const float* wlane = Wsh + S::offset(0, lane);
    const unsigned nchunks = (rows + rpc - 1) / rpc;
    // per-warp scratch behind the filter: the sorted tile (+ the histograms of the bin sort when it runs here)
    const int FP = ((F + 31) / 32) * 32;
    int* hA = reinterpret_cast<int*>(Wsh + (size_t)F * S::FLOATS) + (size_t)warp * (PLANNED ? 128 : sort_smem_ints(F));
    int* hB = hA + (FP + 1 + 3) / 4 * 4;
    unsigned* sOff = PLANNED ? reinterpret_cast<unsigned*>(hA) : reinterpret_cast<unsigned*>(hB + FP);
    int* sCode = reinterpret_cast<int*>(sOff + 64);

    for (unsigned chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const unsigned rbeg = chunk * rpc;
        const unsigned rend = min(rbeg + rpc, rows);
        unsigned row = rbeg + warp;
        if (row >= rend) continue;
        RowCursor cur;
        cur.init(row, M);
        // planned form: count and first tile of the row after this one are loaded one row ahead
        int cnt_n = 0;
        unsigned w0_n = 0, w1_n = 0;
        auto prefetch = [&](unsigned rw) {
            cnt_n = min(__ldg(nn_count + rw), K);
            const int* wr = nn_index + (size_t)rw * K;
            w0_n = (lane < cnt_n) ? (unsigned)__ldg(wr + lane) : 0u;
            w1_n = (32 + lane < cnt_n) ? (unsigned)__ldg(wr + 32 + lane) : 0u;
        };
        if constexpr (PLANNED) prefetch(row);
        for (; row < rend; row += nwarps, cur.advance(nwarps, M)) {
            int cnt;
            unsigned w0_c = 0, w1_c = 0;
            if constexpr (PLANNED) {
                cnt = cnt_n; w0_c = w0_n; w1_c = w1_n;
                if (row + nwarps < rend) prefetch(row + nwarps);
            } else {
                cnt = min(__ldg(nn_count + row), K);
            }
            const char* inb = reinterpret_cast<const char*>(input) + cur.b * cloudB + (size_t)c0ld * 4;
            const int* idxrow = nn_index + (size_t)row * K;
            const int* binrow = bin_index + (size_t)row * K;
            float acc[E];
#pragma unroll
            for (int e = 0; e < E; e++) acc[e] = 0.f;

            for (int kt = 0; kt < cnt; kt += 64) {
                const int k0 = kt + lane, k1 = kt + 32 + lane;
                if constexpr (PLANNED) {
                    unsigned w0 = w0_c, w1 = w1_c;
                    if (kt != 0) {                                // rows longer than one tile (K > 64): later tiles on demand
                        w0 = (k0 < cnt) ? (unsigned)__ldg(idxrow + k0) : 0u;
                        w1 = (k1 < cnt) ? (unsigned)__ldg(idxrow + k1) : 0u;
                    }
                    sOff[lane] = (w0 >> 8) * strideB;      sCode[lane] = (int)(w0 & 255u);
                    sOff[32 + lane] = (w1 >> 8) * strideB; sCode[32 + lane] = (int)(w1 & 255u);
                    __syncwarp();
                } else {
                    unsigned o0 = 0, o1 = 0;
                    int b0 = -1, b1 = -1;
                    if (k0 < cnt) { o0 = (unsigned)__ldg(idxrow + k0) * strideB; b0 = __ldg(binrow + k0); }
                    if (k1 < cnt) { o1 = (unsigned)__ldg(idxrow + k1) * strideB; b1 = __ldg(binrow + k1); }
                    // group the tile's edges by bin (shared-memory counting sort), then one flat gather loop
                    sort_tile_by_bin(o0, b0, o1, b1, FP, lane, hA, hB, sOff, sCode);
                }
                const int nt = min(64, cnt - kt);
                float s[VEC];
#pragma unroll
                for (int v = 0; v < VEC; v++) s[v] = 0.f;
                auto consume = [&](const float (&v)[VEC], int code) {
                    strip_add<VEC>(s, v);
                    if (code & 1) {                               // last edge of its bin: apply the filter strip once
                        float w[E];
                        S::load(w, wlane + (code >> 1) * S::FLOATS, 0);
                        if constexpr (E % 2 == 0) {
#pragma unroll
                            for (int e = 0; e < E; e += 2) {      // FFMA2: (acc[e],acc[e+1]) += (s,s') * (w[e],w[e+1])
                                float2 a = __ffma2_rn(make_float2(s[e / R], s[(e + 1) / R]), make_float2(w[e], w[e + 1]),
                                                      make_float2(acc[e], acc[e + 1]));
                                acc[e] = a.x; acc[e + 1] = a.y;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < E; e++) acc[e] = fmaf(s[e / R], w[e], acc[e]);
                        }
#pragma unroll
                        for (int v2 = 0; v2 < VEC; v2++) s[v2] = 0.f;
                    }
                };
                int p = 0;
                for (; p + 4 <= nt; p += 4) {                     // four independent gathers in flight
                    const uint4 oo = *reinterpret_cast<const uint4*>(sOff + p);
                    const int4 cc = *reinterpret_cast<const int4*>(sCode + p);
                    float v0[VEC], v1[VEC], v2[VEC], v3[VEC];
                    ld_strip<VEC>(v0, inb, oo.x); ld_strip<VEC>(v1, inb, oo.y);
                    ld_strip<VEC>(v2, inb, oo.z); ld_strip<VEC>(v3, inb, oo.w);
                    consume(v0, cc.x); consume(v1, cc.y); consume(v2, cc.z); consume(v3, cc.w);
                }
                for (; p < nt; p++) {
                    float v0[VEC];
                    ld_strip<VEC>(v0, inb, sOff[p]);
                    consume(v0, sCode[p]);
                }
                __syncwarp();                                     // sOff/sCode are rewritten by the next tile/row
            }
            if (active) {
                const float inv = cnt > 0 ? 1.0f / (float)cnt : 0.f;
#pragma unroll
                for (int e = 0; e < E; e++) acc[e] *= inv;
                float* out = output + ((size_t)row * C + c0) * R;
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
}

// Graph-only half of the forward pass: every row's edges grouped by bin (ascending bin, original k order inside a bin;
// tiles of 64 edges are sorted independently), packed as  neighbour id << 8 | bin << 1 | last-edge-of-its-bin  at the
// edge's new position row*K + p.  Words beyond nn_count are not written.  One warp per row, same counting sort as the
// one-call kernel, so the planned and the one-call forward sum in the same order (bit-identical outputs).
__global__ void __launch_bounds__(256)
conv_sort_kernel(unsigned rows, int K, int F, const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                 const int* __restrict__ bin_index, unsigned* __restrict__ words)
{
    extern __shared__ __align__(16) int ssm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int FP = ((F + 31) / 32) * 32;
    int* hA = ssm + (size_t)warp * sort_smem_ints(F);
    int* hB = hA + (FP + 1 + 3) / 4 * 4;
    unsigned* sOff = reinterpret_cast<unsigned*>(hB + FP);
    int* sCode = reinterpret_cast<int*>(sOff + 64);
    for (unsigned row = blockIdx.x * nwarps + warp; row < rows; row += gridDim.x * nwarps) {
        const int cnt = min(__ldg(nn_count + row), K);
        const int* idxrow = nn_index + (size_t)row * K;
        const int* binrow = bin_index + (size_t)row * K;
        unsigned* wrow = words + (size_t)row * K;
        for (int kt = 0; kt < cnt; kt += 64) {
            const int k0 = kt + lane, k1 = kt + 32 + lane;
            unsigned o0 = 0, o1 = 0;
            int b0 = -1, b1 = -1;
            if (k0 < cnt) { o0 = (unsigned)__ldg(idxrow + k0); b0 = __ldg(binrow + k0); }
            if (k1 < cnt) { o1 = (unsigned)__ldg(idxrow + k1); b1 = __ldg(binrow + k1); }
            if ((unsigned)b0 >= (unsigned)F) b0 = (k0 < cnt) ? 0 : -1;      // malformed bins fold into bin 0 (never out of bounds)
            if ((unsigned)b1 >= (unsigned)F) b1 = (k1 < cnt) ? 0 : -1;
            sort_tile_by_bin(o0, b0, o1, b1, FP, lane, hA, hB, sOff, sCode);
            if (k0 < cnt) wrow[k0] = (sOff[lane] << 8) | (unsigned)sCode[lane];
            if (k1 < cnt) wrow[k1] = (sOff[32 + lane] << 8) | (unsigned)sCode[32 + lane];
            __syncwarp();
        }
    }
}

// generic fallback (any r, any F, any size): one thread per output element
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
        int cnt = min(__ldg(nn_count + row), K);
        float acc = 0.f;
        for (int k = 0; k < cnt; k++) {
            int n = __ldg(nn_index + row * K + k), f = __ldg(bin_index + row * K + k);
            acc = fmaf(__ldg(input + ((size_t)b * N + n) * C + ci), __ldg(filter + (size_t)f * Co + co), acc);
        }
        output[t] = cnt > 0 ? acc / (float)cnt : 0.f;
    }
}

static ConvPlan plan_fwd(int B, int N, int M, int F, int C, int r, bool planned)
{
    ConvPlan p{};
    if ((r != 1 && r != 2) || !fits_32bit(B, N, M, C, r) || F > 128) return p;
    if (planned && N >= (1 << 24)) return p;                        // a sorted edge word keeps the neighbour id in 24 bits
    int vec = pick_vec_full_warp(C);
    {   // sweep knob: force a strip width (must divide C)
        int v_env = tunables().fwd_vec > 0 ? tunables().fwd_vec : vec;
        if ((v_env == 1 || v_env == 2 || v_env == 4) && C % v_env == 0) vec = v_env;
    }
    const size_t sort_bytes = (size_t)32 * (planned ? 128 : sort_smem_ints(F)) * sizeof(int);     // 32 warps
    size_t smem = (size_t)F * 32 * vec * r * sizeof(float);
    while (smem + sort_bytes > SMEM_CAP && vec > 1) { vec >>= 1; smem >>= 1; }
    if (smem + sort_bytes > SMEM_CAP) return p;
    smem += sort_bytes;
    if (tunables().fwd_smem_pad_kb > 0 && smem + (size_t)tunables().fwd_smem_pad_kb * 1024 <= SMEM_CAP)
        smem += (size_t)tunables().fwd_smem_pad_kb * 1024;        // sensitivity sweep only: the pad is never touched
    p.vec = vec; p.smem = smem;
    p.chunks = (C + 32 * vec - 1) / (32 * vec);
    const long long rows = (long long)B * M;
    long long want = sm_count();                                   // one persistent 32-warp CTA per SM ...
    if (p.chunks > 1) want = (want + p.chunks - 1) / p.chunks;      // ... shared by the channel chunks
    if (want < 1) want = 1;
    const int rpc = pick_rows_per_chunk(rows, want, 32);           // small problems: shorter chunks, more CTAs
    p.rpc = rpc;
    const long long nchunks = (rows + rpc - 1) / rpc;
    p.grid_x = (int)(nchunks < want ? nchunks : want);
    // small problems: fewer warps per CTA so that more SMs get work
    p.threads = tun(tunables().fwd_threads, 1024);
    if (p.threads > 1024 || p.threads % 32) p.threads = 1024;
    while (p.threads > 128 && (long long)p.grid_x * p.chunks * (p.threads / 32) > rows && p.grid_x * p.chunks < sm_count())
        p.threads >>= 1;
    return p;
}

}  // namespace sph3d

using namespace sph3d;

static int run_fwd(int B, int N, int M, int F, int C, int r, int K, const int* nn_index, const int* nn_count,
                   const int* bin_index, const float* input, const float* filter, float* output, bool planned,
                   cudaStream_t st)
{
    ConvPlan p = plan_fwd(B, N, M, F, C, r, planned);
    if (p.vec == 0) {
        if (planned) return (int)cudaErrorInvalidValue;           // sph3d_conv_sort_bytes returned 0 for this shape
        size_t total = (size_t)B * M * C * r;
        size_t want = (total + 255) / 256, cap = (size_t)sm_count() * 16;
        conv_fwd_generic<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(B, N, M, C, r, K, nn_index, nn_count,
                                                                              bin_index, input, filter, output);
        SPH3D_CHECK_LAUNCH();
        g_last_launch_count = 1;
        return 0;
    }
    dim3 grid(p.grid_x, p.chunks);
    const unsigned rows = (unsigned)((long long)B * M);
    const unsigned rpc = (unsigned)p.rpc;
    cudaError_t e = cudaSuccess;
#define LAUNCH_FWD2(V, RR, PL)                                                                          \
    do {                                                                                                 \
        e = set_smem(conv_fwd_kernel<V, RR, PL>, p.smem);                                                \
        if (e != cudaSuccess) return (int)e;                                                             \
        conv_fwd_kernel<V, RR, PL><<<grid, p.threads, p.smem, st>>>(rows, rpc, N, (unsigned)M, F, C, K,  \
                                                                    nn_index, nn_count, bin_index,       \
                                                                    input, filter, output);              \
    } while (0)
#define LAUNCH_FWD(V, RR)                                                                               \
    do {                                                                                                 \
        if (planned) LAUNCH_FWD2(V, RR, true); else LAUNCH_FWD2(V, RR, false);                           \
    } while (0)
    if (p.vec == 4 && r == 1) LAUNCH_FWD(4, 1);
    else if (p.vec == 4 && r == 2) LAUNCH_FWD(4, 2);
    else if (p.vec == 2 && r == 1) LAUNCH_FWD(2, 1);
    else if (p.vec == 2 && r == 2) LAUNCH_FWD(2, 2);
    else if (p.vec == 1 && r == 1) LAUNCH_FWD(1, 1);
    else LAUNCH_FWD(1, 2);
#undef LAUNCH_FWD
#undef LAUNCH_FWD2
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}

extern "C" int sph3d_depthwise_conv3d(int B, int N, int M, int F, int C, int r, int K,
                                      const int* nn_index, const int* nn_count, const int* bin_index,
                                      const float* input, const float* filter, float* output, void* stream)
{
    g_last_launch_count = 0;
    if (B <= 0 || N <= 0 || M <= 0 || F <= 0 || C <= 0 || r <= 0 || K <= 0 || !nn_index || !nn_count ||
        !bin_index || !input || !filter || !output)
        return (int)cudaErrorInvalidValue;
    return run_fwd(B, N, M, F, C, r, K, nn_index, nn_count, bin_index, input, filter, output, false, (cudaStream_t)stream);
}

// ---- split form of the forward launcher: graph-only sort + planned convolution -------------------------------------
extern "C" size_t sph3d_conv_sort_bytes(int B, int N, int M, int F, int K)
{
    if (B <= 0 || N <= 0 || M <= 0 || F <= 0 || K <= 0 || F > 128 || N >= (1 << 24)) return 0;
    if ((long long)B * M >= (1LL << 31)) return 0;
    return (size_t)B * M * K * sizeof(unsigned);
}

extern "C" int sph3d_conv_sort(int B, int N, int M, int F, int K, const int* nn_index, const int* nn_count,
                               const int* bin_index, void* plan, size_t plan_bytes, void* stream)
{
    g_last_launch_count = 0;
    const size_t need = sph3d_conv_sort_bytes(B, N, M, F, K);
    if (!need || !nn_index || !nn_count || !bin_index || !plan || plan_bytes < need) return (int)cudaErrorInvalidValue;
    const unsigned rows = (unsigned)((long long)B * M);
    const int warps = 8;
    const size_t smem = (size_t)warps * sort_smem_ints(F) * sizeof(int);
    unsigned want = (rows + warps - 1) / warps, cap = (unsigned)sm_count() * 8;
    conv_sort_kernel<<<want < cap ? want : cap, warps * 32, smem, (cudaStream_t)stream>>>(
        rows, K, F, nn_index, nn_count, bin_index, reinterpret_cast<unsigned*>(plan));
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}

extern "C" int sph3d_depthwise_conv3d_planned(int B, int N, int M, int F, int C, int r, int K, const int* nn_count,
                                              const void* plan, size_t plan_bytes, const float* input,
                                              const float* filter, float* output, void* stream)
{
    g_last_launch_count = 0;
    if (B <= 0 || N <= 0 || M <= 0 || F <= 0 || C <= 0 || r <= 0 || K <= 0 || !nn_count || !plan || !input || !filter || !output)
        return (int)cudaErrorInvalidValue;
    const size_t need = sph3d_conv_sort_bytes(B, N, M, F, K);
    if (!need || plan_bytes < need) return (int)cudaErrorInvalidValue;
    return run_fwd(B, N, M, F, C, r, K, reinterpret_cast<const int*>(plan), nn_count, nullptr, input, filter, output, true,
                   (cudaStream_t)stream);
}

extern "C" size_t sph3d_depthwise_conv3d_planned_supported(int B, int N, int M, int F, int C, int r, int K)
{
    if (!sph3d_conv_sort_bytes(B, N, M, F, K) || C <= 0) return 0;
    return plan_fwd(B, N, M, F, C, r, true).vec != 0 ? 1 : 0;
}

extern "C" int sph3d_abi_version(void) { return SPH3D_B200_ABI_VERSION; }
extern "C" void sph3d_reload_tunables(void) { g_tunables = read_tunables(); }
extern "C" int sph3d_last_launch_count(void) { return g_last_launch_count; }
