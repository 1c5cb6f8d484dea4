// sepconv.cu -- the separable SPH3D layer as ONE kernel, sm_100a: depthwise spherical convolution -> pointwise product
// on the tcgen05 tensor cores -> bias / ELU / per-channel affine (folded batch normalisation), hand-written.
//
// Replaces the node chain of /root/reference/utils/sph3gcn_util.py:128-161 (separable_conv3d): tf_conv3d.depthwise_conv3d
// (tf_conv3d_gpu.cu:7-29) -> tf.matmul -> tf.nn.bias_add -> tf.nn.elu -> batch normalisation.  SURVEY.md section 8(f)
// row N2: the (B, M, C*r) intermediate never leaves the SM unless the caller asks for it (training keeps it for the
// weight gradient).
//
// A persistent CTA of 32 warps owns 128 consecutive output points (a TILE) at a time:
//   1. GATHER (all warps; the loop of conv_fwd.cu, same summation order, so the optional depthwise output is bit-identical
//      to sph3d_depthwise_conv3d): a warp per output point sums the neighbour strips bin by bin, applies the spherical
//      filter once per (point, bin), scales by 1/count, and writes its C*r values -- split into three bf16 terms
//      hi + mid + lo -- straight into the A operand of the tensor core in shared memory (K-major, 128-byte swizzle;
//      tc05.cuh).
//   2. PRODUCT (one thread): the pointwise weights arrive as a pre-split bf16 image (sph3d_sepconv_pack_weights: the
//      byte image of the B operand's shared-memory layout, so a unit is one contiguous bulk copy) through a small ring of
//      TMA bulk copies; six of the nine cross products (every term down to 2^-24 of the result: hi*hi, mid*hi, lo*hi,
//      hi*mid, mid*mid, hi*lo) are issued as tcgen05.mma 128 x Cout x 16 instructions accumulating in tensor memory.
//   3. EPILOGUE (all warps): tcgen05.ld the accumulator rows, out = act(x + bias) * scale + shift, 64-byte stores.
// The product and epilogue of a tile take ~10 % of its gather time (the gather is L1-bound), so the three
// phases run back to back on one A buffer; what the fusion removes is the write + read of the intermediate, the read
// of the product's output by the layer tail, and two launches.
#include "conv_common.cuh"
#include "tc05.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

using namespace tc05;

constexpr int SC_TILE = 128;                     // output points per tile = UMMA M
constexpr int SC_WARPS = 32;
constexpr int SC_UNIT_A = SC_TILE * 128;         // bytes of one A unit (128 rows x 64 bf16)
constexpr int SC_MAX_STAGES = 6;

enum { SC_ACT_NONE = 0, SC_ACT_ELU = 1 };

struct SepconvArgs {
    unsigned rows, M;
    int N, F, C, K, Cout;
    int KC;                // 64-wide k-chunks of the product (C*r rounded up)
    int NP;                // Cout rounded up to 16: UMMA N, rows of a B unit
    int stages;            // B ring depth (units of NP x 64 bf16)
    int act;
    const int* nn_index;
    const int* nn_count;
    const int* bin_index;
    const float* input;
    const float* filter;
    const unsigned char* wimage;   // [KC][3][NP rows x 128 B]
    const float* bias;
    const float* scale;
    const float* shift;
    float* depthwise;      // optional (rows, C*r)
    float* output;         // (rows, Cout)
};

// pointwise weights W (K x Cout fp32, row-major) -> image[kc][term][n][swizzled 64 k] of bf16, zero padded
__global__ void __launch_bounds__(256)
sepconv_pack_kernel(int K, int Cout, int KC, int NP, const float* __restrict__ W, unsigned char* __restrict__ image)
{
    const int total = KC * NP * UNIT_K;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int kk = t % UNIT_K, n = (t / UNIT_K) % NP, kc = t / (UNIT_K * NP);
        const int k = kc * UNIT_K + kk;
        const float v = (k < K && n < Cout) ? __ldg(W + (size_t)k * Cout + n) : 0.f;
        const float hi = bf16_round(v), mid = bf16_round(v - hi), lo = bf16_round(v - hi - mid);
        const size_t unit = (size_t)NP * 128;
        unsigned char* base = image + (size_t)kc * 3 * unit + unit_offset(n, kk);
        *reinterpret_cast<__nv_bfloat16*>(base) = __float2bfloat16_rn(hi);
        *reinterpret_cast<__nv_bfloat16*>(base + unit) = __float2bfloat16_rn(mid);
        *reinterpret_cast<__nv_bfloat16*>(base + 2 * unit) = __float2bfloat16_rn(lo);
    }
}

template <int ACT> __device__ __forceinline__ float sc_act(float z)
{
    if constexpr (ACT == SC_ACT_ELU) return z > 0.f ? z : expf(z) - 1.0f;      // as csrc/post.cu evaluates it
    return z;
}

template <int VEC, int R>
__global__ void __launch_bounds__(SC_WARPS * 32, 1)
sepconv_kernel(const SepconvArgs a)
{
    constexpr int E = VEC * R;                      // values a lane contributes to a row of the A operand
    static_assert(E == 2 || E == 4, "a lane writes 4 or 8 bytes per bf16 term");
    using S = SmemStrip<E>;
    extern __shared__ unsigned char sc_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- shared memory carve-up (operand units need 1024-byte alignment) ----
    const uint32_t raw = smem_u32(sc_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = sc_raw + (base - raw);
    const uint32_t unitB = (uint32_t)a.NP * 128u;
    const uint32_t A_off = 0, B_off = A_off + (uint32_t)a.KC * 3u * SC_UNIT_A, W_off = B_off + (uint32_t)a.stages * unitB;
    float* Wsh = reinterpret_cast<float*>(sm + W_off);
    const int FP = ((a.F + 31) / 32) * 32;
    int* sortbase = reinterpret_cast<int*>(Wsh + (size_t)a.F * S::FLOATS);
    int* hA = sortbase + (size_t)warp * sort_smem_ints(a.F);
    int* hB = hA + (FP + 1 + 3) / 4 * 4;
    unsigned* sOff = reinterpret_cast<unsigned*>(hB + FP);
    int* sCode = reinterpret_cast<int*>(sOff + 64);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sortbase + (size_t)SC_WARPS * sort_smem_ints(a.F));   // 8-byte aligned: sort_smem_ints is a multiple of 4
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8u * SC_MAX_STAGES, bar_acc = bar_empty + 8u * SC_MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SC_MAX_STAGES + 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; s++) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    const uint32_t tmem_cols = tmem_cols_pow2((uint32_t)a.NP);
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    stage_filter<VEC, R>(Wsh, a.filter, a.F, a.C, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int c0 = lane * VEC;
    const bool active = c0 < a.C;
    const int c0ld = active ? c0 : 0;
    const unsigned strideB = (unsigned)a.C * 4u;
    const size_t cloudB = (size_t)a.N * a.C * 4;
    const float* wlane = Wsh + S::offset(0, lane);
    const int Co = a.C * R;
    // where this lane's E values go inside a row of the A operand
    const uint32_t k0 = (uint32_t)lane * E;
    const bool writes = k0 < (uint32_t)a.KC * UNIT_K;
    const uint32_t a_unit0 = base + A_off + (k0 >> 6) * 3u * SC_UNIT_A;
    const uint32_t idesc = idesc_bf16_f32(SC_TILE, (uint32_t)a.NP);
    const int NU = a.KC * 3;                        // B units per tile

    const unsigned ntiles = (a.rows + SC_TILE - 1) / SC_TILE;
    unsigned loads = 0, uses = 0;                   // thread 0: B units requested / consumed since the kernel started
    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const unsigned rbeg = tile * SC_TILE;
        // ---- B ring: request this tile's first units now, they land while the tile gathers ----
        int requested = 0;
        auto request = [&](int u) {
            const unsigned s = loads % (unsigned)a.stages;
            if (loads >= (unsigned)a.stages) mbar_wait(bar_empty + 8u * s, ((loads / a.stages) - 1u) & 1u);
            mbar_expect_tx(bar_full + 8u * s, unitB);
            bulk_g2s(base + B_off + s * unitB, a.wimage + (size_t)u * unitB, unitB, bar_full + 8u * s);
            loads++;
        };
        if (threadIdx.x == 0)
            for (; requested < NU && requested < a.stages; requested++) request(requested);

        // ---- 1. gather: rows rbeg + warp, + 32, ... ----
        RowCursor cur;
        cur.init(min(rbeg + warp, a.rows - 1), a.M);
        for (unsigned rt = warp; rt < SC_TILE; rt += SC_WARPS, cur.advance(SC_WARPS, a.M)) {
            const unsigned row = rbeg + rt;
            float acc[E];
#pragma unroll
            for (int e = 0; e < E; e++) acc[e] = 0.f;
            if (row < a.rows) {
                const int cnt = min(__ldg(a.nn_count + row), a.K);
                const char* inb = reinterpret_cast<const char*>(a.input) + cur.b * cloudB + (size_t)c0ld * 4;
                const int* idxrow = a.nn_index + (size_t)row * a.K;
                const int* binrow = a.bin_index + (size_t)row * a.K;
                for (int kt = 0; kt < cnt; kt += 64) {
                    const int e0 = kt + lane, e1 = kt + 32 + lane;
                    unsigned o0 = 0, o1 = 0;
                    int b0 = -1, b1 = -1;
                    if (e0 < cnt) { o0 = (unsigned)__ldg(idxrow + e0) * strideB; b0 = __ldg(binrow + e0); }
                    if (e1 < cnt) { o1 = (unsigned)__ldg(idxrow + e1) * strideB; b1 = __ldg(binrow + e1); }
                    sort_tile_by_bin(o0, b0, o1, b1, FP, lane, hA, hB, sOff, sCode);
                    const int nt = min(64, cnt - kt);
                    float s[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v++) s[v] = 0.f;
                    auto consume = [&](const float (&v)[VEC], int code) {
                        strip_add<VEC>(s, v);
                        if (code & 1) {                           // last edge of its bin: apply the filter strip once
                            float w[E];
                            S::load(w, wlane + (code >> 1) * S::FLOATS, 0);
#pragma unroll
                            for (int e = 0; e < E; e += 2) {
                                float2 t = __ffma2_rn(make_float2(s[e / R], s[(e + 1) / R]), make_float2(w[e], w[e + 1]),
                                                      make_float2(acc[e], acc[e + 1]));
                                acc[e] = t.x; acc[e + 1] = t.y;
                            }
#pragma unroll
                            for (int v2 = 0; v2 < VEC; v2++) s[v2] = 0.f;
                        }
                    };
                    int p = 0;
                    for (; p + 4 <= nt; p += 4) {
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
                    __syncwarp();
                }
                const float inv = cnt > 0 ? 1.0f / (float)cnt : 0.f;
#pragma unroll
                for (int e = 0; e < E; e++) acc[e] = active ? acc[e] * inv : 0.f;
                if (a.depthwise != nullptr && active) {
                    float* out = a.depthwise + (size_t)row * Co + (size_t)c0 * R;
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
            if (writes) {                                         // rows past the end and idle lanes contribute zeros
                const uint32_t dst = a_unit0 + unit_offset(rt, k0 & 63u);
                uint32_t hi[E / 2], mid[E / 2], lo[E / 2];
#pragma unroll
                for (int e = 0; e < E; e += 2) split3_pack2(acc[e], acc[e + 1], hi[e / 2], mid[e / 2], lo[e / 2]);
                if constexpr (E == 4) {
                    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst), "r"(hi[0]), "r"(hi[1]) : "memory");
                    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst + SC_UNIT_A), "r"(mid[0]), "r"(mid[1]) : "memory");
                    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst + 2 * SC_UNIT_A), "r"(lo[0]), "r"(lo[1]) : "memory");
                } else {
                    asm volatile("st.shared.b32 [%0], %1;" :: "r"(dst), "r"(hi[0]) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" :: "r"(dst + SC_UNIT_A), "r"(mid[0]) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" :: "r"(dst + 2 * SC_UNIT_A), "r"(lo[0]) : "memory");
                }
            }
        }
        fence_proxy_async_smem();                                 // the A tile was written with ordinary stores
        __syncthreads();

        // ---- 2. product: one thread feeds the tensor core ----
        if (threadIdx.x == 0) {
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int u = 0; u < NU; u++) {
                const int kc = u / 3, tb = u - kc * 3;           // B term tb meets A terms 0 .. 2 - tb
                const unsigned s = uses % (unsigned)a.stages;
                mbar_wait(bar_full + 8u * s, (uses / a.stages) & 1u);
                tc_fence_after();
                const uint32_t bunit = base + B_off + s * unitB;
                for (int ta = 0; ta + tb <= 2; ta++) {
                    const uint32_t aunit = base + A_off + (uint32_t)(kc * 3 + ta) * SC_UNIT_A;
#pragma unroll
                    for (int j = 0; j < UNIT_K / UMMA_K; j++) {
                        umma_bf16(tmem, smem_desc_sw128(aunit + 32u * j), smem_desc_sw128(bunit + 32u * j), idesc, accumulate);
                        accumulate = 1;
                    }
                }
                umma_commit(bar_empty + 8u * s);                  // the unit's slot is free once these complete
                uses++;
                if (requested < NU) { request(requested); requested++; }
            }
            umma_commit(bar_acc);                                 // accumulator complete
        }

        // ---- 3. epilogue ----
        mbar_wait(bar_acc, it & 1u);
        tc_fence_after();
        {
            const int q = warp & 3, cg = warp >> 2;
            const unsigned row = rbeg + 32u * q + lane;
            const bool live = row < a.rows;
            float* orow = a.output + (size_t)row * a.Cout;
            for (int blk = cg; blk * 16 < a.NP; blk += SC_WARPS / 4) {
                float v[16];
                tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(blk * 16), v);
                const int col0 = blk * 16;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int c = min(col0 + i, a.Cout - 1);
                    float z = v[i] + (a.bias ? __ldg(a.bias + c) : 0.f);
                    z = a.act == SC_ACT_ELU ? sc_act<SC_ACT_ELU>(z) : z;
                    if (a.scale) z *= __ldg(a.scale + c);
                    if (a.shift) z += __ldg(a.shift + c);
                    v[i] = z;
                }
                if (live) {
                    if ((a.Cout & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            if (col0 + i < a.Cout)
                                *reinterpret_cast<float4*>(orow + col0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i++)
                            if (col0 + i < a.Cout) orow[col0 + i] = v[i];
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();                                          // accumulator and A tile are free for the next tile
    }
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

struct SepconvGeom {
    int vec, KC, NP, stages;
    size_t smem;
    bool ok;
};

static SepconvGeom sepconv_geom(int B, int N, int M, int F, int C, int r, int K, int Cout)
{
    SepconvGeom g{};
    if (B <= 0 || N <= 0 || M <= 0 || F <= 0 || C <= 0 || K <= 0 || Cout <= 0) return g;
    if ((r != 1 && r != 2) || F > 128 || !fits_32bit(B, N, M, C, r)) return g;
    const int Kp = C * r;
    if (Kp > 128 || Cout > 256) return g;                        // one A buffer of two k-chunks; UMMA N <= 256
    const int vec = pick_vec_full_warp(C);
    if (32 * vec < C) return g;                                  // single channel chunk
    const int E = vec * r;
    if (E != 2 && E != 4) return g;
    g.vec = vec;
    g.KC = (Kp + UNIT_K - 1) / UNIT_K;
    g.NP = (Cout + 15) / 16 * 16;
    const size_t fixed = 1024 + (size_t)g.KC * 3 * SC_UNIT_A + (size_t)F * 32 * E * sizeof(float) +
                         (size_t)SC_WARPS * sort_smem_ints(F) * sizeof(int) + (2 * SC_MAX_STAGES + 2) * 8;
    const size_t unitB = (size_t)g.NP * 128;
    int stages = g.KC * 3;
    if (stages > SC_MAX_STAGES) stages = SC_MAX_STAGES;
    while (stages > 1 && fixed + stages * unitB > SMEM_CAP) stages--;
    if (fixed + stages * unitB > SMEM_CAP) return g;
    g.stages = stages;
    g.smem = fixed + stages * unitB;
    g.ok = true;
    return g;
}

}  // namespace sph3d

using namespace sph3d;

extern "C" int sph3d_separable_conv3d_supported(int B, int N, int M, int F, int C, int r, int K, int Cout)
{
    return sepconv_geom(B, N, M, F, C, r, K, Cout).ok ? 1 : 0;
}

extern "C" size_t sph3d_sepconv_weight_image_bytes(int Kp, int Cout)
{
    if (Kp <= 0 || Cout <= 0) return 0;
    const size_t KC = (Kp + UNIT_K - 1) / UNIT_K, NP = (Cout + 15) / 16 * 16;
    return KC * 3 * NP * 128;
}

extern "C" int sph3d_sepconv_pack_weights(int Kp, int Cout, const float* weights, void* image, void* stream)
{
    if (Kp <= 0 || Cout <= 0 || !weights || !image) return (int)cudaErrorInvalidValue;
    const int KC = (Kp + UNIT_K - 1) / UNIT_K, NP = (Cout + 15) / 16 * 16;
    const int total = KC * NP * UNIT_K;
    sepconv_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(Kp, Cout, KC, NP, weights,
                                                                               static_cast<unsigned char*>(image));
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}

extern "C" int sph3d_separable_conv3d(int B, int N, int M, int F, int C, int r, int K, int Cout, const int* nn_index,
                                      const int* nn_count, const int* bin_index, const float* input, const float* filter,
                                      const void* weight_image, const float* bias, const float* scale, const float* shift,
                                      int act, float* depthwise_out, float* output, void* stream)
{
    const SepconvGeom g = sepconv_geom(B, N, M, F, C, r, K, Cout);
    if (!g.ok || (act != SC_ACT_NONE && act != SC_ACT_ELU)) return (int)cudaErrorInvalidValue;
    SepconvArgs a{};
    a.rows = (unsigned)((long long)B * M); a.M = (unsigned)M;
    a.N = N; a.F = F; a.C = C; a.K = K; a.Cout = Cout;
    a.KC = g.KC; a.NP = g.NP; a.stages = g.stages; a.act = act;
    a.nn_index = nn_index; a.nn_count = nn_count; a.bin_index = bin_index;
    a.input = input; a.filter = filter; a.wimage = static_cast<const unsigned char*>(weight_image);
    a.bias = bias; a.scale = scale; a.shift = shift; a.depthwise = depthwise_out; a.output = output;
    const unsigned ntiles = (a.rows + SC_TILE - 1) / SC_TILE;
    const unsigned grid = ntiles < (unsigned)sm_count() ? ntiles : (unsigned)sm_count();
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
#define LAUNCH_SC(V, RR)                                                           \
    do {                                                                            \
        e = set_smem(sepconv_kernel<V, RR>, g.smem);                                \
        if (e != cudaSuccess) return (int)e;                                        \
        sepconv_kernel<V, RR><<<grid, SC_WARPS * 32, g.smem, st>>>(a);              \
    } while (0)
    if (g.vec == 4 && r == 1) LAUNCH_SC(4, 1);
    else if (g.vec == 2 && r == 1) LAUNCH_SC(2, 1);
    else if (g.vec == 2 && r == 2) LAUNCH_SC(2, 2);
    else if (g.vec == 1 && r == 2) LAUNCH_SC(1, 2);
    else return (int)cudaErrorInvalidValue;
#undef LAUNCH_SC
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}
