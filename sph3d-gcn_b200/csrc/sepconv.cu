// sepconv.cu -- the separable SPH3D layer as ONE kernel, sm_100a: depthwise spherical convolution -> pointwise product
// on the tcgen05 tensor cores -> bias / ELU / per-channel affine (folded batch normalisation), hand-written.
//
// Replaces the node chain of /root/reference/utils/sph3gcn_util.py:128-161 (separable_conv3d): tf_conv3d.depthwise_conv3d
// (tf_conv3d_gpu.cu:7-29) -> tf.matmul -> tf.nn.bias_add -> tf.nn.elu -> batch normalisation.  SURVEY.md section 8(f)
// row N2: the (B, M, C*r) intermediate never leaves the SM unless the caller asks for it (training keeps it for the
// weight gradient).
//
// The product is computed TRANSPOSED, out^T (Cout x rows) = W^T (Cout x K) * dw^T (K x rows): the pointwise weights are
// the tensor core's A operand (M = 128 output channels per block), a TILE of TR = 64 or 32 gathered output points is its
// B operand (N = TR), so a tile is small enough to be double buffered in shared memory, and in the accumulator a
// thread's tensor-memory lane is an output CHANNEL: bias / scale / shift are per-thread constants and a warp's store of
// one point is 128 contiguous bytes.
//
// A persistent CTA of 32 warps; every warp runs the same loop, nothing but mbarriers in it.  Rows are drawn one at a time
// from a CTA-wide queue (with rows assigned statically, every tile waited for its slowest warp: measured 0.96 ms against
// 0.59 ms for the gather alone at the headline shape):
//   1. GATHER the drawn row of tile t (the loop of conv_fwd.cu, same summation order, so the optional depthwise output is
//      bit-identical to sph3d_depthwise_conv3d): neighbour strips summed bin by bin, the spherical filter applied once per
//      (point, bin), 1/count, and the row's C*r values -- split into three bf16 terms hi + mid + lo -- stored straight into
//      buffer t%2 of the B operand (K-major, 128-byte swizzle; tc05.cuh).  Arrive on the buffer's "tile full" barrier.
//   2. The warp whose row completes the tile issues the tile's product (one lane; issuers take turns in tile order): the weights stream as a pre-split bf16 image
//      (sph3d_sepconv_pack_weights writes the byte image of the operand's shared-memory layout, a unit = one contiguous
//      16 KB bulk copy) through a ring of TMA bulk copies that runs ahead across tiles; six of the nine cross products
//      (everything down to 2^-24 of the result) are tcgen05.mma 128 x TR x 16 instructions into accumulator t%2 of
//      tensor memory; tcgen05.commit signals the tile's "product done" barrier.
//   3. EPILOGUE: every warp owns a fixed slice of each tile's accumulator (tensor-memory lanes are tied to the warp id) and
//      finishes the slices of all tiles up to t-2 before it writes a row of tile t: wait "product done", tcgen05.ld,
//      out = act(x + bias) * scale + shift, store, arrive on "accumulator free" (which the issuer of tile t+2 waits for).
// Warps therefore never wait for the product or for each other inside a tile; the coupling is two tiles of slack.
#include "conv_common.cuh"
#include "tc05.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

using namespace tc05;

constexpr int SC_WARPS = 32;
constexpr int SC_SUPER = 128;                    // rows of consecutive tiles a CTA takes together (L1 locality)
constexpr int SC_UNIT_W = 128 * 128;             // bytes of one weight unit (128 output channels x 64 bf16)
constexpr int SC_MAX_STAGES = 6;

enum { SC_ACT_NONE = 0, SC_ACT_ELU = 1 };

struct SepconvArgs {
    unsigned rows, M;
    int N, F, C, K, Cout;
    int KC;                // 64-wide k-chunks of the product (C*r rounded up)
    int MB;                // blocks of 128 output channels
    int stages;            // weight ring depth (units of 128 x 64 bf16)
    int act;
    const int* nn_index;
    const int* nn_count;
    const int* bin_index;
    const float* input;
    const float* filter;
    const unsigned char* wimage;   // [KC][3 terms][MB][128 rows x 128 B]
    const float* bias;
    const float* scale;
    const float* shift;
    float* depthwise;      // optional (rows, C*r)
    float* output;         // (rows, Cout)
};

// pointwise weights W (K x Cout fp32, row-major) -> image[kc][term][mb][n][swizzled 64 k] of bf16, zero padded:
// row n of block mb is output channel mb*128 + n (the operand is W^T, K-major)
__global__ void __launch_bounds__(256)
sepconv_pack_kernel(int K, int Cout, int KC, int MB, const float* __restrict__ W, unsigned char* __restrict__ image)
{
    const int total = KC * MB * 128 * UNIT_K;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int kk = t % UNIT_K, n = (t / UNIT_K) % 128, mb = (t / (UNIT_K * 128)) % MB, kc = t / (UNIT_K * 128 * MB);
        const int k = kc * UNIT_K + kk, co = mb * 128 + n;
        const float v = (k < K && co < Cout) ? __ldg(W + (size_t)k * Cout + co) : 0.f;
        const float hi = bf16_round(v), mid = bf16_round(v - hi), lo = bf16_round(v - hi - mid);
        unsigned char* base = image + ((size_t)(kc * 3) * MB + mb) * SC_UNIT_W + unit_offset(n, kk);
        *reinterpret_cast<__nv_bfloat16*>(base) = __float2bfloat16_rn(hi);
        *reinterpret_cast<__nv_bfloat16*>(base + (size_t)MB * SC_UNIT_W) = __float2bfloat16_rn(mid);
        *reinterpret_cast<__nv_bfloat16*>(base + (size_t)2 * MB * SC_UNIT_W) = __float2bfloat16_rn(lo);
    }
}

__device__ __forceinline__ float sc_elu(float z) { return z > 0.f ? z : expf(z) - 1.0f; }      // as csrc/post.cu evaluates it

// shared bookkeeping of a CTA (behind the operand buffers)
struct SepconvSync {
    uint64_t w_full[SC_MAX_STAGES];    // weight unit landed (tx bytes)
    uint64_t w_empty[SC_MAX_STAGES];   // the products that read the stage completed (tcgen05.commit)
    uint64_t tile_full[2];             // TR row arrivals: buffer holds the tile
    uint64_t prod_done[2];             // the tile's products completed: accumulator ready, buffer reusable
    uint64_t acc_free[2];              // 32 warp arrivals: every warp has read its slice of the accumulator
    unsigned arrivals[2];              // election of the issuing warp
    unsigned next_row;                 // the CTA's row queue
    unsigned units_used;               // weight units consumed so far by this CTA (ring position)
    uint32_t tmem;
};

template <int VEC, int R, int TR>
__global__ void __launch_bounds__(SC_WARPS * 32, 1)
sepconv_kernel(const SepconvArgs a)
{
    constexpr int E = VEC * R;                      // values a lane contributes to a row of the tile operand
    static_assert(E == 2 || E == 4 || E == 8, "a lane writes 4, 8 or 16 bytes per bf16 term");
    static_assert(TR == 64 || TR == 32, "tile rows");
    constexpr int SUB = SC_SUPER / TR;              // tiles per 128-row super tile
    constexpr int UNIT_T = TR * 128;                // bytes of one tile unit (TR rows x 64 bf16)
    constexpr int EPI = TR / 8;                     // accumulator columns (= points) a warp finishes per tile and block
    using S = SmemStrip<E>;
    extern __shared__ __align__(16) float sc_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- shared memory: filter | per-warp sort scratch | barriers  (as conv_fwd.cu lays them out), then, 1024-byte
    // aligned and touched only through 32-bit shared addresses, the two tile buffers and the weight ring ----
    float* Wsh = sc_smem;
    const int FP = ((a.F + 31) / 32) * 32;
    int* sortbase = reinterpret_cast<int*>(Wsh + (size_t)a.F * S::FLOATS);
    int* hA = sortbase + (size_t)warp * sort_smem_ints(a.F);
    int* hB = hA + (FP + 1 + 3) / 4 * 4;
    unsigned* sOff = reinterpret_cast<unsigned*>(hB + FP);
    int* sCode = reinterpret_cast<int*>(sOff + 64);
    SepconvSync* sy = reinterpret_cast<SepconvSync*>(sortbase + (size_t)SC_WARPS * sort_smem_ints(a.F));   // 16-byte aligned
    const uint32_t w_full = smem_u32(sy->w_full), w_empty = smem_u32(sy->w_empty);
    const uint32_t tile_full = smem_u32(sy->tile_full), prod_done = smem_u32(sy->prod_done), acc_free = smem_u32(sy->acc_free);
    const uint32_t base = (smem_u32(sy + 1) + 1023u) & ~1023u;
    const uint32_t tile_bytes = (uint32_t)a.KC * 3u * UNIT_T;
    const uint32_t T_off = 0, W_off = 2u * tile_bytes;

    // ---- this CTA's tiles: super tiles blockIdx.x, + gridDim.x, ..., SUB consecutive tiles each; tile number tau of the
    // CTA's sequence is global tile (blockIdx.x + (tau / SUB) * gridDim.x) * SUB + tau % SUB ----
    const unsigned ntiles = (a.rows + TR - 1) / TR;
    const unsigned nsuper = (ntiles + SUB - 1) / SUB;
    unsigned my_tiles = 0;
    for (unsigned sp = blockIdx.x; sp < nsuper; sp += gridDim.x) my_tiles += min((unsigned)SUB, ntiles - sp * SUB);
    const unsigned my_rows = my_tiles * TR;
    const int NU = a.KC * 3 * a.MB;                 // weight units per tile
    const unsigned total_units = my_tiles * (unsigned)NU;
    auto tile_first_row = [&](unsigned tau) { return ((blockIdx.x + (tau / SUB) * gridDim.x) * SUB + tau % SUB) * (unsigned)TR; };

    auto request_unit = [&](unsigned L) {           // weight unit number L of this CTA's stream -> stage L % stages
        const unsigned s = L % (unsigned)a.stages;
        if (L >= (unsigned)a.stages) mbar_wait(w_empty + 8u * s, ((L / a.stages) - 1u) & 1u);
        mbar_expect_tx(w_full + 8u * s, SC_UNIT_W);
        bulk_g2s(base + W_off + s * SC_UNIT_W, a.wimage + (size_t)(L % (unsigned)NU) * SC_UNIT_W, SC_UNIT_W, w_full + 8u * s);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; s++) { mbar_init(w_full + 8u * s, 1); mbar_init(w_empty + 8u * s, 1); }
        for (int d = 0; d < 2; d++) {
            mbar_init(tile_full + 8u * d, TR); mbar_init(prod_done + 8u * d, 1); mbar_init(acc_free + 8u * d, SC_WARPS);
        }
        sy->arrivals[0] = sy->arrivals[1] = 0;
        sy->next_row = 0;
        sy->units_used = 0;
        fence_mbar_init();
        for (unsigned L = 0; L + 1 < (unsigned)a.stages && L < total_units; L++) request_unit(L);   // the ring runs stages-1 ahead
    }
    const uint32_t tmem_cols = tmem_cols_pow2(2u * a.MB * TR);
    if (warp == 0) tmem_alloc(smem_u32(&sy->tmem), tmem_cols);
    stage_filter<VEC, R>(Wsh, a.filter, a.F, a.C, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sy->tmem;

    const int c0 = lane * VEC;
    const bool active = c0 < a.C;
    const int c0ld = active ? c0 : 0;
    const unsigned strideB = (unsigned)a.C * 4u;
    const size_t cloudB = (size_t)a.N * a.C * 4;
    const float* wlane = Wsh + S::offset(0, lane);
    const int Co = a.C * R;
    // where this lane's E values go inside a row of the tile operand
    const uint32_t k0 = (uint32_t)lane * E;
    const bool writes = k0 < (uint32_t)a.KC * UNIT_K;
    const uint32_t lane_unit0 = (k0 >> 6) * 3u * UNIT_T;
    const uint32_t idesc = idesc_bf16_f32(128, TR);

    // This warp's slice of the epilogue of the CTA's tile tau: channels 32*(warp&3)+lane of every block, points
    // (warp>>2)*EPI .. +EPI of the tile.  Every warp runs it exactly once per tile, in tile order.
    const bool elu = a.act == SC_ACT_ELU;
    auto epilogue = [&](unsigned tau) {
        const unsigned d = tau & 1u, use = tau >> 1;
        mbar_wait(prod_done + 8u * d, use & 1u);
        tc_fence_after();
        const int q = warp & 3, p0 = (warp >> 2) * EPI;
        const unsigned row0 = tile_first_row(tau) + p0;
        const int nvalid = row0 < a.rows ? (int)min((unsigned)EPI, a.rows - row0) : 0;
        for (int mb = 0; mb < a.MB; mb++) {
            const int c = mb * 128 + 32 * q + lane;
            float v[EPI];
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)((d * a.MB + mb) * TR + p0);
            if constexpr (EPI == 8) tmem_ld8(taddr, v); else tmem_ld4(taddr, v);
            if (c < a.Cout) {
                const float bi = a.bias ? __ldg(a.bias + c) : 0.f;
                const float sc = a.scale ? __ldg(a.scale + c) : 1.f;
                const float sh = a.shift ? __ldg(a.shift + c) : 0.f;
                float* o = a.output + (size_t)row0 * a.Cout + c;
#pragma unroll
                for (int j = 0; j < EPI; j++) {
                    float z = v[j] + bi;
                    if (elu) z = sc_elu(z);
                    z = fmaf(z, sc, sh);
                    if (j < nvalid) o[(size_t)j * a.Cout] = z;
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free + 8u * d);
    };

    // Rows are handed out one at a time from a CTA-wide queue (a warp that drew a long row does not hold the others up);
    // a row of tile tau may be written once tile tau-2 has left its buffer, and before that the warp finishes its
    // epilogue slices of every tile up to tau-2.
    unsigned epi_next = 0;
    for (;;) {
        unsigned r = 0;
        if (lane == 0) r = atomicAdd(&sy->next_row, 1u);
        r = __shfl_sync(FULL_MASK, r, 0);
        if (r >= my_rows) break;
        const unsigned tau = r / TR, rt = r % TR;
        const unsigned d = tau & 1u, use = tau >> 1;
        while (epi_next + 2 <= tau) epilogue(epi_next++);         // includes the wait that frees buffer d
        const uint32_t tbuf = base + T_off + d * tile_bytes;
        const unsigned row = tile_first_row(tau) + rt;

        // ---- 1. gather ----
        float acc[E];
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = 0.f;
        if (row < a.rows) {
            const unsigned b = row / a.M;
            const int cnt = min(__ldg(a.nn_count + row), a.K);
            const char* inb = reinterpret_cast<const char*>(a.input) + b * cloudB + (size_t)c0ld * 4;
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
                    if (code & 1) {                               // last edge of its bin: apply the filter strip once
                        float w[E];
                        S::load(w, wlane + (code >> 1) * S::FLOATS, 0);
#pragma unroll
                        for (int e = 0; e < E; e += 2) {
                            float2 f2 = __ffma2_rn(make_float2(s[e / R], s[(e + 1) / R]), make_float2(w[e], w[e + 1]),
                                                   make_float2(acc[e], acc[e + 1]));
                            acc[e] = f2.x; acc[e + 1] = f2.y;
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
                    float tv[VW];
#pragma unroll
                    for (int u = 0; u < VW; u++) tv[u] = acc[pl * VW + u];
                    VecIO<VW>::st(out + pl * VW, tv);
                }
            }
        }
        if (writes) {                                             // rows past the end and idle lanes contribute zeros
            const uint32_t dst = tbuf + lane_unit0 + unit_offset(rt, k0 & 63u);
            uint32_t hi[E / 2], mid[E / 2], lo[E / 2];
#pragma unroll
            for (int e = 0; e < E; e += 2) split3_pack2(acc[e], acc[e + 1], hi[e / 2], mid[e / 2], lo[e / 2]);
            if constexpr (E == 8) {
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(dst + UNIT_T), "r"(mid[0]), "r"(mid[1]), "r"(mid[2]), "r"(mid[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(dst + 2 * UNIT_T), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
            } else if constexpr (E == 4) {
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst), "r"(hi[0]), "r"(hi[1]) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst + UNIT_T), "r"(mid[0]), "r"(mid[1]) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst + 2 * UNIT_T), "r"(lo[0]), "r"(lo[1]) : "memory");
            } else {
                asm volatile("st.shared.b32 [%0], %1;" :: "r"(dst), "r"(hi[0]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" :: "r"(dst + UNIT_T), "r"(mid[0]) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" :: "r"(dst + 2 * UNIT_T), "r"(lo[0]) : "memory");
            }
        }
        fence_proxy_async_smem();                                 // the row was written with ordinary stores
        __syncwarp();

        // ---- 2. the warp whose row completes the tile issues its products ----
        if (lane == 0) {
            mbar_arrive(tile_full + 8u * d);
            if (atomicAdd(&sy->arrivals[d], 1u) == TR - 1) {
                sy->arrivals[d] = 0;
                mbar_wait(tile_full + 8u * d, use & 1u);
                if (use >= 1) mbar_wait(acc_free + 8u * d, (use - 1u) & 1u);      // tile tau-2 has been read out of accumulator d
                // issuers take turns in tile order (the other warps may complete tile tau while tile tau-1 is still being
                // issued): the ring position doubles as the ticket
                volatile unsigned* ticket = &sy->units_used;
                if (*ticket != tau * (unsigned)NU) {
                    const long long t0 = clock64();
                    while (*ticket != tau * (unsigned)NU)
                        if (clock64() - t0 > 4000000000LL) __trap();
                }
                __threadfence_block();
                tc_fence_after();
                unsigned g = tau * (unsigned)NU;
                for (int kc = 0; kc < a.KC; kc++)
                    for (int tw = 0; tw < 3; tw++)                // weight term tw meets tile terms 0 .. 2 - tw
                        for (int mb = 0; mb < a.MB; mb++, g++) {
                            const unsigned s = g % (unsigned)a.stages;
                            mbar_wait(w_full + 8u * s, (g / a.stages) & 1u);
                            tc_fence_after();
                            const uint32_t wunit = base + W_off + s * SC_UNIT_W;
                            const uint32_t dcol = tmem + (uint32_t)((d * a.MB + mb) * TR);
                            for (int tt = 0; tt + tw <= 2; tt++) {
                                const uint32_t tunit = tbuf + (uint32_t)(kc * 3 + tt) * UNIT_T;
#pragma unroll
                                for (int j = 0; j < UNIT_K / UMMA_K; j++)
                                    umma_bf16(dcol, smem_desc_sw128(wunit + 32u * j), smem_desc_sw128(tunit + 32u * j), idesc,
                                              (kc | tw | tt | j) != 0);
                            }
                            umma_commit(w_empty + 8u * s);        // the stage is free once these complete
                            const unsigned L = g + (unsigned)a.stages - 1u;
                            if (L < total_units) request_unit(L);
                        }
                umma_commit(prod_done + 8u * d);
                __threadfence_block();
                *ticket = g;
            }
        }
        __syncwarp();
    }
    while (epi_next < my_tiles) epilogue(epi_next++);
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

struct SepconvGeom {
    int vec, KC, MB, stages, tr;
    size_t smem;
    bool ok;
};

static SepconvGeom sepconv_geom(int B, int N, int M, int F, int C, int r, int K, int Cout)
{
    SepconvGeom g{};
    if (B <= 0 || N <= 0 || M <= 0 || F <= 0 || C <= 0 || K <= 0 || Cout <= 0) return g;
    if ((r != 1 && r != 2) || F > 128 || !fits_32bit(B, N, M, C, r)) return g;
    const int Kp = C * r;
    if (Kp > 256 || Cout > 256) return g;                        // four k-chunks; two blocks of 128 output channels
    const int vec = pick_vec_full_warp(C);
    if (32 * vec < C) return g;                                  // single channel chunk
    const int E = vec * r;
    if (E != 2 && E != 4 && E != 8) return g;
    g.vec = vec;
    g.KC = (Kp + UNIT_K - 1) / UNIT_K;
    g.MB = (Cout + 127) / 128;
    g.tr = g.KC <= 2 ? 64 : 32;                                  // two tile buffers stay at <= 96 KB
    if (tunables().sepconv_tile == 32) g.tr = 32;
    const size_t fixed = 1024 + 2 * (size_t)g.KC * 3 * g.tr * 128 + (size_t)F * 32 * E * sizeof(float) +
                         (size_t)SC_WARPS * sort_smem_ints(F) * sizeof(int) + sizeof(SepconvSync) + 16;
    // two stages measured fastest (0.758 ms against 0.783 with six at the headline shape): the ring only has to cover one
    // unit's copy latency, and every 16 KB it does not take stays L1 for the gather
    int stages = 2;
    if (tunables().sepconv_stages >= 2 && tunables().sepconv_stages <= SC_MAX_STAGES) stages = tunables().sepconv_stages;
    while (stages > 2 && fixed + (size_t)stages * SC_UNIT_W > SMEM_CAP) stages--;
    if (fixed + (size_t)stages * SC_UNIT_W > SMEM_CAP) return g;
    g.stages = stages;
    g.smem = fixed + (size_t)stages * SC_UNIT_W;
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
    const size_t KC = (Kp + UNIT_K - 1) / UNIT_K, MB = (Cout + 127) / 128;
    return KC * 3 * MB * SC_UNIT_W;
}

extern "C" int sph3d_sepconv_pack_weights(int Kp, int Cout, const float* weights, void* image, void* stream)
{
    if (Kp <= 0 || Cout <= 0 || !weights || !image) return (int)cudaErrorInvalidValue;
    const int KC = (Kp + UNIT_K - 1) / UNIT_K, MB = (Cout + 127) / 128;
    const int total = KC * MB * 128 * UNIT_K;
    sepconv_pack_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(Kp, Cout, KC, MB, weights,
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
    a.KC = g.KC; a.MB = g.MB; a.stages = g.stages; a.act = act;
    a.nn_index = nn_index; a.nn_count = nn_count; a.bin_index = bin_index;
    a.input = input; a.filter = filter; a.wimage = static_cast<const unsigned char*>(weight_image);
    a.bias = bias; a.scale = scale; a.shift = shift; a.depthwise = depthwise_out; a.output = output;
    const unsigned nsuper = (a.rows + SC_SUPER - 1) / SC_SUPER;
    const unsigned grid = nsuper < (unsigned)sm_count() ? nsuper : (unsigned)sm_count();
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
#define LAUNCH_SC(V, RR, TRR)                                                           \
    do {                                                                                 \
        e = set_smem(sepconv_kernel<V, RR, TRR>, g.smem);                                \
        if (e != cudaSuccess) return (int)e;                                             \
        sepconv_kernel<V, RR, TRR><<<grid, SC_WARPS * 32, g.smem, st>>>(a);              \
    } while (0)
#define LAUNCH_SC_TR(V, RR)                                                             \
    do {                                                                                 \
        if (g.tr == 64) LAUNCH_SC(V, RR, 64); else LAUNCH_SC(V, RR, 32);                 \
    } while (0)
    if (g.vec == 4 && r == 1) LAUNCH_SC_TR(4, 1);
    else if (g.vec == 4 && r == 2) LAUNCH_SC_TR(4, 2);
    else if (g.vec == 2 && r == 1) LAUNCH_SC_TR(2, 1);
    else if (g.vec == 2 && r == 2) LAUNCH_SC_TR(2, 2);
    else if (g.vec == 1 && r == 2) LAUNCH_SC_TR(1, 2);
    else return (int)cudaErrorInvalidValue;
#undef LAUNCH_SC_TR
#undef LAUNCH_SC
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}
