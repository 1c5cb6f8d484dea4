// rowsgemm.cu -- the pointwise product of an SPH3D layer over its B*M rows, hand-written for the sm_100a tensor cores.
//
// Replaces the tf.matmul of /root/reference/utils/sph3gcn_util.py:144-146 (separable_conv3d), :203-205 (pointwise_conv3d)
// and :254-256 (fully_connected) together with its input gradient:
//      y  (R x N) = x (R x K) * W   (K x N)                     forward
//      gx (R x K) = g (R x N) * W^T                               backward w.r.t. the rows (the same kernel on W^T)
// fp32 in, fp32 out, fp32 accuracy.  tcgen05 has no fp32 product, so every operand is split into three bf16 terms
// (v = hi + mid + lo to 2^-24) and the six cross products that matter (hi*hi, hi*mid, mid*hi, hi*lo, lo*hi, mid*mid) are
// accumulated in tensor memory -- the scheme csrc/sepconv.cu uses inside the fused layer, here as a streaming kernel:
//
//   * the WEIGHTS are split once per call (sph3d_rows_gemm_pack) into the byte image of the tensor core's B operand
//     (K-major, 128-byte swizzle, units of 128 output channels x 64 k = 16 KB), so a unit is one contiguous bulk copy;
//   * the ROWS cross HBM exactly once, as fp32: sixteen producer warps read a 128-row x 64-k chunk with coalesced
//     16-byte loads, split it in registers and store the three terms straight into the other operand (st.shared at the
//     swizzled address) -- no fp32 staging buffer, no second pass over shared memory.  The registers ARE the prefetch
//     queue: the loads of the next two chunks are in flight while a chunk is converted (64 KB per SM, what it takes to
//     cover the HBM latency at full bandwidth; with one chunk ahead the kernel ran at 3 TB/s);
//   * the product is computed TRANSPOSED, y^T (N x rows) = W^T * x^T: the weights are the tensor core's A operand (M = 128
//     output columns), the row tile its B operand, so in the accumulator a thread's tensor-memory lane is an output COLUMN
//     and a warp's store of one row is 128 contiguous bytes (with rows on the lanes every store instruction touched 32
//     different lines and the epilogue alone cost more L1 cycles than the products);
//   * ONE thread issues the tcgen05.mma 128 x 128 x 16 instructions (24 per chunk and block of 128 output columns) into
//     one of two accumulators in tensor memory, tcgen05.commit releases the operand stages and hands the accumulator to
//   * four epilogue warps, which tcgen05.ld their columns and store them while the producers and the tensor core are
//     already on the next tile;
//   * a weight image of up to six units (K*N <= 128*128) stays RESIDENT in shared memory for the whole kernel; larger ones
//     stream through a four-unit ring from L2 once per row tile.
// A persistent CTA per SM walks the row tiles blockIdx.x, + gridDim.x, ...; blockIdx.y selects a group of up to 256 output
// columns (2 blocks x 128 = the 512 tensor-memory columns of two accumulators).  Nothing but mbarriers in the loops.
//
// Roles by warp: 0-3 epilogue (warp % 4 = the tensor-memory lane quarter it may read), 4 issuer, 5 weight loader,
// 6-21 producers.
#include "conv_common.cuh"
#include "tc05.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

using namespace tc05;

constexpr int RG_TM = 128;                       // rows per tile (UMMA M)
constexpr int RG_UNIT = 128 * 128;               // bytes of one operand unit: 128 rows x 64 bf16
constexpr int RG_EPI = 4, RG_PROD = 16;
constexpr int RG_WARPS = RG_EPI + 2 + RG_PROD;
constexpr int RG_PR = RG_TM / RG_PROD / 2;       // row pairs a producer warp converts per chunk
constexpr int RG_AST_MAX = 3;                    // stages of the row operand (3 terms x 16 KB each)
constexpr int RG_WST_MAX = 8;                    // weight units in shared memory: a ring of 4, or the whole image (<= 8)
constexpr int RG_WST_RING = 4;

// mbarrier wait of this kernel's dedicated warps: plain try_wait polling (no suspend-time hint; measured: with or without
// the 2 us hint of tc05::mbar_wait the kernel runs the same, the roles here never share issue slots with a gather loop).
// Traps after ~2 s instead of hanging the device.
__device__ __forceinline__ void rg_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    long long t0 = 0;
    for (unsigned it = 0;; it++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (it == 64) t0 = clock64();
        if (it > 64 && (it & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}

struct RowsGemmArgs {
    unsigned R;
    int K, N, KC, MB;
    int ast, wst, resident;                      // operand stages; resident: the ring holds the whole weight image
    int mbg;                                     // blocks of 128 output columns per CTA (2, or 1 to spread a small problem)
    int terms;                                   // bf16 terms per operand: 3 (six cross products) or 2 (four)
    const float* x;
    const unsigned char* image;                  // [KC][3 terms][MB][128 x 128 B]
    float* y;
    long long* dbg;                              // optional: clock64 stamps of CTA (0,0), [6 roles: producer, issuer, 4 epilogue warps][64 steps][4] (profiles/check_rowsgemm.py --trace)
};

struct RowsGemmSync {
    uint64_t a_full[RG_AST_MAX], a_empty[RG_AST_MAX];
    uint64_t w_full[RG_WST_MAX], w_empty[RG_WST_MAX];
    uint64_t acc_full[2], acc_free[2];
    uint32_t tmem;
};

// W -> image[kc][term][mb][n][swizzled 64 k] of bf16, zero padded.  trans = 0: W is (K x N) row-major and the product is
// x * W; trans = 1: W is (N x K) row-major and the product is x * W^T.  Row n of block mb is output column mb*128 + n.
// One thread converts the eight k of one 16-byte chunk of a row and writes the chunk of each term with one 16-byte store;
// n is the fastest thread index (trans = 0: coalesced reads; trans = 1: every thread reads one full 32-byte sector).
__device__ __forceinline__ void rows_gemm_pack_chunk(int t, int K, int N, int MB, int trans, const float* __restrict__ W,
                                                     unsigned char* __restrict__ image)
{
    const int n = t % 128, c = (t / 128) % 8, mb = (t / 1024) % MB, kc = t / (1024 * MB);
    const int k0 = kc * UNIT_K + c * 8, co = mb * 128 + n;
    uint32_t hi[4], mid[4], lo[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
        float v0 = 0.f, v1 = 0.f;
        if (co < N) {
            if (k0 + e < K) v0 = trans ? __ldg(W + (size_t)co * K + k0 + e) : __ldg(W + (size_t)(k0 + e) * N + co);
            if (k0 + e + 1 < K) v1 = trans ? __ldg(W + (size_t)co * K + k0 + e + 1) : __ldg(W + (size_t)(k0 + e + 1) * N + co);
        }
        split3_pack2(v0, v1, hi[e / 2], mid[e / 2], lo[e / 2]);
    }
    unsigned char* base = image + ((size_t)(kc * 3) * MB + mb) * RG_UNIT + unit_offset(n, c * 8);
    *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + (size_t)MB * RG_UNIT) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
    *reinterpret_cast<uint4*>(base + (size_t)2 * MB * RG_UNIT) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// image of W for the product it is asked for (K, N, trans) and, when image2 != NULL, in the same launch the image of the
// opposite orientation (N, K, !trans): a training step needs both x * W (forward) and g * W^T (input gradient)
__global__ void __launch_bounds__(256)
rows_gemm_pack_kernel(int K, int N, int trans, const float* __restrict__ W, unsigned char* __restrict__ image,
                      unsigned char* __restrict__ image2)
{
    const int KC = (K + UNIT_K - 1) / UNIT_K, MB = (N + 127) / 128;
    const int KC2 = (N + UNIT_K - 1) / UNIT_K, MB2 = (K + 127) / 128;
    const int total = KC * MB * 1024, total2 = image2 ? KC2 * MB2 * 1024 : 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total + total2; t += gridDim.x * blockDim.x) {
        if (t < total) rows_gemm_pack_chunk(t, K, N, MB, trans, W, image);
        else rows_gemm_pack_chunk(t - total, N, K, MB2, !trans, W, image2);
    }
}

// NT = bf16 terms per operand, MBGT = blocks of 128 output columns per CTA: compile-time so that the issuer's products of a
// chunk are straight-line code (with run-time loop bounds the single issuing warp spent more cycles per tcgen05.mma on loop
// and address arithmetic than the tensor core needs for the product: 0.076 -> 0.11 ms at R = 320 000, K = N = 128)
template <int NT, int MBGT>
__global__ void __launch_bounds__(RG_WARPS * 32, 1)
rows_gemm_kernel(const RowsGemmArgs a)
{
    extern __shared__ __align__(16) unsigned char rg_smem[];
    RowsGemmSync* sy = reinterpret_cast<RowsGemmSync*>(rg_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t a_full = smem_u32(sy->a_full), a_empty = smem_u32(sy->a_empty);
    const uint32_t w_full = smem_u32(sy->w_full), w_empty = smem_u32(sy->w_empty);
    const uint32_t acc_full = smem_u32(sy->acc_full), acc_free = smem_u32(sy->acc_free);
    const uint32_t base = (smem_u32(sy + 1) + 1023u) & ~1023u;
    const unsigned AST = (unsigned)a.ast, WST = (unsigned)a.wst;
    const uint32_t A_off = 0, W_off = AST * NT * RG_UNIT;

    const int mb0 = blockIdx.y * MBGT;                               // my blocks of 128 output columns: mb0 .. mb0 + MBG - 1
    const int MBG = min(MBGT, a.MB - mb0);
    const unsigned ntiles = (a.R + RG_TM - 1) / RG_TM;
    const unsigned my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const unsigned KC = (unsigned)a.KC;
    const unsigned NU = KC * NT * (unsigned)MBG;                    // weight units per tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < RG_AST_MAX; s++) { mbar_init(a_full + 8u * s, RG_PROD); mbar_init(a_empty + 8u * s, 1); }
        for (int s = 0; s < RG_WST_MAX; s++) { mbar_init(w_full + 8u * s, 1); mbar_init(w_empty + 8u * s, 1); }
        for (int d = 0; d < 2; d++) { mbar_init(acc_full + 8u * d, 1); mbar_init(acc_free + 8u * d, RG_EPI); }
        fence_mbar_init();
    }
    const uint32_t tmem_cols = tmem_cols_pow2(2u * MBGT * 128u);
    if (warp == 0) tmem_alloc(smem_u32(&sy->tmem), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sy->tmem;

    if (warp < RG_EPI) {
        // ------------------------------------------------------------------ epilogue: accumulator -> y
        // lane = output column 32*q + lane of the block, tensor-memory column = row of the tile
        const int q = warp;
        for (unsigned tau = 0; tau < my_tiles; tau++) {
            const unsigned d = tau & 1u, use = tau >> 1;
            const bool tr = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && tau < 64;
            if (tr) a.dbg[((2 + q) * 64 + tau) * 4 + 0] = clock64();
            rg_wait(acc_full + 8u * d, use & 1u);
            tc_fence_after();
            if (tr) a.dbg[((2 + q) * 64 + tau) * 4 + 1] = clock64();
            const unsigned row0 = (blockIdx.x + tau * gridDim.x) * RG_TM;
            const int nrows = (int)min((unsigned)RG_TM, a.R - row0);
            for (int j = 0; j < MBG; j++) {
                const int col = (mb0 + j) * 128 + 32 * q + lane;
                if ((mb0 + j) * 128 + 32 * q >= a.N) break;          // warp-uniform: this quarter is beyond the last column
                float* ycol = a.y + (size_t)row0 * a.N + col;
                const uint32_t tbase = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)((d * MBGT + j) * 128);
#pragma unroll 1
                for (int r64 = 0; r64 < nrows; r64 += 64) {         // two queued loads of 32 rows each, ONE wait
                    uint32_t v0[32], v1[32];
                    tmem_ld32_nowait(tbase + (uint32_t)r64, v0);
                    tmem_ld32_nowait(tbase + (uint32_t)r64 + 32u, v1);
                    tmem_ld_wait();
                    if (tr && j == 0 && r64 == 0) a.dbg[((2 + q) * 64 + tau) * 4 + 3] = clock64();
                    if (col < a.N) {
                        float* yc = ycol + (size_t)r64 * a.N;
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (r64 + i < nrows) yc[(size_t)i * a.N] = __uint_as_float(v0[i]);
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (r64 + 32 + i < nrows) yc[(size_t)(32 + i) * a.N] = __uint_as_float(v1[i]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free + 8u * d);
            if (tr) a.dbg[((2 + q) * 64 + tau) * 4 + 2] = clock64();
        }
    } else if (warp == RG_EPI) {
        // ------------------------------------------------------------------ issuer (the whole warp runs the loop with
        // warp-uniform values, one elected lane executes each tcgen05 instruction: tc05.cuh)
        {
            const uint32_t idesc = idesc_bf16_f32(128, 128);         // M = 128 output columns, N = 128 rows of the tile
            unsigned g = 0, L = 0;
            for (unsigned tau = 0; tau < my_tiles; tau++) {
                const unsigned d = tau & 1u, use = tau >> 1;
                if (use >= 1) rg_wait(acc_free + 8u * d, (use - 1u) & 1u);
                tc_fence_after();
                for (unsigned kc = 0; kc < KC; kc++, g++) {
                    const unsigned sa = g % AST;
                    const bool tr = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && g < 64 && lane == 0;
                    if (tr) a.dbg[(1 * 64 + g) * 4 + 0] = clock64();
                    rg_wait(a_full + 8u * sa, (g / AST) & 1u);
                    tc_fence_after();
                    if (tr) a.dbg[(1 * 64 + g) * 4 + 1] = clock64();
                    const uint32_t xlo = smem_desc_lo(base + A_off + sa * NT * RG_UNIT);
                    // three terms: weight term tw meets row terms 0 .. 2 - tw (six products, everything down to 2^-24 of a
                    // product); two terms: all four products of hi + mid (what is dropped is below 2^-17 of a product)
#pragma unroll
                    for (int tw = 0; tw < NT; tw++)
#pragma unroll
                        for (int j = 0; j < MBGT; j++) {
                            if (j >= MBG) continue;                  // odd number of blocks: the last group has one
                            const unsigned sw = a.resident ? L % NU : L % WST;
                            if (!a.resident || tau == 0) {           // a resident image is waited for once, on the first tile
                                rg_wait(w_full + 8u * sw, a.resident ? 0u : (L / WST) & 1u);
                                tc_fence_after();
                            }
                            const uint32_t wlo = smem_desc_lo(base + W_off + sw * RG_UNIT);
                            const uint32_t dcol = tmem + (uint32_t)((d * MBGT + j) * 128);
#pragma unroll
                            for (int tt = 0; tt < NT; tt++) {
                                if (NT == 3 && tt + tw > 2) continue;
                                const uint32_t tlo = xlo + (uint32_t)tt * (RG_UNIT >> 4);
#pragma unroll
                                for (int ks = 0; ks < UNIT_K / UMMA_K; ks++)
                                    umma_bf16_lo(dcol, wlo + 2u * ks, tlo + 2u * ks, idesc, (kc | (unsigned)(tw | tt | ks)) != 0u);
                            }
                            if (!a.resident) umma_commit_elect(w_empty + 8u * sw);
                            L++;
                        }
                    umma_commit_elect(a_empty + 8u * sa);
                    if (tr) a.dbg[(1 * 64 + g) * 4 + 2] = clock64();
                }
                umma_commit_elect(acc_full + 8u * d);
            }
        }
    } else if (warp == RG_EPI + 1) {
        // ------------------------------------------------------------------ weight loader
        if (lane == 0) {
            const unsigned total = a.resident ? min(my_tiles, 1u) * NU : my_tiles * NU;
            for (unsigned L = 0; L < total; L++) {
                const unsigned s = a.resident ? L : L % WST;
                if (!a.resident && L >= WST) rg_wait(w_empty + 8u * s, ((L / WST) - 1u) & 1u);
                const unsigned u = L % NU;
                const unsigned kc = u / (NT * MBG), tw = (u / MBG) % NT, j = u % MBG;
                mbar_expect_tx(w_full + 8u * s, RG_UNIT);
                bulk_g2s(base + W_off + s * RG_UNIT, a.image + ((size_t)(kc * 3u + tw) * a.MB + mb0 + j) * RG_UNIT, RG_UNIT,
                         w_full + 8u * s);
            }
        }
    } else {
        // ------------------------------------------------------------------ producers: rows -> three bf16 terms in the operand
        // warp pw converts rows 8*pw .. +7 of every chunk: lanes 0-15 the even row of a pair, lanes 16-31 the odd one, four
        // consecutive k each (256 contiguous bytes per row and chunk)
        const int pw = warp - (RG_EPI + 2);
        const int half = lane >> 4;
        const unsigned k4 = (unsigned)(lane & 15) * 4u;
        const unsigned total = my_tiles * KC;
        auto load = [&](unsigned g, float4 (&v)[RG_PR]) {
            const unsigned tau = g / KC, kc = g - tau * KC;
            const unsigned row0 = (blockIdx.x + tau * gridDim.x) * RG_TM + (unsigned)pw * (2u * RG_PR) + half;
            const unsigned k = kc * UNIT_K + k4;
#pragma unroll
            for (int p = 0; p < RG_PR; p++) {
                const unsigned r = row0 + 2u * p;
                v[p] = (r < a.R && k < (unsigned)a.K) ? __ldg(reinterpret_cast<const float4*>(a.x + (size_t)r * a.K + k))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto emit = [&](unsigned g, const float4 (&v)[RG_PR]) {
            const unsigned sa = g % AST;
            const bool tr = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && pw == 0 && lane == 0 && g < 64;
            if (tr) a.dbg[(0 * 64 + g) * 4 + 0] = clock64();
            if (g >= AST) rg_wait(a_empty + 8u * sa, ((g / AST) - 1u) & 1u);
            if (tr) a.dbg[(0 * 64 + g) * 4 + 1] = clock64();
            const uint32_t xbase = base + A_off + sa * NT * RG_UNIT;
#pragma unroll
            for (int p = 0; p < RG_PR; p++) {
                const unsigned m = (unsigned)pw * (2u * RG_PR) + 2u * p + half;
                const uint32_t dst = xbase + unit_offset(m, k4);
                uint32_t h0, m0, l0, h1, m1, l1;
                split3_pack2(v[p].x, v[p].y, h0, m0, l0);
                split3_pack2(v[p].z, v[p].w, h1, m1, l1);
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst), "r"(h0), "r"(h1) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst + RG_UNIT), "r"(m0), "r"(m1) : "memory");
                if (NT == 3) asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst + 2 * RG_UNIT), "r"(l0), "r"(l1) : "memory");
            }
            fence_proxy_async_smem();                                // ordinary stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8u * sa);
            if (tr) a.dbg[(0 * 64 + g) * 4 + 2] = clock64();
        };
        // three register buffers: chunks g+1 and g+2 are in flight while chunk g is converted
        float4 v0[RG_PR], v1[RG_PR], v2[RG_PR];
        if (total > 0) load(0, v0);
        if (total > 1) load(1, v1);
        for (unsigned g = 0; g < total; g += 3) {
            if (g + 2 < total) load(g + 2, v2);
            emit(g, v0);
            if (g + 1 < total) {
                if (g + 3 < total) load(g + 3, v0);
                emit(g + 1, v1);
            }
            if (g + 2 < total) {
                if (g + 4 < total) load(g + 4, v1);
                emit(g + 2, v2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

struct RowsGemmGeom { int ast, wst, resident; size_t smem; };
static inline RowsGemmGeom rows_gemm_geom(int KC, int MB, int terms, int mbg_max)
{
    RowsGemmGeom g{};
    const int mbg = MB < mbg_max ? MB : mbg_max;
    const int nu = KC * terms * mbg;                                 // weight units of one column group
    auto bytes = [&](int ast, int wst) { return sizeof(RowsGemmSync) + 1024 + (size_t)ast * terms * RG_UNIT + (size_t)wst * RG_UNIT; };
    // every CTA loads a small image once and keeps it, as long as two stages of the row operand still fit beside it
    g.resident = (nu <= RG_WST_MAX && MB <= mbg_max && bytes(2, nu) <= SMEM_CAP) ? 1 : 0;
    g.wst = g.resident ? nu : RG_WST_RING;
    g.ast = RG_AST_MAX;
    while (g.ast > 2 && bytes(g.ast, g.wst) > SMEM_CAP) g.ast--;
    g.smem = bytes(g.ast, g.wst);
    return g;
}

}  // namespace sph3d

using namespace sph3d;

extern "C" size_t sph3d_rows_gemm_image_bytes(int K, int N)
{
    if (K <= 0 || N <= 0) return 0;
    const size_t KC = (K + UNIT_K - 1) / UNIT_K, MB = (N + 127) / 128;
    return KC * 3 * MB * RG_UNIT;
}

static int rows_gemm_pack_launch(int K, int N, const float* weights, int trans, void* image, void* image2, void* stream)
{
    g_last_launch_count = 0;
    if (K <= 0 || N <= 0 || !weights || !image) return (int)cudaErrorInvalidValue;
    const long long KC = (K + UNIT_K - 1) / UNIT_K, MB = (N + 127) / 128, KC2 = (N + UNIT_K - 1) / UNIT_K, MB2 = (K + 127) / 128;
    const long long total = KC * MB * 1024 + (image2 ? KC2 * MB2 * 1024 : 0);
    if (total >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
    rows_gemm_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(K, N, trans ? 1 : 0, weights, static_cast<unsigned char*>(image),
                                                                 static_cast<unsigned char*>(image2));
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}

extern "C" int sph3d_rows_gemm_pack(int K, int N, const float* weights, int trans, void* image, void* stream)
{
    return rows_gemm_pack_launch(K, N, weights, trans, image, nullptr, stream);
}

extern "C" int sph3d_rows_gemm_pack_pair(int K, int N, const float* weights, void* image, void* image_t, void* stream)
{
    if (!image_t) return (int)cudaErrorInvalidValue;
    return rows_gemm_pack_launch(K, N, weights, 0, image, image_t, stream);
}

static long long* g_rows_gemm_dbg = nullptr;
extern "C" void sph3d_rows_gemm_trace(void* buffer) { g_rows_gemm_dbg = static_cast<long long*>(buffer); }   // 6*64*4 int64, or NULL

extern "C" int sph3d_rows_gemm(int R, int K, int N, int terms, const float* x, const void* image, float* y, void* stream)
{
    g_last_launch_count = 0;
    if (R <= 0 || K <= 0 || N <= 0 || (terms != 2 && terms != 3) || !x || !image || !y) return (int)cudaErrorInvalidValue;
    if ((K & 3) || (N & 3) || (((uintptr_t)x | (uintptr_t)y | (uintptr_t)image) & 15)) return (int)cudaErrorInvalidValue;
    RowsGemmArgs a{};
    a.R = (unsigned)R; a.K = K; a.N = N;
    a.KC = (K + UNIT_K - 1) / UNIT_K; a.MB = (N + 127) / 128;
    a.x = x; a.image = static_cast<const unsigned char*>(image); a.y = y;
    a.dbg = g_rows_gemm_dbg;
    const unsigned ntiles = ((unsigned)R + RG_TM - 1) / RG_TM;
    // two blocks of 128 output columns per CTA (the rows are converted once for both) unless that leaves most SMs idle
    a.mbg = (a.MB >= 2 && (long long)ntiles * ((a.MB + 1) / 2) * 3 < (long long)sm_count() * 2) ? 1 : 2;
    const unsigned groups = (unsigned)(a.MB + a.mbg - 1) / a.mbg;
    unsigned gx = (unsigned)sm_count() / groups;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    const RowsGemmGeom geo = rows_gemm_geom(a.KC, a.MB, terms, a.mbg);
    a.terms = terms;
    a.ast = geo.ast; a.wst = geo.wst; a.resident = geo.resident;
    cudaError_t e = cudaSuccess;
#define LAUNCH_RG(NT_, MBG_)                                                                             \
    do {                                                                                                  \
        e = set_smem(rows_gemm_kernel<NT_, MBG_>, geo.smem);                                              \
        if (e != cudaSuccess) return (int)e;                                                              \
        rows_gemm_kernel<NT_, MBG_><<<dim3(gx, groups), RG_WARPS * 32, geo.smem, (cudaStream_t)stream>>>(a); \
    } while (0)
    if (terms == 3 && a.mbg == 2) LAUNCH_RG(3, 2);
    else if (terms == 3) LAUNCH_RG(3, 1);
    else if (a.mbg == 2) LAUNCH_RG(2, 2);
    else LAUNCH_RG(2, 1);
#undef LAUNCH_RG
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}
