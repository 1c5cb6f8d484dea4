// rowswgrad.cu -- the weight gradient of an SPH3D layer's pointwise product, hand-written for the sm_100a tensor cores:
//      gw (K x N) = x^T (K x R) * g (R x N)             a sum over the R = B*M rows of the layer
// (the gradient of the tf.matmul of /root/reference/utils/sph3gcn_util.py:144-146, :203-205, :254-256 w.r.t. its weights).
// Third member of the family rowsgemm.cu starts (same three-term bf16 split, same products, same tc05.cuh plumbing); what
// is different is that the contraction runs over the ROWS, the slow index of both operands:
//
//   * both operands are MN-major for the tensor core: a tile of 64 rows x 64 columns, stored row by row (128 bytes of
//     bf16 per row, 16-byte chunks XOR-swizzled with the row's position in its group of eight) -- byte for byte the unit
//     rowsgemm.cu writes -- is the canonical "MN-major, 128-byte swizzle" operand whose MN index is the COLUMN and whose
//     K index is the ROW: leading byte offset = distance between the units of neighbouring 64-column chunks, stride
//     byte offset = 1024 (eight rows), and one tcgen05.mma (16 rows) advances the start address by 2048 bytes.  The
//     producers therefore convert x and g exactly as they arrive (coalesced 16-byte loads, split in registers, 8-byte
//     st.shared), and nothing is transposed anywhere;
//   * a CTA owns one 128 x 128 block of gw (tensor-memory accumulator, lane = k, column = n) over one SLAB of rows:
//     grid = (blocks of gw) x (slabs), sized to fill the machine; the slabs' partial blocks are summed in slab order by
//     reduce_partials_kernel (deterministic);
//   * 16 producer warps fill a stage of 64 rows (x: 128 columns of the block's k range, g: 128 columns of its n range,
//     three terms each = 96 KB; two stages), one warp issues 6 products x 4 k-steps per stage and commits the stage
//     back; the next stage's loads are already in flight in the producers' registers.
// Roles by warp: 0 issuer, 1-16 producers; warps 1-4 (four different warp % 4) read the accumulator out at the end.
#include "conv_common.cuh"
#include "tc05.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

using namespace tc05;

constexpr int WG_ROWS = 64;                       // rows per stage
constexpr int WG_UNIT = WG_ROWS * 128;            // bytes of one 64-row x 64-column bf16 unit
constexpr int WG_PROD = 16;
constexpr int WG_WARPS = 1 + WG_PROD;
constexpr int WG_STAGES = 2;

struct RowsWgradArgs {
    unsigned R, rows_per_slab;
    int K, N, KB, NB;                             // KB / NB: 128-blocks of gw along k / n
    const float* x;
    const float* g;
    float* part;                                  // [slabs][K][N]
};

struct RowsWgradSync {
    uint64_t full[WG_STAGES], empty[WG_STAGES];
    uint64_t acc_full;
    uint32_t tmem;
};

__device__ __forceinline__ void wg_wait(uint32_t bar, uint32_t parity)
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

// instruction descriptor of rowsgemm's products with BOTH operands MN-major (bits 15 / 16)
__host__ __device__ __forceinline__ uint32_t idesc_bf16_f32_mn(uint32_t M, uint32_t N)
{
    return idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
}
// low / high words of the MN-major SWIZZLE_128B descriptor: start address >> 4 | leading byte offset >> 4 at [16,30);
// stride byte offset 1024 >> 4 at [32,46), version 1 at [46,48), layout type 2 at [61,64)
__device__ __forceinline__ uint32_t smem_desc_lo_mn(uint32_t addr, uint32_t lbo_bytes)
{
    return ((addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}

template <int NT>
__global__ void __launch_bounds__(WG_WARPS * 32, 1)
rows_wgrad_kernel(const RowsWgradArgs a)
{
    extern __shared__ __align__(16) unsigned char wg_smem[];
    RowsWgradSync* sy = reinterpret_cast<RowsWgradSync*>(wg_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t full = smem_u32(sy->full), empty = smem_u32(sy->empty), acc_full = smem_u32(&sy->acc_full);
    const uint32_t base = (smem_u32(sy + 1) + 1023u) & ~1023u;
    constexpr uint32_t STAGE = NT * 4u * WG_UNIT;              // [term][x chunk 0, x chunk 1, g chunk 0, g chunk 1]

    const int kb = blockIdx.x / a.NB, nb = blockIdx.x - kb * a.NB;
    const unsigned slab = blockIdx.y;
    const unsigned row_beg = slab * a.rows_per_slab;
    const unsigned row_end = min(a.R, row_beg + a.rows_per_slab);
    const unsigned nstage = row_beg < row_end ? (row_end - row_beg + WG_ROWS - 1) / WG_ROWS : 0u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; s++) { mbar_init(full + 8u * s, WG_PROD); mbar_init(empty + 8u * s, 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&sy->tmem), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sy->tmem;

    if (warp == 0) {
        // ------------------------------------------------------------------ issuer (warp-uniform, elected lane issues)
        const uint32_t idesc = idesc_bf16_f32_mn(128, 128);
        for (unsigned st = 0; st < nstage; st++) {
            const unsigned s = st % WG_STAGES;
            wg_wait(full + 8u * s, (st / WG_STAGES) & 1u);
            tc_fence_after();
            const uint32_t sbase = base + s * STAGE;
#pragma unroll
            for (int tx = 0; tx < NT; tx++)
#pragma unroll
                for (int tg = 0; tg < NT; tg++) {
                    if (NT == 3 && tx + tg > 2) continue;
                    const uint32_t alo = smem_desc_lo_mn(sbase + (uint32_t)(tx * 4) * WG_UNIT, WG_UNIT);
                    const uint32_t blo = smem_desc_lo_mn(sbase + (uint32_t)(tg * 4 + 2) * WG_UNIT, WG_UNIT);
#pragma unroll
                    for (int ks = 0; ks < WG_ROWS / UMMA_K; ks++)         // 16 rows = two 1024-byte groups per step
                        umma_bf16_lo(tmem, alo + (2048u >> 4) * ks, blo + (2048u >> 4) * ks, idesc, (st | (unsigned)(tx | tg | ks)) != 0u);
                }
            umma_commit_elect(empty + 8u * s);
        }
        umma_commit_elect(acc_full);
    } else {
        // ------------------------------------------------------------------ producers
        // warp pw converts rows 4*pw .. +3 of every stage: per row one 16-byte load per lane from x (columns kb*128 + 4*lane)
        // and one from g (columns nb*128 + 4*lane)
        const int pw = warp - 1;
        const unsigned cx = (unsigned)kb * 128u + 4u * lane, cg = (unsigned)nb * 128u + 4u * lane;
        const bool okx = cx < (unsigned)a.K, okg = cg < (unsigned)a.N;       // K, N multiples of 4
        const uint32_t unit = (uint32_t)(lane >> 4);                          // which 64-column chunk my four columns are in
        const uint32_t c64 = (4u * lane) & 63u;
        auto load = [&](unsigned st, float4 (&vx)[4], float4 (&vg)[4]) {
            const unsigned r0 = row_beg + st * WG_ROWS + (unsigned)pw * 4u;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const unsigned r = r0 + p;
                const bool in = r < row_end;
                vx[p] = (in && okx) ? __ldg(reinterpret_cast<const float4*>(a.x + (size_t)r * a.K + cx)) : make_float4(0.f, 0.f, 0.f, 0.f);
                vg[p] = (in && okg) ? __ldg(reinterpret_cast<const float4*>(a.g + (size_t)r * a.N + cg)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto put = [&](uint32_t dst, const float4& v) {
            uint32_t h0, m0, l0, h1, m1, l1;
            split3_pack2(v.x, v.y, h0, m0, l0);
            split3_pack2(v.z, v.w, h1, m1, l1);
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst), "r"(h0), "r"(h1) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst + 4u * WG_UNIT), "r"(m0), "r"(m1) : "memory");
            if (NT == 3) asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(dst + 8u * WG_UNIT), "r"(l0), "r"(l1) : "memory");
        };
        auto emit = [&](unsigned st, const float4 (&vx)[4], const float4 (&vg)[4]) {
            const unsigned s = st % WG_STAGES;
            if (st >= WG_STAGES) wg_wait(empty + 8u * s, ((st / WG_STAGES) - 1u) & 1u);
            const uint32_t sbase = base + s * STAGE;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const uint32_t off = unit_offset((uint32_t)pw * 4u + p, c64);
                put(sbase + unit * WG_UNIT + off, vx[p]);
                put(sbase + (2u + unit) * WG_UNIT + off, vg[p]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + 8u * s);
        };
        float4 x0[4], g0[4], x1[4], g1[4];
        if (nstage > 0) load(0, x0, g0);
        for (unsigned st = 0; st < nstage; st += 2) {
            if (st + 1 < nstage) load(st + 1, x1, g1);
            emit(st, x0, g0);
            if (st + 1 < nstage) {
                if (st + 2 < nstage) load(st + 2, x0, g0);
                emit(st + 1, x1, g1);
            }
        }
        // ------------------------------------------------------------------ epilogue (warps 1-4): accumulator -> partial block
        if (warp <= 4) {
            const int q = warp & 3;
            wg_wait(acc_full, 0u);
            tc_fence_after();
            const int k = kb * 128 + 32 * q + lane;
            float* prow = a.part + ((size_t)slab * a.K + (size_t)(k < a.K ? k : 0)) * a.N + (size_t)nb * 128;
            const int ncols = min(128, a.N - nb * 128);
#pragma unroll 1
            for (int c64b = 0; c64b < ncols; c64b += 64) {
                uint32_t v0[32], v1[32];
                const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c64b;
                tmem_ld32_nowait(taddr, v0);
                tmem_ld32_nowait(taddr + 32u, v1);
                tmem_ld_wait();
                if (k < a.K) {
                    if (nstage == 0) {                                  // an empty slab contributes zeros
#pragma unroll
                        for (int i = 0; i < 32; i++) { v0[i] = 0u; v1[i] = 0u; }
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (c64b + i < ncols)
                            *reinterpret_cast<uint4*>(prow + c64b + i) = make_uint4(v0[i], v0[i + 1], v0[i + 2], v0[i + 3]);
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        if (c64b + 32 + i < ncols)
                            *reinterpret_cast<uint4*>(prow + c64b + 32 + i) = make_uint4(v1[i], v1[i + 1], v1[i + 2], v1[i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

struct WgradGeom { int KB, NB; unsigned slabs, rows_per_slab; size_t part_bytes; };
static inline WgradGeom wgrad_geom(int R, int K, int N)
{
    WgradGeom g{};
    g.KB = (K + 127) / 128; g.NB = (N + 127) / 128;
    const long long blocks = (long long)g.KB * g.NB;
    // two waves of CTAs at most, a slab at least 8 stages long (the pipeline's fill and the final read-out are paid per CTA)
    long long slabs = (2LL * sm_count() + blocks - 1) / blocks;
    const long long max_slabs = ((long long)R + 8 * WG_ROWS - 1) / (8 * WG_ROWS);
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    unsigned rps = (unsigned)(((long long)R + slabs - 1) / slabs);
    rps = (rps + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
    g.rows_per_slab = rps;
    g.slabs = (unsigned)(((long long)R + rps - 1) / rps);
    g.part_bytes = (size_t)g.slabs * K * N * sizeof(float);
    return g;
}

}  // namespace sph3d

using namespace sph3d;

extern "C" size_t sph3d_rows_wgrad_workspace_bytes(int R, int K, int N)
{
    if (R <= 0 || K <= 0 || N <= 0) return 0;
    return wgrad_geom(R, K, N).part_bytes;
}

extern "C" int sph3d_rows_wgrad(int R, int K, int N, int terms, const float* x, const float* g, float* gw, void* workspace,
                                size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    if (R <= 0 || K <= 0 || N <= 0 || (terms != 2 && terms != 3) || !x || !g || !gw || !workspace) return (int)cudaErrorInvalidValue;
    if ((K & 3) || (N & 3) || (((uintptr_t)x | (uintptr_t)g | (uintptr_t)gw | (uintptr_t)workspace) & 15)) return (int)cudaErrorInvalidValue;
    const WgradGeom geo = wgrad_geom(R, K, N);
    if (workspace_bytes < geo.part_bytes || (long long)geo.KB * geo.NB > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
    RowsWgradArgs a{};
    a.R = (unsigned)R; a.rows_per_slab = geo.rows_per_slab; a.K = K; a.N = N; a.KB = geo.KB; a.NB = geo.NB;
    a.x = x; a.g = g;
    a.part = geo.slabs == 1 ? gw : static_cast<float*>(workspace);   // a single slab writes the result itself: no partial, no sum
    const size_t smem = sizeof(RowsWgradSync) + 1024 + (size_t)WG_STAGES * terms * 4 * WG_UNIT;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    const dim3 grid((unsigned)(geo.KB * geo.NB), geo.slabs);
    if (terms == 3) {
        e = set_smem(rows_wgrad_kernel<3>, smem);
        if (e != cudaSuccess) return (int)e;
        rows_wgrad_kernel<3><<<grid, WG_WARPS * 32, smem, st>>>(a);
    } else {
        e = set_smem(rows_wgrad_kernel<2>, smem);
        if (e != cudaSuccess) return (int)e;
        rows_wgrad_kernel<2><<<grid, WG_WARPS * 32, smem, st>>>(a);
    }
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    if (geo.slabs > 1) {
        const int rc = launch_reduce_partials((int)geo.slabs, (size_t)K * N, a.part, gw, st);
        if (rc) return rc;
        g_last_launch_count = 2;
    }
    return 0;
}
