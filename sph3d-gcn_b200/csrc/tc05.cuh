// tc05.cuh -- the sm_100a tensor-core plumbing the hand-written kernels of this library share: mbarriers, the bulk
// (TMA, non-tensor) global->shared copy, tensor-memory allocation, UMMA shared-memory / instruction descriptors,
// tcgen05.mma / commit / ld.  Raw PTX only: nothing here comes from a template library.
//
// Operand layout used throughout ("K-major, 128-byte swizzle"): an operand tile is a stack of UNITS, one unit =
// rows x 64 bf16 (128 bytes per row, rows a multiple of 8), unit base 1024-byte aligned.  Element (row m, k) of a unit
// lives at byte   m*128 + (((k>>3) ^ (m&7)) << 4) + (k&7)*2 :   the 16-byte chunk index is XORed with the row's position
// inside its 8-row group, which is what the tensor core undoes when the descriptor says SWIZZLE_128B.  One tcgen05.mma
// of kind::f16 consumes 16 k-values = 32 bytes of every row: k-step j of a unit starts at unit base + 32*j.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace sph3d {
namespace tc05 {

constexpr int UNIT_K = 64;                      // bf16 elements per 128-byte swizzled row
constexpr int UMMA_K = 16;                      // k-values one kind::f16 instruction consumes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (m, k) inside a unit, k < 64
__host__ __device__ __forceinline__ uint32_t unit_offset(uint32_t m, uint32_t k)
{
    return m * 128u + ((((k >> 3) ^ (m & 7u)) & 7u) << 4) + (k & 7u) * 2u;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    // the last operand is the suspend-time hint (ns): a waiting thread sleeps in hardware instead of spinning through
    // issue slots the gathering warps need
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(2000u) : "memory");
    return ok != 0;
}
// A barrier that never completes would hang the device; a wait that lasts over ~2 s traps instead (the launch then
// fails with an error the caller sees).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}

// ---- bulk copy global -> shared (TMA engine, 1-D: no tensor map), completion counted on an mbarrier ----
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// whole-warp calls; ncols a power of two in [32, 512]; the base address lands in *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
static inline __host__ __device__ uint32_t tmem_cols_pow2(uint32_t n)
{
    uint32_t c = 32;
    while (c < n) c <<= 1;
    return c;
}

// ---- descriptors ----
// shared-memory matrix descriptor of a K-major SWIZZLE_128B unit whose 8-row groups are 1024 bytes apart:
// start address >> 4 in [0,14), leading byte offset (unused for swizzled K-major) = 1 in [16,30), stride byte offset
// 1024 >> 4 in [32,46), descriptor version 1 (Blackwell) in [46,48), layout type 2 = SWIZZLE_128B in [61,64)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr)
{
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D = fp32 (1 at [4,6)), A = B = bf16 (1 at [7,10) and [10,13)), both K-major
// (0 at 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ __forceinline__ uint32_t idesc_bf16_f32(uint32_t M, uint32_t N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Warp-uniform issue.  Inside `if (lane == 0)` the compiler treats the descriptors as per-lane values: every UTCHMMA is
// wrapped in an ELECT / BRA.U.ANY loop behind R2UR moves and four uniform ALU operations that rebuild the descriptor
// (~13 dependent instructions of ONE warp per MMA = more cycles than the 64 the tensor core needs for a 128 x 128 x 16
// product; measured in rowsgemm.cu: 135 cycles per MMA).  Here the whole warp runs the loop with warp-uniform values,
// one elected lane executes the instruction, and a k-step only adds 2 to the low descriptor word (32 bytes >> 4).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t SMEM_DESC_HI_SW128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
// D[tmem] (+)= A * B with the descriptors given by their low words (high word = SMEM_DESC_HI_SW128); whole-warp call,
// one elected lane issues
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate)
{
    if (elect_one())
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                     :: "r"(d_tmem), "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"(SMEM_DESC_HI_SW128) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar)
{
    if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// the mbarrier receives one arrival when every tcgen05.mma issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// 16 consecutive fp32 columns of this thread's lane (warp w reads lanes 32*(w%4) .. +31): whole-warp call
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive columns WITHOUT the wait: several loads can be queued before one tmem_ld_wait().  (A tcgen05.ld issued
// while tcgen05.mma instructions are queued completes only ~1 300 cycles later -- measured in rowsgemm.cu with clock64
// stamps -- so an epilogue that waits after every 16 columns pays that latency eight times per accumulator.)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 8 / 4 consecutive columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4])
{
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __uint_as_float(r[i]);
}

// ---- fp32 -> three bf16 terms (v = hi + mid + lo to ~2^-24 relative) ----
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// two values -> packed bf16x2 words of their hi / mid / lo terms (element 0 in the low half)
__device__ __forceinline__ void split3_pack2(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float ra = a - __low2float(h), rb = b - __high2float(h);
    __nv_bfloat162 m = __floats2bfloat162_rn(ra, rb);
    __nv_bfloat162 l = __floats2bfloat162_rn(ra - __low2float(m), rb - __high2float(m));
    hi = *reinterpret_cast<uint32_t*>(&h);
    mid = *reinterpret_cast<uint32_t*>(&m);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

}  // namespace tc05
}  // namespace sph3d
