// sample.cu -- farthest point sampling, sm_100a.
//
// Replaces farthestPointSampleLauncher (/root/reference/tf_ops/sampling/tf_sample_gpu.cu:77-80,
// kernel :7-73).  FPS is `npoint` strictly sequential rounds, so the only thing that matters is
// the latency of one round.  The reference keeps the running min-distance array in GLOBAL memory
// (re-read and re-written every round), reads the picked point's coordinates from global memory
// and reduces with a 10-level shared-memory tree (11 __syncthreads per round), one CTA per cloud.
//
// Here a cloud lives entirely in REGISTERS of one thread-block cluster: CS CTAs x 1024 threads,
// each thread owns P points (xyz + running min distance = 4P registers, loaded once).  A round is
//   P x (3 FADD, FMUL, 2 FFMA, FMNMX, compare/select)                      -- no memory traffic
//   two REDUX.SYNC per warp (max of the float bits, then min of the tie key)
//   one 20-byte record per warp written straight into EVERY cluster CTA's shared memory (DSMEM)
//   ONE barrier (bar.sync for CS=1, barrier.cluster for CS>1), double-buffered records
//   every warp re-reduces the 32*CS records itself (no second barrier, no broadcast step); the
//   record carries the winner's coordinates, so the next round starts without a global load.
//
// Exact selection rule of the reference: argmax of the running min distance, ties to
// the smallest (k mod 1024), then to the smallest k.  Thread `tid` of every CTA owns only points
// with k mod 1024 == tid, in ascending k, with a strict '>' scan -- the reference thread's own rule --
// and the cross-thread rule is carried by the key (tid << 21 | k >> 10), minimised among maxima.
#include <cooperative_groups.h>
#include "conv_common.cuh"
#include "../../include/sph3d_b200.h"

namespace cg = cooperative_groups;

namespace sph3d {

constexpr int FPS_THREADS = 1024;

template <int CS>
struct FpsSlots {
    int bits[2][CS * 32];
    int key[2][CS * 32];
    float x[2][CS * 32], y[2][CS * 32], z[2][CS * 32];
};

// reduce (bits,key) records: max bits, then min key.  Returns winner key; bits via reference.
__device__ __forceinline__ void warp_argmax(int bits, int key, int& wbits, int& wkey)
{
    wbits = __reduce_max_sync(FULL_MASK, bits);
    wkey = __reduce_min_sync(FULL_MASK, bits == wbits ? key : 0x7fffffff);
}

// ---------------------------------------------------------------------------------------------------
// Single-CTA kernel (clouds of up to 1024*P points, P <= 10): the fastest path, because one round then
// costs a bar.sync instead of a cluster barrier + DSMEM traffic.  Thread t owns points t, t+1024, ...
// in registers (xyz + running min distance); the whole cloud additionally sits in shared memory so the
// coordinates of the round's winner are three broadcast LDS away (no global load on the critical path,
// and no per-thread "which of my points won" select chain).  Distance updates use Blackwell's packed
// fp32x2 pipe (FADD2 / FMUL2 / FFMA2) on pairs of points -- two independent IEEE round-to-nearest
// operations per instruction, so results are bit-identical to the scalar reference order (Q6/Q12).
template <int P>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_single_kernel(int B, int N, int npoint, const float* __restrict__ xyz, int* __restrict__ out)
{
    extern __shared__ __align__(16) float fps_smem[];
    float* sxyz = fps_smem;                                   // [N*3]
    int* sbits = reinterpret_cast<int*>(fps_smem + ((N * 3 + 3) / 4) * 4);   // [2][32]
    int* skey = sbits + 64;                                   // [2][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const float* pts = xyz + (size_t)cloud * N * 3;
    for (int i = tid; i < N * 3; i += FPS_THREADS) sxyz[i] = __ldg(pts + i);
    __syncthreads();

    constexpr int P2 = (P + 1) / 2;                           // point pairs
    float2 px[P2], py[P2], pz[P2], td[P2];
#pragma unroll
    for (int h = 0; h < P2; h++) {
        float c[2][4];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            int k = (2 * h + u) * FPS_THREADS + tid;
            bool ok = (2 * h + u < P) && (k < N);
            c[u][0] = ok ? sxyz[3 * k] : 0.f; c[u][1] = ok ? sxyz[3 * k + 1] : 0.f; c[u][2] = ok ? sxyz[3 * k + 2] : 0.f;
            c[u][3] = ok ? 1e38f : -1.f;                      // -1 never beats best = -1 (strict >)
        }
        px[h] = make_float2(c[0][0], c[1][0]); py[h] = make_float2(c[0][1], c[1][1]);
        pz[h] = make_float2(c[0][2], c[1][2]); td[h] = make_float2(c[0][3], c[1][3]);
    }
    float x1 = sxyz[0], y1 = sxyz[1], z1 = sxyz[2];
    if (tid == 0) out[(size_t)cloud * npoint] = 0;

    for (int j = 1; j < npoint; j++) {
        const int buf = (j & 1) * 32;
        const float2 nx = make_float2(-x1, -x1), ny = make_float2(-y1, -y1), nz = make_float2(-z1, -z1);
        float best = -1.f; int bp = 0;
#pragma unroll
        for (int h = 0; h < P2; h++) {
            const float2 dx = __fadd2_rn(px[h], nx), dy = __fadd2_rn(py[h], ny), dz = __fadd2_rn(pz[h], nz);
            float2 d = __fmul2_rn(dy, dy);                    // y product rounded alone (Q6) ...
            d = __ffma2_rn(dx, dx, d);                        // ... x and z products fused
            d = __ffma2_rn(dz, dz, d);
            const float a = fminf(d.x, td[h].x), b = fminf(d.y, td[h].y);
            td[h] = make_float2(a, b);
            if (a > best) { best = a; bp = 2 * h; }           // ascending k inside the thread, strict '>'
            if (2 * h + 1 < P) { if (b > best) { best = b; bp = 2 * h + 1; } }
        }
        const int bits = __float_as_int(best);
        const int key = (tid << 21) | bp;                     // (k mod 1024, k / 1024): the reference's tie order
        int wbits, wkey;
        warp_argmax(bits, key, wbits, wkey);
        if (lane == 0) { sbits[buf + warp] = wbits; skey[buf + warp] = wkey; }
        __syncthreads();
        int gb, gk;
        warp_argmax(sbits[buf + lane], skey[buf + lane], gb, gk);
        const int k = ((gk & 0x1fffff) << 10) | (gk >> 21);
        x1 = sxyz[3 * k]; y1 = sxyz[3 * k + 1]; z1 = sxyz[3 * k + 2];
        if (tid == 0) out[(size_t)cloud * npoint + j] = k;
    }
}

// ---------------------------------------------------------------------------------------------------
// Cluster kernel (clouds of 10 241 .. 81 920 points): CS CTAs of one thread-block cluster share a cloud,
// thread t of CTA `rank` owns points (p*CS + rank)*1024 + t in registers; the CTA's points also sit in its
// shared memory.  A round is two-level: (1) 32 warp records -> bar.sync -> warp 0 picks the CTA winner and
// reads its coordinates from shared memory; (2) warp 0 pushes ONE 20-byte record per CTA into every
// cluster CTA's shared memory (DSMEM), one cluster barrier, everybody reduces the CS records.
// (Pushing all 32 warp records of every CTA through DSMEM instead costs 2.4 us/round; measured.)
//
// Round handshake (HS = true, default).  A cluster barrier costs ~2 000 cycles per round with the DSMEM stores in front
// of it -- most of the 1.41 us round at N = 65 536 -- and release / acquire at cluster scope compiles to MEMBAR.ALL.GPU +
// CCTL.IVALL.  Instead a record travels as TWO 16-byte vectors that each carry the round number:
//      v0 = {distance bits, tie key, x, round}      v1 = {y, z, round, 0}
// written with one st.shared::cluster.v4 each (a 16-byte aligned vector store reaches the peer's shared memory as one
// transaction) and polled by the receiver with one volatile ld.shared.v4 each until both show the round number: no
// fences, no barrier, and the polling lane already holds the record when the spin ends.  Two buffers suffice: a CTA
// can write round j+2 into a peer only after it has read that peer's round j+1 record, which the peer sent after all
// of its warps had finished reading round j.  The spin is bounded (a peer that never answers ends the kernel with
// garbage instead of hanging the GPU; the parity tests would see it).  HS = false keeps the cluster barrier
// (SPH3D_FPS_HANDSHAKE=0).
template <int CS>
struct ClusterSlots {
    int lbits[2][32], lkey[2][32];                      // local warp records
    int cbits[2][CS], ckey[2][CS];                      // barrier form: one record per cluster CTA (written remotely)
    float cx[2][CS], cy[2][CS], cz[2][CS];
    alignas(16) int4 v0[2][CS], v1[2][CS];              // handshake form
};

__device__ __forceinline__ void st_peer_v4(const void* local_slot, int peer, int a, int b, int c, int d)
{
    const unsigned la = (unsigned)__cvta_generic_to_shared(local_slot);
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(peer));
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(ra), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int4 ld_volatile_v4(const void* local_slot)
{
    const unsigned la = (unsigned)__cvta_generic_to_shared(local_slot);
    int4 v;
    asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(la) : "memory");
    return v;
}

template <int CS, int P, bool HS>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_cluster_kernel(int B, int N, int npoint, const float* __restrict__ xyz, int* __restrict__ out, int poll_all)
{
    extern __shared__ __align__(16) float fps_smem[];
    float* lxyz = fps_smem;                             // [P*1024*3] this CTA's points, local index p*1024 + t
    __shared__ ClusterSlots<CS> slots;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / CS;
    const float* pts = xyz + (size_t)cloud * N * 3;

    constexpr int P2 = (P + 1) / 2;
    float2 px[P2], py[P2], pz[P2], td[P2];
#pragma unroll
    for (int h = 0; h < P2; h++) {
        float c[2][4];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int p = 2 * h + u;
            const int k = (p * CS + rank) * FPS_THREADS + tid;
            const bool ok = (p < P) && (k < N);
            c[u][0] = ok ? __ldg(pts + 3 * k) : 0.f; c[u][1] = ok ? __ldg(pts + 3 * k + 1) : 0.f;
            c[u][2] = ok ? __ldg(pts + 3 * k + 2) : 0.f; c[u][3] = ok ? 1e38f : -1.f;
            if (p < P) {
                const int li = p * FPS_THREADS + tid;
                lxyz[3 * li] = c[u][0]; lxyz[3 * li + 1] = c[u][1]; lxyz[3 * li + 2] = c[u][2];
            }
        }
        px[h] = make_float2(c[0][0], c[1][0]); py[h] = make_float2(c[0][1], c[1][1]);
        pz[h] = make_float2(c[0][2], c[1][2]); td[h] = make_float2(c[0][3], c[1][3]);
    }
    float x1 = __ldg(pts), y1 = __ldg(pts + 1), z1 = __ldg(pts + 2);
    if (rank == 0 && tid == 0) out[(size_t)cloud * npoint] = 0;
    if (tid < 2 * CS) { (&slots.v0[0][0])[tid] = make_int4(0, 0, 0, 0); (&slots.v1[0][0])[tid] = make_int4(0, 0, 0, 0); }
    cluster.sync();                                     // peers' shared memory exists (and is initialised) before anyone writes to it

    for (int j = 1; j < npoint; j++) {
        const int buf = j & 1;
        const float2 nx = make_float2(-x1, -x1), ny = make_float2(-y1, -y1), nz = make_float2(-z1, -z1);
        float best = -1.f; int bp = 0;
#pragma unroll
        for (int h = 0; h < P2; h++) {
            const float2 dx = __fadd2_rn(px[h], nx), dy = __fadd2_rn(py[h], ny), dz = __fadd2_rn(pz[h], nz);
            float2 d = __fmul2_rn(dy, dy);
            d = __ffma2_rn(dx, dx, d);
            d = __ffma2_rn(dz, dz, d);
            const float a = fminf(d.x, td[h].x), b = fminf(d.y, td[h].y);
            td[h] = make_float2(a, b);
            if (a > best) { best = a; bp = 2 * h; }
            if (2 * h + 1 < P) { if (b > best) { best = b; bp = 2 * h + 1; } }
        }
        int wbits, wkey;
        warp_argmax(__float_as_int(best), (tid << 21) | (bp * CS + rank), wbits, wkey);
        if (lane == 0) { slots.lbits[buf][warp] = wbits; slots.lkey[buf][warp] = wkey; }
        __syncthreads();
        if (warp == 0) {
            int gb, gk;
            warp_argmax(slots.lbits[buf][lane], slots.lkey[buf][lane], gb, gk);
            const int kk = ((gk & 0x1fffff) << 10) | (gk >> 21);            // global point id of the CTA winner
            const int li = ((kk >> 10) / CS) * FPS_THREADS + (kk & (FPS_THREADS - 1));
            const float wx = lxyz[3 * li], wy = lxyz[3 * li + 1], wz = lxyz[3 * li + 2];
            if (lane < CS) {
                if constexpr (HS) {                     // my record into slot [buf][rank] of peer `lane` (my own CTA included)
                    st_peer_v4(&slots.v0[buf][rank], lane, gb, gk, __float_as_int(wx), j);
                    st_peer_v4(&slots.v1[buf][rank], lane, __float_as_int(wy), __float_as_int(wz), j, 0);
                } else {
                    ClusterSlots<CS>* rs = cluster.map_shared_rank(&slots, lane);
                    rs->cbits[buf][rank] = gb; rs->ckey[buf][rank] = gk;
                    rs->cx[buf][rank] = wx; rs->cy[buf][rank] = wy; rs->cz[buf][rank] = wz;
                }
            }
        }
        if constexpr (HS) {
            int4 r0 = make_int4((int)0x80000000, 0x7fffffff, 0, 0), r1 = make_int4(0, 0, 0, 0);
            // POLL_ALL: every warp spins on the slots itself (no second barrier, but 31 spinning warps compete for issue
            // slots with the warps still updating distances).  Otherwise warp 0 alone spins and releases the others
            // through a second bar.sync.
            if (poll_all || warp == 0) {
                int spins = 0;
                bool ready;
                do {                                    // lanes < CS each watch one slot of this round's buffer
                    ready = true;
                    if (lane < CS) {
                        r0 = ld_volatile_v4(&slots.v0[buf][lane]);
                        r1 = ld_volatile_v4(&slots.v1[buf][lane]);
                        ready = (r0.w == j) && (r1.z == j);
                    }
                } while (!__all_sync(FULL_MASK, ready) && ++spins < (1 << 22));
            }
            if (!poll_all) {
                __syncthreads();
                if (warp != 0 && lane < CS) { r0 = ld_volatile_v4(&slots.v0[buf][lane]); r1 = ld_volatile_v4(&slots.v1[buf][lane]); }
            }
            int fb, fk;
            warp_argmax(lane < CS ? r0.x : (int)0x80000000, lane < CS ? r0.y : 0x7fffffff, fb, fk);
            const int wl = __ffs(__ballot_sync(FULL_MASK, lane < CS && r0.x == fb && r0.y == fk)) - 1;   // lowest rank among equals
            x1 = __int_as_float(__shfl_sync(FULL_MASK, r0.z, wl));
            y1 = __int_as_float(__shfl_sync(FULL_MASK, r1.x, wl));
            z1 = __int_as_float(__shfl_sync(FULL_MASK, r1.y, wl));
            if (rank == 0 && tid == 0) out[(size_t)cloud * npoint + j] = ((fk & 0x1fffff) << 10) | (fk >> 21);
            continue;
        }
        cluster.sync();
        int fb, fk;
        warp_argmax(lane < CS ? slots.cbits[buf][lane] : (int)0x80000000, lane < CS ? slots.ckey[buf][lane] : 0x7fffffff, fb, fk);
        int wl = 0;
#pragma unroll
        for (int c = 1; c < CS; c++) if (slots.cbits[buf][c] == fb && slots.ckey[buf][c] == fk) wl = c;
        if (slots.cbits[buf][0] == fb && slots.ckey[buf][0] == fk) wl = 0;
        x1 = slots.cx[buf][wl]; y1 = slots.cy[buf][wl]; z1 = slots.cz[buf][wl];
        if (rank == 0 && tid == 0) out[(size_t)cloud * npoint + j] = ((fk & 0x1fffff) << 10) | (fk >> 21);
    }
    cluster.sync();                                     // no CTA may exit while peers can still write its smem
}

// Any-N fallback: one CTA per cloud, running min distances in a global workspace.
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_global_kernel(int B, int N, int npoint, const float* __restrict__ xyz, float* __restrict__ temp,
                  int* __restrict__ out)
{
    __shared__ FpsSlots<1> slots;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int cloud = blockIdx.x; cloud < B; cloud += gridDim.x) {
        const float* pts = xyz + (size_t)cloud * N * 3;
        float* td = temp + (size_t)cloud * N;
        for (int k = tid; k < N; k += FPS_THREADS) td[k] = 1e38f;
        float x1 = __ldg(pts), y1 = __ldg(pts + 1), z1 = __ldg(pts + 2);
        if (tid == 0) out[(size_t)cloud * npoint] = 0;
        __syncthreads();
        for (int j = 1; j < npoint; j++) {
            const int buf = j & 1;
            float best = -1.f; int bk = 0; float bx = 0.f, by = 0.f, bz = 0.f;
            for (int k = tid; k < N; k += FPS_THREADS) {
                float x2 = __ldg(pts + 3 * k), y2 = __ldg(pts + 3 * k + 1), z2 = __ldg(pts + 3 * k + 2);
                float d = sqdist_ref(__fsub_rn(x2, x1), __fsub_rn(y2, y1), __fsub_rn(z2, z1));
                float d2 = fminf(d, td[k]);
                td[k] = d2;
                if (d2 > best) { best = d2; bk = k; bx = x2; by = y2; bz = z2; }
            }
            const int bits = __float_as_int(best);
            const int key = (tid << 21) | (bk >> 10);
            int wbits, wkey;
            warp_argmax(bits, key, wbits, wkey);
            const int src = __ffs(__ballot_sync(FULL_MASK, bits == wbits && key == wkey)) - 1;
            float wx = __shfl_sync(FULL_MASK, bx, src), wy = __shfl_sync(FULL_MASK, by, src), wz = __shfl_sync(FULL_MASK, bz, src);
            if (lane == 0) {
                slots.bits[buf][warp] = wbits; slots.key[buf][warp] = wkey;
                slots.x[buf][warp] = wx; slots.y[buf][warp] = wy; slots.z[buf][warp] = wz;
            }
            __syncthreads();
            int gb, gk;
            const int mb = slots.bits[buf][lane], mk = slots.key[buf][lane];
            warp_argmax(mb, mk, gb, gk);
            const int wl = __ffs(__ballot_sync(FULL_MASK, mb == gb && mk == gk)) - 1;
            x1 = slots.x[buf][wl]; y1 = slots.y[buf][wl]; z1 = slots.z[buf][wl];
            if (tid == 0) out[(size_t)cloud * npoint + j] = ((gk & 0x1fffff) << 10) | (gk >> 21);
        }
        __syncthreads();
    }
}

struct FpsPlan { int cs, p; };

static FpsPlan plan_fps(int n)
{
    // clouds of up to 10 240 points: one CTA (SPH3D_FPS_CLUSTER_MIN_N lowers the switch-over point for sweeps)
    const int single_max = tunables().fps_cluster_min_n > 0 ? tunables().fps_cluster_min_n - 1 : 10 * FPS_THREADS;
    if (n <= single_max && n <= 10 * FPS_THREADS) return FpsPlan{1, (n + FPS_THREADS - 1) / FPS_THREADS};   // single-CTA kernel
    const int cs_opts[4] = {1, 2, 4, 8};
    for (int i = (n <= 10 * FPS_THREADS ? 1 : 0); i < 4; i++) {
        int cs = cs_opts[i];
        int p = (n + cs * FPS_THREADS - 1) / (cs * FPS_THREADS);
        if (p <= 8) return FpsPlan{cs, p};
    }
    int p = (n + 8 * FPS_THREADS - 1) / (8 * FPS_THREADS);
    if (p <= 10) return FpsPlan{8, p};
    return FpsPlan{0, 0};
}

template <int CS, int P, bool HS>
static cudaError_t launch_fps_hs(int B, int N, int npoint, const float* xyz, int* out, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * CS);
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = (size_t)P * FPS_THREADS * 3 * sizeof(float);
    if (cfg.dynamicSmemBytes + sizeof(ClusterSlots<CS>) > 48 * 1024) {      // static slots count against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute(fps_cluster_kernel<CS, P, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)cfg.dynamicSmemBytes);
        if (e != cudaSuccess) return e;
    }
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fps_cluster_kernel<CS, P, HS>, B, N, npoint, xyz, out, tunables().fps_handshake == 2 ? 1 : 0);
}

template <int CS, int P>
static cudaError_t launch_fps(int B, int N, int npoint, const float* xyz, int* out, cudaStream_t st)
{
    if (tunables().fps_handshake == 0) return launch_fps_hs<CS, P, false>(B, N, npoint, xyz, out, st);
    return launch_fps_hs<CS, P, true>(B, N, npoint, xyz, out, st);
}

template <int P>
static cudaError_t launch_fps_single(int B, int N, int npoint, const float* xyz, int* out, cudaStream_t st)
{
    const size_t smem = (size_t)((N * 3 + 3) / 4) * 4 * sizeof(float) + 128 * sizeof(int);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fps_single_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    fps_single_kernel<P><<<B, FPS_THREADS, smem, st>>>(B, N, npoint, xyz, out);
    return cudaPeekAtLastError();
}

static cudaError_t launch_fps_single_p(int p, int B, int N, int npoint, const float* xyz, int* out, cudaStream_t st)
{
    switch (p) {
        case 1: return launch_fps_single<1>(B, N, npoint, xyz, out, st);
        case 2: return launch_fps_single<2>(B, N, npoint, xyz, out, st);
        case 3: case 4: return launch_fps_single<4>(B, N, npoint, xyz, out, st);
        case 5: case 6: return launch_fps_single<6>(B, N, npoint, xyz, out, st);
        case 7: case 8: return launch_fps_single<8>(B, N, npoint, xyz, out, st);
        default: return launch_fps_single<10>(B, N, npoint, xyz, out, st);
    }
}

template <int CS>
static cudaError_t launch_fps_p(int p, int B, int N, int npoint, const float* xyz, int* out, cudaStream_t st)
{
    switch (p) {
        case 1: return launch_fps<CS, 1>(B, N, npoint, xyz, out, st);
        case 2: return launch_fps<CS, 2>(B, N, npoint, xyz, out, st);
        case 3: return launch_fps<CS, 3>(B, N, npoint, xyz, out, st);
        case 4: return launch_fps<CS, 4>(B, N, npoint, xyz, out, st);
        case 5: return launch_fps<CS, 5>(B, N, npoint, xyz, out, st);
        case 6: return launch_fps<CS, 6>(B, N, npoint, xyz, out, st);
        case 7: return launch_fps<CS, 7>(B, N, npoint, xyz, out, st);
        case 8: return launch_fps<CS, 8>(B, N, npoint, xyz, out, st);
        default: break;
    }
    if (CS == 8 && p == 9) return launch_fps<8, 9>(B, N, npoint, xyz, out, st);
    if (CS == 8 && p == 10) return launch_fps<8, 10>(B, N, npoint, xyz, out, st);
    return cudaErrorInvalidValue;
}

}  // namespace sph3d

using namespace sph3d;

extern "C" size_t sph3d_farthest_point_sample_workspace_bytes(int b, int n, int m)
{
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    FpsPlan p = plan_fps(n);
    return p.cs ? 0 : sizeof(float) * (size_t)b * n;
}

extern "C" int sph3d_farthest_point_sample(int b, int n, int m, const float* inp, void* temp, size_t temp_bytes,
                                           int* out, void* stream)
{
    g_last_launch_count = 0;
    if (b <= 0 || n <= 0 || m <= 0 || !inp || !out) return (int)cudaErrorInvalidValue;   // tf_sample.cpp:34 npoint>0
    if (n > (1 << 30)) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    FpsPlan p = plan_fps(n);
    cudaError_t e;
    if (p.cs == 0) {
        if (!temp || temp_bytes < sizeof(float) * (size_t)b * n) return (int)cudaErrorInvalidValue;
        int grid = b < sm_count() ? b : sm_count();
        fps_global_kernel<<<grid, FPS_THREADS, 0, st>>>(b, n, m, inp, (float*)temp, out);
        e = cudaPeekAtLastError();
    } else if (p.cs == 1) e = launch_fps_single_p(p.p, b, n, m, inp, out, st);
    else if (p.cs == 2) e = launch_fps_p<2>(p.p, b, n, m, inp, out, st);
    else if (p.cs == 4) e = launch_fps_p<4>(p.p, b, n, m, inp, out, st);
    else e = launch_fps_p<8>(p.p, b, n, m, inp, out, st);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    g_last_launch_count = 1;
    return 0;
}
