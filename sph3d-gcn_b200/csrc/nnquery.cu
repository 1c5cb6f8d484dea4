// nnquery.cu -- range (ball) and cube neighbour queries for sm_100a.
//
// Replaces buildSphereNeighborLauncher / buildCubeNeighborLauncher
// (/root/reference/tf_ops/nnquery/tf_nnquery_gpu.cu:115-127) and the cudaMemset zero fill of
// tf_nnquery.cpp:100-102,155-156.  Semantics are reproduced
// bit for bit; the implementation is not the reference's one-thread-per-query serial scan:
//
//  * one WARP scans the database for QPW queries at once: lane l tests point base+l against the
//    QPW query points held in (uniform) registers, a ballot turns the hits into an ordered
//    compaction, so neighbours come out in ascending database index (Q4) with coalesced writes
//    and the scan stops as soon as every query of the warp has K hits;
//  * the reference predicate  d=sqrtf(d2); d<r && (double)fabsf(d-r)>1e-6  is monotone in d2, so
//    it is folded into ONE exact float threshold per query (found by bisection over the float
//    bit pattern with the literal predicate): the inner loop is 3 FADD + FMUL + 2 FFMA + FSETP
//    per test, no sqrt, no fp64;
//  * the growing radius (Q1): the reference thread (blockIdx=i%32, threadIdx=j%1024) carries its
//    radius from query to query, +0.05 per pass.  Phase 1 gives every query the radius it has
//    when no earlier query of its chain needed a retry (closed form in the chain step t).
//    Phase 2 (one warp per chain, exits at once when the chain has no empty query) replays only
//    the chains in which some query found nothing, with the exact carried radius.
#include "common.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

__device__ __forceinline__ bool in_range_ref(float d2, float radius)
{
    float d = __fsqrt_rn(d2);
    return (d < radius) && ((double)fabsf(__fsub_rn(d, radius)) > 1e-6);
}

// largest float t for which in_range_ref(t, radius) holds, -1 if there is none
__device__ __noinline__ float range_threshold(float radius)
{
    if (!in_range_ref(0.0f, radius)) return -1.0f;
    unsigned lo = 0u, hi = 0x7f800000u;     // predicate true at lo, false at hi (+inf)
    while (hi - lo > 1u) {
        unsigned mid = lo + ((hi - lo) >> 1);
        if (in_range_ref(__uint_as_float(mid), radius)) lo = mid; else hi = mid;
    }
    return __uint_as_float(lo);
}

__device__ __forceinline__ float next_radius(float r)      // radius += 0.05  (double literal)
{
    return __double2float_rn(__dadd_rn((double)r, 0.05));
}

template <int QPW>
__global__ void __launch_bounds__(256)
sphere_query_kernel(int B, int N, int M, int K, float radius0,
                    const float* __restrict__ database, const float* __restrict__ query,
                    int* __restrict__ nn_index, int* __restrict__ nn_count,
                    float* __restrict__ nn_dist)
{
    const int lane = threadIdx.x & 31;
    const int groups = (M + QPW - 1) / QPW;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= (long long)B * groups) return;
    const int b = (int)(gw / groups);
    const int j0 = (int)(gw % groups) * QPW;

    // per-query constants; lane q prepares query q, then broadcast
    float myT = -1.0f;
    if (lane < QPW && j0 + lane < M) {
        int j = j0 + lane;
        int tx = j % REF_BLOCK;
        int per_i = (M - tx + REF_BLOCK - 1) / REF_BLOCK;       // queries of this chain per cloud
        int t = (b / REF_GRID) * per_i + j / REF_BLOCK;          // chain step of query (b,j)
        float r = radius0;
        for (int s = 0; s < t; s++) r = next_radius(r);
        myT = range_threshold(r);
    }
    float qx[QPW], qy[QPW], qz[QPW], thr[QPW];
    int cnt[QPW];
#pragma unroll
    for (int q = 0; q < QPW; q++) {
        int j = min(j0 + q, M - 1);
        const float* qp = query + ((size_t)b * M + j) * 3;
        qx[q] = __ldg(qp); qy[q] = __ldg(qp + 1); qz[q] = __ldg(qp + 2);
        thr[q] = __shfl_sync(FULL_MASK, myT, q);
        cnt[q] = (j0 + q < M) ? 0 : K;                           // padding queries are "done"
    }

    const float* db = database + (size_t)b * N * 3;
    const unsigned lt = (1u << lane) - 1u;
    const float qnan = __int_as_float(0x7fc00000);
    for (int base = 0; base < N; base += 32) {
        int k = base + lane;
        float px = qnan, py = qnan, pz = qnan;
        if (k < N) { px = __ldg(db + 3 * k); py = __ldg(db + 3 * k + 1); pz = __ldg(db + 3 * k + 2); }
        bool all_done = true;
#pragma unroll
        for (int q = 0; q < QPW; q++) {
            if (cnt[q] < K) {                                    // warp-uniform
                float d2 = sqdist_ref(__fsub_rn(px, qx[q]), __fsub_rn(py, qy[q]), __fsub_rn(pz, qz[q]));
                bool hit = d2 <= thr[q];
                unsigned m = __ballot_sync(FULL_MASK, hit);
                if (m) {
                    int slot = cnt[q] + __popc(m & lt);
                    if (hit && slot < K) {
                        size_t o = ((size_t)b * M + (j0 + q)) * K + slot;
                        nn_index[o] = k;
                        nn_dist[o] = __fsqrt_rn(__fsqrt_rn(d2));        // Q2: sqrt of the distance
                    }
                    cnt[q] += __popc(m);
                }
                all_done &= (cnt[q] >= K);
            }
        }
        if (all_done) break;
    }
#pragma unroll
    for (int q = 0; q < QPW; q++) {
        if (j0 + q < M) {
            int c = min(cnt[q], K);
            size_t row = ((size_t)b * M + (j0 + q)) * K;
            for (int s = c + lane; s < K; s += 32) { nn_index[row + s] = 0; nn_dist[row + s] = 0.0f; }
            if (lane == 0) nn_count[(size_t)b * M + j0 + q] = c;   // 0 == "found nothing": phase 2
        }
    }
}

// warp-cooperative scan of one query; returns the number of hits (saturating at >= K)
__device__ __forceinline__ int scan_one(const float* __restrict__ db, int N, int K, float qx, float qy,
                                        float qz, float thr, int* __restrict__ oi, float* __restrict__ od)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    for (int base = 0; base < N && cnt < K; base += 32) {
        int k = base + lane;
        bool hit = false; float d2 = 0.f;
        if (k < N) {
            d2 = sqdist_ref(__fsub_rn(__ldg(db + 3 * k), qx), __fsub_rn(__ldg(db + 3 * k + 1), qy),
                            __fsub_rn(__ldg(db + 3 * k + 2), qz));
            hit = d2 <= thr;
        }
        unsigned m = __ballot_sync(FULL_MASK, hit);
        int slot = cnt + __popc(m & lt);
        if (hit && slot < K) { oi[slot] = k; od[slot] = __fsqrt_rn(__fsqrt_rn(d2)); }
        cnt += __popc(m);
    }
    return cnt;
}

// Phase 2: replay the reference chains that contain an empty query.  One warp per chain.
__global__ void __launch_bounds__(128)
sphere_fixup_kernel(int B, int N, int M, int K, float radius0,
                    const float* __restrict__ database, const float* __restrict__ query,
                    int* __restrict__ nn_index, int* __restrict__ nn_count, float* __restrict__ nn_dist)
{
    const int lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int bx = chain / REF_BLOCK, tx = chain % REF_BLOCK;
    if (bx >= REF_GRID || bx >= B || tx >= M) return;
    const int per_i = (M - tx + REF_BLOCK - 1) / REF_BLOCK;
    const int n_i = (B - bx + REF_GRID - 1) / REF_GRID;
    const int steps = per_i * n_i;

    bool any = false;
    for (int s = lane; s < steps; s += 32) {
        int i = bx + (s / per_i) * REF_GRID, j = tx + (s % per_i) * REF_BLOCK;
        any |= (nn_count[(size_t)i * M + j] == 0);
    }
    if (!__any_sync(FULL_MASK, any)) return;

    float r = radius0;
    bool deviated = false;          // true once the carried radius differs from phase 1's
    for (int s = 0; s < steps; s++) {
        int i = bx + (s / per_i) * REF_GRID, j = tx + (s % per_i) * REF_BLOCK;
        int c = nn_count[(size_t)i * M + j];
        if (!deviated && c != 0) { r = next_radius(r); continue; }
        if (!deviated) { r = next_radius(r); deviated = true; }      // phase 1's pass at r found nothing
        const float* qp = query + ((size_t)i * M + j) * 3;
        float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        size_t row = ((size_t)i * M + j) * K;
        int found = 0;
        for (int guard = 0; guard < (1 << 16); guard++) {            // reference: while(s==0)
            float thr = range_threshold(r);
            found = scan_one(database + (size_t)i * N * 3, N, K, qx, qy, qz, thr, nn_index + row, nn_dist + row);
            r = next_radius(r);
            if (found) break;
        }
        int cfin = min(found, K);
        __syncwarp();
        for (int t = cfin + lane; t < K; t += 32) { nn_index[row + t] = 0; nn_dist[row + t] = 0.0f; }
        if (lane == 0) nn_count[(size_t)i * M + j] = cfin;
        __syncwarp();
    }
}

template <int QPW>
__global__ void __launch_bounds__(256)
cube_query_kernel(int B, int N, int M, int grid, int K, float length,
                  const float* __restrict__ database, const float* __restrict__ query,
                  int* __restrict__ nn_index, int* __restrict__ nn_count)
{
    const int lane = threadIdx.x & 31;
    const int groups = (M + QPW - 1) / QPW;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= (long long)B * groups) return;
    const int b = (int)(gw / groups);
    const int j0 = (int)(gw % groups) * QPW;
    const float half = __fdiv_rn(length, 2.0f);                  // length/2      (float/int)
    const float cell = __fdiv_rn(length, (float)grid);           // length/gridSize
    float qx[QPW], qy[QPW], qz[QPW];
    int cnt[QPW];
#pragma unroll
    for (int q = 0; q < QPW; q++) {
        int j = min(j0 + q, M - 1);
        const float* qp = query + ((size_t)b * M + j) * 3;
        qx[q] = __ldg(qp); qy[q] = __ldg(qp + 1); qz[q] = __ldg(qp + 2);
        cnt[q] = (j0 + q < M) ? 0 : K;
    }
    const float* db = database + (size_t)b * N * 3;
    const unsigned lt = (1u << lane) - 1u;
    const float qnan = __int_as_float(0x7fc00000);
    for (int base = 0; base < N; base += 32) {
        int k = base + lane;
        float px = qnan, py = qnan, pz = qnan;
        if (k < N) { px = __ldg(db + 3 * k); py = __ldg(db + 3 * k + 1); pz = __ldg(db + 3 * k + 2); }
        bool all_done = true;
#pragma unroll
        for (int q = 0; q < QPW; q++) {
            if (cnt[q] < K) {
                float dx = __fsub_rn(px, qx[q]), dy = __fsub_rn(py, qy[q]), dz = __fsub_rn(pz, qz[q]);
                bool hit = fabsf(dx) < half && fabsf(dy) < half && fabsf(dz) < half;
                unsigned m = __ballot_sync(FULL_MASK, hit);
                if (m) {
                    int slot = cnt[q] + __popc(m & lt);
                    if (hit && slot < K) {
                        int xi = (int)__fdiv_rn(__fadd_rn(dx, half), cell);
                        int yi = (int)__fdiv_rn(__fadd_rn(dy, half), cell);
                        int zi = (int)__fdiv_rn(__fadd_rn(dz, half), cell);
                        size_t o = (((size_t)b * M + (j0 + q)) * K + slot) * 2;
                        *reinterpret_cast<int2*>(nn_index + o) = make_int2(k, xi * grid * grid + yi * grid + zi);
                    }
                    cnt[q] = min(K, cnt[q] + __popc(m));         // the reference stops counting at K
                }
                all_done &= (cnt[q] >= K);
            }
        }
        if (all_done) break;
    }
#pragma unroll
    for (int q = 0; q < QPW; q++) {
        if (j0 + q < M) {
            size_t row = ((size_t)b * M + (j0 + q)) * K * 2;
            for (int s = 2 * cnt[q] + lane; s < 2 * K; s += 32) nn_index[row + s] = 0;
            if (lane == 0) nn_count[(size_t)b * M + j0 + q] = cnt[q];
        }
    }
}

}  // namespace sph3d

using namespace sph3d;

extern "C" int sph3d_build_sphere_neighbor(int B, int N, int M, int K, float radius,
                                           const float* database, const float* query,
                                           int* nn_index, int* nn_count, float* nn_dist, void* stream)
{
    g_last_launch_count = 0;
    if (B <= 0 || N <= 0 || M <= 0 || K <= 0 || !(radius > 0.0f) || !database || !query ||
        !nn_index || !nn_count || !nn_dist)
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int QPW = 8;
    long long warps = (long long)B * ((M + QPW - 1) / QPW);
    long long ctas = (warps + 7) / 8;
    if (ctas > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
    sphere_query_kernel<QPW><<<(unsigned)ctas, 256, 0, st>>>(B, N, M, K, radius, database, query,
                                                              nn_index, nn_count, nn_dist);
    SPH3D_CHECK_LAUNCH();
    int chains = REF_GRID * REF_BLOCK;
    sphere_fixup_kernel<<<chains / 4, 128, 0, st>>>(B, N, M, K, radius, database, query,
                                                    nn_index, nn_count, nn_dist);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 2;
    return 0;
}

extern "C" int sph3d_build_cube_neighbor(int B, int N, int M, int grid_size, int K, float length,
                                         const float* database, const float* query,
                                         int* nn_index, int* nn_count, void* stream)
{
    g_last_launch_count = 0;
    if (B <= 0 || N <= 0 || M <= 0 || K <= 0 || grid_size <= 0 || !(length > 0.0f) || !database ||
        !query || !nn_index || !nn_count)
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int QPW = 8;
    long long warps = (long long)B * ((M + QPW - 1) / QPW);
    long long ctas = (warps + 7) / 8;
    if (ctas > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
    cube_query_kernel<QPW><<<(unsigned)ctas, 256, 0, st>>>(B, N, M, grid_size, K, length, database,
                                                            query, nn_index, nn_count);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}
