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
#include <cstdlib>
#include "conv_common.cuh"
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

// ---------------------------------------------------------------------------------------------------
// Uniform cell grid over each database cloud (optional accelerator).  A query whose search
// radius spans at most CELL_REACH cells is answered from the (2R+1)^3 cell stencil instead of scanning the
// whole cloud: the in-range candidates are marked in a per-warp bitmap indexed by point id, and the bitmap
// is then read out in ascending id order -- exactly the reference's "first K in-range points by index" (Q4),
// with the same distance arithmetic and the same threshold, so results stay bit-identical.  Queries with a
// larger (chain-grown) radius keep the brute-force scan, which exits early for them anyway.
constexpr int GRID_MAX = 32;                    // cells per axis
constexpr int GRID_CELLS = GRID_MAX * GRID_MAX * GRID_MAX;
constexpr int CELL_REACH = 2;                   // grid path when the radius spans <= 2 cells (<= 125 cells)
// Measured (profiles/r1_nnquery_grid_vs_scan.txt): the 8-queries-per-warp scan with early exit wins up to
// N = 10^4 (1.19 vs 1.70 ms at B=32, N=10^4, r=0.1) because the growing radius (Q1) lets most rows stop after a
// few hundred points; the grid wins from N ~ 3*10^4 (0.96 vs 1.55 ms at B=4, N=65536, r=0.05).
// below this the scan is faster than building + walking a grid.  Round 2 (profiles/r2_nnquery.json): with the packed
// fp32x2 scan the grid loses at every BASELINE shape, unsaturated config radii included (ModelNet level 1, r = 0.1: 1.62
// vs 1.83 ms; cfg5, N = 65 536: 0.78 vs 0.99 ms), so it is kept for clouds beyond those (the scan grows with N^2).
constexpr int GRID_MIN_N = 98304;

struct GridInfo {
    float ox, oy, oz, inv_h;                    // cell = floor((p - o) * inv_h)
    float h;
    int nx, ny, nz;
};

__device__ __forceinline__ int cell_coord(float p, float o, float inv_h) { return (int)floorf((p - o) * inv_h); }

// stencil half-width for radius r; callers use the grid path iff the result is <= CELL_REACH.
// |p-q| <= r  =>  |cell(p) - cell(q)| <= floor(r/h + eps) + 1  (floor-based cells, eps covers float rounding)
__device__ __forceinline__ int cell_reach(float r, float h) { return (int)floorf(r / h + 1e-3f) + 1; }

__device__ __forceinline__ float chain_radius(int b, int j, int M, float radius0)
{
    const int tx = j % REF_BLOCK;
    const int per_i = (M - tx + REF_BLOCK - 1) / REF_BLOCK;     // queries of this reference chain per cloud
    const int t = (b / REF_GRID) * per_i + j / REF_BLOCK;        // chain step of query (b,j)
    float r = radius0;
    for (int s = 0; s < t; s++) r = next_radius(r);
    return r;
}

// one CTA per cloud: bounding box -> grid geometry; zero the cloud's cell counters
__global__ void __launch_bounds__(256)
grid_setup_kernel(int N, float radius0, const float* __restrict__ database, GridInfo* __restrict__ ginfo,
                  int* __restrict__ cell_count)
{
    __shared__ float smin[3][8], smax[3][8];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* db = database + (size_t)b * N * 3;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int k = threadIdx.x; k < N; k += blockDim.x)
#pragma unroll
        for (int a = 0; a < 3; a++) { float v = __ldg(db + 3 * k + a); mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(FULL_MASK, mn[a], d));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL_MASK, mx[a], d));
        }
        if (lane == 0) { smin[a][warp] = mn[a]; smax[a][warp] = mx[a]; }
    }
    __syncthreads();
    __shared__ GridInfo gi;
    if (threadIdx.x == 0) {
        float lo[3], ext = 0.f;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float l = smin[a][0], u = smax[a][0];
            for (int w = 1; w < 8; w++) { l = fminf(l, smin[a][w]); u = fmaxf(u, smax[a][w]); }
            lo[a] = l; ext = fmaxf(ext, u - l);
            smax[a][0] = u - l;
        }
        // cell a little larger than the base radius (so a base-radius query reaches exactly one ring of
        // cells), but never more than GRID_MAX cells per axis
        float h = fmaxf(radius0 * 1.001f, ext / (float)GRID_MAX * 1.0001f);
        if (!(h > 0.f) || !isfinite(h)) h = 1.f;
        gi.ox = lo[0]; gi.oy = lo[1]; gi.oz = lo[2]; gi.h = h; gi.inv_h = 1.0f / h;
        gi.nx = min(GRID_MAX, (int)floorf(smax[0][0] * gi.inv_h) + 1);
        gi.ny = min(GRID_MAX, (int)floorf(smax[1][0] * gi.inv_h) + 1);
        gi.nz = min(GRID_MAX, (int)floorf(smax[2][0] * gi.inv_h) + 1);
        ginfo[b] = gi;
    }
    __syncthreads();
    const int ncell = gi.nx * gi.ny * gi.nz;
    int* cc = cell_count + (size_t)b * (GRID_CELLS + 1);
    for (int c = threadIdx.x; c <= ncell; c += blockDim.x) cc[c] = 0;
}

__device__ __forceinline__ int point_cell(const GridInfo& g, float x, float y, float z)
{
    const int cx = min(g.nx - 1, max(0, cell_coord(x, g.ox, g.inv_h)));
    const int cy = min(g.ny - 1, max(0, cell_coord(y, g.oy, g.inv_h)));
    const int cz = min(g.nz - 1, max(0, cell_coord(z, g.oz, g.inv_h)));
    return (cz * g.ny + cy) * g.nx + cx;
}

// pass 0: count points per cell; pass 1 (after the scan): scatter point ids into cell order
__global__ void __launch_bounds__(256)
grid_bin_kernel(int B, int N, int pass, const float* __restrict__ database, const GridInfo* __restrict__ ginfo,
                int* __restrict__ cell_count, int* __restrict__ cell_cursor, int* __restrict__ cell_pts)
{
    const size_t total = (size_t)B * N;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / N), k = (int)(t - (size_t)b * N);
        const GridInfo g = ginfo[b];
        const float* p = database + t * 3;
        const int c = point_cell(g, __ldg(p), __ldg(p + 1), __ldg(p + 2));
        if (pass == 0) atomicAdd(cell_count + (size_t)b * (GRID_CELLS + 1) + c, 1);
        else cell_pts[(size_t)b * N + atomicAdd(cell_cursor + (size_t)b * (GRID_CELLS + 1) + c, 1)] = k;
    }
}

// one CTA per cloud: exclusive scan of the cell counters -> cell_start (in place) and a cursor copy
__global__ void __launch_bounds__(1024)
grid_scan_kernel(const GridInfo* __restrict__ ginfo, int* __restrict__ cell_count, int* __restrict__ cell_cursor)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ncell = ginfo[b].nx * ginfo[b].ny * ginfo[b].nz;
    int* cc = cell_count + (size_t)b * (GRID_CELLS + 1);
    int* cur = cell_cursor + (size_t)b * (GRID_CELLS + 1);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += 1024) {
        const int c = base + threadIdx.x;
        const int v = c < ncell ? cc[c] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int s = wsum[lane], si = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(FULL_MASK, si, d); if (lane >= d) si += y; }
            wsum[lane] = si - s;                                   // exclusive prefix of the warp sums
        }
        __syncthreads();
        const int excl = carry + wsum[warp] + incl - v;
        if (c < ncell) { cc[c] = excl; cur[c] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cc[ncell] = carry;
}

// one warp per grid-eligible query
__global__ void __launch_bounds__(256)
sphere_query_grid_kernel(int B, int N, int M, int K, float radius0,
                         const float* __restrict__ database, const float* __restrict__ query,
                         const GridInfo* __restrict__ ginfo, const int* __restrict__ cell_start,
                         const int* __restrict__ cell_pts,
                         int* __restrict__ nn_index, int* __restrict__ nn_count, float* __restrict__ nn_dist)
{
    extern __shared__ unsigned bitmaps[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int words = (N + 31) / 32;
    unsigned* bm = bitmaps + (size_t)warp * words;
    for (int w = lane; w < words; w += 32) bm[w] = 0u;
    __syncwarp();
    const long long total = (long long)B * M;
    const unsigned lt = (1u << lane) - 1u;
    for (long long qid = (long long)blockIdx.x * (blockDim.x >> 5) + warp; qid < total;
         qid += (long long)gridDim.x * (blockDim.x >> 5)) {
        const int b = (int)(qid / M), j = (int)(qid - (long long)b * M);
        const GridInfo g = ginfo[b];
        const float r = chain_radius(b, j, M, radius0);
        const int R = cell_reach(r, g.h);
        if (R > CELL_REACH) continue;                              // brute-force kernel owns this query
        const float thr = range_threshold(r);
        const float* qp = query + (size_t)qid * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        const float* db = database + (size_t)b * N * 3;
        const int* cs = cell_start + (size_t)b * (GRID_CELLS + 1);
        const int* cp = cell_pts + (size_t)b * N;
        const int cx = cell_coord(qx, g.ox, g.inv_h), cy = cell_coord(qy, g.oy, g.inv_h), cz = cell_coord(qz, g.oz, g.inv_h);
        // points are stored in CLAMPED cells, and clamping is monotone: an in-range point with (unclamped)
        // cell in [c-R, c+R] is stored in [clamp(c-R), clamp(c+R)] -- also for queries outside the box
        const int x0 = min(g.nx - 1, max(0, cx - R)), x1 = min(g.nx - 1, max(0, cx + R));
        const int y0 = min(g.ny - 1, max(0, cy - R)), y1 = min(g.ny - 1, max(0, cy + R));
        const int z0 = min(g.nz - 1, max(0, cz - R)), z1 = min(g.nz - 1, max(0, cz + R));
        {
            for (int z = z0; z <= z1; z++)
                for (int y = y0; y <= y1; y++) {
                    const int rowc = (z * g.ny + y) * g.nx;
                    const int beg = __ldg(cs + rowc + x0), end = __ldg(cs + rowc + x1 + 1);   // x-run is contiguous
                    for (int p = beg + lane; p < end; p += 32) {
                        const int id = __ldg(cp + p);
                        const float d2 = sqdist_ref(__fsub_rn(__ldg(db + 3 * id), qx), __fsub_rn(__ldg(db + 3 * id + 1), qy),
                                                    __fsub_rn(__ldg(db + 3 * id + 2), qz));
                        if (d2 <= thr) atomicOr(bm + (id >> 5), 1u << (id & 31));
                    }
                }
        }
        __syncwarp();
        // read the bitmap out in ascending id order; clear it on the way
        int cnt = 0;
        const size_t row = (size_t)qid * K;
        for (int w0 = 0; w0 < words; w0 += 32) {
            const int w = w0 + lane;
            unsigned word = 0u;
            if (w < words) { word = bm[w]; if (word) bm[w] = 0u; }
            unsigned nz = __ballot_sync(FULL_MASK, word != 0u);
            while (nz) {
                const int l = __ffs(nz) - 1;
                nz &= nz - 1;
                const unsigned wv = __shfl_sync(FULL_MASK, word, l);
                if (cnt < K) {
                    const int nb = __popc(wv);
                    if (lane < nb) {
                        const int slot = cnt + lane;
                        if (slot < K) {
                            const int id = (w0 + l) * 32 + (int)__fns(wv, 0, lane + 1);
                            const float d2 = sqdist_ref(__fsub_rn(__ldg(db + 3 * id), qx), __fsub_rn(__ldg(db + 3 * id + 1), qy),
                                                        __fsub_rn(__ldg(db + 3 * id + 2), qz));
                            nn_index[row + slot] = id;
                            nn_dist[row + slot] = __fsqrt_rn(__fsqrt_rn(d2));
                        }
                    }
                    cnt += nb;
                }
            }
        }
        const int c = min(cnt, K);
        for (int s = c + lane; s < K; s += 32) { nn_index[row + s] = 0; nn_dist[row + s] = 0.0f; }
        if (lane == 0) nn_count[qid] = c;                          // 0 == "found nothing": phase 2 replays the chain
        __syncwarp();
    }
}

template <int QPW>
__global__ void __launch_bounds__(256)
sphere_query_kernel(int B, int N, int M, int K, float radius0,
                    const float* __restrict__ database, const float* __restrict__ query,
                    const GridInfo* __restrict__ ginfo,
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
    int mine = 0;                                                // 1 = this kernel answers the query
    if (lane < QPW && j0 + lane < M) {
        const float r = chain_radius(b, j0 + lane, M, radius0);
        mine = (ginfo == nullptr) || (cell_reach(r, ginfo[b].h) > CELL_REACH);
        if (mine) myT = range_threshold(r);
    }
    if (!__any_sync(FULL_MASK, mine)) return;                    // the grid kernel owns the whole group
    // negated query coordinates, duplicated into both halves of a register pair: p - q == p + (-q) exactly, and the
    // distance of TWO database points per lane is then formed with packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2:
    // the same IEEE operations in the same order as sqdist_ref, half the issue slots -- the scan is issue-bound)
    float2 nqx[QPW], nqy[QPW], nqz[QPW];
    float thr[QPW];
    int cnt[QPW];
    bool own[QPW];
#pragma unroll
    for (int q = 0; q < QPW; q++) {
        int j = min(j0 + q, M - 1);
        const float* qp = query + ((size_t)b * M + j) * 3;
        const float x = -__ldg(qp), y = -__ldg(qp + 1), z = -__ldg(qp + 2);
        nqx[q] = make_float2(x, x); nqy[q] = make_float2(y, y); nqz[q] = make_float2(z, z);
        thr[q] = __shfl_sync(FULL_MASK, myT, q);
        own[q] = __shfl_sync(FULL_MASK, mine, q) != 0;
        cnt[q] = own[q] ? 0 : K;                                 // padding / grid-owned queries are "done"
    }

    const float* db = database + (size_t)b * N * 3;
    const unsigned lt = (1u << lane) - 1u;
    const float qnan = __int_as_float(0x7fc00000);
    for (int base = 0; base < N; base += 64) {                   // lane l: points base+l (.x) and base+32+l (.y)
        const int k0 = base + lane, k1 = base + 32 + lane;
        float2 px = make_float2(qnan, qnan), py = px, pz = px;
        if (k0 < N) { px.x = __ldg(db + 3 * k0); py.x = __ldg(db + 3 * k0 + 1); pz.x = __ldg(db + 3 * k0 + 2); }
        if (k1 < N) { px.y = __ldg(db + 3 * k1); py.y = __ldg(db + 3 * k1 + 1); pz.y = __ldg(db + 3 * k1 + 2); }
        bool all_done = true;
#pragma unroll
        for (int q = 0; q < QPW; q++) {
            if (cnt[q] < K) {                                    // warp-uniform
                const float2 dx = __fadd2_rn(px, nqx[q]), dy = __fadd2_rn(py, nqy[q]), dz = __fadd2_rn(pz, nqz[q]);
                float2 d2 = __fmul2_rn(dy, dy);                  // Q6: the y product is rounded alone ...
                d2 = __ffma2_rn(dx, dx, d2);                     // ... the x and z products are fused
                d2 = __ffma2_rn(dz, dz, d2);
                const bool hit0 = d2.x <= thr[q], hit1 = d2.y <= thr[q];
                const unsigned m0 = __ballot_sync(FULL_MASK, hit0), m1 = __ballot_sync(FULL_MASK, hit1);
                if (m0 | m1) {
                    const size_t o = ((size_t)b * M + (j0 + q)) * K;
                    int c = cnt[q];
                    const int slot0 = c + __popc(m0 & lt);        // ascending index: all of the first half first
                    if (hit0 && slot0 < K) {
                        nn_index[o + slot0] = k0;
                        nn_dist[o + slot0] = __fsqrt_rn(__fsqrt_rn(d2.x));   // Q2: sqrt of the distance
                    }
                    c += __popc(m0);
                    const int slot1 = c + __popc(m1 & lt);
                    if (hit1 && slot1 < K) {
                        nn_index[o + slot1] = k1;
                        nn_dist[o + slot1] = __fsqrt_rn(__fsqrt_rn(d2.y));
                    }
                    cnt[q] = c + __popc(m1);
                }
                all_done &= (cnt[q] >= K);
            }
        }
        if (all_done) break;
    }
#pragma unroll
    for (int q = 0; q < QPW; q++) {
        if (own[q]) {
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

// workspace layout: GridInfo[B] | cell_start[B][GRID_CELLS+1] | cell_cursor[B][GRID_CELLS+1] | cell_pts[B][N]
static bool grid_applicable(int N)
{
    const int v = tunables().nnquery_grid;                 // SPH3D_NNQUERY_GRID: 0 = never, 2 = whenever possible (tests), else auto
    if (v == 0) return false;
    const int min_n = (v == 2) ? 32 : GRID_MIN_N;
    // per-warp bitmap of N bits, 8 warps per CTA, within the shared-memory budget
    return N >= min_n && (size_t)((N + 31) / 32) * 4 * 8 <= 200 * 1024;
}

extern "C" size_t sph3d_build_sphere_neighbor_workspace_bytes(int B, int N, int M, int K)
{
    if (B <= 0 || N <= 0 || M <= 0 || K <= 0 || !grid_applicable(N)) return 0;
    size_t info = ((size_t)B * sizeof(GridInfo) + 255) / 256 * 256;
    return info + (size_t)B * (GRID_CELLS + 1) * sizeof(int) * 2 + (size_t)B * N * sizeof(int);
}

extern "C" int sph3d_build_sphere_neighbor(int B, int N, int M, int K, float radius,
                                           const float* database, const float* query,
                                           int* nn_index, int* nn_count, float* nn_dist,
                                           void* workspace, size_t workspace_bytes, void* stream)
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
    int launches = 0;
    const GridInfo* ginfo = nullptr;
    const size_t need = sph3d_build_sphere_neighbor_workspace_bytes(B, N, M, K);
    if (need && workspace && workspace_bytes >= need) {
        // cell grid over every database cloud + stencil queries for the small-radius chain steps
        GridInfo* gi = (GridInfo*)workspace;
        int* cell_start = (int*)((char*)workspace + ((size_t)B * sizeof(GridInfo) + 255) / 256 * 256);
        int* cell_cursor = cell_start + (size_t)B * (GRID_CELLS + 1);
        int* cell_pts = cell_cursor + (size_t)B * (GRID_CELLS + 1);
        grid_setup_kernel<<<B, 256, 0, st>>>(N, radius, database, gi, cell_start);
        SPH3D_CHECK_LAUNCH();
        size_t pts = (size_t)B * N;
        unsigned gb = (unsigned)((pts + 255) / 256 < (size_t)sm_count() * 16 ? (pts + 255) / 256 : (size_t)sm_count() * 16);
        grid_bin_kernel<<<gb, 256, 0, st>>>(B, N, 0, database, gi, cell_start, cell_cursor, cell_pts);
        SPH3D_CHECK_LAUNCH();
        grid_scan_kernel<<<B, 1024, 0, st>>>(gi, cell_start, cell_cursor);
        SPH3D_CHECK_LAUNCH();
        grid_bin_kernel<<<gb, 256, 0, st>>>(B, N, 1, database, gi, cell_start, cell_cursor, cell_pts);
        SPH3D_CHECK_LAUNCH();
        const size_t bm_bytes = (size_t)((N + 31) / 32) * 4 * 8;
        if (bm_bytes > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(sphere_query_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_bytes);
            if (e != cudaSuccess) return (int)e;
        }
        long long qctas = ((long long)B * M + 7) / 8;
        long long cap = (long long)sm_count() * 8;
        sphere_query_grid_kernel<<<(unsigned)(qctas < cap ? qctas : cap), 256, bm_bytes, st>>>(
            B, N, M, K, radius, database, query, gi, cell_start, cell_pts, nn_index, nn_count, nn_dist);
        SPH3D_CHECK_LAUNCH();
        ginfo = gi;
        launches += 5;
    }
    sphere_query_kernel<QPW><<<(unsigned)ctas, 256, 0, st>>>(B, N, M, K, radius, database, query, ginfo,
                                                              nn_index, nn_count, nn_dist);
    SPH3D_CHECK_LAUNCH();
    int chains = REF_GRID * REF_BLOCK;
    sphere_fixup_kernel<<<chains / 4, 128, 0, st>>>(B, N, M, K, radius, database, query,
                                                    nn_index, nn_count, nn_dist);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = launches + 2;
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
