// conv_bwd_t.cu -- depthwise spherical graph convolution, backward, "transposed" form, sm_100a.
//
// Same contract as conv_bwd.cu (replaces depthwiseConv3dGradLauncher,
// /root/reference/tf_ops/convolution/tf_conv3d_gpu.cu:115-140, kernels :32-101), different algorithm.
//
// The reference (and conv_bwd.cu) walk the graph row by row of the OUTPUT points m and scatter into
// grad_input with one float reduction per edge and channel; on B200 those E*C reductions are the floor
// of that form (L2 reduction throughput).  Here the graph is first TRANSPOSED: for every
// INPUT point n the list of (output point m, bin f) pairs that reference it, grouped by bin.  With
//      T[n,f,:] = sum_{m : (m -> n) in bin f} gO[b,m,:] / cnt[b,m]              (one gather per edge)
// both gradients are plain gathers and register sums -- no per-edge atomics at all:
//      grad_input [b,n,c]   = sum_f sum_j W[f,c,j] * T[n,f,c*r+j]
//      grad_filter[f,c,j]   = sum_{b,n} in[b,n,c] * T[n,f,c*r+j]
// so ONE pass over the edges (E strips of C*r floats gathered from gO) replaces one gather pass plus one
// reduction pass.
//
// Pipeline of one call (all on the caller's stream, caller-owned workspace):
//   1. transpose_edges<count>: rank = seg[(b,n), f']++ per edge (one returning integer atomic), f' = (f % G)*SLOTS + f / G
//   2. exclusive scan of seg (two small kernels): seg[s] = START of segment s, seg[B*N*G*SLOTS] = number of edges
//   3. transpose_edges<fill>: entries[seg[s] + rank] = (b*M+m) << 8 | (f / G), no atomics
//   4. (optional, default on) per-segment insertion sort of the entries by m => run-to-run deterministic
//   5. scale_rows: gs[b,m,:] = gO[b,m,:] / cnt[b,m] into a (B, M+1, C*r) buffer whose row M is zero
//      (the landing row of the padding edges that round every tile up to a multiple of four)
//   6. conv_bwd_t_kernel: a group of G warps shares an input point; warp w of the group owns the bins
//      f = w, w+G, ... (SLOTS = 9 of them), i.e. ONE contiguous, bin-sorted sub-list of the point's
//      entries.  Flat 4-deep gather loop over the sub-list (LDG.128 + FADD2); at the end of each bin
//      segment: grad_input strip += W[f]*T (filter strip from shared memory, FFMA2) and the warp's REGISTER
//      accumulator of that bin += in[n]*T (warp-uniform switch => static register indexing, no atomics, no
//      shared-memory accumulation).  The G partial grad_input strips of a point meet in global memory with
//      G vector reductions per point (E/K*G instead of E).
//   7. reduce_partials_kernel (conv_bwd.cu) sums the per-group filter partials in a fixed order.
// Steps 1-4 depend only on the graph: sph3d_conv_transpose() exposes them so that a caller who reuses a
// graph (two convolutions per level, every training step of a static graph) builds the plan once.
//
// Channels are handled in FLAT output-channel space i = c*r + j (Co = C*r): lane l owns VEC consecutive
// flat channels, so the kernel is the r=1 kernel plus an expansion of in[n,c] to flat channels and a
// pairwise sum when grad_input is written (r = 2).
#include "conv_common.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

// CTA shape of the gather kernel (one CTA per SM): 768 threads x 8 gathers in flight per warp by default;
// SPH3D_BWDT_THREADS = 512 / 640 / 1024 and SPH3D_BWDT_DEPTH select the other compiled shapes (sweeps in profiles/)
static inline int t_warps()
{
    const int t = tun(tunables().bwdt_threads, 768);
    return (t == 512 || t == 640 || t == 1024) ? t / 32 : 24;
}
static inline int t_depth(int warps)
{
    const int d = tunables().bwdt_depth;
    if (warps == 16) return d == 4 ? 4 : 3;
    if (warps == 20) return d == 3 ? 3 : 2;
    if (warps == 32) return 1;
    return 2;
}
constexpr int SCAN_TILE = 4096;          // ints per CTA in the scan kernels (256 threads x 16)

// ------------------------------------------------------------------------------------------- plan
struct TGeom {
    int G, SLOTS, FP;           // bin classes (f % G), bins per class, G*SLOTS segments per point
    size_t nseg, nseg_pad;      // B*N*FP; nseg+1 (the total sits behind the last start) rounded up to SCAN_TILE
    int rank_bytes;             // 2 or 4: width of the per-edge rank scratch
    int sb, cb;                 // entry = row << (sb+cb) | (cnt-1) << sb | slot: bits of the slot and of the folded 1/cnt code
    bool fold;                  // cb > 0: the entry carries nn_count of its output row, the kernel gathers grad_output itself
    int scan_blocks;
    size_t seg_off, sums_off, ent_off, rank_off, total;   // byte offsets inside the plan
    bool ok;
};

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static TGeom t_geom(int B, int N, int M, int F, int K, bool want_fold)
{
    TGeom g{};
    g.ok = false;
    if (B <= 0 || N <= 0 || M <= 0 || F <= 0 || K <= 0) return g;
    // smallest number of bin classes G (a divisor of the CTA's warp count) whose per-warp accumulators
    // (ceil(F/G) strips of 512 B, sized for 16-byte strips so that the plan does not depend on C) fit in shared
    // memory beside the filter and the staging arrays: F=33 -> G=4 x 9 bins, F=17 -> 2 x 9, F=49 -> 4 x 13
    int G = 0, SL = 0;
    const int nw = t_warps();
    const int g_min = tun(tunables().bwdt_g, 1);
    // Bin 0 (the query point itself) holds one edge per point; the other F-1 bins split evenly over the classes
    // when G divides F-1 (n*p*q with n = 8 azimuth bins: G = 2, 4, 8), which keeps the classes' edge counts equal.
    for (int pass = 0; pass < 2 && !G; pass++)
        for (int d = g_min; d <= nw; d++) {
            if (nw % d) continue;
            if (pass == 0 && F > 1 && (F - 1) % d) continue;
            const int sl = (F + d - 1) / d;
            if (sl > 127) continue;
            if ((size_t)nw * sl * 512 + (size_t)F * 512 + (size_t)nw * 768 <= 210 * 1024) { G = d; SL = sl; break; }
        }
    if (!G) return g;                                             // very large F: conv_bwd.cu handles it
    // entry layout: the slot (bin inside its class) in the low sb bits; above it, when they fit beside the output row
    // b*M+m, the cb bits of nn_count-1 of that row, so that the gather kernel scales by 1/cnt itself and no scaled copy
    // of grad_output is ever written (fold).  Otherwise cb = 0 and the kernel gathers a pre-scaled copy.
    int sb = 1;
    while ((1 << sb) < SL) sb++;
    int cb = 1;
    while ((1 << cb) < K) cb++;
    const long long BM = (long long)B * M;
    // want_fold: the pool / unpool gather always folds (its scale is per edge); the convolution folds only on request
    // (SPH3D_BWDT_FOLD=1): the extra staged word costs the gather kernel more than the scaled copy saves
    bool fold = want_fold && BM <= (1LL << (32 - sb - cb));
    if (!fold) cb = 0;
    if (BM > (1LL << (32 - sb))) return g;
    g.sb = sb; g.cb = cb; g.fold = fold;
    if ((long long)B * M * K >= (1LL << 31) || (long long)B * N * G * SL >= (1LL << 31)) return g;
    g.G = G; g.SLOTS = SL; g.FP = G * SL;
    g.nseg = (size_t)B * N * g.FP;
    g.nseg_pad = (g.nseg + 1 + SCAN_TILE - 1) / SCAN_TILE * SCAN_TILE;
    if (g.nseg_pad / SCAN_TILE > 65536) return g;
    g.scan_blocks = (int)(g.nseg_pad / SCAN_TILE);
    g.seg_off = 0;
    g.sums_off = align256(g.nseg_pad * sizeof(int));
    g.ent_off = g.sums_off + align256((size_t)g.scan_blocks * sizeof(int));
    g.rank_bytes = (M <= 65535) ? 2 : 4;                          // a segment holds distinct output points of one cloud
    g.rank_off = g.ent_off + align256((size_t)B * M * K * sizeof(int));
    g.total = g.rank_off + align256((size_t)B * M * K * g.rank_bytes);
    g.ok = true;
    return g;
}

// ------------------------------------------------------------------------------ transpose kernels
// one thread per edge slot (b,m,k); edges beyond nn_count and malformed ids are skipped.
// Pass 1 (FILL = false): rank of the edge inside its segment from ONE returning integer atomic; seg ends up holding the
// segment sizes.  Pass 2 (FILL = true, after the exclusive scan): position = start + rank, no atomics.
// code_k: the cb code bits of an entry hold the edge's slot k in its row (weighted interpolation: the kernel fetches
// weight[row, k]) instead of nn_count - 1.
template <bool FILL, typename RankT>
__global__ void __launch_bounds__(256)
transpose_edges_kernel(size_t slots, unsigned M, unsigned N, int K, int F, int G, int SLOTS, int sb, int cb, int code_k,
                       const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                       const int* __restrict__ bin_index, int* __restrict__ seg, RankT* __restrict__ ranks,
                       unsigned* __restrict__ entries)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < slots; t += (size_t)gridDim.x * blockDim.x) {
        const size_t row = t / (unsigned)K;
        const int k = (int)(t - row * (unsigned)K);
        const int cnt = min(__ldg(nn_count + row), K);
        if (k >= cnt) continue;
        const int n = __ldg(nn_index + t), f = bin_index ? __ldg(bin_index + t) : 0;
        if ((unsigned)n >= N || (unsigned)f >= (unsigned)F) continue;
        const unsigned b = (unsigned)(row / M);
        const size_t s = ((size_t)b * N + n) * (G * SLOTS) + (f % G) * SLOTS + f / G;
        if constexpr (FILL) {
            const int pos = __ldg(seg + s) + (int)ranks[t];
            // row = b*M + m: the row of grad_output this edge gathers; cb bits of cnt-1 when the 1/cnt scale is folded in
            entries[pos] = ((unsigned)row << (sb + cb)) | (cb ? ((unsigned)(code_k ? k : cnt - 1) << sb) : 0u) | (unsigned)(f / G);
        } else {
            ranks[t] = (RankT)atomicAdd(seg + s, 1);
        }
    }
}

__device__ __forceinline__ int block_sum_256(int v, int* sm)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    if (lane == 0) sm[w] = v;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) tot += sm[i];
    __syncthreads();
    return tot;
}

__global__ void __launch_bounds__(256)
scan_reduce_kernel(const int4* __restrict__ seg4, int* __restrict__ sums)
{
    __shared__ int sm[8];
    const int4* p = seg4 + (size_t)blockIdx.x * (SCAN_TILE / 4);
    int s = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int4 v = p[j * 256 + threadIdx.x];
        s += v.x + v.y + v.z + v.w;
    }
    const int tot = block_sum_256(s, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// in-place exclusive scan of the CTA's 4096 counters, offset by the sum of all earlier CTAs
__global__ void __launch_bounds__(256)
scan_apply_kernel(int4* __restrict__ seg4, const int* __restrict__ sums)
{
    __shared__ int sm[8];
    __shared__ int wtot[8];
    int pre = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += 256) pre += sums[i];
    const int base = block_sum_256(pre, sm);
    int4* p = seg4 + (size_t)blockIdx.x * (SCAN_TILE / 4) + threadIdx.x * 4;     // 16 consecutive ints per thread
    int4 v[4];
    int tsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { v[j] = p[j]; tsum += v[j].x + v[j].y + v[j].z + v[j].w; }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    int run = base + incl - tsum;
    for (int i = 0; i < w; i++) run += wtot[i];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int4 o;
        o.x = run; run += v[j].x;
        o.y = run; run += v[j].y;
        o.z = run; run += v[j].z;
        o.w = run; run += v[j].w;
        p[j] = o;
    }
}

// canonical order inside every segment (ascending m): makes the float sums independent of the order in which the
// fill kernel's atomic cursor handed out positions.  Segments are short (mean E / (#non-empty segments) ~ 4).
__global__ void __launch_bounds__(256)
sort_segments_kernel(size_t nseg, const int* __restrict__ seg, unsigned* __restrict__ entries)
{
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += (size_t)gridDim.x * blockDim.x) {
        const int beg = seg[s], end = seg[s + 1];
        for (int i = beg + 1; i < end; i++) {
            const unsigned x = entries[i];
            int j = i;
            while (j > beg) {
                const unsigned y = entries[j - 1];
                if (y <= x) break;
                entries[j] = y;
                j--;
            }
            entries[j] = x;
        }
    }
}

// Sub-lists (one point, one bin class) longer than T_PART entries are cut into parts: part 0 stays with the point, parts
// 1.. become OVERFLOW items {point, first entry, one past the last entry} of their class, which the gather kernel hands
// out before the points.  The table lives in the plan's rank scratch, which is dead once the fill pass has run:
// [G counters | G tables of `cap` items].  One thread per (point, class).
constexpr int T_PART = 2048;

__global__ void __launch_bounds__(256)
build_overflow_kernel(size_t lists /* B*N*G */, int G, int SLOTS, int cap, const int* __restrict__ seg,
                      int* __restrict__ hdr, int4* __restrict__ tab)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < lists; t += (size_t)gridDim.x * blockDim.x) {
        const size_t pt = t / (unsigned)G;
        const int cls = (int)(t - pt * (unsigned)G);
        const size_t s0 = pt * ((size_t)G * SLOTS) + (size_t)cls * SLOTS;
        const int beg = __ldg(seg + s0), end = __ldg(seg + s0 + SLOTS);
        const int extra = (end - beg - 1) / T_PART;                   // parts beyond the first
        if (extra <= 0) continue;
        const int base = atomicAdd(hdr + cls, extra);
        for (int j = 0; j < extra && base + j < cap; j++) {           // cap = slots / T_PART + 1 can never be exceeded
            const int b0 = beg + (j + 1) * T_PART;
            tab[(size_t)cls * cap + base + j] = make_int4((int)pt, b0, min(end, b0 + T_PART), 0);
        }
    }
}

constexpr int T_SPLIT_MIN_ROWS = 16384;          // below this no sub-list of a ball-query graph gets long enough to matter
struct TOver { bool split; int cap; size_t hdr_off, tab_off; };
static inline TOver t_over(int B, int M, int K, const TGeom& g)
{
    TOver o{};
    const size_t slots = (size_t)B * M * K;
    o.cap = (int)(slots / T_PART + 1);
    o.hdr_off = g.rank_off;
    o.tab_off = g.rank_off + 256;
    const size_t region = align256(slots * g.rank_bytes);
    o.split = M > T_SPLIT_MIN_ROWS && g.FP > 1 && 256 + (size_t)g.G * o.cap * sizeof(int4) <= region && g.G <= 64;
    return o;
}

// gs[row, :] = gO[row, :] / cnt[row]   (row = b*M + m < B*M);   gs[B*M, :] = 0      -- gs is (B*M + 1, Co)
template <int V>
__global__ void __launch_bounds__(256)
scale_rows_kernel(size_t total /* (B*M+1)*Co/V */, size_t rows /* B*M */, unsigned CoV, int K,
                  const int* __restrict__ nn_count, const float* __restrict__ go, float* __restrict__ gs)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t row = t / CoV;
        float v[V];
#pragma unroll
        for (int u = 0; u < V; u++) v[u] = 0.f;
        if (row < rows) {
            const int cnt = min(__ldg(nn_count + row), K);
            if (cnt > 0) {
                const float inv = 1.0f / (float)cnt;
                VecIO<V>::ld(v, go + t * V, true);
#pragma unroll
                for (int u = 0; u < V; u++) v[u] *= inv;
            }
        }
        VecIO<V>::st(gs + t * V, v);
    }
}

// -------------------------------------------------------------------------------------- main kernel
template <int VEC>
__device__ __forceinline__ void strip_fma(float (&a)[VEC], const float (&x)[VEC], const float (&y)[VEC])
{
    if constexpr (VEC % 2 == 0) {
#pragma unroll
        for (int e = 0; e < VEC; e += 2) {
            const float2 t = __ffma2_rn(make_float2(x[e], x[e + 1]), make_float2(y[e], y[e + 1]), make_float2(a[e], a[e + 1]));
            a[e] = t.x; a[e + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int e = 0; e < VEC; e++) a[e] = fmaf(x[e], y[e], a[e]);
    }
}

template <int VEC>
__device__ __forceinline__ void ld_strip_smem(float (&v)[VEC], const float* p)
{
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (VEC == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = *p;
    }
}

template <int VEC>
__device__ __forceinline__ void st_strip_smem(float* p, const float (&v)[VEC])
{
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else *p = v[0];
}

// Work distribution: warp w of every CTA belongs to bin class w % G; the warps of a class take the input points
// round-robin (point = slot, slot + W, ... with W = warps of that class in the grid).  The in-degrees are very
// uneven (the "first K by index" rule sends most edges to the low-index points of every cloud) but each warp's
// points are an even sample of all of them, so the static split balances, all warps sweep the batch cloud by
// cloud together (the gathered cloud stays L2-resident), and the assignment is reproducible.
// Per point the warp runs a three-stage software pipeline, two points ahead: segment boundaries (i+2), entry
// list + input strip (i+1), gathers (i); inside a point 4*DEPTH feature-strip gathers are in flight.
// FOLD = false: `gs` is the pre-scaled copy grad_output / cnt with one zero row appended at row index `zrow` (the landing
// row of the padding edges) and the segment sums are plain adds.  FOLD = true: `gs` is grad_output itself, every entry
// carries nn_count - 1 of its row in cb bits and the sums are FMAs with 1/cnt (padding edges: scale 0); no scaled copy
// is written, at the price of one more staged word per edge (measured slower at Cfg-T: DESIGN.md 4.3).
template <int VEC, int R, int THREADS, int DEPTH, bool FOLD, bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
conv_bwd_t_kernel(unsigned rows /* B*N */, unsigned zrow, int sb, int cb, int F, int C, int G, int SLOTS,
                  int part, int over_cap, const int* __restrict__ over_hdr, const int4* __restrict__ over_tab,
                  const int* __restrict__ seg, const unsigned* __restrict__ entries, const float* __restrict__ gs,
                  const float* __restrict__ input, const float* __restrict__ filter,
                  float* __restrict__ grad_input, float* __restrict__ gw_partial)
{
    static_assert(VEC % R == 0, "a lane's flat strip must cover whole input channels");
    constexpr int VI = VEC / R;                          // input channels per lane
    constexpr int STRIP = 32 * VEC;                      // floats in a warp-wide strip
    constexpr int NWARPS = THREADS / 32;
    const int Co = C * R;
    extern __shared__ __align__(16) float smem[];
    float* Wsh = smem;                                   // [F][STRIP] flat-channel filter strips of this chunk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cls = warp % G;                            // my bins: f = s*G + cls, s < SLOTS
    const int cbase = blockIdx.y * STRIP;                // first flat channel of this chunk
    for (int t = threadIdx.x; t < F * STRIP; t += THREADS) {
        const int f = t / STRIP, ch = cbase + t % STRIP;
        Wsh[t] = (ch < Co) ? __ldg(filter + (size_t)f * Co + ch) : 0.f;
    }
    float* accS = Wsh + (size_t)F * STRIP + (size_t)warp * SLOTS * STRIP + lane * VEC;   // my accumulators [SLOTS][strip]
    unsigned* sOff = reinterpret_cast<unsigned*>(Wsh + (size_t)F * STRIP + (size_t)NWARPS * SLOTS * STRIP) + warp * 192;
    int* sCode = reinterpret_cast<int*>(sOff + 64);
    float* sScale = reinterpret_cast<float*>(sOff + 128);   // 1/cnt of the edge's output row (1 when gs is pre-scaled, 0 for padding)
    {
        float z[VEC];
#pragma unroll
        for (int e = 0; e < VEC; e++) z[e] = 0.f;
        for (int s = 0; s < SLOTS; s++) st_strip_smem<VEC>(accS + s * STRIP, z);
    }
    __syncthreads();

    const int i0 = cbase + lane * VEC;                   // my first flat channel
    const bool active = i0 < Co;
    const int i0ld = active ? i0 : 0;                    // idle lanes load a valid strip, never store
    const unsigned gsStrideB = (unsigned)Co * 4u;
    const unsigned smask = (1u << sb) - 1u, cmask = (1u << cb) - 1u;
    const int sh = sb + cb;
    const unsigned zoff = zrow * gsStrideB;              // FOLD = false: padding edges gather the zero row
    const char* gb = reinterpret_cast<const char*>(gs) + (size_t)i0ld * 4;
    const float* inl = input + i0ld / R;                 // my first input channel
    float* gil = grad_input + i0ld / R;
    const float* wlane = Wsh + (size_t)cls * STRIP + lane * VEC;      // strip of my first bin; next bin: + G*STRIP
    const int wstep = G * STRIP;
    const unsigned FP = (unsigned)(G * SLOTS);
    const unsigned W = gridDim.x * (NWARPS / G);         // warps of my class in the grid
    const unsigned segoff = (unsigned)(cls * SLOTS);

    // SPLIT (graphs of more than T_SPLIT_MIN_ROWS rows per cloud).  Work items of my class: first the OVERFLOW items --
    // parts 1, 2, ... of the sub-lists longer than `part` entries (build_overflow_kernel; a hub referenced by tens of
    // thousands of rows, as in the ScanNet-stress graph whose radius chain ends up keeping the same 64 lowest-index points
    // for 90 % of its rows, is otherwise ONE warp's serial walk and the whole kernel's critical path) --, then every
    // point's sub-list, capped at its first `part` entries.  Partial sums of a point meet in grad_input through the
    // vector reductions the bin classes already use.  Without SPLIT an item is a point and the loop is the one measured
    // at the headline shape (the three extra live values of the item form cost it 6 %: register spills).
    const unsigned nov = (SPLIT && over_hdr) ? (unsigned)min(__ldg(over_hdr + cls), over_cap) : 0u;
    const int4* otab = over_tab + (size_t)cls * over_cap;
    const unsigned nitems = SPLIT ? nov + rows : rows;
    // stage "boundaries": lane 0 loads the start, lane 1 the end of the item's entry range (SPLIT: lane 2 its point)
    auto load_b = [&](unsigned it) {
        int bv = 0;
        if constexpr (SPLIT) {
            if (it < nov) {
                if (lane < 3) {
                    const int4 o = __ldg(otab + it);
                    bv = lane == 0 ? o.y : (lane == 1 ? o.z : o.x);
                }
            } else if (it < nitems) {
                const unsigned pt = it - nov;
                const unsigned sb = pt * FP + segoff;
                if (lane == 0) bv = __ldg(seg + sb);
                if (lane == 1) bv = __ldg(seg + sb + SLOTS);
                if (lane == 2) bv = (int)pt;
            }
        } else {
            if (it < rows) {
                const unsigned sb = it * FP + segoff;         // seg[s] = start of segment s; seg[nseg] = number of edges
                if (lane == 0) bv = __ldg(seg + sb);
                if (lane == 1) bv = __ldg(seg + sb + SLOTS);
            }
        }
        return bv;
    };
    unsigned row = blockIdx.x * (NWARPS / G) + warp / G;  // item index (without SPLIT: the point)
    unsigned row1 = row + W;
    // item i+1: boundaries resolved, entries + input strip in flight; item i+2: boundaries in flight
    int bv2 = load_b(row);
    int beg1, end1; unsigned pt1 = 0, e0_1 = 0, e1_1 = 0; float in1[VI];
    auto stage_e = [&](unsigned it, int bv) {
        beg1 = __shfl_sync(FULL_MASK, bv, 0); end1 = __shfl_sync(FULL_MASK, bv, 1);
        unsigned pt = it;
        if constexpr (SPLIT) {
            pt = (unsigned)__shfl_sync(FULL_MASK, bv, 2);
            pt1 = pt;
            if (it >= nov) end1 = min(end1, beg1 + part);    // the rest of a long sub-list is among the overflow items
        }
        e0_1 = 0; e1_1 = 0;
        if (end1 > beg1) {
            if (beg1 + lane < end1) e0_1 = __ldg(entries + beg1 + lane);
            if (beg1 + 32 + lane < end1) e1_1 = __ldg(entries + beg1 + 32 + lane);
            VecIO<VI>::ld(in1, inl + (size_t)pt * C, true);
        }
    };
#pragma unroll
    for (int v = 0; v < VI; v++) in1[v] = 0.f;
    stage_e(row, bv2);
    bv2 = load_b(row1);

    for (; row < nitems;) {
        const unsigned crow = SPLIT ? pt1 : row;
        const int beg = beg1, end = end1;
        const unsigned ce0 = e0_1, ce1 = e1_1;
        float inx[VEC];                                   // in[b,n,c] expanded to my flat channels
#pragma unroll
        for (int e = 0; e < VEC; e++) inx[e] = in1[e / R];
        stage_e(row1, bv2);                               // loads for item i+1 (its boundaries arrived during item i-1)
        row = row1;
        row1 += W;
        bv2 = load_b(row1);                               // boundary loads for item i+2
        if (end <= beg) continue;

        float gi[VEC], T[VEC];
#pragma unroll
        for (int e = 0; e < VEC; e++) { gi[e] = 0.f; T[e] = 0.f; }

        auto consume = [&](const float (&v)[VEC], int code, float sc) {
            if constexpr (FOLD) {
                float scv[VEC];
#pragma unroll
                for (int e = 0; e < VEC; e++) scv[e] = sc;
                strip_fma<VEC>(T, v, scv);
            } else {
                strip_add<VEC>(T, v);
            }
            if (code & 1) {                                // last edge of its bin segment (warp-uniform)
                const int s = code >> 1;
                float w[VEC], a[VEC];
                ld_strip_smem<VEC>(w, wlane + s * wstep);
                ld_strip_smem<VEC>(a, accS + s * STRIP);
                strip_fma<VEC>(gi, w, T);
                strip_fma<VEC>(a, inx, T);
                st_strip_smem<VEC>(accS + s * STRIP, a);
#pragma unroll
                for (int e = 0; e < VEC; e++) T[e] = 0.f;
            }
        };
        auto load4 = [&](int p, float (&v)[4][VEC]) {
            const uint4 oo = *reinterpret_cast<const uint4*>(sOff + p);
            ld_strip<VEC>(v[0], gb, oo.x); ld_strip<VEC>(v[1], gb, oo.y);
            ld_strip<VEC>(v[2], gb, oo.z); ld_strip<VEC>(v[3], gb, oo.w);
        };
        auto consume4 = [&](int p, const float (&v)[4][VEC]) {
            const int4 cc = *reinterpret_cast<const int4*>(sCode + p);
            float4 ss = make_float4(1.f, 1.f, 1.f, 1.f);
            if constexpr (FOLD) ss = *reinterpret_cast<const float4*>(sScale + p);
            if (((cc.x | cc.y | cc.z | cc.w) & 1) == 0) {  // no segment ends inside this batch (the common case)
                if constexpr (FOLD) {
                    float s0[VEC], s1[VEC], s2[VEC], s3[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; e++) { s0[e] = ss.x; s1[e] = ss.y; s2[e] = ss.z; s3[e] = ss.w; }
                    strip_fma<VEC>(T, v[0], s0); strip_fma<VEC>(T, v[1], s1);
                    strip_fma<VEC>(T, v[2], s2); strip_fma<VEC>(T, v[3], s3);
                } else {
                    float u[VEC], w2[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; e++) { u[e] = v[0][e]; w2[e] = v[2][e]; }
                    strip_add<VEC>(u, v[1]); strip_add<VEC>(w2, v[3]);
                    strip_add<VEC>(T, u); strip_add<VEC>(T, w2);
                }
            } else {
                consume(v[0], cc.x, ss.x); consume(v[1], cc.y, ss.y); consume(v[2], cc.z, ss.z); consume(v[3], cc.w, ss.w);
            }
        };

        for (int kt = beg; kt < end; kt += 64) {
            const int nt = min(64, end - kt);
            const int nt4 = (nt + 3) & ~3;
            const int p0 = lane, p1 = 32 + lane;
            unsigned e0 = ce0, e1 = ce1;
            if (kt != beg) {                               // only the first tile was prefetched
                e0 = 0; e1 = 0;
                if (p0 < nt) e0 = __ldg(entries + kt + p0);
                if (p1 < nt) e1 = __ldg(entries + kt + p1);
            }
            const int s0 = (int)(e0 & smask), s1 = (int)(e1 & smask);
            // padding edges (tiles are rounded up to a multiple of four): FOLD gathers the tile's first row with scale 0
            unsigned off_pad = zoff;
            if constexpr (FOLD) off_pad = __shfl_sync(FULL_MASK, (e0 >> sh) * gsStrideB, 0);
            int nx0 = __shfl_down_sync(FULL_MASK, s0, 1);
            const int nx1 = __shfl_down_sync(FULL_MASK, s1, 1);
            const int first1 = __shfl_sync(FULL_MASK, s1, 0);
            if (lane == 31) nx0 = first1;
            if (p0 < nt4) {
                const bool real = p0 < nt;
                sOff[p0] = real ? (e0 >> sh) * gsStrideB : off_pad;
                sCode[p0] = real ? ((s0 << 1) | ((p0 == nt - 1 || nx0 != s0) ? 1 : 0)) : 0;
                if constexpr (FOLD) sScale[p0] = real ? 1.0f / (float)(((e0 >> sb) & cmask) + 1u) : 0.f;
            }
            if (p1 < nt4) {
                const bool real = p1 < nt;
                sOff[p1] = real ? (e1 >> sh) * gsStrideB : off_pad;
                sCode[p1] = real ? ((s1 << 1) | ((p1 == nt - 1 || nx1 != s1) ? 1 : 0)) : 0;
                if constexpr (FOLD) sScale[p1] = real ? 1.0f / (float)(((e1 >> sb) & cmask) + 1u) : 0.f;
            }
            __syncwarp();
            // software pipeline over batches of four gathers, 4*DEPTH strips in flight
            float v[DEPTH][4][VEC];
#pragma unroll
            for (int d = 0; d < DEPTH - 1; d++)
                if (d * 4 < nt4) load4(d * 4, v[d]);
            for (int p = 0; p < nt4; p += 4 * DEPTH) {
#pragma unroll
                for (int d = 0; d < DEPTH; d++) {
                    const int pp = p + 4 * d;
                    if (pp < nt4) {
                        const int pn = pp + 4 * (DEPTH - 1);
                        if (pn < nt4) load4(pn, v[(d + DEPTH - 1) % DEPTH]);
                        consume4(pp, v[d]);
                    }
                }
            }
            __syncwarp();                                      // sOff/sCode are rewritten by the next tile/point
        }
        if (active) {
            float o[VI];
#pragma unroll
            for (int v = 0; v < VI; v++) {
                float t = 0.f;
#pragma unroll
                for (int j = 0; j < R; j++) t += gi[v * R + j];
                o[v] = t;
            }
            VecIO<VI>::red(gil + (size_t)crow * C, o);
        }
    }
    // partial [blockIdx.x][warp / G][f][flat channel]: every (f, channel) of a partial is written by exactly one warp
    __syncwarp();
    if (active) {
        float* part = gw_partial + ((size_t)blockIdx.x * (NWARPS / G) + warp / G) * F * Co;
        for (int s = 0; s < SLOTS; s++) {
            const int f = s * G + cls;
            if (f < F) {
                float a[VEC];
                ld_strip_smem<VEC>(a, accS + s * STRIP);
                VecIO<VEC>::st(part + (size_t)f * Co + i0, a);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ host helpers
struct TPlanMain {
    int vec, chunks, grid_x, threads, per_cta;     // per_cta: filter partials one CTA writes (warps / G)
    size_t smem;
};

static bool t_plan_main(int B, int N, int M, int F, int C, int r, const TGeom& g, TPlanMain* out)
{
    if (!g.ok || (r != 1 && r != 2)) return false;
    const long long Co = (long long)C * r;
    if (((long long)B * M + 1) * Co * 4 >= (1LL << 32)) return false;     // 32-bit byte offsets into gs
    int vec = pick_vec_full_warp((int)Co);
    if (vec % r != 0) vec = r;                                            // a lane's strip covers whole input channels (Co % r == 0)
    TPlanMain p{};
    p.vec = vec;
    p.chunks = (int)((Co + 32 * vec - 1) / (32 * vec));
    const int nw = t_warps();
    p.threads = nw * 32;
    if (nw % g.G) return false;
    p.per_cta = nw / g.G;
    p.smem = ((size_t)F + (size_t)nw * g.SLOTS) * 32 * vec * sizeof(float) + (size_t)nw * 192 * sizeof(int);
    if (p.smem > SMEM_CAP) return false;
    const long long rows = (long long)B * N;
    long long want = sm_count();
    if (p.chunks > 1) want = (want + p.chunks - 1) / p.chunks;
    const long long need = (rows + p.per_cta - 1) / p.per_cta;             // one point per warp of a class at least
    if (want > need) want = need;
    if (want < 1) want = 1;
    p.grid_x = (int)want;
    *out = p;
    return true;
}

static inline unsigned grid_for(size_t work, int threads, int per_sm)
{
    size_t want = (work + threads - 1) / threads, cap = (size_t)sm_count() * per_sm;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

static int t_build_plan(int B, int N, int M, int F, int K, const TGeom& g, const int* nn_index, const int* nn_count,
                        const int* bin_index, char* plan, cudaStream_t st, int* launches, int code_k = 0)
{
    int* seg = reinterpret_cast<int*>(plan + g.seg_off);
    int* sums = reinterpret_cast<int*>(plan + g.sums_off);
    unsigned* ent = reinterpret_cast<unsigned*>(plan + g.ent_off);
    cudaError_t e = cudaMemsetAsync(seg, 0, g.nseg_pad * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    const size_t slots = (size_t)B * M * K;
    const unsigned ge = grid_for(slots, 256, 16);
    void* ranks = plan + g.rank_off;
#define EDGES(FILL, RT)                                                                                                    \
    transpose_edges_kernel<FILL, RT><<<ge, 256, 0, st>>>(slots, (unsigned)M, (unsigned)N, K, F, g.G, g.SLOTS, g.sb, g.cb, code_k, nn_index, \
                                                         nn_count, bin_index, seg, reinterpret_cast<RT*>(ranks), ent)
    if (g.rank_bytes == 2) EDGES(false, unsigned short); else EDGES(false, unsigned);
    SPH3D_CHECK_LAUNCH();
    scan_reduce_kernel<<<g.scan_blocks, 256, 0, st>>>(reinterpret_cast<const int4*>(seg), sums);
    SPH3D_CHECK_LAUNCH();
    scan_apply_kernel<<<g.scan_blocks, 256, 0, st>>>(reinterpret_cast<int4*>(seg), sums);
    SPH3D_CHECK_LAUNCH();
    if (g.rank_bytes == 2) EDGES(true, unsigned short); else EDGES(true, unsigned);
#undef EDGES
    SPH3D_CHECK_LAUNCH();
    *launches += 4;
    {   // overflow items of the gather kernel (in the rank scratch, which the fill pass has finished with)
        const TOver o = t_over(B, M, K, g);
        if (o.split) {
            e = cudaMemsetAsync(plan + o.hdr_off, 0, 256, st);
            if (e != cudaSuccess) return (int)e;
            const size_t lists = (size_t)B * N * g.G;
            build_overflow_kernel<<<grid_for(lists, 256, 8), 256, 0, st>>>(lists, g.G, g.SLOTS, o.cap, seg,
                                                                            reinterpret_cast<int*>(plan + o.hdr_off),
                                                                            reinterpret_cast<int4*>(plan + o.tab_off));
            SPH3D_CHECK_LAUNCH();
            *launches += 1;
        }
    }
    if (tunables().bwdt_sort == 1) {             // opt-in: canonical order inside every segment
        sort_segments_kernel<<<grid_for(g.nseg, 256, 16), 256, 0, st>>>(g.nseg, seg, ent);
        SPH3D_CHECK_LAUNCH();
        *launches += 1;
    }
    return 0;
}

// workspace of the gradient call proper (scaled grad_output + filter partials), after an optional plan
struct TWork { size_t gs_off, part_off, total; size_t P; };

static TWork t_work(int B, int M, int F, int C, int r, const TGeom& g, const TPlanMain& p)
{
    TWork w{};
    const size_t Co = (size_t)C * r;
    w.gs_off = 0;
    w.part_off = g.fold ? 0 : align256(((size_t)B * M + 1) * Co * sizeof(float));   // no scaled copy when 1/cnt rides in the plan
    w.P = (size_t)p.grid_x * p.per_cta;
    w.total = w.part_off + align256(w.P * F * Co * sizeof(float));
    return w;
}

static int t_run_main(int B, int N, int M, int F, int C, int r, int K, const TGeom& g, const TPlanMain& p, const char* plan,
                      const int* nn_count, const float* input, const float* filter, const float* grad_output,
                      float* grad_input, float* grad_filter, char* work, cudaStream_t st, int* launches)
{
    const TWork w = t_work(B, M, F, C, r, g, p);
    const size_t Co = (size_t)C * r;
    float* gs = reinterpret_cast<float*>(work + w.gs_off);
    float* part = reinterpret_cast<float*>(work + w.part_off);
    const int* seg = reinterpret_cast<const int*>(plan + g.seg_off);
    const unsigned* ent = reinterpret_cast<const unsigned*>(plan + g.ent_off);
    cudaError_t e = cudaMemsetAsync(grad_input, 0, sizeof(float) * (size_t)B * N * C, st);
    if (e != cudaSuccess) return (int)e;
    const float* gsrc = grad_output;
    const unsigned zrow = (unsigned)((long long)B * M);
    if (!g.fold) {
        gsrc = gs;
        const int V = (Co % 4 == 0) ? 4 : ((Co % 2 == 0) ? 2 : 1);
        const size_t orows = (size_t)B * M;
        const size_t total = (orows + 1) * Co / V;
        const unsigned gr = grid_for(total, 256, 16);
        if (V == 4) scale_rows_kernel<4><<<gr, 256, 0, st>>>(total, orows, (unsigned)(Co / 4), K, nn_count, grad_output, gs);
        else if (V == 2) scale_rows_kernel<2><<<gr, 256, 0, st>>>(total, orows, (unsigned)(Co / 2), K, nn_count, grad_output, gs);
        else scale_rows_kernel<1><<<gr, 256, 0, st>>>(total, orows, (unsigned)Co, K, nn_count, grad_output, gs);
        SPH3D_CHECK_LAUNCH();
    }
    dim3 grid(p.grid_x, p.chunks);
    const unsigned rows = (unsigned)((long long)B * N);
    const TOver ov = t_over(B, M, K, g);
    const int* ohdr = reinterpret_cast<const int*>(plan + ov.hdr_off);
    const int4* otab = reinterpret_cast<const int4*>(plan + ov.tab_off);
    const int opart = T_PART;
#define LAUNCH_T3(V, RR, TH, DP, FO, SP)                                                                         \
    do {                                                                                                         \
        e = set_smem(conv_bwd_t_kernel<V, RR, TH, DP, FO, SP>, p.smem);                                          \
        if (e != cudaSuccess) return (int)e;                                                                     \
        conv_bwd_t_kernel<V, RR, TH, DP, FO, SP><<<grid, TH, p.smem, st>>>(rows, zrow, g.sb, g.cb, F, C, g.G, g.SLOTS, opart, \
                                                                           ov.cap, ohdr, otab, seg, ent, gsrc, input, filter, \
                                                                           grad_input, part);                    \
    } while (0)
#define LAUNCH_T2(V, RR, TH, DP)                                                                                 \
    do {                                                                                                         \
        if (g.fold && ov.split) LAUNCH_T3(V, RR, TH, DP, true, true);                                            \
        else if (g.fold) LAUNCH_T3(V, RR, TH, DP, true, false);                                                  \
        else if (ov.split) LAUNCH_T3(V, RR, TH, DP, false, true);                                                \
        else LAUNCH_T3(V, RR, TH, DP, false, false);                                                             \
    } while (0)
    const int depth = t_depth(p.threads / 32);
#define LAUNCH_T(V, RR)                                                                                          \
    do {                                                                                                         \
        if (p.threads == 1024) LAUNCH_T2(V, RR, 1024, 1);                                                        \
        else if (p.threads == 512 && depth == 4) LAUNCH_T2(V, RR, 512, 4);                                       \
        else if (p.threads == 512) LAUNCH_T2(V, RR, 512, 3);                                                     \
        else if (p.threads == 640 && depth == 3) LAUNCH_T2(V, RR, 640, 3);                                       \
        else if (p.threads == 640) LAUNCH_T2(V, RR, 640, 2);                                                     \
        else LAUNCH_T2(V, RR, 768, 2);                                                                           \
    } while (0)
    if (p.vec == 4 && r == 1) LAUNCH_T(4, 1);
    else if (p.vec == 4 && r == 2) LAUNCH_T(4, 2);
    else if (p.vec == 2 && r == 1) LAUNCH_T(2, 1);
    else if (p.vec == 2 && r == 2) LAUNCH_T(2, 2);
    else if (p.vec == 1 && r == 1) LAUNCH_T(1, 1);
    else return (int)cudaErrorInvalidValue;
#undef LAUNCH_T2
#undef LAUNCH_T3
#undef LAUNCH_T
    SPH3D_CHECK_LAUNCH();
    int rc = launch_reduce_partials((int)w.P, (size_t)F * Co, part, grad_filter, st);
    if (rc) return rc;
    *launches += g.fold ? 2 : 3;
    return 0;
}

// used by conv_bwd.cu to route sph3d_depthwise_conv3d_grad
bool bwd_transposed_supported(int B, int N, int M, int F, int C, int r, int K)
{
    // SPH3D_BWD_ALGO: unset = auto, 1 = row-owned form (conv_bwd.cu) everywhere, 2 = transposed form wherever it applies.
    // Auto: the transposed form gathers C*r floats per edge where the row-owned form gathers C and reduces C, so it
    // wins for r = 1 (measured) and is left to the planned entry points for r = 2.
    const int algo = tunables().bwd_algo;
    if (algo == 1) return false;
    if (algo != 2 && r != 1) return false;
    TPlanMain p;
    return t_plan_main(B, N, M, F, C, r, t_geom(B, N, M, F, K, tunables().bwdt_fold == 1), &p);
}

size_t bwd_transposed_workspace_bytes(int B, int N, int M, int F, int C, int r, int K)
{
    const TGeom g = t_geom(B, N, M, F, K, tunables().bwdt_fold == 1);
    TPlanMain p;
    if (!t_plan_main(B, N, M, F, C, r, g, &p)) return 0;
    return g.total + t_work(B, M, F, C, r, g, p).total;
}

int bwd_transposed_run(int B, int N, int M, int F, int C, int r, int K, const int* nn_index, const int* nn_count,
                       const int* bin_index, const float* input, const float* filter, const float* grad_output,
                       float* grad_input, float* grad_filter, void* workspace, size_t workspace_bytes, cudaStream_t st)
{
    const TGeom g = t_geom(B, N, M, F, K, tunables().bwdt_fold == 1);
    TPlanMain p;
    if (!t_plan_main(B, N, M, F, C, r, g, &p)) return (int)cudaErrorInvalidValue;
    if (!workspace || workspace_bytes < g.total + t_work(B, M, F, C, r, g, p).total) return (int)cudaErrorInvalidValue;
    char* ws = reinterpret_cast<char*>(workspace);
    int launches = 0;
    int rc = t_build_plan(B, N, M, F, K, g, nn_index, nn_count, bin_index, ws, st, &launches);
    if (rc) return rc;
    rc = t_run_main(B, N, M, F, C, r, K, g, p, ws, nn_count, input, filter, grad_output, grad_input, grad_filter,
                    ws + g.total, st, &launches);
    g_last_launch_count = launches;
    return rc;
}

// ---------------------------------------------------------------- gather form of the pool / unpool gradients
// avg-pool, mean- and weighted-interpolate backward are the same sum without bins:
//      grad_input[b, s, :] = sum_{rows m, slots k : nn[b,m,k] = s} grad_output[b, m, :] * scale(m, k)
// with scale = 1/nn_count[b,m] (avg / mean) or weight[b,m,k] (weighted).  The reference scatters with one float atomic
// per edge and channel (tf_pool3d_gpu.cu:73-90, tf_unpool3d_gpu.cu:25-42, :66-84).  Here the graph is transposed with the
// same three kernels as the convolution's (one bin, one class), then one warp per SOURCE point gathers the rows that
// reference it: register sums, plain stores, no atomics, no zero fill.
template <int VEC, bool WEIGHTED>
__global__ void __launch_bounds__(256)
t_gather_kernel(unsigned rows /* B*S */, int sh, int sb, unsigned cmask, int C, int K, const int* __restrict__ seg,
                const unsigned* __restrict__ entries, const float* __restrict__ weight,
                const float* __restrict__ grad_output, float* __restrict__ grad_input)
{
    const int lane = threadIdx.x & 31;
    const int c0 = blockIdx.y * 32 * VEC + lane * VEC;
    const bool active = c0 < C;
    const int c0ld = active ? c0 : 0;
    const unsigned strideB = (unsigned)C * 4u;
    const char* gb = reinterpret_cast<const char*>(grad_output) + (size_t)c0ld * 4;
    const unsigned nw = gridDim.x * (blockDim.x >> 5);
    for (unsigned row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += nw) {
        const int beg = __ldg(seg + row), end = __ldg(seg + row + 1);
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[v] = 0.f;
        for (int kt = beg; kt < end; kt += 32) {
            const int nk = min(32, end - kt);
            unsigned myo = 0; float mys = 0.f;
            if (lane < nk) {
                const unsigned e = __ldg(entries + kt + lane);
                const unsigned r = e >> sh, code = (e >> sb) & cmask;
                myo = r * strideB;
                mys = WEIGHTED ? __ldg(weight + (size_t)r * K + code) : 1.0f / (float)(code + 1u);
            }
            for (int kk = 0; kk < nk; kk += 8) {                    // eight row gathers in flight
                unsigned o[8]; float sc[8]; float x[8][VEC];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int src = min(kk + u, nk - 1);
                    o[u] = __shfl_sync(FULL_MASK, myo, src);
                    sc[u] = (kk + u < nk) ? __shfl_sync(FULL_MASK, mys, src) : 0.f;   // tail slots re-read the last row with scale 0
                }
#pragma unroll
                for (int u = 0; u < 8; u++) ld_strip<VEC>(x[u], gb, o[u]);
#pragma unroll
                for (int u = 0; u < 8; u++)
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[v] = fmaf(x[u][v], sc[u], acc[v]);
            }
        }
        if (active) VecIO<VEC>::st(grad_input + (size_t)row * C + c0, acc);
    }
}

// ---- streaming form of the same gather: four points per warp, 32 channels per CTA column, points in degree order ----
// The warp-per-point kernel above is bound by L2: consecutive source points share no rows, every 512-byte gather
// misses L1 (10.5 GB through L2 at the Cfg-T unpool shape), and a point referenced by thousands of rows (the reference's
// growing-radius ball query makes the low-index points of every cloud such hubs) keeps one warp busy long after the
// others have finished.  Here
//   * a CTA column handles 32 channels (blockIdx.y), so a gathered row piece is ONE 128-byte line and L1 holds four times
//     as many rows;
//   * a warp is four groups of eight lanes, each group streaming the list of its own point: one LDG.128 per lane fetches
//     four different rows (four full lines), so the instruction count per edge is a quarter of the warp-per-point form;
//   * the points of a cloud are taken in descending order of in-degree (degree_order_kernel: a counting sort into 124
//     degree classes), 128 consecutive points per CTA round: the four lists of a warp have lengths within 25 % of each
//     other (little lock-step padding), and the 128 lists of a round -- entries roughly ascending in the referencing row,
//     the order in which the transposition ranked them -- sweep the same rows at the same time, which is what makes the
//     gathers hit L1 (simulated 77-80 % with a 200 KB L1).
// Entries are staged per group in shared memory (32 at a time: byte offset + scale), read back four steps at a time.
// CTA shape: TQ_WARPS warps x DEPTH steps in flight per warp.  Every step of a warp is one LDG.128 (four 128-byte lines);
// with ~60 % L1 hits almost every batch of DEPTH steps waits for an L2 round trip.  Measured at the Cfg-T unpool shape
// (profiles/r2_pool_stream.json): the gather itself 0.9-1.0 ms at 32 x 8 and at 16 x 16 (L1 hit 61 %, L1 data pipe 51 %,
// 380 M warp instructions) against 1.5 ms for the warp-per-point form; the whole gradient 1.35 ms against 1.96 ms for the
// scatter form.  (A first 16 x 16 measurement of 0.35 ms was a launch-shape bug that skipped half of the work items.)

// Work items.  A point referenced by thousands of rows would keep one lane group busy for thousands of steps while the rest
// of the machine has finished (measured: 1.23 ms, nothing saturated), so a list is cut into PARTS of TQ_PART entries and the
// parts of all points are the work items; partial sums of multi-part points meet in grad_input through vector reductions
// (grad_input is zero-filled first).  Item order: part index major, points by descending degree inside a part index --
// because the points are sorted by their number of parts, "the points that have a part j" are a PREFIX of the sorted
// order, so item q resolves to (part j, point order[q - cum[j]]) from the small table cum[] alone.  A round of 128 items is
// then 128 different points' j-th parts: equal lengths (no lock-step padding) over the same range of referencing rows.
constexpr int TQ_PART = 256;
constexpr int TQ_KEYS = 512;                                       // 32 degree classes for single-part points + parts counts

// ascending in degree; single-part points: 0..31 (degree classes), others 32 + number of parts, at most pmax_cap parts
// (a longer list -- only possible when rows repeat a neighbour -- puts its surplus into the last part)
__device__ __forceinline__ int degree_key(int deg, int pmax_cap)
{
    if (deg <= TQ_PART) {
        if (deg < 8) return deg;
        const int e = 31 - __clz(deg);
        return min(31, 8 + (e - 3) * 4 + ((deg >> (e - 2)) & 3));
    }
    return 32 + min(pmax_cap, (deg + TQ_PART - 1) / TQ_PART);      // pmax_cap <= TQ_KEYS - 33
}

// per cloud (one CTA): order[] = points by descending degree_key (counting sort), cum[j] = number of items whose part index
// is < j (cum[0] = 0, cum[pmax] = all items), written as cum[b*(pmax_cap+1) + j]
__global__ void __launch_bounds__(1024)
degree_order_kernel(int S, int pmax_cap, const int* __restrict__ seg, int* __restrict__ order, int* __restrict__ cum)
{
    __shared__ int hist[TQ_KEYS];
    __shared__ int start[TQ_KEYS];
    const int b = blockIdx.x;
    const int* sg = seg + (size_t)b * S;
    for (int i = threadIdx.x; i < TQ_KEYS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) atomicAdd(&hist[degree_key(sg[i + 1] - sg[i], pmax_cap)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {                                        // descending key order; 512 steps, once per cloud
        int run = 0;
        for (int k = TQ_KEYS - 1; k >= 0; k--) { start[k] = run; run += hist[k]; }
        // points with more than j parts: keys >= 32 + (j + 1) for j >= 1, everything for j = 0
        int* cm = cum + (size_t)b * (pmax_cap + 1);
        int acc = 0;
        cm[0] = 0;
        for (int j = 0; j < pmax_cap; j++) {
            acc += (j == 0) ? S : start[32 + j];                   // points with more than j parts <=> key > 32 + j: sorted before it
            cm[j + 1] = acc;
        }
    }
    __syncthreads();
    int* ord = order + (size_t)b * S;
    for (int i = threadIdx.x; i < S; i += blockDim.x) ord[atomicAdd(&start[degree_key(sg[i + 1] - sg[i], pmax_cap)], 1)] = i;
}

template <bool WEIGHTED, int TQ_WARPS, int DEPTH>
__global__ void __launch_bounds__(TQ_WARPS * 32, 1)
tq_gather_kernel(int B, int S, int C, int K, int sh, int sb, unsigned cmask, int pmax_cap, int rounds,
                 const int* __restrict__ seg, const unsigned* __restrict__ entries, const int* __restrict__ order,
                 const int* __restrict__ cum, const float* __restrict__ weight,
                 const float* __restrict__ grad_output, float* __restrict__ grad_input)
{
    __shared__ __align__(16) unsigned sOff[TQ_WARPS][4][32];
    __shared__ __align__(16) float sScale[TQ_WARPS][4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 3, l8 = lane & 7;
    const int c0 = blockIdx.y * 32 + l8 * 4;
    const bool chan_ok = c0 < C;                                   // C % 4 == 0 (host-checked)
    const unsigned strideB = (unsigned)C * 4u;
    const char* gb = reinterpret_cast<const char*>(grad_output) + (size_t)(chan_ok ? c0 : 0) * 4;
    const int tiles = B * rounds;
    unsigned* myOff = sOff[warp][grp];
    float* myScale = sScale[warp][grp];
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int b = t / rounds, r = t - b * rounds;
        const int* cm = cum + (size_t)b * (pmax_cap + 1);
        const int total = __ldg(cm + pmax_cap);
        if (r * (TQ_WARPS * 4) >= total) continue;                 // (CTA-uniform) this cloud has fewer rounds
        const int q = r * (TQ_WARPS * 4) + warp * 4 + grp;        // my work item
        int pt = -1, beg = 0, len = 0;
        if (q < total) {
            int lo = 0, hi = pmax_cap;                             // largest j with cum[j] <= q
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(cm + mid) <= q) lo = mid; else hi = mid;
            }
            pt = __ldg(order + (size_t)b * S + (q - __ldg(cm + lo)));
            const int pbeg = __ldg(seg + (size_t)b * S + pt), pend = __ldg(seg + (size_t)b * S + pt + 1);
            beg = pbeg + lo * TQ_PART;
            len = (lo == pmax_cap - 1) ? pend - beg : min(TQ_PART, pend - beg);     // the last possible part takes any surplus
        }
        int steps = len;                                           // the warp walks max(len) steps; short groups pad with scale 0
        steps = max(steps, __shfl_xor_sync(FULL_MASK, steps, 8));
        steps = max(steps, __shfl_xor_sync(FULL_MASK, steps, 16));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = 0; base < steps; base += 32) {
            // stage this group's next 32 entries: lane l8 converts entries base + 4*l8 .. + 3
            unsigned o4[4]; float s4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = base + l8 * 4 + u;
                o4[u] = 0u; s4[u] = 0.f;
                if (e < len) {
                    const unsigned ent = __ldg(entries + beg + e);
                    const unsigned row = ent >> sh, code = (ent >> sb) & cmask;
                    o4[u] = row * strideB;
                    s4[u] = WEIGHTED ? __ldg(weight + (size_t)row * K + code) : 1.0f / (float)(code + 1u);
                }
            }
            *reinterpret_cast<uint4*>(myOff + l8 * 4) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            *reinterpret_cast<float4*>(myScale + l8 * 4) = make_float4(s4[0], s4[1], s4[2], s4[3]);
            __syncwarp();
            const int n = min(32, steps - base);
            for (int qq = 0; qq < n; qq += DEPTH) {                // DEPTH steps = DEPTH rows per group in flight (pads: scale 0)
                float4 v[DEPTH];
                float sc[DEPTH];
#pragma unroll
                for (int d4 = 0; d4 < DEPTH; d4 += 4) {
                    const uint4 oo = *reinterpret_cast<const uint4*>(myOff + qq + d4);
                    const float4 ss = *reinterpret_cast<const float4*>(myScale + qq + d4);
                    v[d4] = __ldg(reinterpret_cast<const float4*>(gb + oo.x));
                    v[d4 + 1] = __ldg(reinterpret_cast<const float4*>(gb + oo.y));
                    v[d4 + 2] = __ldg(reinterpret_cast<const float4*>(gb + oo.z));
                    v[d4 + 3] = __ldg(reinterpret_cast<const float4*>(gb + oo.w));
                    sc[d4] = ss.x; sc[d4 + 1] = ss.y; sc[d4 + 2] = ss.z; sc[d4 + 3] = ss.w;
                }
#pragma unroll
                for (int d = 0; d < DEPTH; d += 2) {               // two partial sums per pair of steps shorten the FMA chain
                    float4 a;
                    a.x = v[d].x * sc[d]; a.y = v[d].y * sc[d]; a.z = v[d].z * sc[d]; a.w = v[d].w * sc[d];
                    a.x = fmaf(v[d + 1].x, sc[d + 1], a.x); a.y = fmaf(v[d + 1].y, sc[d + 1], a.y);
                    a.z = fmaf(v[d + 1].z, sc[d + 1], a.z); a.w = fmaf(v[d + 1].w, sc[d + 1], a.w);
                    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
                }
            }
            __syncwarp();                                          // the staging arrays are rewritten by the next refill
        }
        if (pt >= 0 && chan_ok && len > 0)                         // parts of one point meet in the zero-filled grad_input
            red_add_v4(grad_input + ((size_t)b * S + pt) * C + c0, acc.x, acc.y, acc.z, acc.w);
        __syncthreads();                                           // the CTA's warps enter the next round together
    }
}

// scratch of the streaming form behind the plan: order[B*S] | cum[B*(pmax_cap+1)]
struct TqGeom { int pmax_cap; long long max_items; size_t cum_off, bytes; };
static TqGeom tq_geom(int B, int S, int R, int K)
{
    TqGeom q{};
    long long pm = ((long long)R + TQ_PART - 1) / TQ_PART;         // a point is referenced at most once per row (ball queries)
    if (pm < 1) pm = 1;
    if (pm > TQ_KEYS - 33) pm = TQ_KEYS - 33;
    q.pmax_cap = (int)pm;
    q.max_items = (long long)S + ((long long)R * K + TQ_PART - 1) / TQ_PART;       // one item per point + one per full part
    q.cum_off = align256((size_t)B * S * sizeof(int));
    q.bytes = q.cum_off + align256((size_t)B * (q.pmax_cap + 1) * sizeof(int));
    return q;
}

// S = source points per cloud (rows of grad_input), R = referencing rows per cloud (rows of grad_output)
static TGeom pool_geom(int B, int S, int R, int K)
{
    TGeom g = t_geom(B, S, R, 1, K, true);
    if (g.ok && !g.fold) g.ok = false;                              // the scale code must ride in the entry
    if (g.ok && ((long long)B * R * 4 >= (1LL << 30))) g.ok = false;
    return g;
}

size_t pool_scatter_workspace_bytes(int B, int S, int R, int C, int K)
{
    if (B <= 0 || S <= 0 || R <= 0 || C <= 0 || K <= 0) return 0;
    if ((long long)B * R * C * 4 >= (1LL << 32)) return 0;          // 32-bit byte offsets into grad_output
    const TGeom g = pool_geom(B, S, R, K);
    return g.ok ? g.total + tq_geom(B, S, R, K).bytes : 0;         // plan + degree order and item table of the streaming form
}

int pool_scatter_run(int B, int S, int R, int C, int K, const int* nn_index, const int* nn_count, const float* weight,
                     const float* grad_output, float* grad_input, void* workspace, size_t workspace_bytes, cudaStream_t st)
{
    const TGeom g = pool_geom(B, S, R, K);
    if (!g.ok || !workspace || workspace_bytes < g.total + tq_geom(B, S, R, K).bytes) return (int)cudaErrorInvalidValue;
    char* plan = reinterpret_cast<char*>(workspace);
    int launches = 0;
    int rc = t_build_plan(B, S, R, 1, K, g, nn_index, nn_count, nullptr, plan, st, &launches, weight ? 1 : 0);
    if (rc) return rc;
    const int* seg = reinterpret_cast<const int*>(plan + g.seg_off);
    const unsigned* ent = reinterpret_cast<const unsigned*>(plan + g.ent_off);
    if (C % 4 == 0 && tunables().pool_stream != 0) {               // streaming form (SPH3D_POOL_STREAM=0: warp-per-point form)
        const TqGeom q = tq_geom(B, S, R, K);
        int* order = reinterpret_cast<int*>(plan + g.total);
        int* cum = reinterpret_cast<int*>(plan + g.total + q.cum_off);
        cudaError_t e = cudaMemsetAsync(grad_input, 0, sizeof(float) * (size_t)B * S * C, st);
        if (e != cudaSuccess) return (int)e;
        degree_order_kernel<<<B, 1024, 0, st>>>(S, q.pmax_cap, seg, order, cum);
        SPH3D_CHECK_LAUNCH();
        const int shape = tunables().pool_stream > 0 ? tunables().pool_stream : 328;     // 324 / 328: 32 warps x 4 / 8 steps in flight; 168 / 1616: 16 warps x 8 / 16
        const int tq_warps = (shape == 168 || shape == 1616) ? 16 : 32;
        const int rounds = (int)((q.max_items + tq_warps * 4 - 1) / (tq_warps * 4));
        const long long tiles = (long long)B * rounds;
        const int chunks = (C + 31) / 32;
        long long want = sm_count();
        dim3 grid((unsigned)(tiles < want ? tiles : want), (unsigned)chunks);
        const int sh = g.sb + g.cb;
        const unsigned cmask = (1u << g.cb) - 1u;
#define LAUNCH_TQ(WW, DD)                                                                                                          \
    do {                                                                                                                           \
        if (weight) tq_gather_kernel<true, WW, DD><<<grid, WW * 32, 0, st>>>(B, S, C, K, sh, g.sb, cmask, q.pmax_cap, rounds, seg, ent, order, cum, weight, grad_output, grad_input); \
        else tq_gather_kernel<false, WW, DD><<<grid, WW * 32, 0, st>>>(B, S, C, K, sh, g.sb, cmask, q.pmax_cap, rounds, seg, ent, order, cum, weight, grad_output, grad_input);      \
    } while (0)
        if (shape == 324) LAUNCH_TQ(32, 4);
        else if (shape == 1616) LAUNCH_TQ(16, 16);
        else if (shape == 168) LAUNCH_TQ(16, 8);
        else LAUNCH_TQ(32, 8);
#undef LAUNCH_TQ
        SPH3D_CHECK_LAUNCH();
        g_last_launch_count = launches + 2;
        return 0;
    }
    int vec = pick_vec_full_warp(C);
    const int chunks = (C + 32 * vec - 1) / (32 * vec);
    const unsigned rows = (unsigned)((long long)B * S);
    long long want = (long long)sm_count() * 8 / chunks, tiles = ((long long)rows + 7) / 8;
    if (want < 1) want = 1;
    dim3 grid((unsigned)(tiles < want ? tiles : want), (unsigned)chunks);
    const int sh = g.sb + g.cb;
    const unsigned cmask = (1u << g.cb) - 1u;
#define LAUNCH_TG(V)                                                                                                   \
    do {                                                                                                               \
        if (weight) t_gather_kernel<V, true><<<grid, 256, 0, st>>>(rows, sh, g.sb, cmask, C, K, seg, ent, weight, grad_output, grad_input); \
        else t_gather_kernel<V, false><<<grid, 256, 0, st>>>(rows, sh, g.sb, cmask, C, K, seg, ent, weight, grad_output, grad_input);      \
    } while (0)
    if (vec == 4) LAUNCH_TG(4); else if (vec == 2) LAUNCH_TG(2); else LAUNCH_TG(1);
#undef LAUNCH_TG
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = launches + 1;
    return 0;
}

}  // namespace sph3d

using namespace sph3d;

extern "C" size_t sph3d_conv_transpose_bytes(int B, int N, int M, int F, int K)
{
    const TGeom g = t_geom(B, N, M, F, K, tunables().bwdt_fold == 1);
    return g.ok ? g.total : 0;
}

extern "C" int sph3d_conv_transpose(int B, int N, int M, int F, int K, const int* nn_index, const int* nn_count,
                                    const int* bin_index, void* plan, size_t plan_bytes, void* stream)
{
    g_last_launch_count = 0;
    const TGeom g = t_geom(B, N, M, F, K, tunables().bwdt_fold == 1);
    if (!g.ok || !nn_index || !nn_count || !bin_index || !plan || plan_bytes < g.total) return (int)cudaErrorInvalidValue;
    int launches = 0;
    int rc = t_build_plan(B, N, M, F, K, g, nn_index, nn_count, bin_index, reinterpret_cast<char*>(plan),
                          (cudaStream_t)stream, &launches);
    g_last_launch_count = launches;
    return rc;
}

extern "C" size_t sph3d_depthwise_conv3d_grad_planned_workspace_bytes(int B, int N, int M, int F, int C, int r, int K)
{
    const TGeom g = t_geom(B, N, M, F, K, tunables().bwdt_fold == 1);
    TPlanMain p;
    if (!t_plan_main(B, N, M, F, C, r, g, &p)) return 0;
    return t_work(B, M, F, C, r, g, p).total;
}

extern "C" int sph3d_depthwise_conv3d_grad_planned(int B, int N, int M, int F, int C, int r, int K, const int* nn_count,
                                                   const void* plan, size_t plan_bytes, const float* input,
                                                   const float* filter, const float* grad_output, float* grad_input,
                                                   float* grad_filter, void* workspace, size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    const TGeom g = t_geom(B, N, M, F, K, tunables().bwdt_fold == 1);
    TPlanMain p;
    if (!t_plan_main(B, N, M, F, C, r, g, &p)) return (int)cudaErrorInvalidValue;
    if (!nn_count || !plan || plan_bytes < g.total || !input || !filter || !grad_output || !grad_input || !grad_filter ||
        !workspace || workspace_bytes < t_work(B, M, F, C, r, g, p).total)
        return (int)cudaErrorInvalidValue;
    int launches = 0;
    int rc = t_run_main(B, N, M, F, C, r, K, g, p, reinterpret_cast<const char*>(plan), nn_count, input, filter, grad_output,
                        grad_input, grad_filter, reinterpret_cast<char*>(workspace), (cudaStream_t)stream, &launches);
    g_last_launch_count = launches;
    return rc;
}
