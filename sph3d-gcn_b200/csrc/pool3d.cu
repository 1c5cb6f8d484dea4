// pool3d.cu -- graph max/avg pooling and mean/weighted interpolation (unpooling), sm_100a.
//
// Replaces maxPool3dLauncher, maxPool3dGradLauncher, avgPool3dLauncher, avgPool3dGradLauncher
// (/root/reference/tf_ops/pooling/tf_pool3d_gpu.cu:93-119; kernels :5-90), meanInterpolateLauncher,
// meanInterpolateGradLauncher, weightedInterpolateLauncher, weightedInterpolateGradLauncher
// (/root/reference/tf_ops/unpooling/tf_unpool3d_gpu.cu:87-113; kernels :5-84) and the cudaMemset
// zero fills in tf_pool3d.cpp / tf_unpool3d.cpp.  avg-pool and mean-interpolate are the same
// arithmetic (the glue only swaps which tensor is called N and which M), so they share kernels.
//
// Same work unit as the convolution (rowwarp.cuh): a warp owns one output point and 32*VEC
// channels; neighbour ids are read once per warp (coalesced) and broadcast by shuffle, every
// feature-row gather is one coalesced 128*VEC-byte warp load, four gathers are kept in flight,
// the reduction lives in registers (the reference read-modify-writes global memory per edge).
// Backward passes scatter with 16-byte vector reductions (REDG.ADD.F32x4).
//
// max-pool keeps the reference's selection rule exactly (Q11): neighbour 0 initialises, a later
// neighbour replaces only if strictly greater, max_index is the database point id.
#include "conv_common.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

constexpr int POOL_WARPS = 8;
enum GatherOp { OP_MAX = 0, OP_MEAN = 1, OP_WEIGHTED = 2 };

template <int VEC, int OP>
__global__ void __launch_bounds__(POOL_WARPS * 32)
row_gather_kernel(int B, int S, int Rr, int C, int K,      // S = source points, Rr = output rows per cloud
                  const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                  const float* __restrict__ weight, const float* __restrict__ input,
                  float* __restrict__ output, int* __restrict__ max_index)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.y * 32 * VEC + lane * VEC;
    const bool active = c0 < C;
    const long long rows = (long long)B * Rr;
    for (long long row = (long long)blockIdx.x * POOL_WARPS + warp; row < rows;
         row += (long long)gridDim.x * POOL_WARPS) {
        const int b = (int)(row / Rr);
        const int cnt = min(__ldg(nn_count + row), K);
        const float* inb = input + (size_t)b * S * C + c0;
        const int* idxrow = nn_index + (size_t)row * K;
        float acc[VEC];
        int arg[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) { acc[v] = 0.f; arg[v] = 0; }
        for (int kt = 0; kt < cnt; kt += 32) {
            const int k = kt + lane;
            int myi = 0; float myw = 0.f;
            if (k < cnt) {
                myi = __ldg(idxrow + k);
                if (OP == OP_WEIGHTED) myw = __ldg(weight + (size_t)row * K + k);
            }
            const int nk = min(32, cnt - kt);
            for (int kk = 0; kk < nk; kk += 4) {
                int n[4]; float w[4]; float x[4][VEC];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int src = min(kk + u, nk - 1);
                    n[u] = __shfl_sync(FULL_MASK, myi, src);
                    if (OP == OP_WEIGHTED) w[u] = __shfl_sync(FULL_MASK, myw, src);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) VecIO<VEC>::ld(x[u], inb + (size_t)n[u] * C, active && (kk + u < nk));
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (kk + u < nk) {                                 // warp-uniform
#pragma unroll
                        for (int v = 0; v < VEC; v++) {
                            if (OP == OP_MAX) {
                                bool take = (kt + kk + u == 0) || (x[u][v] > acc[v]);
                                if (take) { acc[v] = x[u][v]; arg[v] = n[u]; }
                            } else if (OP == OP_MEAN) {
                                acc[v] += x[u][v];
                            } else {
                                acc[v] = fmaf(x[u][v], w[u], acc[v]);   // the reference's contracted FMA, in k order
                            }
                        }
                    }
                }
            }
        }
        if (active) {
            if (OP == OP_MEAN) {
                const float inv = cnt > 0 ? 1.0f / (float)cnt : 0.f;
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[v] *= inv;
            }
            VecIO<VEC>::st(output + (size_t)row * C + c0, acc);
            if (OP == OP_MAX) VecIO<VEC>::sti(max_index + (size_t)row * C + c0, arg);
        }
    }
}

// grad_input[b, nn[row,k], c] += grad_output[row, c] * (1/cnt | weight[row,k])
template <int VEC, int OP>
__global__ void __launch_bounds__(POOL_WARPS * 32)
row_scatter_kernel(int B, int S, int Rr, int C, int K,
                   const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                   const float* __restrict__ weight, const float* __restrict__ grad_output,
                   float* __restrict__ grad_input)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.y * 32 * VEC + lane * VEC;
    const bool active = c0 < C;
    const long long rows = (long long)B * Rr;
    for (long long row = (long long)blockIdx.x * POOL_WARPS + warp; row < rows;
         row += (long long)gridDim.x * POOL_WARPS) {
        const int b = (int)(row / Rr);
        const int cnt = min(__ldg(nn_count + row), K);
        if (cnt <= 0) continue;
        float* gib = grad_input + (size_t)b * S * C + c0;
        const int* idxrow = nn_index + (size_t)row * K;
        float g[VEC];
        VecIO<VEC>::ld(g, grad_output + (size_t)row * C + c0, active);
        if (OP == OP_MEAN) {
            // reference: atomicAdd(gradInput, gradOutput/nnSize) -- one IEEE division per element
#pragma unroll
            for (int v = 0; v < VEC; v++) g[v] = __fdiv_rn(g[v], (float)cnt);
        }
        for (int kt = 0; kt < cnt; kt += 32) {
            const int k = kt + lane;
            int myi = 0; float myw = 0.f;
            if (k < cnt) {
                myi = __ldg(idxrow + k);
                if (OP == OP_WEIGHTED) myw = __ldg(weight + (size_t)row * K + k);
            }
            const int nk = min(32, cnt - kt);
            for (int kk = 0; kk < nk; kk++) {
                int n = __shfl_sync(FULL_MASK, myi, kk);
                float d[VEC];
                if (OP == OP_WEIGHTED) {
                    float w = __shfl_sync(FULL_MASK, myw, kk);
#pragma unroll
                    for (int v = 0; v < VEC; v++) d[v] = __fmul_rn(g[v], w);
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; v++) d[v] = g[v];
                }
                if (active) VecIO<VEC>::red(gib + (size_t)n * C, d);
            }
        }
    }
}

// max-pool backward: grad_input[b, max_index[b,m,c], c] += grad_output[b,m,c]
// (tf_pool3d_gpu.cu:38-50: no nn_count test -- rows with no neighbour hit point 0, as there).
__global__ void __launch_bounds__(256)
max_pool_grad_kernel(int B, int N, int M, int C, const int* __restrict__ max_index,
                     const float* __restrict__ grad_output, float* __restrict__ grad_input)
{
    const size_t total = (size_t)B * M * C;
    const size_t per = (size_t)M * C;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(t / per);
        int c = (int)(t % C);
        int n = __ldg(max_index + t);
        atomicAdd(grad_input + ((size_t)b * N + n) * C + c, __ldg(grad_output + t));
    }
}

static dim3 row_grid(int B, int rows_per_cloud, int C, int vec)
{
    int chunks = (C + 32 * vec - 1) / (32 * vec);
    long long tiles = ((long long)B * rows_per_cloud + POOL_WARPS - 1) / POOL_WARPS;
    long long want = (long long)sm_count() * 8 / chunks;
    if (want < 1) want = 1;
    return dim3((unsigned)(tiles < want ? tiles : want), (unsigned)chunks);
}

template <int OP>
static int launch_gather(int B, int S, int Rr, int C, int K, const int* nn_index, const int* nn_count,
                         const float* weight, const float* input, float* output, int* max_index, cudaStream_t st)
{
    int vec = pick_vec_full_warp(C);
    dim3 grid = row_grid(B, Rr, C, vec);
    if (vec == 4) row_gather_kernel<4, OP><<<grid, POOL_WARPS * 32, 0, st>>>(B, S, Rr, C, K, nn_index, nn_count, weight, input, output, max_index);
    else if (vec == 2) row_gather_kernel<2, OP><<<grid, POOL_WARPS * 32, 0, st>>>(B, S, Rr, C, K, nn_index, nn_count, weight, input, output, max_index);
    else row_gather_kernel<1, OP><<<grid, POOL_WARPS * 32, 0, st>>>(B, S, Rr, C, K, nn_index, nn_count, weight, input, output, max_index);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}

template <int OP>
static int launch_scatter(int B, int S, int Rr, int C, int K, const int* nn_index, const int* nn_count,
                          const float* weight, const float* grad_output, float* grad_input, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(grad_input, 0, sizeof(float) * (size_t)B * S * C, st);
    if (e != cudaSuccess) return (int)e;
    int vec = pick_vec_full_warp(C);
    dim3 grid = row_grid(B, Rr, C, vec);
    if (vec == 4) row_scatter_kernel<4, OP><<<grid, POOL_WARPS * 32, 0, st>>>(B, S, Rr, C, K, nn_index, nn_count, weight, grad_output, grad_input);
    else if (vec == 2) row_scatter_kernel<2, OP><<<grid, POOL_WARPS * 32, 0, st>>>(B, S, Rr, C, K, nn_index, nn_count, weight, grad_output, grad_input);
    else row_scatter_kernel<1, OP><<<grid, POOL_WARPS * 32, 0, st>>>(B, S, Rr, C, K, nn_index, nn_count, weight, grad_output, grad_input);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}

}  // namespace sph3d

using namespace sph3d;

#define BAD5(B, N, M, C, K) ((B) <= 0 || (N) <= 0 || (M) <= 0 || (C) <= 0 || (K) <= 0)

extern "C" int sph3d_max_pool3d(int B, int N, int M, int C, int K, const int* nn_index, const int* nn_count,
                                const float* input, float* output, int* max_index, void* stream)
{
    g_last_launch_count = 0;
    if (BAD5(B, N, M, C, K) || !nn_index || !nn_count || !input || !output || !max_index) return (int)cudaErrorInvalidValue;
    return launch_gather<OP_MAX>(B, N, M, C, K, nn_index, nn_count, nullptr, input, output, max_index, (cudaStream_t)stream);
}

extern "C" int sph3d_max_pool3d_grad(int B, int N, int M, int C, const int* max_index,
                                     const float* grad_output, float* grad_input, void* stream)
{
    g_last_launch_count = 0;
    if (BAD5(B, N, M, C, 1) || !max_index || !grad_output || !grad_input) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_input, 0, sizeof(float) * (size_t)B * N * C, st);
    if (e != cudaSuccess) return (int)e;
    size_t total = (size_t)B * M * C, want = (total + 255) / 256, cap = (size_t)sm_count() * 16;
    max_pool_grad_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(B, N, M, C, max_index, grad_output, grad_input);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}

extern "C" int sph3d_avg_pool3d(int B, int N, int M, int C, int K, const int* nn_index, const int* nn_count,
                                const float* input, float* output, void* stream)
{
    g_last_launch_count = 0;
    if (BAD5(B, N, M, C, K) || !nn_index || !nn_count || !input || !output) return (int)cudaErrorInvalidValue;
    return launch_gather<OP_MEAN>(B, N, M, C, K, nn_index, nn_count, nullptr, input, output, nullptr, (cudaStream_t)stream);
}

extern "C" size_t sph3d_avg_pool3d_grad_workspace_bytes(int B, int N, int M, int C, int K)
{
    return pool_scatter_workspace_bytes(B, N, M, C, K);
}
extern "C" size_t sph3d_interpolate_grad_workspace_bytes(int B, int N, int M, int C, int K)
{
    return pool_scatter_workspace_bytes(B, M, N, C, K);
}

extern "C" int sph3d_avg_pool3d_grad(int B, int N, int M, int C, int K, const int* nn_index, const int* nn_count,
                                     const float* grad_output, float* grad_input,
                                     void* workspace, size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    if (BAD5(B, N, M, C, K) || !nn_index || !nn_count || !grad_output || !grad_input) return (int)cudaErrorInvalidValue;
    if (workspace && workspace_bytes && pool_scatter_workspace_bytes(B, N, M, C, K))
        return pool_scatter_run(B, N, M, C, K, nn_index, nn_count, nullptr, grad_output, grad_input, workspace,
                                workspace_bytes, (cudaStream_t)stream);
    return launch_scatter<OP_MEAN>(B, N, M, C, K, nn_index, nn_count, nullptr, grad_output, grad_input, (cudaStream_t)stream);
}

// unpooling: N = fine/output points (rows), M = coarse/input points (source)  (tf_unpool3d.cpp:76-80)
extern "C" int sph3d_mean_interpolate(int B, int N, int M, int C, int K, const int* nn_index, const int* nn_count,
                                      const float* input, float* output, void* stream)
{
    g_last_launch_count = 0;
    if (BAD5(B, N, M, C, K) || !nn_index || !nn_count || !input || !output) return (int)cudaErrorInvalidValue;
    return launch_gather<OP_MEAN>(B, M, N, C, K, nn_index, nn_count, nullptr, input, output, nullptr, (cudaStream_t)stream);
}

extern "C" int sph3d_mean_interpolate_grad(int B, int N, int M, int C, int K, const int* nn_index,
                                           const int* nn_count, const float* grad_output, float* grad_input,
                                           void* workspace, size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    if (BAD5(B, N, M, C, K) || !nn_index || !nn_count || !grad_output || !grad_input) return (int)cudaErrorInvalidValue;
    if (workspace && workspace_bytes && pool_scatter_workspace_bytes(B, M, N, C, K))
        return pool_scatter_run(B, M, N, C, K, nn_index, nn_count, nullptr, grad_output, grad_input, workspace,
                                workspace_bytes, (cudaStream_t)stream);
    return launch_scatter<OP_MEAN>(B, M, N, C, K, nn_index, nn_count, nullptr, grad_output, grad_input, (cudaStream_t)stream);
}

extern "C" int sph3d_weighted_interpolate(int B, int N, int M, int C, int K, const int* nn_index,
                                          const int* nn_count, const float* input, const float* weight,
                                          float* output, void* stream)
{
    g_last_launch_count = 0;
    if (BAD5(B, N, M, C, K) || !nn_index || !nn_count || !input || !weight || !output) return (int)cudaErrorInvalidValue;
    return launch_gather<OP_WEIGHTED>(B, M, N, C, K, nn_index, nn_count, weight, input, output, nullptr, (cudaStream_t)stream);
}

extern "C" int sph3d_weighted_interpolate_grad(int B, int N, int M, int C, int K, const int* nn_index,
                                               const int* nn_count, const float* grad_output, const float* weight,
                                               float* grad_input, void* workspace, size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    if (BAD5(B, N, M, C, K) || !nn_index || !nn_count || !grad_output || !weight || !grad_input) return (int)cudaErrorInvalidValue;
    if (workspace && workspace_bytes && pool_scatter_workspace_bytes(B, M, N, C, K))
        return pool_scatter_run(B, M, N, C, K, nn_index, nn_count, weight, grad_output, grad_input, workspace,
                                workspace_bytes, (cudaStream_t)stream);
    return launch_scatter<OP_WEIGHTED>(B, M, N, C, K, nn_index, nn_count, weight, grad_output, grad_input, (cudaStream_t)stream);
}
