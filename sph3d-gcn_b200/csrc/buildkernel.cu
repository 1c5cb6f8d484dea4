// buildkernel.cu -- spherical-bin index of every graph edge, sm_100a.
//
// Replaces sphericalKernelLauncher (/root/reference/tf_ops/buildkernel/tf_buildkernel_gpu.cu:83-89,
// kernel :20-79) and the zero fill of tf_buildkernel.cpp:89.  The reference runs one thread per
// QUERY (serial over its K neighbours, <=32 CTAs); here one thread owns one EDGE slot (b,m,k), so
// nn_index / nn_dist / filt_index move as fully coalesced streams and only the 12-byte neighbour
// coordinates are gathered (from an L2-resident cloud).  HBM-bound: 4*(3BN+3BM+3E+BM) bytes.
//
// Bit-exactness: the mixed fp32/fp64 expression is restated operation by
// operation with explicit round-to-nearest intrinsics (nothing for nvcc to re-contract); M_PI is
// glibc's double; atan2f is the same CUDA 12.9 libdevice routine the reference kernel inlines.
#include "common.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

__device__ __forceinline__ int spherical_bin(float dx, float dy, float dz, float dist, float radius,
                                             int n, int p, int q)
{
    const double PI = 3.14159265358979323846;
    const float EPS = 1.01e-3F;
    if (!(dist > EPS && (double)fabsf(__fsub_rn(dist, EPS)) > 1e-6)) return 0;
    float dist2d = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    float theta = atan2f(dy, dx);
    float phi = atan2f(dz, dist2d);
    theta = __double2float_rn(((double)theta < PI) ? (double)theta : -PI);
    theta = __double2float_rn(((double)theta > -PI) ? (double)theta : -PI);
    theta = __double2float_rn(__dadd_rn((double)theta, PI));
    phi = __double2float_rn(((double)phi < PI / 2) ? (double)phi : PI / 2);
    phi = __double2float_rn(((double)phi > -PI / 2) ? (double)phi : -PI / 2);
    phi = __double2float_rn(__dadd_rn((double)phi, PI / 2));
    float alpha = __double2float_rn(__ddiv_rn((double)__fdiv_rn(__fmul_rn(theta, (float)n), 2.0f), PI));
    float beta = __double2float_rn(__ddiv_rn((double)__fmul_rn(phi, (float)p), PI));
    float gamma = __fdiv_rn(__fmul_rn(dist, (float)q), __fadd_rn(radius, 1e-6F));
    int nID = min(n - 1, (int)alpha);
    int pID = min(p - 1, (int)beta);
    int qID = min(q - 1, (int)gamma);
    return qID * p * n + pID * n + nID + 1;
}

__global__ void __launch_bounds__(256)
spherical_kernel_kernel(int B, int N, int M, int K, int n, int p, int q, float radius,
                        const float* __restrict__ database, const float* __restrict__ query,
                        const int* __restrict__ nn_index, const int* __restrict__ nn_count,
                        const float* __restrict__ nn_dist, int* __restrict__ filt_index)
{
    const size_t total = (size_t)B * M * K;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        size_t row = e / K;                       // = b*M + m
        int k = (int)(e - row * K);
        int bin = 0;
        if (k < __ldg(nn_count + row)) {
            int b = (int)(row / M);
            int id = __ldg(nn_index + e);
            const float* pt = database + ((size_t)b * N + id) * 3;
            const float* qp = query + row * 3;
            float dx = __fsub_rn(__ldg(pt), __ldg(qp));
            float dy = __fsub_rn(__ldg(pt + 1), __ldg(qp + 1));
            float dz = __fsub_rn(__ldg(pt + 2), __ldg(qp + 2));
            bin = spherical_bin(dx, dy, dz, __ldg(nn_dist + e), radius, n, p, q);
        }
        filt_index[e] = bin;
    }
}

}  // namespace sph3d

using namespace sph3d;

extern "C" int sph3d_spherical_kernel(int B, int N, int M, int K, int n, int p, int q, float radius,
                                      const float* database, const float* query, const int* nn_index,
                                      const int* nn_count, const float* nn_dist, int* filt_index,
                                      void* stream)
{
    g_last_launch_count = 0;
    // attribute checks of tf_buildkernel.cpp:39-49
    if (B <= 0 || N <= 0 || M <= 0 || K <= 0 || !(radius > 0.0f) || !(n > 2 && n % 2 == 0) ||
        !(p > 0 && p % 2 == 0) || !(q > 0) || !database || !query || !nn_index || !nn_count ||
        !nn_dist || !filt_index)
        return (int)cudaErrorInvalidValue;
    size_t total = (size_t)B * M * K;
    size_t want = (total + 255) / 256;
    unsigned grid = (unsigned)(want < (size_t)sm_count() * 32 ? want : (size_t)sm_count() * 32);
    spherical_kernel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B, N, M, K, n, p, q, radius, database, query,
                                                                    nn_index, nn_count, nn_dist, filt_index);
    SPH3D_CHECK_LAUNCH();
    g_last_launch_count = 1;
    return 0;
}
