// dense_abi.cu -- C ABI of the tensor-core pointwise product (dense_gemm.cuh).
#include "common.cuh"
#include "dense_gemm.cuh"
#include "../../include/sph3d_b200.h"

extern "C" size_t sph3d_dense_gemm_workspace_bytes(int op, int M, int N, int K, int L)
{
    if (M <= 0 || N <= 0 || K <= 0 || L <= 0) return 0;
    switch (op) {
    case 0: return sph3d_dense_nn::workspace(M, N, K, L);
    case 1: return sph3d_dense_nt::workspace(M, N, K, L);
    case 2: return sph3d_dense_tn::workspace(M, N, K, L);
    case 3: return sph3d_dense_nn2::workspace(M, N, K, L);
    case 4: return sph3d_dense_nt2::workspace(M, N, K, L);
    default: return 0;
    }
}

extern "C" int sph3d_dense_gemm(int op, int M, int N, int K, int L, const float* A, const float* B, float* D,
                                void* workspace, size_t workspace_bytes, void* stream)
{
    sph3d::g_last_launch_count = 0;
    if (M <= 0 || N <= 0 || K <= 0 || L <= 0 || !A || !B || !D) return (int)cudaErrorInvalidValue;
    // TMA: 16-byte aligned bases and leading dimensions
    if ((((uintptr_t)A | (uintptr_t)B | (uintptr_t)D) & 15) != 0) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (op) {
    case 0: rc = sph3d_dense_nn::run(M, N, K, L, A, B, D, workspace, workspace_bytes, st); break;
    case 1: rc = sph3d_dense_nt::run(M, N, K, L, A, B, D, workspace, workspace_bytes, st); break;
    case 2: rc = sph3d_dense_tn::run(M, N, K, L, A, B, D, workspace, workspace_bytes, st); break;
    case 3: rc = sph3d_dense_nn2::run(M, N, K, L, A, B, D, workspace, workspace_bytes, st); break;
    case 4: rc = sph3d_dense_nt2::run(M, N, K, L, A, B, D, workspace, workspace_bytes, st); break;
    default: return (int)cudaErrorInvalidValue;
    }
    if (rc == 0) sph3d::g_last_launch_count = 1;
    return rc;
}
