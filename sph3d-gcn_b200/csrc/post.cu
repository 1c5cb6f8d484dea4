// post.cu -- the tail of every SPH3D layer (bias -> ELU -> batch normalisation), forward and backward, sm_100a.
//
// Replaces the chain of TensorFlow graph nodes that follows the pointwise matmul in
// /root/reference/utils/sph3gcn_util.py:147-161 (separable_conv3d), :206-220 (pointwise_conv3d), :257-271
// (fully_connected): tf.nn.bias_add -> activation_fn (tf.nn.elu) -> tf.layers.batch_normalization(momentum=0.99)
// (:328-332).  SURVEY.md section 8(f) row N2: "the step either side of the conv".
//
// x is the (R, C) row-major matmul result, R = B*M rows.  With z = x + bias, y = act(z):
//   forward   out = gamma * (y - mean) * invstd + beta        mean / var over the R rows (training) or the moving
//                                                             statistics (inference); moving statistics updated
//   backward  dbeta = sum_r g,  dgamma = sum_r g * yhat,  dy = gamma*invstd*(g - dbeta/R - yhat*dgamma/R),
//             dx = dy * act'(z),  dbias = sum_r dx
// y and yhat are never written to memory: both passes of a direction re-derive them from x (HBM-bound column
// reductions and streams; eager execution of the same chain moves ~4x the bytes and launches ~6x the kernels).
//
// Work unit: a CTA is 8 warps; lane l of every warp owns VEC consecutive channels of a 32*VEC-channel chunk
// (blockIdx.x), warp w walks rows w, w+8*gridDim.y, ... of row block blockIdx.y, four rows in flight.  A warp load
// is one contiguous 128*VEC-byte piece of a row.  Column sums leave a CTA as ONE partial per channel (the 8 warps
// meet in shared memory) and the partials are folded in a fixed order by a second tiny kernel: results are
// bit-reproducible run to run.  The variance uses per-thread shifted sums joined with Chan's parallel formula (no
// E[y^2] - mean^2 cancellation), the cross-CTA join runs in fp64.
#include "rowwarp.cuh"
#include "../../include/sph3d_b200.h"

namespace sph3d {

enum { POST_ACT_NONE = 0, POST_ACT_ELU = 1 };
enum { PASS_STATS = 0, PASS_APPLY = 1, PASS_BSUMS = 2, PASS_BAPPLY = 3 };

struct PostArgs {
    int R, C;
    int has_bias, has_bn, training;
    const float* x;
    const float* bias;
    const float* gamma;
    const float* beta;
    const float* mean;
    const float* invstd;
    const float* g;           // grad_out (backward)
    const float* dgamma;      // column sums produced by PASS_BSUMS + fold (backward apply)
    const float* dbeta;
    float* out;               // forward: out; backward: grad_x
    float* part;              // [gridDim.y][NQ][C] partials
};

// ELU as TensorFlow evaluates it (tensorflow/core/kernels/relu_op_functor.h: features < 0 ? exp(features) - 1 : features;
// gradient (activations + 1) * gradients = exp(features) * gradients): one expf serves value and derivative.  The kernels
// are instruction-issue bound on this function (ncu: 75 % issue slots busy with expm1f, profiles/r1_layer_tail_ncu.md), so
// the 10-instruction expf is used rather than the 25-instruction expm1f; |exp(z)-1 - expm1(z)| <= 6e-8.
template <int ACT> __device__ __forceinline__ void act_eval(float z, float& y, float& der)
{
    if constexpr (ACT == POST_ACT_ELU) {
        const float e = expf(fminf(z, 0.f));
        const bool pos = z > 0.f;
        y = pos ? z : e - 1.0f;
        der = pos ? 1.0f : e;
    } else {
        y = z;
        der = 1.0f;
    }
}
template <int ACT> __device__ __forceinline__ float act_fwd(float z)
{
    float y, d;
    act_eval<ACT>(z, y, d);
    return y;
}

constexpr int POST_WARPS = 8;
constexpr int POST_UNROLL = 4;

// number of per-channel quantities a pass leaves per CTA
__host__ __device__ constexpr int post_nq(int pass) { return pass == PASS_STATS ? 3 : (pass == PASS_BSUMS ? 2 : (pass == PASS_BAPPLY ? 1 : 0)); }

template <int VEC, int PASS, int ACT>
__global__ void __launch_bounds__(POST_WARPS * 32)
post_pass_kernel(const PostArgs a)
{
    constexpr int NQ = post_nq(PASS);
    __shared__ float sh[(NQ > 0 ? NQ : 1) * POST_WARPS * 32 * VEC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + lane) * VEC;
    const bool active = c0 < a.C;                       // C % VEC == 0 (host), so an active lane owns VEC real channels
    const int cl = active ? c0 : 0;
    const size_t stride = (size_t)gridDim.y * POST_WARPS;
    const size_t R = (size_t)a.R;

    float b[VEC], mu[VEC], sc[VEC], be[VEC], is[VEC], kg[VEC], kb[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        b[v] = a.has_bias ? __ldg(a.bias + cl + v) : 0.f;
        mu[v] = 0.f; sc[v] = 1.f; be[v] = 0.f; is[v] = 1.f; kg[v] = 0.f; kb[v] = 0.f;
        if (PASS != PASS_STATS && a.has_bn) {
            mu[v] = __ldg(a.mean + cl + v);
            is[v] = __ldg(a.invstd + cl + v);
            sc[v] = __ldg(a.gamma + cl + v) * is[v];
            if (PASS == PASS_APPLY) be[v] = __ldg(a.beta + cl + v);
            if (PASS == PASS_BAPPLY && a.training) {
                const float invR = 1.0f / (float)a.R;
                kg[v] = __ldg(a.dgamma + cl + v) * invR;
                kb[v] = __ldg(a.dbeta + cl + v) * invR;
            }
        }
    }

    // per-thread accumulators: STATS {shift k, sum(y-k), sum (y-k)^2}; BSUMS {sum g, sum g*yhat}; BAPPLY {sum dx}
    float acc0[VEC], acc1[VEC], kshift[VEC];
    int nrows = 0;
#pragma unroll
    for (int v = 0; v < VEC; v++) { acc0[v] = 0.f; acc1[v] = 0.f; kshift[v] = 0.f; }

    size_t r = (size_t)blockIdx.y * POST_WARPS + warp;
    if (PASS == PASS_STATS && r < R) {                  // the shift is this thread's first sample
        float xv[VEC];
        VecIO<VEC>::ld(xv, a.x + r * a.C + cl, true);
#pragma unroll
        for (int v = 0; v < VEC; v++) kshift[v] = act_fwd<ACT>(xv[v] + b[v]);
    }

    for (; r < R; r += stride * POST_UNROLL) {
        float xv[POST_UNROLL][VEC], gv[POST_UNROLL][VEC];
#pragma unroll
        for (int u = 0; u < POST_UNROLL; u++) {
            const size_t ru = r + u * stride;
            const bool ok = ru < R;
            const size_t off = (ok ? ru : r) * a.C + cl;
            VecIO<VEC>::ld(xv[u], a.x + off, true);
            if (PASS == PASS_BSUMS || PASS == PASS_BAPPLY) VecIO<VEC>::ld(gv[u], a.g + off, true);
        }
#pragma unroll
        for (int u = 0; u < POST_UNROLL; u++) {
            const size_t ru = r + u * stride;
            if (ru >= R) break;
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const float z = xv[u][v] + b[v];
                float y, der;
                act_eval<ACT>(z, y, der);
                if (PASS == PASS_STATS) {
                    const float d = y - kshift[v];
                    acc0[v] += d;
                    acc1[v] = fmaf(d, d, acc1[v]);
                } else if (PASS == PASS_APPLY) {
                    o[v] = a.has_bn ? fmaf(y - mu[v], sc[v], be[v]) : y;
                } else if (PASS == PASS_BSUMS) {
                    const float yh = (y - mu[v]) * is[v];
                    acc0[v] += gv[u][v];
                    acc1[v] = fmaf(gv[u][v], yh, acc1[v]);
                } else {
                    float dy = gv[u][v];
                    if (a.has_bn) {
                        const float yh = (y - mu[v]) * is[v];
                        dy = sc[v] * (dy - kb[v] - yh * kg[v]);
                    }
                    const float dz = dy * der;
                    o[v] = dz;
                    acc0[v] += dz;
                }
            }
            if ((PASS == PASS_APPLY || PASS == PASS_BAPPLY) && active) VecIO<VEC>::st(a.out + ru * a.C + c0, o);
            nrows++;
        }
    }

    if constexpr (NQ > 0) {
        // the 8 warps of the CTA meet in shared memory; warp 0 writes one partial per channel
        float q[3][VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            if (PASS == PASS_STATS) {
                const float n = (float)nrows;
                q[0][v] = n;
                q[1][v] = nrows ? kshift[v] + acc0[v] / n : 0.f;                 // local mean
                q[2][v] = nrows ? fmaxf(acc1[v] - acc0[v] * acc0[v] / n, 0.f) : 0.f;   // local sum of squared deviations
            } else {
                q[0][v] = acc0[v]; q[1][v] = acc1[v]; q[2][v] = 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < NQ; k++)
#pragma unroll
            for (int v = 0; v < VEC; v++) sh[((k * POST_WARPS + warp) * 32 + lane) * VEC + v] = q[k][v];
        __syncthreads();
        if (warp == 0 && active && a.part != nullptr) {
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                float* dst = a.part + ((size_t)blockIdx.y * NQ) * a.C + c0 + v;
                if (PASS == PASS_STATS) {
                    float n = 0.f, m = 0.f, M2 = 0.f;
                    for (int w = 0; w < POST_WARPS; w++) {
                        const float nb = sh[((0 * POST_WARPS + w) * 32 + lane) * VEC + v];
                        if (nb > 0.f) {
                            const float mb = sh[((1 * POST_WARPS + w) * 32 + lane) * VEC + v];
                            const float Mb = sh[((2 * POST_WARPS + w) * 32 + lane) * VEC + v];
                            const float nn = n + nb, d = mb - m;
                            m += d * (nb / nn);
                            M2 += Mb + d * d * (n * nb / nn);
                            n = nn;
                        }
                    }
                    dst[0] = n; dst[a.C] = m; dst[2 * (size_t)a.C] = M2;
                } else {
#pragma unroll
                    for (int k = 0; k < NQ; k++) {
                        float s = 0.f;
                        for (int w = 0; w < POST_WARPS; w++) s += sh[((k * POST_WARPS + w) * 32 + lane) * VEC + v];
                        dst[(size_t)k * a.C] = s;
                    }
                }
            }
        }
    }
}

// Fold the P per-CTA (n, mean, M2) records of every channel (fixed order, fp64), emit mean / invstd and move the
// moving statistics (tf.layers.batch_normalization: moving = moving*momentum + batch*(1-momentum), biased variance).
// The records are ~1 MB spread over P rows; a fold is latency-bound, so it wants many loads in flight: a CTA owns only
// FOLD_CH = 8 channels (one 32-byte sector per record row) and cuts the record list into FOLD_SL = 128 slices.
// Two sweeps instead of a chain of Chan joins: mean = sum n_p m_p / sum n_p, then M2 = sum [M2_p + n_p (m_p - mean)^2]
// -- no division inside the loops.
constexpr int FOLD_CH = 8;
constexpr int FOLD_SL = 128;

// sum of v over the FOLD_SL slices of a channel, same value (and same summation order) in every thread of the channel
__device__ __forceinline__ double fold_slices(double v, double (*sh)[FOLD_CH], int ch, int warp)
{
    v += __shfl_xor_sync(FULL_MASK, v, 8);               // a warp holds 4 slices x 8 channels
    v += __shfl_xor_sync(FULL_MASK, v, 16);
    if ((threadIdx.x & 31) < FOLD_CH) sh[warp][ch] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < FOLD_SL / 4; w++) t += sh[w][ch];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(FOLD_CH * FOLD_SL)
post_fold_stats_kernel(int P, int C, float eps, float momentum, const float* __restrict__ part,
                       float* __restrict__ mean, float* __restrict__ invstd,
                       float* __restrict__ moving_mean, float* __restrict__ moving_var)
{
    __shared__ double sh[FOLD_SL / 4][FOLD_CH];
    const int ch = threadIdx.x & (FOLD_CH - 1), sl = threadIdx.x / FOLD_CH, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * FOLD_CH + ch;
    const int cc = c < C ? c : 0;
    double n = 0.0, nm = 0.0;
    for (int p = sl; p < P; p += FOLD_SL) {
        const float* rec = part + (size_t)p * 3 * C + cc;
        const double nb = rec[0];
        n += nb;
        nm += nb * (double)rec[C];
    }
    n = fold_slices(n, sh, ch, warp);
    nm = fold_slices(nm, sh, ch, warp);
    const double m = n > 0.0 ? nm / n : 0.0;
    double M2 = 0.0;
    for (int p = sl; p < P; p += FOLD_SL) {
        const float* rec = part + (size_t)p * 3 * C + cc;
        const double d = (double)rec[C] - m;
        M2 += (double)rec[2 * (size_t)C] + (double)rec[0] * d * d;
    }
    M2 = fold_slices(M2, sh, ch, warp);
    if (sl == 0 && c < C) {
        const double var = n > 0.0 ? M2 / n : 0.0;
        mean[c] = (float)m;
        invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (moving_mean) moving_mean[c] = moving_mean[c] * momentum + (float)m * (1.0f - momentum);
        if (moving_var) moving_var[c] = moving_var[c] * momentum + (float)var * (1.0f - momentum);
    }
}

// inference: mean / invstd come from the moving statistics
__global__ void post_eval_stats_kernel(int C, float eps, const float* __restrict__ moving_mean,
                                       const float* __restrict__ moving_var, float* __restrict__ mean,
                                       float* __restrict__ invstd)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        mean[c] = moving_mean[c];
        invstd[c] = (float)(1.0 / sqrt((double)moving_var[c] + (double)eps));
    }
}

// out_k[c] = sum_p part[p][k][c], fixed order, fp64 accumulation; NQ in {1, 2}
__global__ void __launch_bounds__(FOLD_CH * FOLD_SL)
post_fold_sums_kernel(int P, int C, int NQ, const float* __restrict__ part, float* __restrict__ out0,
                      float* __restrict__ out1)
{
    __shared__ double sh[FOLD_SL / 4][FOLD_CH];
    const int ch = threadIdx.x & (FOLD_CH - 1), sl = threadIdx.x / FOLD_CH, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * FOLD_CH + ch;
    const int cc = c < C ? c : 0;
    double a0 = 0.0, a1 = 0.0;
    for (int p = sl; p < P; p += FOLD_SL) {
        const float* rec = part + (size_t)p * NQ * C + cc;
        a0 += rec[0];
        if (NQ > 1) a1 += rec[C];
    }
    a0 = fold_slices(a0, sh, ch, warp);
    if (NQ > 1) a1 = fold_slices(a1, sh, ch, warp);
    if (sl == 0 && c < C) {
        if (out0) out0[c] = (float)a0;
        if (NQ > 1 && out1) out1[c] = (float)a1;
    }
}

struct PostPlan { int vec, chunks, P; };

static PostPlan post_plan(int R, int C, const void* p0, const void* p1, const void* p2)
{
    PostPlan p;
    int vec = (C % 4 == 0) ? 4 : ((C % 2 == 0) ? 2 : 1);
    for (int v = 1; v <= vec; v <<= 1)                       // narrowest strip that still covers C with one chunk
        if (C % v == 0 && C <= 32 * v) { vec = v; break; }
    const uintptr_t al = (uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2;
    while (vec > 1 && (al % (vec * sizeof(float))) != 0) vec >>= 1;
    p.vec = vec;
    p.chunks = (C + 32 * vec - 1) / (32 * vec);
    long long want = ((long long)sm_count() * 4 + p.chunks - 1) / p.chunks;     // ~4 CTAs of 8 warps per SM
    const long long need = ((long long)R + POST_WARPS - 1) / POST_WARPS;
    if (want > need) want = need;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    p.P = (int)want;
    return p;
}

// P does not depend on pointer alignment, so the workspace query and the calls agree
static size_t post_workspace_floats(int R, int C)
{
    long long maxP = (long long)sm_count() * 4;               // chunks >= 1
    const long long need = ((long long)R + POST_WARPS - 1) / POST_WARPS;
    if (maxP > need) maxP = need;
    if (maxP < 1) maxP = 1;
    return (size_t)maxP * 3 * (size_t)C;
}

template <int PASS>
static int launch_pass(const PostPlan& p, int act, const PostArgs& a, cudaStream_t st)
{
    dim3 grid(p.chunks, p.P);
#define POST_GO(V, A) post_pass_kernel<V, PASS, A><<<grid, POST_WARPS * 32, 0, st>>>(a)
    if (act == POST_ACT_ELU) {
        if (p.vec == 4) POST_GO(4, POST_ACT_ELU); else if (p.vec == 2) POST_GO(2, POST_ACT_ELU); else POST_GO(1, POST_ACT_ELU);
    } else {
        if (p.vec == 4) POST_GO(4, POST_ACT_NONE); else if (p.vec == 2) POST_GO(2, POST_ACT_NONE); else POST_GO(1, POST_ACT_NONE);
    }
#undef POST_GO
    SPH3D_CHECK_LAUNCH();
    return 0;
}

}  // namespace sph3d

using namespace sph3d;

extern "C" size_t sph3d_bias_act_bn_workspace_bytes(int R, int C)
{
    if (R <= 0 || C <= 0) return 0;
    return post_workspace_floats(R, C) * sizeof(float);
}

extern "C" int sph3d_bias_act_bn(int R, int C, int act, int training, float eps, float momentum,
                                 const float* x, const float* bias, const float* gamma, const float* beta,
                                 float* moving_mean, float* moving_var, float* out, float* save_mean,
                                 float* save_invstd, void* workspace, size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    const bool has_bn = gamma != nullptr;
    if (R <= 0 || C <= 0 || !x || !out || (act != POST_ACT_NONE && act != POST_ACT_ELU)) return (int)cudaErrorInvalidValue;
    if (has_bn && (!beta || !save_mean || !save_invstd || !moving_mean || !moving_var)) return (int)cudaErrorInvalidValue;
    if (has_bn && training && (!workspace || workspace_bytes < post_workspace_floats(R, C) * sizeof(float)))
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    const PostPlan p = post_plan(R, C, x, out, nullptr);
    PostArgs a{};
    a.R = R; a.C = C; a.has_bias = bias != nullptr; a.has_bn = has_bn; a.training = training;
    a.x = x; a.bias = bias; a.gamma = gamma; a.beta = beta; a.mean = save_mean; a.invstd = save_invstd;
    a.out = out; a.part = reinterpret_cast<float*>(workspace);
    int launches = 0, rc;
    if (has_bn) {
        if (training) {
            if ((rc = launch_pass<PASS_STATS>(p, act, a, st)) != 0) return rc;
            post_fold_stats_kernel<<<(C + FOLD_CH - 1) / FOLD_CH, FOLD_CH * FOLD_SL, 0, st>>>(p.P, C, eps, momentum, a.part, save_mean, save_invstd,
                                                                  moving_mean, moving_var);
            SPH3D_CHECK_LAUNCH();
            launches += 2;
        } else {
            post_eval_stats_kernel<<<(C + 127) / 128, 128, 0, st>>>(C, eps, moving_mean, moving_var, save_mean, save_invstd);
            SPH3D_CHECK_LAUNCH();
            launches += 1;
        }
    }
    if ((rc = launch_pass<PASS_APPLY>(p, act, a, st)) != 0) return rc;
    g_last_launch_count = launches + 1;
    return 0;
}

extern "C" int sph3d_bias_act_bn_grad(int R, int C, int act, int training,
                                      const float* x, const float* bias, const float* gamma,
                                      const float* save_mean, const float* save_invstd, const float* grad_out,
                                      float* grad_x, float* grad_bias, float* grad_gamma, float* grad_beta,
                                      void* workspace, size_t workspace_bytes, void* stream)
{
    g_last_launch_count = 0;
    const bool has_bn = gamma != nullptr;
    if (R <= 0 || C <= 0 || !x || !grad_out || !grad_x || (act != POST_ACT_NONE && act != POST_ACT_ELU))
        return (int)cudaErrorInvalidValue;
    if (has_bn && (!save_mean || !save_invstd || !grad_gamma || !grad_beta)) return (int)cudaErrorInvalidValue;
    if ((bias != nullptr) != (grad_bias != nullptr)) return (int)cudaErrorInvalidValue;
    if ((has_bn || bias) && (!workspace || workspace_bytes < post_workspace_floats(R, C) * sizeof(float)))
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    const PostPlan p = post_plan(R, C, x, grad_out, grad_x);
    PostArgs a{};
    a.R = R; a.C = C; a.has_bias = bias != nullptr; a.has_bn = has_bn; a.training = training;
    a.x = x; a.bias = bias; a.gamma = gamma; a.mean = save_mean; a.invstd = save_invstd; a.g = grad_out;
    a.dgamma = grad_gamma; a.dbeta = grad_beta; a.out = grad_x; a.part = reinterpret_cast<float*>(workspace);   // null: no column sums wanted
    int launches = 0, rc;
    if (has_bn) {
        if ((rc = launch_pass<PASS_BSUMS>(p, act, a, st)) != 0) return rc;
        post_fold_sums_kernel<<<(C + FOLD_CH - 1) / FOLD_CH, FOLD_CH * FOLD_SL, 0, st>>>(p.P, C, 2, a.part, grad_beta, grad_gamma);
        SPH3D_CHECK_LAUNCH();
        launches += 2;
    }
    if (!bias) a.part = nullptr;                          // no column sum of dx wanted
    if ((rc = launch_pass<PASS_BAPPLY>(p, act, a, st)) != 0) return rc;
    launches += 1;
    if (bias) {
        post_fold_sums_kernel<<<(C + FOLD_CH - 1) / FOLD_CH, FOLD_CH * FOLD_SL, 0, st>>>(p.P, C, 1, a.part, grad_bias, nullptr);
        SPH3D_CHECK_LAUNCH();
        launches += 1;
    }
    g_last_launch_count = launches;
    return 0;
}
