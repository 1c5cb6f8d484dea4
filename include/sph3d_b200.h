/*
 * sph3d_b200.h -- C ABI of libsph3d_b200.so: the SPH3D-GCN per-layer hot path on B200 (sm_100a).
 *
 * One entry point per host launcher of the reference (14 launchers in
 * /root/reference/tf_ops/<op>/tf_<op>_gpu.cu); each prototype cites the launcher it replaces and
 * keeps that launcher's argument ORDER, followed by (workspace,) stream.  Differences that are
 * deliberate:
 *   - every pointer is a DEVICE pointer owned by the caller; nothing is allocated or freed here;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), never synchronises
 *     the host, never touches the legacy default stream, and is CUDA-graph capturable;
 *   - outputs are written completely by the call (the zero fill the TensorFlow glue did with
 *     cudaMemset before each launch -- tf_<op>.cpp -- is folded into the kernels), so the caller
 *     may pass uninitialised buffers;
 *   - the return value is a cudaError_t as int (0 = cudaSuccess; 1 = cudaErrorInvalidValue for a
 *     bad dimension/attribute, the analogue of the glue's errors::InvalidArgument).
 *
 * Layouts are the reference's: contiguous row-major, int32 indices, fp32 features.
 *   database (B,N,3)  query (B,M,3)  nn_index (B,M,K)  nn_count (B,M)  nn_dist (B,M,K)
 *   input (B,N,C)  filter (F,C,r)  output (B,M,C*r) with cout = c*r + j
 */
#ifndef SPH3D_B200_H_
#define SPH3D_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 5: sph3d_dense_gemm (CUTLASS instantiation) removed; sph3d_rows_gemm*, sph3d_rows_wgrad* added */
#define SPH3D_B200_ABI_VERSION 5
int sph3d_abi_version(void);

/* Number of KERNELS (memsets excluded) the calling thread's most recent entry-point call enqueued (bench.py
 * gpu_launches).  The counter is thread-local: every entry point is reentrant. */
int sph3d_last_launch_count(void);

/* The SPH3D_* launch tunables are read from the environment once, when the library is loaded.
 * A process that changes one afterwards (sweep scripts, tests) calls this to re-read them. */
void sph3d_reload_tunables(void);

/* ---- a1: buildSphereNeighborLauncher, tf_nnquery_gpu.cu:115-121 (kernel :15-65) --------------
 * Ball query with the reference's growing-radius chain semantics: bit-exact
 * nn_index / nn_count / nn_dist.  radius > 0, K > 0.
 * workspace (optional): sph3d_build_sphere_neighbor_workspace_bytes(...) bytes of device scratch for a
 * uniform cell grid over the database clouds; with it, queries whose radius spans <= 2 cells are answered
 * from the cell stencil instead of a full scan (same results).  Pass NULL/0 to always scan. */
size_t sph3d_build_sphere_neighbor_workspace_bytes(int B, int N, int M, int K);
int sph3d_build_sphere_neighbor(int B, int N, int M, int K, float radius,
                                const float* database, const float* query,
                                int* nn_index, int* nn_count, float* nn_dist,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- a2: buildCubeNeighborLauncher, tf_nnquery_gpu.cu:123-127 (kernel :72-113) ---------------
 * nn_index is (B,M,K,2): (point id, grid bin) interleaved; nn_count may be 0. */
int sph3d_build_cube_neighbor(int B, int N, int M, int grid_size, int K, float length,
                              const float* database, const float* query,
                              int* nn_index, int* nn_count, void* stream);

/* ---- a3: sphericalKernelLauncher, tf_buildkernel_gpu.cu:83-89 (kernel :20-79) ----------------
 * n>2 even, p>0 even, q>0, radius>0 (tf_buildkernel.cpp:39-49). filt_index in [0, n*p*q]. */
int sph3d_spherical_kernel(int B, int N, int M, int K, int n, int p, int q, float radius,
                           const float* database, const float* query, const int* nn_index,
                           const int* nn_count, const float* nn_dist, int* filt_index, void* stream);

/* ---- a4: depthwiseConv3dLauncher, tf_conv3d_gpu.cu:107-113 (kernel :7-29) --------------------
 * F (number of filter bins) is needed to stage the filter; the reference read it only in the
 * gradient launcher. */
int sph3d_depthwise_conv3d(int B, int N, int M, int F, int C, int r, int K,
                           const int* nn_index, const int* nn_count, const int* bin_index,
                           const float* input, const float* filter, float* output, void* stream);

/* ---- a4, split form: the same launcher with its graph-only work hoisted out ---------------------------------
 * sph3d_depthwise_conv3d groups every 64-edge tile of a row by filter bin before it gathers (so that the filter is
 * applied once per (row, bin) instead of once per edge).  That grouping depends only on (nn_index, nn_count,
 * bin_index, F): sph3d_conv_sort writes it once per graph -- (B,M,K) words, neighbour id << 8 | bin << 1 | last --
 * and sph3d_depthwise_conv3d_planned convolves from those words; outputs are bit-identical to the one-call form.
 * The reference's models apply two convolutions per graph (models/SPH3D_*.py call separable_conv3d twice per level);
 * the graph-build side (tf_buildkernel.spherical_kernel in the Python layer) emits the words next to filt_index.
 * sph3d_conv_sort_bytes returns 0 where the planned form does not apply (F > 128, N >= 2^24);
 * sph3d_depthwise_conv3d_planned_supported returns 0 when (C, r) is not covered (r > 2). */
size_t sph3d_conv_sort_bytes(int B, int N, int M, int F, int K);
int sph3d_conv_sort(int B, int N, int M, int F, int K,
                    const int* nn_index, const int* nn_count, const int* bin_index,
                    void* plan, size_t plan_bytes, void* stream);
size_t sph3d_depthwise_conv3d_planned_supported(int B, int N, int M, int F, int C, int r, int K);
int sph3d_depthwise_conv3d_planned(int B, int N, int M, int F, int C, int r, int K,
                                   const int* nn_count, const void* plan, size_t plan_bytes,
                                   const float* input, const float* filter, float* output, void* stream);

/* ---- a5: depthwiseConv3dGradLauncher, tf_conv3d_gpu.cu:115-140 (kernels :32-101) -------------
 * workspace: sph3d_depthwise_conv3d_grad_workspace_bytes(...) bytes of device scratch (transposed graph,
 * per-CTA filter-gradient partials reduced in a fixed order).  Summation order: the row-owned form (r = 2 one-call,
 * F > 130) is run-to-run deterministic in grad_filter; the transposed form orders the edges of a segment by an
 * atomic rank, so grad_input / grad_filter may differ in the last ulps between runs (as the reference's atomics
 * do) unless SPH3D_BWDT_SORT=1 asks for the canonical (ascending output point) order. */
size_t sph3d_depthwise_conv3d_grad_workspace_bytes(int B, int N, int M, int F, int C, int r, int K);
int sph3d_depthwise_conv3d_grad(int B, int N, int M, int F, int C, int r, int K,
                                const int* nn_index, const int* nn_count, const int* bin_index,
                                const float* input, const float* filter, const float* grad_output,
                                float* grad_input, float* grad_filter,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- a5, split form: the same launcher with its graph-only work hoisted out -------------------------------
 * sph3d_depthwise_conv3d_grad first transposes the graph (per input point n: the (output point, bin) pairs that
 * reference it, grouped by bin) and then runs one gather pass.  The transposed graph depends only on
 * (nn_index, nn_count, bin_index, F); a caller that applies several convolutions over one graph (the reference's
 * models run two per level: sph3gcn_util.py:88-161 called twice per level in models/SPH3D_*.py) or
 * steps a static graph builds it once with sph3d_conv_transpose and passes it to ..._grad_planned.
 * The plan entry of an edge carries nn_count of its output row when the bits allow (B*M*K*slots <= 2^32), so the gather
 * pass scales by 1/cnt itself and reads grad_output directly (no scaled copy).
 * sph3d_conv_transpose_bytes returns 0 when the planned form does not apply (more than 127 bins per warp class or
 * accumulators beyond shared memory: F > ~380; B*M*slots > 2^32);
 * sph3d_depthwise_conv3d_grad_planned_workspace_bytes returns 0 when (C, r) is not covered (r > 2). */
size_t sph3d_conv_transpose_bytes(int B, int N, int M, int F, int K);
int sph3d_conv_transpose(int B, int N, int M, int F, int K,
                         const int* nn_index, const int* nn_count, const int* bin_index,
                         void* plan, size_t plan_bytes, void* stream);
size_t sph3d_depthwise_conv3d_grad_planned_workspace_bytes(int B, int N, int M, int F, int C, int r, int K);
int sph3d_depthwise_conv3d_grad_planned(int B, int N, int M, int F, int C, int r, int K,
                                        const int* nn_count, const void* plan, size_t plan_bytes,
                                        const float* input, const float* filter, const float* grad_output,
                                        float* grad_input, float* grad_filter,
                                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- a6: farthestPointSampleLauncher, tf_sample_gpu.cu:77-80 (kernel :7-73) ------------------
 * temp: the reference's (32,n) scratch (tf_sample.cpp:50) becomes a caller-owned workspace of
 * sph3d_farthest_point_sample_workspace_bytes(b,n,m) bytes (0 when the cloud fits on chip). */
size_t sph3d_farthest_point_sample_workspace_bytes(int b, int n, int m);
int sph3d_farthest_point_sample(int b, int n, int m, const float* inp, void* temp, size_t temp_bytes,
                                int* out, void* stream);

/* ---- a8: maxPool3dLauncher / maxPool3dGradLauncher, tf_pool3d_gpu.cu:93-105 ------------------ */
int sph3d_max_pool3d(int B, int N, int M, int C, int K, const int* nn_index, const int* nn_count,
                     const float* input, float* output, int* max_index, void* stream);
int sph3d_max_pool3d_grad(int B, int N, int M, int C, const int* max_index,
                          const float* grad_output, float* grad_input, void* stream);

/* ---- a9: avgPool3dLauncher / avgPool3dGradLauncher, tf_pool3d_gpu.cu:107-119 ----------------- */
int sph3d_avg_pool3d(int B, int N, int M, int C, int K, const int* nn_index, const int* nn_count,
                     const float* input, float* output, void* stream);
/* Gradient, two forms.  With workspace == NULL the neighbour rows scatter into grad_input with vector reductions (one
 * per edge).  With sph3d_avg_pool3d_grad_workspace_bytes(...) bytes of scratch (0 = shape not covered) the graph is
 * transposed first and every input point GATHERS the rows that reference it: no atomics, no zero fill. */
size_t sph3d_avg_pool3d_grad_workspace_bytes(int B, int N, int M, int C, int K);
int sph3d_avg_pool3d_grad(int B, int N, int M, int C, int K, const int* nn_index, const int* nn_count,
                          const float* grad_output, float* grad_input,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- a10: meanInterpolateLauncher / GradLauncher, tf_unpool3d_gpu.cu:87-99 -------------------
 * As in the reference, N = fine (output) points, M = coarse (input) points:
 * input (B,M,C), nn_index (B,N,K) into the coarse cloud, output (B,N,C). */
int sph3d_mean_interpolate(int B, int N, int M, int C, int K, const int* nn_index, const int* nn_count,
                           const float* input, float* output, void* stream);
/* gradients: scatter form (workspace == NULL) or gather form over the transposed graph, as for avg-pool;
 * sph3d_interpolate_grad_workspace_bytes serves both the mean and the weighted gradient. */
size_t sph3d_interpolate_grad_workspace_bytes(int B, int N, int M, int C, int K);
int sph3d_mean_interpolate_grad(int B, int N, int M, int C, int K, const int* nn_index,
                                const int* nn_count, const float* grad_output, float* grad_input,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ---- a11: weightedInterpolateLauncher / GradLauncher, tf_unpool3d_gpu.cu:101-113 ------------- */
int sph3d_weighted_interpolate(int B, int N, int M, int C, int K, const int* nn_index,
                               const int* nn_count, const float* input, const float* weight,
                               float* output, void* stream);
int sph3d_weighted_interpolate_grad(int B, int N, int M, int C, int K, const int* nn_index,
                                    const int* nn_count, const float* grad_output, const float* weight,
                                    float* grad_input, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a12 (layer tail): bias -> activation -> batch normalisation, utils/sph3gcn_util.py:147-161, :206-220, :257-271
 * (tf.nn.bias_add, activation_fn = tf.nn.elu, tf.layers.batch_normalization(momentum=0.99) at :328-332).  The reference
 * runs these as separate TensorFlow graph nodes over the (B*M, C) matmul result; here they are one op whose
 * intermediate y = act(x + bias) never goes to memory.
 *   x, out, grad_* : (R, C) row-major, R = B*M.   act: 0 = none, 1 = ELU (alpha 1).
 *   bias == NULL: no bias.   gamma == NULL: no batch normalisation (beta / moving_* / save_* ignored).
 *   training != 0: batch statistics (biased variance), moving_* <- moving_* * momentum + batch * (1 - momentum);
 *   training == 0: the moving statistics normalise.  save_mean / save_invstd (C each) are outputs the gradient needs.
 * Gradient: grad_bias iff bias, grad_gamma / grad_beta iff gamma; column sums are folded in a fixed order
 * (bit-reproducible).  workspace: sph3d_bias_act_bn_workspace_bytes(R, C) bytes for either call. */
size_t sph3d_bias_act_bn_workspace_bytes(int R, int C);
int sph3d_bias_act_bn(int R, int C, int act, int training, float eps, float momentum,
                      const float* x, const float* bias, const float* gamma, const float* beta,
                      float* moving_mean, float* moving_var, float* out, float* save_mean, float* save_invstd,
                      void* workspace, size_t workspace_bytes, void* stream);
int sph3d_bias_act_bn_grad(int R, int C, int act, int training,
                           const float* x, const float* bias, const float* gamma,
                           const float* save_mean, const float* save_invstd, const float* grad_out,
                           float* grad_x, float* grad_bias, float* grad_gamma, float* grad_beta,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- N2: the separable layer as one kernel, utils/sph3gcn_util.py:128-161 (separable_conv3d) --------------------------
 * depthwise_conv3d (tf_conv3d_gpu.cu:7-29) -> tf.matmul with the pointwise weights -> bias_add -> elu -> per-channel
 * affine, fused (csrc/sepconv.cu, hand-written tcgen05): the (B, M, C*r) depthwise result goes from the gathering warps
 * into the tensor core's shared-memory operand and never to memory, unless depthwise_out != NULL (training keeps it for
 * the weight gradient; it is then bit-identical to sph3d_depthwise_conv3d's output).
 *   output (B*M, Cout) = act(depthwise(input) * W + bias) * scale + shift      act: 0 = none, 1 = ELU
 *   bias / scale / shift: Cout floats each or NULL (inference-mode batch normalisation folds into scale / shift:
 *   scale = gamma * invstd, shift = beta - moving_mean * scale; NULL, NULL, NULL, act 0 returns the raw product for the
 *   training-mode layer tail sph3d_bias_act_bn).
 *   weight_image: the pointwise weights W (C*r x Cout, fp32 row-major) re-packed by sph3d_sepconv_pack_weights into
 *   sph3d_sepconv_weight_image_bytes(C*r, Cout) bytes (three bf16 terms per weight in the tensor core's operand
 *   layout); re-pack whenever W changes.  fp32 accuracy: every cross term down to 2^-24 of a product is accumulated.
 * sph3d_separable_conv3d_supported: 1 where the fused kernel applies (r in {1,2}, C <= 128 and even, C*r <= 256,
 * Cout <= 256, F <= 128); elsewhere the caller composes the three ops.  Arguments up to `filter` are
 * sph3d_depthwise_conv3d's. */
int sph3d_separable_conv3d_supported(int B, int N, int M, int F, int C, int r, int K, int Cout);
size_t sph3d_sepconv_weight_image_bytes(int Cr, int Cout);
int sph3d_sepconv_pack_weights(int Cr, int Cout, const float* weights, void* weight_image, void* stream);
int sph3d_separable_conv3d(int B, int N, int M, int F, int C, int r, int K, int Cout,
                           const int* nn_index, const int* nn_count, const int* bin_index,
                           const float* input, const float* filter, const void* weight_image,
                           const float* bias, const float* scale, const float* shift, int act,
                           float* depthwise_out, float* output, void* stream);

/* ---- a12 (pointwise product, hand-written): y (R x N) = x (R x K) * W, utils/sph3gcn_util.py:144-146, :203-205, :254-256
 * csrc/rowsgemm.cu: the rows stream through HBM once as fp32, are split into three bf16 terms in registers and written
 * straight into the tensor core's shared-memory operand; six cross products per k-step (everything down to 2^-24 of a
 * product) accumulate in tensor memory (tcgen05.mma, one issuing thread), four epilogue warps store the finished tile while
 * the next one is produced.  No template library.
 *   image: the weights re-packed by sph3d_rows_gemm_pack into sph3d_rows_gemm_image_bytes(K, N) bytes (three bf16 terms in
 *   the operand's swizzled layout).  trans = 0: weights is (K x N) row-major (y = x * W); trans = 1: weights is (N x K)
 *   row-major (y = x * W^T: the input gradient gx = g * w^T of the same layer packs w with trans = 1).
 *   terms: bf16 terms per operand.  3 = the six cross products down to 2^-24 of a product; 2 = the four products of
 *   hi + mid, every dropped term below 2^-17 of a product (the rounding of an fp32 dot product of 128 terms is of that
 *   size), two thirds of the tensor-core work and of the operand traffic.
 *   K and N multiples of 4, x / y / image 16-byte aligned, else cudaErrorInvalidValue (1). */
size_t sph3d_rows_gemm_image_bytes(int K, int N);
int sph3d_rows_gemm_pack(int K, int N, const float* weights, int trans, void* image, void* stream);
/* weights (K x N): `image` for x * W (sph3d_rows_gemm_image_bytes(K, N) bytes) and `image_t` for g * W^T
 * (sph3d_rows_gemm_image_bytes(N, K) bytes) in one launch -- what a training step needs of a layer */
int sph3d_rows_gemm_pack_pair(int K, int N, const float* weights, void* image, void* image_t, void* stream);
int sph3d_rows_gemm(int R, int K, int N, int terms, const float* x, const void* image, float* y, void* stream);
/* gw (K x N) = x^T * g, the weight gradient of the same product, summed over the R rows (csrc/rowswgrad.cu: both operands
 * MN-major for the tensor core, so x (R x K) and g (R x N) are converted exactly as they lie in memory; one CTA per
 * 128 x 128 block of gw and slab of rows, partial blocks in `workspace` summed in slab order -> deterministic).
 * Same `terms` and alignment rules as sph3d_rows_gemm; workspace of sph3d_rows_wgrad_workspace_bytes(R, K, N) bytes. */
size_t sph3d_rows_wgrad_workspace_bytes(int R, int K, int N);
int sph3d_rows_wgrad(int R, int K, int N, int terms, const float* x, const float* g, float* gw,
                     void* workspace, size_t workspace_bytes, void* stream);
/* diagnostics: while `buffer` (6*64*4 int64 of device memory) is set, CTA (0,0) of every sph3d_rows_gemm launch records
 * clock64 stamps of its producer / issuer / epilogue steps there; NULL switches it off (the default). */
void sph3d_rows_gemm_trace(void* buffer);

#ifdef __cplusplus
}
#endif
#endif /* SPH3D_B200_H_ */
