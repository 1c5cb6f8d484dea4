"""The separable layer as one kernel (csrc/sepconv.cu: gather -> tcgen05 product -> epilogue) against the oracle's
depthwise convolution followed by an fp64 pointwise product, and against the three-op composition it replaces
(/root/reference/utils/sph3gcn_util.py:128-161).  Tolerance: 1e-5 of sum|terms| per ELEMENT (common.assert_close_terms)."""
import numpy as np
import pytest
import torch

from common import assert_close, assert_close_terms, assert_equal, features, make_cloud, saturating_radius

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def T(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def A(t):
    return t.detach().cpu().numpy()


def _elu(z):
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))


def test_identity_graph_is_a_plain_product(pkg):
    """every point its own single neighbour in bin 0 with an all-ones filter: the depthwise result IS the input, so the
    kernel reduces to out = x @ W -- isolates the operand layout / descriptors / tensor-memory read-back."""
    B, N, C, cout = 2, 300, 128, 128
    x = features(1, B, N, C)
    W = features(2, C, cout)
    idx = np.tile(np.arange(N, dtype=np.int32)[None, :, None], (B, 1, 1))
    cnt = np.ones((B, N), np.int32)
    bins = np.zeros((B, N, 1), np.int32)
    filt = np.ones((3, C, 1), np.float32)
    out, dw = pkg.tf_sepconv.separable_conv3d(T(x), T(filt), T(W), T(idx), T(cnt), T(bins), keep_depthwise=True)
    assert_equal(A(dw), x, "depthwise of the identity graph")
    truth = x.astype(np.float64) @ W.astype(np.float64)
    terms = np.abs(x).astype(np.float64) @ np.abs(W).astype(np.float64)
    assert_close_terms(A(out), truth, terms, 1e-5, "identity graph product")
    # one-hot rows: column k of the operand meets row k of W and nothing else (pins the swizzle and the k order)
    eye = np.zeros((1, 128, 128), np.float32)
    eye[0, np.arange(128), np.arange(128)] = 1.0
    out = pkg.tf_sepconv.separable_conv3d(T(eye), T(filt), T(W), T(idx[:1, :128]), T(cnt[:1, :128]), T(bins[:1, :128]))[0]
    assert_close(A(out)[0], W, 1e-6, "one-hot rows read W back")


CASES = [
    # name, B, N, K, C, r, Cout, bias, affine, act
    ("cfgT_like_c128_r1_o128", 2, 1200, 64, 128, 1, 128, True, True, 1),
    ("s3dis_l1_c64_r2_o64", 2, 1000, 32, 64, 2, 64, True, False, 1),
    ("c64_r1_o32_one_chunk", 2, 700, 32, 64, 1, 32, False, False, 0),
    ("c36_r2_o64_padded_k", 2, 800, 64, 36, 2, 64, True, True, 1),
    ("c68_r1_o128", 2, 600, 48, 68, 1, 128, False, True, 1),
    ("c32_r2_o256_ring_wraps", 1, 900, 32, 32, 2, 256, True, True, 0),
    ("c128_r1_o13_odd_outputs", 2, 500, 32, 128, 1, 13, True, False, 0),
    ("many_tiles_per_cta", 4, 6000, 16, 64, 2, 128, True, True, 1),
    ("tiny_one_partial_tile", 1, 37, 8, 128, 1, 64, False, False, 1),
    ("s3dis_l1b_c128_r2_o128_tiles_of_32", 2, 1100, 64, 128, 2, 128, True, True, 1),
    ("c128_r2_o256_two_blocks", 1, 900, 32, 128, 2, 256, True, False, 1),
    ("c96_r2_o200_three_chunks", 2, 700, 32, 96, 2, 200, False, True, 0),
]


def _inputs(oracle, case):
    name, B, N, K, C, r, cout, with_bias, affine, act = case
    xyz = make_cloud(71, B, N)
    radius = saturating_radius(N, min(K, 64)) * 0.9
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, xyz, radius, None, K)
    bins = oracle.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, [8, 2, 2])
    x = features(72, B, N, C)
    filt = (features(73, 33, C, r) * 0.5).astype(np.float32)
    W = (features(74, C * r, cout) * 0.3).astype(np.float32)
    bias = features(75, cout) if with_bias else None
    scale = (1.0 + 0.2 * features(76, cout)).astype(np.float32) if affine else None
    shift = features(77, cout) if affine else None
    return x, filt, W, idx, cnt, bins, bias, scale, shift


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_layer_vs_oracle(case, pkg, oracle):
    x, filt, W, idx, cnt, bins, bias, scale, shift = _inputs(oracle, case)
    act = case[-1]
    out, dw = pkg.tf_sepconv.separable_conv3d(T(x), T(filt), T(W), T(idx), T(cnt), T(bins), bias=T(bias), scale=T(scale),
                                              shift=T(shift), act=act, keep_depthwise=True)
    # the depthwise result the kernel optionally keeps is the forward op's, bit for bit
    assert_equal(A(dw), A(pkg.tf_conv3d.depthwise_conv3d(T(x), T(filt), T(idx), T(cnt), T(bins))), "kept depthwise output")
    d = oracle.depthwise_conv3d(x, filt, idx, cnt, bins, mode=1).astype(np.float64)
    dabs = oracle.depthwise_conv3d(np.abs(x), np.abs(filt), idx, cnt, bins, mode=1).astype(np.float64)
    B, M, Kp = d.shape
    z = d.reshape(-1, Kp) @ W.astype(np.float64)
    zabs = dabs.reshape(-1, Kp) @ np.abs(W).astype(np.float64)
    if bias is not None:
        z, zabs = z + bias, zabs + np.abs(bias)
    if act == 1:
        z = _elu(z)                                            # 1-Lipschitz: the bound on z carries over
    if scale is not None:
        z, zabs = z * scale + shift, zabs * np.abs(scale) + np.abs(shift)
    assert_close_terms(A(out).reshape(-1, W.shape[1]), z, zabs, 1e-5, case[0])
    # without keeping the intermediate: same product
    out2, none = pkg.tf_sepconv.separable_conv3d(T(x), T(filt), T(W), T(idx), T(cnt), T(bins), bias=T(bias), scale=T(scale),
                                                 shift=T(shift), act=act)
    assert none is None
    assert_equal(A(out2), A(out), "output does not depend on keeping the depthwise result")


def test_unsupported_shapes_are_refused(pkg):
    L = pkg._lib.lib()
    assert L.sph3d_separable_conv3d_supported(2, 1000, 1000, 33, 128, 1, 64, 128) == 1
    assert L.sph3d_separable_conv3d_supported(2, 1000, 1000, 33, 128, 2, 64, 128) == 1      # C*r = 256: tiles of 32 points
    assert L.sph3d_separable_conv3d_supported(2, 1000, 1000, 33, 256, 1, 64, 128) == 0      # two channel chunks
    assert L.sph3d_separable_conv3d_supported(2, 1000, 1000, 33, 64, 3, 64, 128) == 0       # multiplier 3
    assert L.sph3d_separable_conv3d_supported(2, 1000, 1000, 33, 64, 1, 64, 512) == 0       # Cout > 256
    x = torch.zeros(1, 10, 256, device=DEV)
    with pytest.raises(ValueError):
        pkg.tf_sepconv.separable_conv3d(x, torch.zeros(3, 256, 1, device=DEV), torch.zeros(256, 8, device=DEV),
                                        torch.zeros(1, 10, 1, dtype=torch.int32, device=DEV),
                                        torch.ones(1, 10, dtype=torch.int32, device=DEV),
                                        torch.zeros(1, 10, 1, dtype=torch.int32, device=DEV))


@pytest.mark.parametrize("training", [True, False], ids=["training", "inference"])
def test_layer_routes_through_the_fused_kernel_and_matches_the_composition(training, pkg, oracle):
    """utils.sph3gcn_util.separable_conv3d with FUSE_SEPARABLE on / off: same outputs, same gradients, same variables"""
    s3g = pkg.utils.sph3gcn_util
    B, N, K, C, r, cout = 2, 900, 32, 64, 2, 128
    xyz = make_cloud(81, B, N)
    radius = saturating_radius(N, K) * 0.9
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, xyz, radius, None, K)
    bins = oracle.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, [8, 2, 2])
    x = features(82, B, N, C)
    g = features(83, B, N, cout)

    def run(fuse):
        s3g.reset_variables()
        s3g.clear_collections()
        torch.manual_seed(5)
        old = s3g.FUSE_SEPARABLE, s3g.FUSE_SEPARABLE_TRAINING
        s3g.FUSE_SEPARABLE = s3g.FUSE_SEPARABLE_TRAINING = fuse
        try:
            xt = T(x).requires_grad_(training)
            with torch.set_grad_enabled(training):
                out = s3g.separable_conv3d(xt, cout, 33, r, 'layer', T(idx), T(cnt), T(bins), weight_decay=1e-4,
                                           with_bn=True, with_bias=True, is_training=training)
            launches = pkg._lib.lib().sph3d_last_launch_count()
            grads = {}
            if training:
                out.backward(T(g))
                grads = {k: A(v.grad) for k, v in s3g.named_variables().items() if v.grad is not None}
                grads['input'] = A(xt.grad)
            names = list(s3g.named_variables())
            return A(out), grads, names, launches
        finally:
            s3g.FUSE_SEPARABLE, s3g.FUSE_SEPARABLE_TRAINING = old

    out_f, grads_f, names_f, _ = run(True)
    out_c, grads_c, names_c, _ = run(False)
    assert names_f == names_c
    assert_close(out_f, out_c, 2e-5, "fused vs composed layer output")
    assert set(grads_f) == set(grads_c)
    for k in grads_c:
        assert_close(grads_f[k], grads_c[k], 1e-4, "gradient of " + k)
