"""Parity at BASELINE.json's full per-cloud sizes (N = 10 000 / 65 536, K = 64, Cin = 128): direct oracle
comparison on a few clouds (the oracle is per-cloud independent, so a B=2..4 slice of the headline batch is the
same code path as B=32), plus size-independent properties on the whole Cfg-T batch."""
import numpy as np
import pytest
import torch

from common import assert_close, assert_equal, features, make_cloud, saturating_radius

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
A = lambda t: t.detach().cpu().numpy()


def test_cfgT_graph_and_conv_vs_oracle(pkg, oracle):
    """N=M=10000 (ten radius-chain steps per reference thread, Q1), K=64, C=128: graph bit-exact, conv 1e-5."""
    B, N, K, C = 2, 10000, 64, 128
    xyz = make_cloud(901, B, N, "cube")
    r = saturating_radius(N, K)
    oi, oc, od = oracle.build_sphere_neighbor(xyz, xyz, r, None, K)
    gi, gc, gd = pkg.tf_nnquery.build_sphere_neighbor(T(xyz), T(xyz), radius=r, nnsample=K)
    assert_equal(A(gi), oi); assert_equal(A(gc), oc); assert_equal(A(gd), od)
    of = oracle.spherical_kernel(xyz, xyz, oi, oc, od, r, [8, 2, 2])
    gf = pkg.tf_buildkernel.spherical_kernel(T(xyz), T(xyz), gi, gc, gd, r, kernel=[8, 2, 2])
    assert_equal(A(gf), of)
    for mult in (1, 2):
        x, W, go = features(902, B, N, C), (0.1 * features(903, 33, C, mult)).astype(np.float32), features(904, B, N, C * mult)
        xt, Wt = T(x).requires_grad_(True), T(W).requires_grad_(True)
        out = pkg.tf_conv3d.depthwise_conv3d(xt, Wt, gi, gc, gf)
        assert_close(A(out), oracle.depthwise_conv3d(x, W, oi, oc, of, 1), 1e-5, "conv fwd r=%d" % mult)
        out.backward(T(go))
        ti, tf = oracle.depthwise_conv3d_grad(x, W, go, oi, oc, of)
        assert_close(A(xt.grad), ti, 1e-5, "grad_input r=%d" % mult)
        assert_close(A(Wt.grad), tf, 1e-5, "grad_filter r=%d" % mult)


def test_cfgT_full_batch_adjoint_identity(pkg):
    """whole headline batch (B=32): the op is bilinear, so <gO, out> = <grad_input, x> = <grad_filter, W>."""
    B, N, K, C = 32, 10000, 64, 128
    g = torch.Generator().manual_seed(905)
    xyz = torch.rand(B, N, 3, generator=g).to(DEV)
    r = saturating_radius(N, K)
    idx, cnt, dst = pkg.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=r, nnsample=K)
    filt = pkg.tf_buildkernel.spherical_kernel(xyz, xyz, idx, cnt, dst, r, kernel=[8, 2, 2])
    assert int(cnt.min()) >= 1 and int(cnt.max()) <= K and int(filt.max()) <= 32 and int(filt.min()) >= 0
    x = torch.randn(B, N, C, generator=g).to(DEV).requires_grad_(True)
    W = (0.1 * torch.randn(33, C, 1, generator=g)).to(DEV).requires_grad_(True)
    go = torch.randn(B, N, C, generator=g).to(DEV)
    out = pkg.tf_conv3d.depthwise_conv3d(x, W, idx, cnt, filt)
    out.backward(go)
    s = float((go.double() * out.detach().double()).sum())
    si = float((x.grad.double() * x.detach().double()).sum())
    sw = float((W.grad.double() * W.detach().double()).sum())
    assert abs(si - s) <= 1e-5 * abs(s) + 1e-2 and abs(sw - s) <= 1e-5 * abs(s) + 1e-2, (s, si, sw)
    # rows are convex-ish combinations: |out| <= max|x| * max|W| summed over nothing more than cnt terms / cnt
    assert float(out.detach().abs().max()) <= float(x.detach().abs().max()) * float(W.detach().abs().max()) * 1.0001


def test_fps_full_sizes_vs_oracle(pkg, oracle):
    for name, B, N, S, kind in (("modelnet_l1", 2, 10000, 2500, "shell"), ("s3dis_l1", 2, 8192, 2048, "cube"),
                                ("scannet_cluster8", 1, 65536, 512, "cube"), ("cluster2", 1, 16384, 700, "grid")):
        xyz = make_cloud(906, B, N, kind)
        assert_equal(A(pkg.tf_sample.farthest_point_sample(S, T(xyz))), oracle.farthest_point_sample(S, xyz), name)


def test_pool_unpool_at_cfgT_vs_oracle(pkg, oracle):
    B, N, K, C, S = 2, 10000, 64, 128, 2500
    xyz = make_cloud(907, B, N, "cube")
    r = saturating_radius(N, K)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, xyz, r, None, K)
    sel = oracle.farthest_point_sample(S, xyz)
    bi = np.arange(B)[:, None]
    pidx, pcnt = idx[bi, sel], cnt[bi, sel]
    x = np.round(features(908, B, N, C) * 4).astype(np.float32)
    mo, mi = pkg.tf_pool3d.max_pool3d(T(x), T(pidx), T(pcnt))
    wo, wi = oracle.max_pool3d(x, pidx, pcnt)
    assert_equal(A(mo), wo); assert_equal(A(mi), wi)
    coarse = xyz[bi, sel]
    uidx, ucnt, udst = oracle.build_sphere_neighbor(coarse, xyz, 2 * r, None, K)
    gi, gc, gd = pkg.tf_nnquery.build_sphere_neighbor(T(coarse), T(xyz), radius=2 * r, nnsample=K)
    assert_equal(A(gi), uidx); assert_equal(A(gc), ucnt); assert_equal(A(gd), udst)
    xc = features(909, B, S, C)
    assert_close(A(pkg.tf_unpool3d.mean_interpolate(T(xc), gi, gc)), oracle.mean_interpolate(xc, uidx, ucnt, 1), 1e-5)
