"""GPU tests of the round-2 paths: the planned forward (graph-side bin sort), plans emitted by spherical_kernel, the
1/cnt fold of the transposed plan, the gather-form gradients of avg-pool / interpolation, the row-owned backward on
tiny problems, and the elementwise (sum-of-magnitudes) accuracy bound next to the scale-relative one."""
import numpy as np
import pytest
import torch

from common import assert_close, assert_close_terms, assert_equal, features
from test_parity_gpu import (A, CONV_CASES, POOL_CASES, T, UNPOOL_CASES, _conv_inputs, _graph, _pool_graph,
                             _unpool_graph)

pytestmark = pytest.mark.gpu

PLANNABLE = [c for c in CONV_CASES if c[5] in (1, 2) and c[6][0] * c[6][1] * c[6][2] + 1 <= 128]


@pytest.mark.parametrize("case", PLANNABLE, ids=[c[0] for c in PLANNABLE])
def test_conv_sort_words(case, pkg, oracle):
    """every 64-edge tile of a row holds its edges once, ascending bin, k order inside a bin, last-of-bin flagged"""
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    cnt = cnt.copy(); cnt[:, 3::13] = 0
    B, M, K = idx.shape
    plan = pkg.tf_conv3d.conv_sort(T(idx), T(cnt), T(filt), W.shape[0], x.shape[1])
    assert plan is not None
    words = A(plan).view(np.uint32).reshape(B, M, K)
    for b in range(B):
        for m in range(0, M, max(1, M // 97)):
            c = min(int(cnt[b, m]), K)
            for kt in range(0, c, 64):
                nt = min(64, c - kt)
                order = np.argsort(filt[b, m, kt:kt + nt], kind="stable")
                n, f = idx[b, m, kt:kt + nt][order].astype(np.uint32), filt[b, m, kt:kt + nt][order].astype(np.uint32)
                last = np.concatenate([f[1:] != f[:-1], [True]]).astype(np.uint32)
                assert_equal(words[b, m, kt:kt + nt], (n << 8) | (f << 1) | last, "%s words row (%d,%d)" % (case[0], b, m))


@pytest.mark.parametrize("case", PLANNABLE, ids=[c[0] for c in PLANNABLE])
def test_planned_forward_is_bit_identical(case, pkg, oracle, monkeypatch):
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    cnt = cnt.copy(); cnt[:, 5::11] = 0
    tx, tW, ti, tc, tf = T(x), T(W), T(idx), T(cnt), T(filt)
    one_call = pkg.tf_conv3d._forward(tx, tW, ti, tc, tf)                    # graph=None: sph3d_depthwise_conv3d
    plan = pkg.tf_conv3d.conv_sort(ti, tc, tf, W.shape[0], x.shape[1])
    for rep in range(2):
        planned = pkg.tf_conv3d.depthwise_conv3d_planned(tx, tW, tc, plan, idx.shape[2])
        assert_equal(A(planned), A(one_call), case[0] + " planned vs one-call forward")
    monkeypatch.setattr(pkg.tf_conv3d, "FORWARD_PLANS", True)
    through_op = pkg.tf_conv3d.depthwise_conv3d(tx, tW, ti, tc, tf)           # SHARE_PLANS + FORWARD_PLANS: builds + caches the words
    assert any(k[0] == "fwd" for k in tf._sph3d_plans)
    assert_equal(A(through_op), A(one_call), case[0] + " op (shared plan) vs one-call forward")
    assert_close(A(planned), oracle.depthwise_conv3d(x, W, idx, cnt, filt, mode=1), 1e-5, case[0] + " planned vs oracle")


def test_spherical_kernel_emits_the_plans(pkg, oracle, monkeypatch):
    """graph-build side: with EMIT_PLANS = "train" the bins come back with both plans attached and the convolutions
    over that graph use them (no plan kernels inside the conv calls)"""
    B, N, K, C = 2, 900, 32, 64
    xyz, q, radius, idx, cnt, dst, filt = _graph(oracle, 131, B, N, K)
    monkeypatch.setattr(pkg.tf_buildkernel, "EMIT_PLANS", "train")
    monkeypatch.setattr(pkg.tf_conv3d, "FORWARD_PLANS", True)
    ti, tc, td = T(idx), T(cnt), T(dst)
    tfilt = pkg.tf_buildkernel.spherical_kernel(T(xyz), T(xyz), ti, tc, td, radius, kernel=[8, 2, 2])
    assert_equal(A(tfilt), filt)
    kinds = sorted(k[0] for k in tfilt._sph3d_plans)
    assert kinds == ["bwd", "fwd"]
    x, W, go = features(132, B, N, C), features(133, 33, C, 1), features(134, B, N, C)
    xt, Wt = T(x).requires_grad_(True), T(W).requires_grad_(True)
    L = pkg._lib.lib()
    out = pkg.tf_conv3d.depthwise_conv3d(xt, Wt, ti, tc, tfilt)
    assert L.sph3d_last_launch_count() == 1                                   # the gather kernel alone
    out.backward(T(go))                                                       # (autograd thread: its launch counter is its own)
    assert len(tfilt._sph3d_plans) == 2                                       # nothing was rebuilt
    bplan = [v for k, v in tfilt._sph3d_plans.items() if k[0] == "bwd"][0]
    pkg.tf_conv3d.depthwise_conv3d_grad_planned(T(x), T(W), T(go), tc, bplan, K)
    assert L.sph3d_last_launch_count() == 3                                   # scaled copy + gather pass + partial reduction
    assert_close(A(out), oracle.depthwise_conv3d(x, W, idx, cnt, filt, 1), 1e-5)
    gi, gf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    assert_close(A(xt.grad), gi, 1e-5); assert_close(A(Wt.grad), gf, 1e-5)


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[5] in (1, 2)][:8], ids=[c[0] for c in CONV_CASES if c[5] in (1, 2)][:8])
def test_conv_elementwise_error_bound(case, pkg, oracle):
    """forward and both gradients within 1e-5 of the sum of the magnitudes of each element's own summands"""
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    go = features(64, x.shape[0], idx.shape[1], x.shape[2] * W.shape[2])
    ax, aW, ago = np.abs(x), np.abs(W), np.abs(go)
    out = pkg.tf_conv3d.depthwise_conv3d(T(x), T(W), T(idx), T(cnt), T(filt))
    w, r = assert_close_terms(A(out), oracle.depthwise_conv3d(x, W, idx, cnt, filt, 1),
                              oracle.depthwise_conv3d(ax, aW, idx, cnt, filt, 1), 1e-5, case[0] + " forward")
    ti, tf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    ai, af = oracle.depthwise_conv3d_grad(ax, aW, ago, idx, cnt, filt)
    gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    wi, ri = assert_close_terms(A(gi), ti, ai, 1e-5, case[0] + " grad_input")
    wf, rf = assert_close_terms(A(gf), tf, af, 1e-5, case[0] + " grad_filter")
    print("%s: err/sum|terms| fwd %.1e gI %.1e gW %.1e; plain relative (elements > 1e-3 of scale) fwd %.1e gI %.1e gW %.1e"
          % (case[0], w, wi, wf, r, ri, rf))


@pytest.mark.parametrize("bm", [1, 2, 3, 5])
def test_row_owned_backward_on_tiny_problems(bm, pkg, oracle, tune):
    """B*M in {1,2,3,5} rows, C = 64, r = 2: the CTA must stay a whole number of warp groups"""
    tune(SPH3D_BWD_ALGO="1")
    B, N, K, C, r = 1, 200, 32, 64, 2
    xyz, q, radius, idx, cnt, dst, filt = _graph(oracle, 141, B, N, K, radius=0.3)
    idx, cnt, filt = idx[:, :bm].copy(), cnt[:, :bm].copy(), filt[:, :bm].copy()
    x, W, go = features(142, B, N, C), features(143, 33, C, r), features(144, B, bm, C * r)
    ti, tf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    for rep in range(2):
        gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
        assert_close(A(gi), ti, 1e-5, "tiny grad_input"); assert_close(A(gf), tf, 1e-5, "tiny grad_filter")


@pytest.mark.parametrize("gather", [True, False], ids=["gather_form", "scatter_form"])
@pytest.mark.parametrize("case", POOL_CASES, ids=[c[0] for c in POOL_CASES])
def test_avg_pool_grad_forms(case, gather, pkg, oracle, monkeypatch, ref):
    name, B, N, M, K, C = case
    monkeypatch.setattr(pkg.tf_pool3d, "GATHER_FORM_GRAD", gather)
    idx, cnt, dst = _pool_graph(oracle, 101, B, N, M, K)
    cnt = cnt.copy(); cnt[:, 2::9] = 0
    x, go = features(102, B, N, C), features(103, B, M, C)
    want = oracle.avg_pool3d_grad(x, go, idx, cnt)
    for rep in range(2):
        got = pkg.tf_pool3d.avg_pool3d_grad(T(x), T(go), T(idx), T(cnt))
        assert_close(A(got), want, 1e-5, name + " avg-pool grad")
    assert pkg._lib.lib().sph3d_last_launch_count() == ((6 if C % 4 == 0 else 5) if gather else 1)   # plan (4) + degree order + streaming gather
    if ref is not None:
        assert_close(A(ref.avg_pool3d_grad(T(x), T(go), T(idx), T(cnt))), want, 1e-5, "reference kernel vs oracle")


@pytest.mark.parametrize("gather", [True, False], ids=["gather_form", "scatter_form"])
@pytest.mark.parametrize("case", UNPOOL_CASES, ids=[c[0] for c in UNPOOL_CASES])
def test_interpolate_grad_forms(case, gather, pkg, oracle, monkeypatch):
    name, B, Mc, Nf, K, C = case
    monkeypatch.setattr(pkg.tf_unpool3d, "GATHER_FORM_GRAD", gather)
    idx, cnt, dst = _unpool_graph(oracle, 111, B, Mc, Nf, K)
    x, go = features(112, B, Mc, C), features(113, B, Nf, C)
    w = ((dst + 1e-7) / (dst.sum(-1, keepdims=True) + 1e-7)).astype(np.float32)
    want_m = oracle.mean_interpolate_grad(x, go, idx, cnt)
    want_w = oracle.weighted_interpolate_grad(x, go, w, idx, cnt)
    for rep in range(2):
        assert_close(A(pkg.tf_unpool3d.mean_interpolate_grad(T(x), T(go), T(idx), T(cnt))), want_m, 1e-5, name + " mean grad")
        assert_close(A(pkg.tf_unpool3d.weighted_interpolate_grad(T(x), T(go), T(w), T(idx), T(cnt))), want_w, 1e-5,
                     name + " weighted grad")


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[0] in ("c128_r1", "c64_r2", "k156_tiles")], ids=["c128_r1", "c64_r2", "k156_tiles"])
def test_transposed_backward_with_folded_scale(case, pkg, oracle, tune):
    """SPH3D_BWDT_FOLD=1: the plan entries carry nn_count, the gather kernel reads grad_output directly"""
    tune(SPH3D_BWDT_FOLD="1", SPH3D_BWD_ALGO="2")
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    cnt = cnt.copy(); cnt[:, 4::7] = np.maximum(cnt[:, 4::7] // 3, 1)
    go = features(65, x.shape[0], idx.shape[1], x.shape[2] * W.shape[2])
    ti, tf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_close(A(gi), ti, 1e-5, case[0] + " folded grad_input"); assert_close(A(gf), tf, 1e-5, case[0] + " folded grad_filter")
