"""CPU oracle self-consistency: closed forms, quirks of SURVEY.md section 0, adjoint identities."""
import math

import numpy as np
import pytest

from common import assert_close, assert_equal, features, make_cloud, saturating_radius


def test_cuda_atan2f_restatement(oracle):
    rng = np.random.default_rng(0)
    for y, x in rng.standard_normal((4000, 2)):
        assert abs(oracle.atan2f(y, x) - math.atan2(np.float32(y), np.float32(x))) < 4e-7     # <= 2 ulp of pi
    f32 = lambda v: float(np.float32(v))
    assert oracle.atan2f(0.0, 1.0) == 0.0 and oracle.atan2f(0.0, -1.0) == f32(math.pi)
    assert oracle.atan2f(-0.0, -1.0) == -f32(math.pi)
    assert oracle.atan2f(1.0, 0.0) == f32(math.pi / 2) and oracle.atan2f(-1.0, 0.0) == -f32(math.pi / 2)
    assert oracle.atan2f(float("inf"), float("inf")) == f32(math.pi / 4)
    assert oracle.atan2f(float("inf"), float("-inf")) == f32(3 * math.pi / 4)
    assert math.isnan(oracle.atan2f(float("nan"), 1.0))


def _brute_sphere(xyz, q, radius_of, K):
    """pure-numpy restatement with an explicit per-query radius (small cases only)."""
    B, N, _ = xyz.shape
    M = q.shape[1]
    idx = np.zeros((B, M, K), np.int32); cnt = np.zeros((B, M), np.int32); dst = np.zeros((B, M, K), np.float32)
    for b in range(B):
        for j in range(M):
            d = xyz[b] - q[b, j]
            t = (d[:, 1] * d[:, 1]).astype(np.float32)
            t = np.float32(1) * (d[:, 0].astype(np.float64) * d[:, 0].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
            # (fma emulated in fp64: exact for fp32 inputs up to one final rounding)
            t = (d[:, 2].astype(np.float64) * d[:, 2].astype(np.float64) + t.astype(np.float64)).astype(np.float32)
            dist = np.sqrt(t).astype(np.float32)
            r = np.float32(radius_of(b, j))
            ok = (dist < r) & (np.abs((dist - r).astype(np.float32)).astype(np.float64) > 1e-6)
            hit = np.nonzero(ok)[0][:K]
            idx[b, j, :len(hit)] = hit; cnt[b, j] = len(hit); dst[b, j, :len(hit)] = np.sqrt(dist[hit]).astype(np.float32)
    return idx, cnt, dst


def test_sphere_radius_chain_closed_form(oracle):
    """Q1: query (i,j) searches with radius0 after t = (i//32)*ceil((M - j%1024)/1024) + j//1024 increments."""
    B, N, M, K = 34, 60, 1100, 6
    xyz, q = make_cloud(1, B, N), make_cloud(2, B, M)
    r0 = np.float32(0.6)                                   # big enough that no query ever retries

    def radius_of(b, j):
        tx = j % 1024
        t = (b // 32) * ((M - tx + 1023) // 1024) + j // 1024
        r = np.float32(r0)
        for _ in range(t):
            r = np.float32(np.float64(r) + 0.05)
        return r
    want = _brute_sphere(xyz, q, radius_of, K)
    got = oracle.build_sphere_neighbor(xyz, q, float(r0), None, K)
    for g, w, n in zip(got, want, ("idx", "cnt", "dst")):
        assert_equal(g, w, n)


def test_sphere_retry_grows_radius_and_carries_over(oracle):
    """A query with nothing in range retries with +0.05 per pass; later queries of the same CUDA thread
    (same j%1024) inherit the grown radius; other chains do not."""
    N, M, K = 4, 1100, 4
    xyz = np.zeros((1, N, 3), np.float32); xyz[0, :, 0] = [0.0, 0.01, 0.02, 0.03]
    q = np.zeros((1, M, 3), np.float32)
    q[0, 5, 0] = 0.33              # chain tx=5, step 0: nearest point 0.30 away -> empty passes until r ~ 0.35
    q[0, 5 + 1024, 0] = 0.26       # chain tx=5, step 1: inherits r ~ 0.40 -> all four points (0.23..0.26) in range
    q[0, 6 + 1024, 0] = 0.26       # chain tx=6, step 1: r = 0.15, retries to ~0.25 -> only 0.23 and 0.24 in range
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, q, 0.1, None, K)
    assert cnt[0, 5] == 4                                   # found only once r reached ~0.35
    assert cnt[0, 6] == 4                                   # unaffected chain, radius 0.1
    assert cnt[0, 5 + 1024] == 4
    assert cnt[0, 6 + 1024] == 2 and list(idx[0, 6 + 1024, :2]) == [2, 3]


def test_sphere_padding_and_sqrt_dist(oracle):
    xyz = make_cloud(3, 2, 300)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, xyz, 0.12, None, 32)
    k = np.arange(32)[None, None, :]
    pad = k >= cnt[..., None]
    assert (idx[pad] == 0).all() and (dst[pad] == 0).all()            # Q5
    b, m = 1, 17
    n = idx[b, m, cnt[b, m] - 1]
    d = np.linalg.norm(xyz[b, n].astype(np.float64) - xyz[b, m].astype(np.float64))
    assert abs(dst[b, m, cnt[b, m] - 1] - math.sqrt(d)) < 1e-6        # Q2: sqrt of the distance


def test_spherical_kernel_bins(oracle):
    """hand-placed neighbours: self -> bin 0; azimuth/elevation octants; radial bin saturates (Q8)."""
    q = np.zeros((1, 1, 3), np.float32)
    pts = np.array([[0, 0, 0], [0.1, 0.01, 0.05], [-0.1, 0.01, 0.05], [0.1, -0.01, -0.05], [0.001, 0.0005, 0.2]], np.float32)[None]
    idx = np.arange(5, dtype=np.int32)[None, None]; cnt = np.array([[5]], np.int32)
    d = np.linalg.norm(pts[0], axis=1).astype(np.float32)
    dst = np.sqrt(d).astype(np.float32)[None, None]
    f = oracle.spherical_kernel(pts, q, idx, cnt, dst, 0.3, [8, 2, 2])[0, 0]
    n, p = 8, 2
    def want(x, y, z, dist):
        th = math.atan2(y, x) + math.pi; ph = math.atan2(z, math.hypot(x, y)) + math.pi / 2
        nid = min(n - 1, int(th * n / 2 / math.pi)); pid = min(p - 1, int(ph * p / math.pi))
        qid = min(1, int(math.sqrt(dist) * 2 / 0.3))
        return qid * p * n + pid * n + nid + 1
    assert f[0] == 0
    for i in range(1, 5):
        assert f[i] == want(*pts[0, i], d[i]), i
    assert (f >= 0).all() and (f <= 32).all()


def test_fps_first_is_zero_and_distinct(oracle):
    xyz = make_cloud(4, 3, 500)
    s = oracle.farthest_point_sample(100, xyz)
    assert (s[:, 0] == 0).all()
    assert all(len(set(r)) == 100 for r in s)
    # second pick = farthest from point 0 (no ties in random data)
    d = ((xyz - xyz[:, :1]) ** 2).sum(-1)
    assert (s[:, 1] == d.argmax(1)).all()


def test_fps_tie_rule(oracle):
    """Q12: equal distances -> smallest (k mod 1024), then smallest k."""
    N = 2100
    xyz = np.zeros((1, N, 3), np.float32)
    xyz[0, 1500] = [1, 0, 0]; xyz[0, 1030] = [-1, 0, 0]; xyz[0, 2054] = [0, 1, 0]      # tids 476, 6, 6
    s = oracle.farthest_point_sample(2, xyz)
    assert s[0, 1] == 1030                                   # tid 6 beats tid 476; k=1030 beats k=2054


def test_conv_modes_agree_and_adjoint(oracle):
    B, N, K, C, r = 2, 400, 24, 5, 2
    xyz = make_cloud(5, B, N)
    rad = saturating_radius(N, K)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, xyz, rad, None, K)
    filt = oracle.spherical_kernel(xyz, xyz, idx, cnt, dst, rad, [8, 2, 2])
    x, W, go = features(6, B, N, C), features(7, 33, C, r), features(8, B, N, C * r)
    o0, o1 = oracle.depthwise_conv3d(x, W, idx, cnt, filt, 0), oracle.depthwise_conv3d(x, W, idx, cnt, filt, 1)
    assert_close(o0, o1, 1e-5)
    gi, gf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    # the op is bilinear: <go, out> = <gi, x> = <gf, W>
    s = float((go.astype(np.float64) * o1).sum())
    assert abs(float((gi.astype(np.float64) * x).sum()) - s) < 1e-4 * abs(s) + 1e-3
    assert abs(float((gf.astype(np.float64) * W).sum()) - s) < 1e-4 * abs(s) + 1e-3


def test_pool_unpool_oracle(oracle):
    B, N, K, C = 2, 300, 16, 4
    xyz = make_cloud(9, B, N)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, xyz, 0.2, None, K)
    x = np.round(features(10, B, N, C)).astype(np.float32)
    out, mi = oracle.max_pool3d(x, idx, cnt)
    b, m, c = 1, 7, 2
    nb = idx[b, m, :cnt[b, m]]
    assert out[b, m, c] == x[b, nb, c].max()
    assert mi[b, m, c] == nb[np.argmax(x[b, nb, c])]          # first maximum in k order (Q11)
    avg = oracle.avg_pool3d(x, idx, cnt, 1)
    assert abs(avg[b, m, c] - x[b, nb, c].mean()) < 1e-6
    go = features(11, B, N, C)
    ga = oracle.avg_pool3d_grad(x, go, idx, cnt)
    assert abs(float((ga.astype(np.float64) * x).sum()) - float((go.astype(np.float64) * avg).sum())) < 1e-3
    w = ((dst + 1e-7) / (dst.sum(-1, keepdims=True) + 1e-7)).astype(np.float32)
    wo = oracle.weighted_interpolate(x, w, idx, cnt, 1)
    assert abs(wo[b, m, c] - float((x[b, nb, c].astype(np.float64) * w[b, m, :len(nb)]).sum())) < 1e-6


def test_cube_neighbor_oracle(oracle):
    xyz = make_cloud(12, 1, 200)
    idx, cnt = oracle.build_cube_neighbor(xyz, xyz, 0.3, None, 16, 3)
    assert idx.shape == (1, 200, 16, 2)
    assert (idx[..., 1] >= 0).all() and (idx[..., 1] < 27).all()
    m = 3
    d = np.abs(xyz[0] - xyz[0, m])
    inside = np.nonzero((d < 0.15).all(1))[0][:16]
    assert cnt[0, m] == len(inside) and (idx[0, m, :len(inside), 0] == inside).all()
