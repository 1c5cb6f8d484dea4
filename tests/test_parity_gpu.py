"""GPU parity tests: the sm_100a kernels (through the host mirrors -> ctypes -> C ABI) against the CPU
oracle (oracle/sph3d_oracle.c) and, when oracle/_ref travelled to the box, against the UNMODIFIED
reference kernels, on the same seeded inputs.  Bar (BASELINE.json north_star): bit-exact for
nn_index / nn_count / nn_dist / filt_index / FPS picks / max_index; feature outputs within 1e-5
relative fp32 (tolerance in common.assert_close)."""
import numpy as np
import pytest
import torch

from common import assert_close, assert_equal, features, make_cloud, saturating_radius

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def A(t):
    return t.detach().cpu().numpy()


def _graph(oracle, seed, B, N, K, radius=None, kind="cube", M=None, kernel=(8, 2, 2)):
    xyz = make_cloud(seed, B, N, kind)
    q = xyz if M is None else make_cloud(seed + 1000, B, M, kind)
    radius = radius or saturating_radius(N, K)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, q, radius, None, K)
    filt = oracle.spherical_kernel(xyz, q, idx, cnt, dst, radius, list(kernel))
    return xyz, q, radius, idx, cnt, dst, filt


# ------------------------------------------------------------------------------------ a1 nnquery
SPHERE_CASES = [
    # name, B, N, M(None=query is database), K, radius, kind
    ("cfg1", 2, 1024, None, 20, 0.15, "cube"),
    ("saturated", 3, 2000, None, 64, None, "cube"),
    ("chain_b_gt_32_m_gt_1024", 34, 300, 1500, 8, 0.2, "cube"),      # Q1: i/32 and j/1024 chain steps
    ("shell", 2, 3000, None, 32, 0.1, "shell"),
    ("grid_ties_on_radius", 2, 700, None, 16, 0.25, "grid"),         # Q3: points exactly at radius / in the band
    ("decoder_retries", 3, 40, 2500, 16, 0.05, "cube"),              # Q1: sparse database -> empty passes, retries
    ("tiny", 1, 5, 3, 4, 0.5, "cube"),
    ("k1", 2, 100, None, 1, 0.3, "cube"),
    ("k_gt_n", 2, 50, None, 80, 10.0, "cube"),
]


@pytest.mark.parametrize("case", SPHERE_CASES, ids=[c[0] for c in SPHERE_CASES])
def test_build_sphere_neighbor(case, pkg, oracle, ref):
    name, B, N, M, K, radius, kind = case
    xyz = make_cloud(11, B, N, kind)
    q = xyz if M is None else make_cloud(12, B, M, kind)
    radius = radius or saturating_radius(N, K)
    oi, oc, od = oracle.build_sphere_neighbor(xyz, q, radius, None, K)
    gi, gc, gd = pkg.tf_nnquery.build_sphere_neighbor(T(xyz), T(q), radius=radius, nnsample=K)
    assert_equal(A(gc), oc, name + " nn_count vs oracle")
    assert_equal(A(gi), oi, name + " nn_index vs oracle")
    assert_equal(A(gd), od, name + " nn_dist vs oracle")
    assert (oc >= 1).all()
    if ref is not None:
        ri, rc, rd = ref.build_sphere_neighbor(T(xyz), T(q), radius, None, K)
        assert_equal(A(gc), A(rc), name + " nn_count vs reference kernel")
        assert_equal(A(gi), A(ri), name + " nn_index vs reference kernel")
        assert_equal(A(gd), A(rd), name + " nn_dist vs reference kernel")


def test_sphere_dilation_and_extra_columns(pkg, oracle):
    xyz = make_cloud(5, 2, 500, "cube")
    feat = np.concatenate([xyz, features(6, 2, 500, 4)], axis=2)          # (B,N,3+x): sliced to xyz
    oi, oc, od = oracle.build_sphere_neighbor(feat, feat, 0.1, 2.0, 24)
    gi, gc, gd = pkg.tf_nnquery.build_sphere_neighbor(T(feat), T(feat), radius=0.1, dilation_rate=2.0, nnsample=24)
    assert_equal(A(gi), oi); assert_equal(A(gc), oc); assert_equal(A(gd), od)


def test_sphere_rows_sorted_and_padded(pkg):
    """size-independent properties at a larger size: ascending neighbour ids, zero padding, self included."""
    B, N, K = 4, 10000, 64
    xyz = T(make_cloud(21, B, N, "cube"))
    r = saturating_radius(N, K) * 0.8
    idx, cnt, dst = pkg.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=r, nnsample=K)
    idx, cnt, dst = A(idx), A(cnt), A(dst)
    k = np.arange(K)[None, None, :]
    valid = k < cnt[..., None]
    assert (cnt >= 1).all() and (cnt <= K).all()
    assert (idx[~valid] == 0).all() and (dst[~valid] == 0).all()
    d = np.diff(idx, axis=-1)
    assert (d[valid[..., 1:]] > 0).all()
    # when the row is not truncated the point itself is a neighbour at distance 0
    full = cnt < K
    me = np.broadcast_to(np.arange(N)[None, :, None], idx.shape)
    assert ((idx == me) & valid).any(axis=-1)[full].all()


CUBE_CASES = [("cube", 2, 800, None, 27, 0.3, 3), ("cube_grid", 2, 600, 300, 10, 0.25, 4), ("cube_empty", 1, 50, 200, 8, 0.01, 3)]


@pytest.mark.parametrize("case", CUBE_CASES, ids=[c[0] for c in CUBE_CASES])
def test_build_cube_neighbor(case, pkg, oracle, ref):
    name, B, N, M, K, length, grid = case
    xyz = make_cloud(31, B, N, "grid" if "grid" in name else "cube")
    q = xyz if M is None else make_cloud(32, B, M, "cube")
    oi, oc = oracle.build_cube_neighbor(xyz, q, length, None, K, grid)
    gi, gc = pkg.tf_nnquery.build_cube_neighbor(T(xyz), T(q), length=length, nnsample=K, gridsize=grid)
    assert_equal(A(gc), oc, name + " count"); assert_equal(A(gi), oi, name + " index")
    if ref is not None:
        ri, rc = ref.build_cube_neighbor(T(xyz), T(q), length, None, K, grid)
        assert_equal(A(gc), A(rc)); assert_equal(A(gi), A(ri))


# --------------------------------------------------------------------------------- a3 buildkernel
KERNEL_CASES = [("k822", 2, 1500, 32, None, (8, 2, 2), "cube"), ("k821", 2, 1000, 20, 0.2, (8, 2, 1), "shell"),
                ("k823_default", 2, 1000, 24, 0.2, (8, 2, 3), "cube"), ("k_12_4_5", 1, 800, 40, 0.3, (12, 4, 5), "cube"),
                ("grid_axes", 2, 600, 30, 0.3, (8, 2, 2), "grid")]     # neighbours exactly on axes / bin boundaries


@pytest.mark.parametrize("case", KERNEL_CASES, ids=[c[0] for c in KERNEL_CASES])
def test_spherical_kernel(case, pkg, oracle, ref):
    name, B, N, K, radius, kernel, kind = case
    xyz, q, radius, idx, cnt, dst, filt = _graph(oracle, 41, B, N, K, radius, kind, None, kernel)
    g = pkg.tf_buildkernel.spherical_kernel(T(xyz), T(q), T(idx), T(cnt), T(dst), radius, kernel=list(kernel))
    if ref is not None:
        r = ref.spherical_kernel(T(xyz), T(q), T(idx), T(cnt), T(dst), radius, list(kernel))
        assert_equal(A(g), A(r), name + " filt_index vs reference kernel")
        assert_equal(filt, A(r), name + " ORACLE filt_index vs reference kernel")
    assert_equal(A(g), filt, name + " filt_index vs oracle")
    F = kernel[0] * kernel[1] * kernel[2] + 1
    assert A(g).min() >= 0 and A(g).max() < F


def test_spherical_kernel_global_graph(pkg, oracle, ref):
    """classification head: one centroid query, K = N, radius 100, kernel [8,2,1] (SPH3D_modelnet.py:86-93)."""
    B, N = 3, 156
    xyz = make_cloud(51, B, N, "shell")
    q = xyz.mean(axis=1, keepdims=True).astype(np.float32)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, q, 100.0, None, N)
    gi, gc, gd = pkg.sph3gcn_util.build_global_graph(T(xyz), T(q), 100.0)
    assert_equal(A(gi), idx); assert_equal(A(gc), cnt); assert_equal(A(gd), dst)
    filt = oracle.spherical_kernel(xyz, q, idx, cnt, dst, 100.0, [8, 2, 1])
    g = pkg.tf_buildkernel.spherical_kernel(T(xyz), T(q), gi, gc, gd, 100.0, kernel=[8, 2, 1])
    assert_equal(A(g), filt)
    if ref is not None:
        assert_equal(A(g), A(ref.spherical_kernel(T(xyz), T(q), gi, gc, gd, 100.0, [8, 2, 1])))


# ------------------------------------------------------------------------------------- a4/a5 conv
CONV_CASES = [
    # name, B, N, K, C, r, kernel
    ("cfg1_c3_r2", 2, 1024, 20, 3, 2, (8, 2, 2)),
    ("c128_r1", 2, 1200, 64, 128, 1, (8, 2, 2)),
    ("c128_r2", 2, 900, 64, 128, 2, (8, 2, 2)),
    ("c64_r2", 2, 1000, 32, 64, 2, (8, 2, 2)),
    ("c256_r1_chunks", 1, 700, 64, 256, 1, (8, 2, 2)),
    ("c35_r2_odd", 2, 800, 64, 35, 2, (8, 2, 2)),
    ("c67_r1_odd", 2, 600, 48, 67, 1, (8, 2, 2)),
    ("c6_r1_vec2", 2, 500, 16, 6, 1, (8, 2, 2)),
    ("c32_r3_generic", 2, 400, 24, 32, 3, (8, 2, 2)),
    ("k156_tiles", 2, 300, 156, 64, 2, (8, 2, 1)),
    ("f121_bins_gt_64", 1, 900, 64, 32, 1, (12, 2, 5)),
    ("f193_generic_kernels", 1, 500, 32, 8, 1, (16, 4, 3)),
    ("c1_r1_single_channel", 2, 300, 16, 1, 1, (8, 2, 2)),
    ("c512_r2_four_chunks", 1, 200, 32, 512, 2, (8, 2, 2)),
]


def _conv_inputs(oracle, case):
    name, B, N, K, C, r, kernel = case
    radius = saturating_radius(N, min(K, 64)) * (0.9 if "k156" not in name else 3.0)
    xyz, q, radius, idx, cnt, dst, filt = _graph(oracle, 61, B, N, K, radius, "cube", None, kernel)
    F = kernel[0] * kernel[1] * kernel[2] + 1
    x = features(62, B, N, C)
    W = (features(63, F, C, r) * 0.5).astype(np.float32)
    return x, W, idx, cnt, filt


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_depthwise_conv3d_forward(case, pkg, oracle, ref):
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    truth = oracle.depthwise_conv3d(x, W, idx, cnt, filt, mode=1)
    out = pkg.tf_conv3d.depthwise_conv3d(T(x), T(W), T(idx), T(cnt), T(filt))
    assert_close(A(out), truth, 1e-5, case[0] + " conv fwd vs fp64 oracle")
    if ref is not None:
        assert_close(A(ref.depthwise_conv3d(T(x), T(W), T(idx), T(cnt), T(filt))), truth, 1e-5, "reference kernel vs oracle")


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_depthwise_conv3d_backward(case, pkg, oracle, ref):
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    go = features(64, x.shape[0], idx.shape[1], x.shape[2] * W.shape[2])
    ti, tf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    xt, Wt = T(x).requires_grad_(True), T(W).requires_grad_(True)
    out = pkg.tf_conv3d.depthwise_conv3d(xt, Wt, T(idx), T(cnt), T(filt))
    out.backward(T(go))
    assert_close(A(xt.grad), ti, 1e-5, case[0] + " grad_input vs fp64 oracle")
    assert_close(A(Wt.grad), tf, 1e-5, case[0] + " grad_filter vs fp64 oracle")
    gi2, gf2 = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_close(A(gi2), ti, 1e-5, case[0] + " grad_input, second call")
    assert_close(A(gf2), tf, 1e-5, case[0] + " grad_filter, second call")
    if ref is not None:
        ri, rf = ref.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
        assert_close(A(ri), ti, 1e-5, "reference grad_input vs oracle")
        assert_close(A(rf), tf, 1e-5, "reference grad_filter vs oracle")


def test_conv_zero_count_rows_and_linearity(pkg, oracle):
    """rows with nn_count == 0 (possible with cube queries) give 0 and no gradient; the op is linear."""
    B, N, K, C, r = 2, 600, 32, 64, 1
    xyz, q, radius, idx, cnt, dst, filt = _graph(oracle, 71, B, N, K)
    cnt = cnt.copy(); cnt[:, ::7] = 0
    x1, x2 = features(72, B, N, C), features(73, B, N, C)
    W = features(74, 33, C, r)
    f = lambda x: A(pkg.tf_conv3d.depthwise_conv3d(T(x), T(W), T(idx), T(cnt), T(filt)))
    o1, o2, o12 = f(x1), f(x2), f(x1 + 2 * x2)
    assert (o1[:, ::7] == 0).all()
    assert_close(o12, o1.astype(np.float64) + 2 * o2.astype(np.float64), 2e-5, "linearity")
    assert_close(o1, oracle.depthwise_conv3d(x1, W, idx, cnt, filt, 1), 1e-5)
    go = features(75, B, N, C * r)
    ti, tf = oracle.depthwise_conv3d_grad(x1, W, go, idx, cnt, filt)
    gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(T(x1), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_close(A(gi), ti, 1e-5); assert_close(A(gf), tf, 1e-5)


def _plan_layout(B, N, M, F, K):
    """layout of the sph3d_conv_transpose plan (csrc/conv_bwd_t.cu, t_geom): int32 words"""
    fits = lambda G: -(-F // G) <= 127 and 24 * -(-F // G) * 512 + F * 512 + 24 * 512 <= 210 * 1024
    divs = (1, 2, 3, 4, 6, 8, 12, 24)
    G = next((G for G in divs if fits(G) and (F == 1 or (F - 1) % G == 0)), None) or next(G for G in divs if fits(G))
    SL = -(-F // G)
    FP = SL * G
    nseg = B * N * FP
    nseg_pad = -(-(nseg + 1) // 4096) * 4096
    a256 = lambda x: (x + 255) // 256 * 256
    sums_off = a256(nseg_pad * 4)
    ent_off = sums_off + a256(nseg_pad // 4096 * 4)
    sb = max(1, (SL - 1).bit_length())
    cb = 0                                       # the 1/cnt code rides in the entries only with SPH3D_BWDT_FOLD=1
    return G, SL, FP, nseg, ent_off // 4, sb, cb


TRANSPOSE_CASES = [c for c in CONV_CASES if c[0] in ("c128_r1", "c64_r2", "k156_tiles", "c6_r1_vec2")] + [
    ("f49_default_kernel_g6", 2, 700, 48, 16, 1, (8, 2, 3)), ("f9_g1", 1, 300, 16, 8, 1, (4, 2, 1))]


@pytest.mark.parametrize("canonical", [0, 1])
@pytest.mark.parametrize("case", TRANSPOSE_CASES, ids=[c[0] for c in TRANSPOSE_CASES])
def test_conv_transpose_plan(case, canonical, pkg, oracle, monkeypatch, tune):
    """the transposed graph holds every edge exactly once, grouped by (input point, bin class, bin); with
    SPH3D_BWDT_SORT=1 additionally in ascending m inside a segment"""
    tune(SPH3D_BWDT_SORT=str(canonical))
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    B, N, _ = x.shape
    M, K, F = idx.shape[1], idx.shape[2], W.shape[0]
    cnt = cnt.copy(); cnt[:, 5::11] = 0                                   # some empty rows
    plan = pkg.tf_conv3d.conv_transpose(T(idx), T(cnt), T(filt), F, N)
    assert plan is not None
    plan = A(plan)
    G, SL, FP, nseg, ent_w, sb, cb = _plan_layout(B, N, M, F, K)
    seg_start = plan[:nseg + 1].astype(np.int64)
    b, m, k = np.nonzero(np.arange(K)[None, None, :] < np.minimum(cnt, K)[:, :, None])
    n, f = idx[b, m, k].astype(np.int64), filt[b, m, k].astype(np.int64)
    key = (b * N + n) * FP + (f % G) * SL + f // G
    order = np.lexsort((m, key))
    # entry = output row << (sb+cb) | (nn_count-1) << sb | bin slot: the 1/cnt scale rides in the plan when the bits allow
    cm1 = (np.minimum(cnt, K)[b[order], m[order]].astype(np.int64) - 1) << sb if cb else 0
    want_entries = (((b[order] * M + m[order]).astype(np.int64) << (sb + cb)) | cm1 | (f[order] // G)).astype(np.uint32)
    counts = np.bincount(key, minlength=nseg)
    assert_equal(seg_start, np.concatenate([[0], np.cumsum(counts)]), case[0] + " segment starts + total")
    got = plan[ent_w:ent_w + len(want_entries)].view(np.uint32)
    if not canonical:                                                      # any order inside a segment
        seg_of = np.repeat(np.arange(nseg), counts)
        got = got[np.lexsort((got, seg_of))]
    assert_equal(got, want_entries, case[0] + " entries")


@pytest.mark.parametrize("case", TRANSPOSE_CASES, ids=[c[0] for c in TRANSPOSE_CASES])
def test_depthwise_conv3d_backward_planned(case, pkg, oracle, monkeypatch, tune):
    """the split form (plan built once, reused) equals the one-call form; with canonical segment order
    (SPH3D_BWDT_SORT=1; the point-to-warp assignment is static) bit for bit in grad_filter"""
    tune(SPH3D_BWDT_SORT="1")
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    go = features(65, x.shape[0], idx.shape[1], x.shape[2] * W.shape[2])
    ti, tf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    plan = pkg.tf_conv3d.conv_transpose(T(idx), T(cnt), T(filt), W.shape[0], x.shape[1])
    for rep in range(2):                                                   # the plan is not consumed by a call
        gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad_planned(T(x), T(W), T(go), T(cnt), plan, idx.shape[2])
        assert_close(A(gi), ti, 1e-5, case[0] + " planned grad_input")
        assert_close(A(gf), tf, 1e-5, case[0] + " planned grad_filter")
    tune(SPH3D_BWD_ALGO="2")
    gi1, gf1 = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_equal(A(gf), A(gf1), case[0] + " planned vs one-call grad_filter")


@pytest.mark.parametrize("case", TRANSPOSE_CASES[:4], ids=[c[0] for c in TRANSPOSE_CASES[:4]])
def test_transposed_backward_32_warp_configuration(case, pkg, oracle, monkeypatch, tune):
    """SPH3D_BWDT_THREADS=1024: 32 warps per CTA, 4 gathers in flight, bin classes dividing 32"""
    tune(SPH3D_BWDT_THREADS="1024")
    tune(SPH3D_BWD_ALGO="2")
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    go = features(65, x.shape[0], idx.shape[1], x.shape[2] * W.shape[2])
    ti, tf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_close(A(gi), ti, 1e-5, case[0] + " grad_input"); assert_close(A(gf), tf, 1e-5, case[0] + " grad_filter")


@pytest.mark.parametrize("case", CONV_CASES[:6], ids=[c[0] for c in CONV_CASES[:6]])
def test_depthwise_conv3d_backward_row_owned_form(case, pkg, oracle, monkeypatch, tune):
    """SPH3D_BWD_ALGO=1 selects the row-owned kernel of conv_bwd.cu everywhere: same results, and its
    grad_filter (fixed-order reduction of register partials) is bit-reproducible run to run"""
    tune(SPH3D_BWD_ALGO="1")
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    go = features(66, x.shape[0], idx.shape[1], x.shape[2] * W.shape[2])
    ti, tf = oracle.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_close(A(gi), ti, 1e-5, case[0] + " row-owned grad_input")
    assert_close(A(gf), tf, 1e-5, case[0] + " row-owned grad_filter")
    gi2, gf2 = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_equal(A(gf2), A(gf), "row-owned grad_filter determinism")


@pytest.mark.parametrize("case", TRANSPOSE_CASES[:3], ids=[c[0] for c in TRANSPOSE_CASES[:3]])
def test_transposed_backward_canonical_order_is_deterministic(case, pkg, oracle, monkeypatch, tune):
    tune(SPH3D_BWDT_SORT="1")
    tune(SPH3D_BWD_ALGO="2")
    x, W, idx, cnt, filt = _conv_inputs(oracle, case)
    go = features(66, x.shape[0], idx.shape[1], x.shape[2] * W.shape[2])
    gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    gi2, gf2 = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_equal(A(gf2), A(gf), "grad_filter determinism with canonical segment order")


def test_conv_backward_unequal_clouds_and_garbage_padding(pkg, oracle):
    """M != N (strided queries, as the inter-level graphs), and index slots beyond nn_count holding garbage"""
    B, N, K, C, r = 2, 900, 32, 64, 2
    xyz = make_cloud(67, B, N, "cube")
    q = np.ascontiguousarray(xyz[:, ::3])
    radius = saturating_radius(N, K)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, q, radius, None, K)
    filt = oracle.spherical_kernel(xyz, q, idx, cnt, dst, radius, [8, 2, 2])
    pad = np.arange(K)[None, None, :] >= cnt[:, :, None]
    idx = idx.copy(); filt = filt.copy()
    idx[pad] = 2 ** 30; filt[pad] = -7                                     # never dereferenced
    x, W, go = features(68, B, N, C), features(69, 33, C, r), features(70, B, q.shape[1], C * r)
    ti, tf = oracle.depthwise_conv3d_grad(x, W, go, np.where(pad, 0, idx), cnt, np.where(pad, 0, filt))
    gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_close(A(gi), ti, 1e-5, "grad_input M != N"); assert_close(A(gf), tf, 1e-5, "grad_filter M != N")


def test_plan_sharing_between_the_convolutions_of_a_level(pkg, oracle):
    """two convolutions over ONE graph (what the models do per level) share one transposed graph, built at the first
    backward that needs it; an in-place edit of the graph invalidates it"""
    assert pkg.tf_conv3d.SHARE_PLANS
    B, N, K, C = 2, 700, 32, 64
    xyz, q, radius, idx, cnt, dst, filt = _graph(oracle, 76, B, N, K)
    x, W1, W2 = features(77, B, N, C), features(78, 33, C, 1), features(79, 33, C, 2)
    g1, g2 = features(80, B, N, C), features(81, B, N, 2 * C)
    ti1, tf1 = oracle.depthwise_conv3d_grad(x, W1, g1, idx, cnt, filt)
    ti2, tf2 = oracle.depthwise_conv3d_grad(x, W2, g2, idx, cnt, filt)
    tidx, tcnt, tfilt = T(idx), T(cnt), T(filt)
    xt, W1t, W2t = T(x).requires_grad_(True), T(W1).requires_grad_(True), T(W2).requires_grad_(True)
    o1 = pkg.tf_conv3d.depthwise_conv3d(xt, W1t, tidx, tcnt, tfilt)
    o2 = pkg.tf_conv3d.depthwise_conv3d(xt, W2t, tidx, tcnt, tfilt)
    (o1 * T(g1)).sum().backward(retain_graph=True)
    bwd_plans = lambda: [v for k, v in getattr(tfilt, "_sph3d_plans").items() if k[0] == "bwd"]
    assert len(bwd_plans()) == 1                                             # (the forward's sorted edge list sits beside it)
    plan_obj = bwd_plans()[0]
    gx1 = xt.grad.clone(); xt.grad = None
    (o2 * T(g2)).sum().backward()
    assert len(bwd_plans()) == 1 and bwd_plans()[0] is plan_obj             # reused, not rebuilt (C*r = 128: planned form)
    assert_close(A(gx1), ti1, 1e-5, "shared plan: grad_input conv 1"); assert_close(A(W1t.grad), tf1, 1e-5, "grad_filter conv 1")
    assert_close(A(xt.grad), ti2, 1e-5, "shared plan: grad_input conv 2"); assert_close(A(W2t.grad), tf2, 1e-5, "grad_filter conv 2")
    # edit the graph in place: the stale plan must not be used
    cnt2 = cnt.copy(); cnt2[:, ::3] = np.maximum(cnt2[:, ::3] // 2, 1)
    tcnt.copy_(T(cnt2))
    ti3, tf3 = oracle.depthwise_conv3d_grad(x, W1, g1, idx, cnt2, filt)
    xt3 = T(x).requires_grad_(True); W3t = T(W1).requires_grad_(True)
    (pkg.tf_conv3d.depthwise_conv3d(xt3, W3t, tidx, tcnt, tfilt) * T(g1)).sum().backward()
    assert len(bwd_plans()) == 1 and bwd_plans()[0] is not plan_obj
    assert_close(A(xt3.grad), ti3, 1e-5, "rebuilt plan: grad_input"); assert_close(A(W3t.grad), tf3, 1e-5, "rebuilt plan: grad_filter")


# ------------------------------------------------------------------------------------------ a6 FPS
FPS_CASES = [("n1024", 3, 1024, 256, "cube"), ("n2048_p2", 2, 2048, 512, "shell"), ("n5000_p5", 2, 5000, 300, "cube"),
             ("n8192_p8", 2, 8192, 2048, "cube"), ("n10000_cs2", 2, 10000, 625, "shell"), ("n20000_cs4", 1, 20000, 200, "cube"),
             ("n70000_cs8", 1, 70000, 100, "cube"), ("n100000_global", 1, 100000, 40, "cube"),
             ("grid_ties", 2, 3000, 400, "grid"), ("b33", 33, 700, 64, "cube"), ("m_gt_n", 2, 40, 60, "cube"), ("n7", 1, 7, 5, "cube")]


@pytest.mark.parametrize("case", FPS_CASES, ids=[c[0] for c in FPS_CASES])
def test_farthest_point_sample(case, pkg, oracle, ref):
    name, B, N, S, kind = case
    xyz = make_cloud(81, B, N, kind)
    want = oracle.farthest_point_sample(S, xyz)
    got = A(pkg.tf_sample.farthest_point_sample(S, T(xyz)))
    assert_equal(got, want, name + " FPS vs oracle")
    assert (got[:, 0] == 0).all()
    if ref is not None:
        assert_equal(got, A(ref.farthest_point_sample(S, T(xyz))), name + " FPS vs reference kernel")


# ------------------------------------------------------------------------------ a8-a11 pool/unpool
POOL_CASES = [("c128", 2, 1500, 400, 64, 128), ("c64", 2, 1000, 300, 32, 64), ("c35", 2, 800, 200, 40, 35), ("c3", 2, 500, 100, 20, 3),
              ("c256", 1, 600, 150, 64, 256), ("k100", 1, 400, 50, 100, 16)]


def _pool_graph(oracle, seed, B, N, M, K):
    """inter graph the way the models build it: FPS rows of the intra graph (SPH3D_s3dis.py:68-72)."""
    xyz = make_cloud(seed, B, N, "cube")
    r = saturating_radius(N, min(K, 64)) * (1.0 if K <= 64 else 2.0)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, xyz, r, None, K)
    sel = oracle.farthest_point_sample(M, xyz)
    bi = np.arange(B)[:, None]
    return idx[bi, sel], cnt[bi, sel], dst[bi, sel]


@pytest.mark.parametrize("case", POOL_CASES, ids=[c[0] for c in POOL_CASES])
def test_max_pool3d(case, pkg, oracle, ref):
    name, B, N, M, K, C = case
    idx, cnt, dst = _pool_graph(oracle, 91, B, N, M, K)
    # coarse-valued features -> many exact ties: argmax must follow the earliest-neighbour rule (Q11)
    x = np.round(features(92, B, N, C) * 2).astype(np.float32)
    wo, wi = oracle.max_pool3d(x, idx, cnt)
    xt = T(x).requires_grad_(True)
    out, mi = pkg.tf_pool3d.max_pool3d(xt, T(idx), T(cnt))
    assert_equal(A(out), wo, name + " max value"); assert_equal(A(mi), wi, name + " max_index")
    go = features(93, B, M, C)
    out.backward(T(go))
    assert_close(A(xt.grad), oracle.max_pool3d_grad(x, go, wi), 1e-5, name + " max grad")
    if ref is not None:
        ro, ri = ref.max_pool3d(T(x), T(idx), T(cnt))
        assert_equal(A(out), A(ro)); assert_equal(A(mi), A(ri))


@pytest.mark.parametrize("case", POOL_CASES, ids=[c[0] for c in POOL_CASES])
def test_avg_pool3d(case, pkg, oracle, ref):
    name, B, N, M, K, C = case
    idx, cnt, dst = _pool_graph(oracle, 101, B, N, M, K)
    x, go = features(102, B, N, C), features(103, B, M, C)
    xt = T(x).requires_grad_(True)
    out = pkg.tf_pool3d.avg_pool3d(xt, T(idx), T(cnt))
    assert_close(A(out), oracle.avg_pool3d(x, idx, cnt, 1), 1e-5, name + " avg fwd")
    out.backward(T(go))
    assert_close(A(xt.grad), oracle.avg_pool3d_grad(x, go, idx, cnt), 1e-5, name + " avg grad")
    if ref is not None:
        assert_close(A(ref.avg_pool3d(T(x), T(idx), T(cnt))), oracle.avg_pool3d(x, idx, cnt, 1), 1e-5)


UNPOOL_CASES = [("c128", 2, 300, 1200, 32, 128), ("c64", 2, 128, 384, 64, 64), ("c35", 2, 100, 500, 16, 35), ("c512", 1, 64, 200, 32, 512)]


def _unpool_graph(oracle, seed, B, Mc, Nf, K):
    """decoder inter graph: database = coarse cloud, query = fine cloud (sph3gcn_util.py:55)."""
    fine = make_cloud(seed, B, Nf, "cube")
    sel = oracle.farthest_point_sample(Mc, fine)
    coarse = fine[np.arange(B)[:, None], sel]
    r = saturating_radius(Mc, min(K, Mc // 4))
    return oracle.build_sphere_neighbor(coarse, fine, r, None, K)


@pytest.mark.parametrize("case", UNPOOL_CASES, ids=[c[0] for c in UNPOOL_CASES])
def test_unpool3d(case, pkg, oracle, ref):
    name, B, Mc, Nf, K, C = case
    idx, cnt, dst = _unpool_graph(oracle, 111, B, Mc, Nf, K)
    x, go = features(112, B, Mc, C), features(113, B, Nf, C)
    # mean
    xt = T(x).requires_grad_(True)
    out = pkg.tf_unpool3d.mean_interpolate(xt, T(idx), T(cnt))
    assert_close(A(out), oracle.mean_interpolate(x, idx, cnt, 1), 1e-5, name + " mean fwd")
    out.backward(T(go))
    assert_close(A(xt.grad), oracle.mean_interpolate_grad(x, go, idx, cnt), 1e-5, name + " mean grad")
    # weighted, weights as unpool3d computes them (sph3gcn_util.py:317-321)
    w = ((dst + 1e-7) / (dst.sum(-1, keepdims=True) + 1e-7)).astype(np.float32)
    xt2 = T(x).requires_grad_(True)
    out2 = pkg.tf_unpool3d.weighted_interpolate(xt2, T(w), T(idx), T(cnt))
    assert_close(A(out2), oracle.weighted_interpolate(x, w, idx, cnt, 1), 1e-5, name + " weighted fwd")
    assert_equal(A(out2), oracle.weighted_interpolate(x, w, idx, cnt, 0), name + " weighted fwd bit-exact vs fp32 in-order")
    out2.backward(T(go))
    assert_close(A(xt2.grad), oracle.weighted_interpolate_grad(x, go, w, idx, cnt), 1e-5, name + " weighted grad")
    if ref is not None:
        assert_close(A(ref.mean_interpolate(T(x), T(idx), T(cnt))), oracle.mean_interpolate(x, idx, cnt, 1), 1e-5)
        assert_equal(A(out2), A(ref.weighted_interpolate(T(x), T(w), T(idx), T(cnt))), "weighted vs reference kernel bits")


# ------------------------------------------------------------------------- a12 layer library slice
def test_layer_pipeline_matches_oracle(pkg, oracle):
    """One encoder level through the reference call sequence (SPH3D_s3dis.py:53-76): build_graph ->
    spherical_kernel -> separable_conv3d -> gather_nd -> pool3d, forward and backward, vs the oracle."""
    u = pkg.sph3gcn_util
    u.reset_variables()
    torch.manual_seed(0)
    B, N, K, S, C, r, Cout = 2, 1024, 20, 256, 8, 2, 16
    xyz = make_cloud(121, B, N, "cube")
    x = features(122, B, N, C)
    xyz_t, x_t = T(xyz), T(x).requires_grad_(True)
    intra_idx, intra_cnt, intra_dst, indices = u.build_graph(xyz_t, 0.15, K, S, sample_method='FPS')
    filt = u.spherical_kernel(xyz_t, xyz_t, intra_idx, intra_cnt, intra_dst, 0.15, kernel=[8, 2, 2])
    net = u.separable_conv3d(x_t, Cout, 33, r, 'conv1_1', intra_idx, intra_cnt, filt, with_bn=True, is_training=True)
    inter_idx, inter_cnt = u.gather_nd(intra_idx, indices), u.gather_nd(intra_cnt, indices)
    pooled = u.pool3d(net, inter_idx, inter_cnt, scope='pool1', method='max')
    assert pooled.shape == (B, S, Cout)
    loss = (pooled ** 2).sum()
    loss.backward()

    # oracle replay with the same weights
    oi, oc, od = oracle.build_sphere_neighbor(xyz, xyz, 0.15, None, K)
    sel = oracle.farthest_point_sample(S, xyz)
    assert_equal(A(indices[..., 1]), sel); assert_equal(A(intra_idx), oi)
    of = oracle.spherical_kernel(xyz, xyz, oi, oc, od, 0.15, [8, 2, 2])
    assert_equal(A(filt), of)
    vars_ = u.named_variables()
    Wd, Wp = A(vars_['conv1_1/depthwise_weights']), A(vars_['conv1_1/weights'])
    dw = oracle.depthwise_conv3d(x, Wd, oi, oc, of, 1)
    y = torch.from_numpy(dw).double().reshape(-1, C * r) @ torch.from_numpy(Wp).double()
    y = torch.nn.functional.elu(y)
    y = (y - y.mean(0)) / torch.sqrt(y.var(0, unbiased=False) + 1e-3)
    assert_close(A(net).reshape(-1, Cout), y.numpy(), 2e-5, "separable_conv3d (ELU before BN)")
    assert x_t.grad is not None and torch.isfinite(x_t.grad).all()
    assert vars_['conv1_1/depthwise_weights'].grad is not None


# ------------------------------------------------------------------------- stream / graph semantics
def test_ops_are_cuda_graph_capturable(pkg, oracle):
    """include/sph3d_b200.h promises stream-ordered, allocation-free, sync-free entry points: capture the
    whole layer slice (ball query, bins, FPS, conv fwd+bwd, max-pool) in a CUDA graph, replay it on new
    inputs, and compare with the oracle."""
    B, N, K, C, r, S = 2, 1024, 20, 16, 2, 128
    L = pkg._lib.lib()
    dev = torch.device(DEV)
    xyz = torch.empty(B, N, 3, device=dev); x = torch.empty(B, N, C, device=dev)
    W = T(features(131, 33, C, r) * 0.3); go = T(features(132, B, N, C * r))
    idx = torch.empty(B, N, K, dtype=torch.int32, device=dev); cnt = torch.empty(B, N, dtype=torch.int32, device=dev)
    dst = torch.empty(B, N, K, device=dev); filt = torch.empty_like(idx)
    sel = torch.empty(B, S, dtype=torch.int32, device=dev)
    out = torch.empty(B, N, C * r, device=dev); gi = torch.empty(B, N, C, device=dev); gf = torch.empty(33, C, r, device=dev)
    ws_bytes = L.sph3d_depthwise_conv3d_grad_workspace_bytes(B, N, N, 33, C, r, K)
    ws = torch.empty(max(ws_bytes // 4, 1), device=dev)
    nb = L.sph3d_build_sphere_neighbor_workspace_bytes(B, N, N, K)
    nws = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
    p = lambda t: t.data_ptr()

    def enqueue(st):
        rc = L.sph3d_build_sphere_neighbor(B, N, N, K, 0.15, p(xyz), p(xyz), p(idx), p(cnt), p(dst), p(nws), nb, st)
        rc |= L.sph3d_spherical_kernel(B, N, N, K, 8, 2, 2, 0.15, p(xyz), p(xyz), p(idx), p(cnt), p(dst), p(filt), st)
        rc |= L.sph3d_farthest_point_sample(B, N, S, p(xyz), None, 0, p(sel), st)
        rc |= L.sph3d_depthwise_conv3d(B, N, N, 33, C, r, K, p(idx), p(cnt), p(filt), p(x), p(W), p(out), st)
        rc |= L.sph3d_depthwise_conv3d_grad(B, N, N, 33, C, r, K, p(idx), p(cnt), p(filt), p(x), p(W), p(go), p(gi), p(gf),
                                            p(ws), ws_bytes, st)
        assert rc == 0

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        enqueue(side.cuda_stream)                       # warm-up outside capture (sets the smem attributes once)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        enqueue(torch.cuda.current_stream().cuda_stream)
    for seed in (141, 142):
        xyz_np, x_np = make_cloud(seed, B, N, "cube"), features(seed + 10, B, N, C)
        xyz.copy_(T(xyz_np)); x.copy_(T(x_np))
        graph.replay()
        torch.cuda.synchronize()
        oi, oc, od = oracle.build_sphere_neighbor(xyz_np, xyz_np, 0.15, None, K)
        of = oracle.spherical_kernel(xyz_np, xyz_np, oi, oc, od, 0.15, [8, 2, 2])
        assert_equal(A(idx), oi); assert_equal(A(cnt), oc); assert_equal(A(filt), of)
        assert_equal(A(sel), oracle.farthest_point_sample(S, xyz_np))
        assert_close(A(out), oracle.depthwise_conv3d(x_np, A(W), oi, oc, of, 1), 1e-5)
        ti, tf = oracle.depthwise_conv3d_grad(x_np, A(W), A(go), oi, oc, of)
        assert_close(A(gi), ti, 1e-5); assert_close(A(gf), tf, 1e-5)


@pytest.mark.parametrize("r", [1, 2])
def test_transposed_backward_cuts_hub_lists_into_parts(r, pkg, tune):
    """More than 16 384 rows per cloud: sub-lists longer than 2 048 entries are cut into overflow items (conv_bwd_t.cu,
    build_overflow_kernel).  A synthetic hub graph -- 20 000 rows that all reference the same 4 points, so every
    (point, bin class) list holds ~3 000 entries -- against a float64 evaluation of the gradient formulas, and one-call
    against planned (the table lives in the plan)."""
    B, N, M, K, C, F = 2, 4, 20000, 4, 32, 33
    rng = np.random.default_rng(77)
    idx = np.stack([np.stack([rng.permutation(N)[:K] for _ in range(M)]) for _ in range(B)]).astype(np.int32)
    cnt = rng.integers(1, K + 1, size=(B, M)).astype(np.int32)
    filt = rng.integers(0, F, size=(B, M, K)).astype(np.int32)
    x = features(78, B, N, C)
    W = (0.2 * features(79, F, C, r)).astype(np.float32)
    go = features(80, B, M, C * r)
    # gradient formulas in float64: out[b,m,c*r+j] = sum_k in[b,n_k,c] W[f_k,c,j] / cnt
    gi = np.zeros((B, N, C)); gf = np.zeros((F, C, r))
    g3 = go.reshape(B, M, C, r).astype(np.float64)
    for b in range(B):
        for k in range(K):
            live = cnt[b] > k
            m = np.nonzero(live)[0]
            n, f = idx[b, m, k], filt[b, m, k]
            gs = g3[b, m] / cnt[b, m, None, None]                       # (rows, C, r)
            np.add.at(gi[b], n, (gs * W[f].astype(np.float64)).sum(-1))
            np.add.at(gf, f, x[b, n].astype(np.float64)[:, :, None] * gs)
    tune(SPH3D_BWD_ALGO="2")
    gi1, gf1 = pkg.tf_conv3d.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    assert_close(A(gi1), gi, 1e-5, "grad_input with hub lists cut into parts")
    assert_close(A(gf1), gf, 1e-5, "grad_filter with hub lists cut into parts")
    plan = pkg.tf_conv3d.conv_transpose(T(idx), T(cnt), T(filt), F, N)
    gi2, gf2 = pkg.tf_conv3d.depthwise_conv3d_grad_planned(T(x), T(W), T(go), T(cnt), plan, K)
    assert_close(A(gi2), gi, 1e-5, "planned grad_input with hub lists cut into parts")
    assert_close(A(gf2), gf, 1e-5, "planned grad_filter with hub lists cut into parts")
