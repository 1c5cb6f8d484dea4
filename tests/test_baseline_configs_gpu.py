"""Parity on BASELINE.json's configurations at FULL size against the unmodified reference kernels (oracle/_ref) and the
fp64 CPU oracle: the whole Cfg-T batch (B=32), the ScanNet stress shape cfg5 (B=4, N=65 536, C=256: grid ball query,
bins, convolution forward / backward, the 65 536 -> 16 384 FPS level), and the RNG-driven samplers of build_graph
('IDS' / 'random') checked against their closed-form distributions."""
import numpy as np
import pytest
import torch

from common import assert_close, assert_close_terms, assert_equal, saturating_radius

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
A = lambda t: t.detach().cpu().numpy()


def _graph_both(pkg, ref, xyz, radius, K, kernel):
    idx, cnt, dst = pkg.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=radius, nnsample=K)
    ri, rc, rd = ref.build_sphere_neighbor(xyz, xyz, radius, None, K)
    assert_equal(A(cnt), A(rc), "nn_count vs reference kernel")
    assert_equal(A(idx), A(ri), "nn_index vs reference kernel")
    assert_equal(A(dst), A(rd), "nn_dist vs reference kernel")
    filt = pkg.tf_buildkernel.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, kernel=kernel)
    assert_equal(A(filt), A(ref.spherical_kernel(xyz, xyz, ri, rc, rd, radius, kernel)), "filt_index vs reference kernel")
    return idx, cnt, dst, filt


def _conv_both(pkg, ref, oracle, x, W, go, idx, cnt, filt, what, with_oracle=True):
    out = pkg.tf_conv3d._forward(x, W, idx, cnt, filt)
    gi, gf = pkg.tf_conv3d.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    rout = ref.depthwise_conv3d(x, W, idx, cnt, filt)
    rgi, rgf = ref.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    if with_oracle:                                   # fp64 truth + the per-element sum of magnitudes of its summands
        n = lambda t: A(t)
        tout = oracle.depthwise_conv3d(n(x), n(W), n(idx), n(cnt), n(filt), 1)
        aout = oracle.depthwise_conv3d(np.abs(n(x)), np.abs(n(W)), n(idx), n(cnt), n(filt), 1)
        ti, tf = oracle.depthwise_conv3d_grad(n(x), n(W), n(go), n(idx), n(cnt), n(filt))
        ai, af = oracle.depthwise_conv3d_grad(np.abs(n(x)), np.abs(n(W)), np.abs(n(go)), n(idx), n(cnt), n(filt))
        for name, new, old, truth, terms in (("forward", out, rout, tout, aout), ("grad_input", gi, rgi, ti, ai),
                                             ("grad_filter", gf, rgf, tf, af)):
            w, r = assert_close_terms(A(new), truth, terms, 1e-5, "%s %s (this library)" % (what, name))
            rw, rr = assert_close_terms(A(old), truth, terms, 1e-5, "%s %s (reference kernel)" % (what, name))
            assert_close(A(new), truth, 1e-5, "%s %s vs fp64 oracle" % (what, name))
            print("%s %s: err / sum|terms| %.1e (reference kernel %.1e); plain relative error on elements > 1e-3 of the "
                  "scale %.1e (reference kernel %.1e)" % (what, name, w, rw, r, rr))
    else:
        assert_close(A(out), A(rout), 1e-5, what + " forward vs reference kernel")
        assert_close(A(gi), A(rgi), 2e-5, what + " grad_input vs reference kernel")
        assert_close(A(gf), A(rgf), 2e-5, what + " grad_filter vs reference kernel")


def test_cfgT_whole_batch_vs_reference_kernels_and_oracle(pkg, ref, oracle):
    """the headline batch itself: B=32, N=M=10000, K=64, C=128, r=1 -- graph bit-exact, conv fwd/bwd within 1e-5"""
    B, N, K, C = 32, 10000, 64, 128
    g = torch.Generator().manual_seed(1234 + 2)
    xyz = torch.rand(B, N, 3, generator=g).to(DEV)
    r = saturating_radius(N, K)
    idx, cnt, dst, filt = _graph_both(pkg, ref, xyz, r, K, [8, 2, 2])
    x = torch.randn(B, N, C, generator=g).to(DEV)
    W = (0.1 * torch.randn(33, C, 1, generator=g)).to(DEV)
    go = torch.randn(B, N, C, generator=g).to(DEV)
    _conv_both(pkg, ref, oracle, x, W, go, idx, cnt, filt, "Cfg-T B=32")


def test_cfg5_scannet_stress_vs_reference_kernels(pkg, ref, oracle, tune):
    """B=4, N=65536, K=64, C=256: ball query by the packed scan AND (forced) by the cell grid; convolution through 2 channel chunks"""
    B, N, K, C = 4, 65536, 64, 256
    g = torch.Generator().manual_seed(1234 + 5)
    xyz = torch.rand(B, N, 3, generator=g).to(DEV)
    r = saturating_radius(N, K)
    idx, cnt, dst, filt = _graph_both(pkg, ref, xyz, r, K, [8, 2, 2])
    tune(SPH3D_NNQUERY_GRID="2")                                                             # the same graph through the grid path
    assert pkg._lib.lib().sph3d_build_sphere_neighbor_workspace_bytes(B, N, N, K) > 0
    gidx, gcnt, gdst = pkg.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=r, nnsample=K)
    assert_equal(A(gidx), A(idx), "grid path nn_index"); assert_equal(A(gcnt), A(cnt)); assert_equal(A(gdst), A(dst))
    tune(SPH3D_NNQUERY_GRID=None)
    x = torch.randn(B, N, C, generator=g).to(DEV)
    W = (0.1 * torch.randn(33, C, 1, generator=g)).to(DEV)
    go = torch.randn(B, N, C, generator=g).to(DEV)
    _conv_both(pkg, ref, oracle, x, W, go, idx, cnt, filt, "cfg5 B=4 N=65536")
    sel = pkg.tf_sample.farthest_point_sample(16384, xyz)
    assert_equal(A(sel), A(ref.farthest_point_sample(16384, xyz)), "FPS 65536 -> 16384 vs reference kernel")
    # the next level of the hierarchy on the sampled cloud (16384 -> 4096), pooled with the reference's argmax rule
    bi = torch.arange(B, device=DEV)[:, None]
    pidx, pcnt = idx[bi, sel.long()].contiguous(), cnt[bi, sel.long()].contiguous()
    po, pi = pkg.tf_pool3d.max_pool3d(x, pidx, pcnt)
    ro, ri = ref.max_pool3d(x, pidx, pcnt)
    assert_equal(A(po), A(ro), "max-pool values"); assert_equal(A(pi), A(ri), "max-pool argmax")


@pytest.mark.parametrize("method", ["IDS", "random"])
def test_build_graph_with_random_samplers(pkg, oracle, method):
    """build_graph(sample_method='IDS' | 'random') on the GPU: shapes, ranges, batch ids, and the graph it returns"""
    u = pkg.sph3gcn_util
    B, N, K, S = 3, 2000, 32, 500
    g = torch.Generator().manual_seed(321)
    xyz = torch.rand(B, N, 3, generator=g).to(DEV)
    torch.manual_seed(99)
    idx, cnt, dst, indices = u.build_graph(xyz, 0.12, K, S, sample_method=method)
    oi, oc, od = oracle.build_sphere_neighbor(A(xyz), A(xyz), 0.12, None, K)
    assert_equal(A(idx), oi); assert_equal(A(cnt), oc); assert_equal(A(dst), od)
    ind = A(indices)
    assert ind.shape == (B, S, 2) and ind.dtype == np.int32
    assert (ind[..., 0] == np.arange(B)[:, None]).all()
    assert ind[..., 1].min() >= 0 and ind[..., 1].max() < N
    if method == "IDS":                                       # Gumbel top-k draws WITHOUT replacement
        assert all(len(np.unique(ind[b, :, 1])) == S for b in range(B))
    else:                                                     # uniform ints WITH replacement: 500 of 2000 -> ~58 repeats expected
        assert any(len(np.unique(ind[b, :, 1])) < S for b in range(B))
    pooled = u.gather_nd(xyz, indices)                        # the models' row selection works on these indices
    assert_equal(A(pooled), A(xyz)[np.arange(B)[:, None], ind[..., 1]])


def _chi2_ok(counts, probs, draws, sigmas=6.0):
    """Pearson chi-square against expected = draws * probs; accepted within `sigmas` standard deviations of its mean"""
    exp = draws * probs
    chi2 = float(((counts - exp) ** 2 / exp).sum())
    dof = len(probs) - 1
    return abs(chi2 - dof) <= sigmas * np.sqrt(2.0 * dof), chi2, dof


def test_inverse_density_sample_distribution(pkg):
    """tf_sample.py:27-41: Gumbel-max top-k on log(p).  Closed form (Plackett-Luce): the first pick is i with probability
    p_i / sum(p); given the first pick i the second is j with probability p_j / (sum(p) - p_i)."""
    n, draws = 12, 200000
    p = torch.tensor([1.0, 2.0, 0.5, 4.0, 1.5, 3.0, 0.25, 2.5, 1.0, 0.75, 5.0, 1.25], device=DEV)
    torch.manual_seed(2024)
    picks = A(pkg.tf_sample.inverse_density_sample(3, p[None, :].expand(draws, n).contiguous()))
    assert picks.shape == (draws, 3) and picks.dtype == np.int32
    assert (picks[:, 0] != picks[:, 1]).all() and (picks[:, 1] != picks[:, 2]).all() and (picks[:, 0] != picks[:, 2]).all()
    pn = A(p).astype(np.float64)
    ok, chi2, dof = _chi2_ok(np.bincount(picks[:, 0], minlength=n), pn / pn.sum(), draws)
    assert ok, ("first pick", chi2, dof)
    first = 10                                                # condition on the most likely first pick
    sub = picks[picks[:, 0] == first]
    rest = np.delete(np.arange(n), first)
    ok, chi2, dof = _chi2_ok(np.bincount(sub[:, 1], minlength=n)[rest], pn[rest] / pn[rest].sum(), len(sub))
    assert ok, ("second pick | first", chi2, dof)


def test_random_sample_distribution(pkg):
    """tf_sample.py:44-49: uniform integers in [0, N) with replacement"""
    B, N, S = 64, 50, 4000
    torch.manual_seed(7)
    picks = A(pkg.tf_sample.random_sample(S, torch.zeros(B, N, 3, device=DEV)))
    assert picks.shape == (B, S) and picks.dtype == np.int32 and picks.min() >= 0 and picks.max() < N
    ok, chi2, dof = _chi2_ok(np.bincount(picks.reshape(-1), minlength=N), np.full(N, 1.0 / N), B * S)
    assert ok, (chi2, dof)
    # rows are independent draws: two rows are not identical, and consecutive picks are uncorrelated
    assert not (picks[0] == picks[1]).all()
    a, b = picks[:, :-1].reshape(-1).astype(np.float64), picks[:, 1:].reshape(-1).astype(np.float64)
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.01
