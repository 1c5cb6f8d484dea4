"""B < G (BASELINE configs[4]): the ranks that share a cloud split its query points
(utils/dist_util.query_sharded).  On one GPU the shards run one after the other: the output rows are exact slices of the
unsharded op and the shards' grad_input / grad_filter partial sums add up to the unsharded gradients -- what the group
all-reduce delivers on several ranks (the collective itself is covered by tests/test_dist_gloo_cpu.py on gloo)."""
import numpy as np
import pytest
import torch

from common import assert_close, features, make_cloud, saturating_radius

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nshards", [2, 3])
def test_query_sharded_conv_matches_the_whole_op(pkg, nshards):
    du = pkg.utils.dist_util
    C3 = pkg.tf_conv3d
    B, N, K, C, r = 2, 3001, 32, 64, 1
    dev = "cuda"
    xyz = torch.from_numpy(make_cloud(91, B, N)).to(dev)
    radius = saturating_radius(N, K)
    idx, cnt, dst = pkg.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=radius, nnsample=K)
    filt = pkg.tf_buildkernel.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, kernel=[8, 2, 2])
    x = torch.from_numpy(features(92, B, N, C)).to(dev)
    W = torch.from_numpy(0.1 * features(93, 33, C, r)).to(dev)
    go = torch.from_numpy(features(94, B, N, C * r)).to(dev)

    xf, Wf = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    full = C3.depthwise_conv3d(xf, Wf, idx, cnt, filt)
    full.backward(go)

    gi = torch.zeros_like(x)
    gw = torch.zeros_like(W)
    covered = 0
    for s in range(nshards):
        xs, Ws = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
        out, (m0, m1) = du.query_sharded(lambda i, a, b, c: C3.depthwise_conv3d(i, Ws, a, b, c), xs, [idx, cnt, filt], s, nshards)
        assert out.shape == (B, m1 - m0, C * r)
        assert torch.equal(out.detach(), full.detach()[:, m0:m1]), "a row slice of the convolution is not bit-identical"
        out.backward(go[:, m0:m1].contiguous())
        gi += xs.grad
        gw += Ws.grad
        covered += m1 - m0
    assert covered == N
    assert_close(gi.cpu().numpy(), xf.grad.cpu().numpy(), 1e-5, "grad_input summed over query shards")
    assert_close(gw.cpu().numpy(), Wf.grad.cpu().numpy(), 1e-5, "grad_filter summed over query shards")


def test_cloud_shard_table(pkg):
    du = pkg.utils.dist_util
    assert [du.cloud_shard(r, 8, 4) for r in range(8)] == [(0, 1, 0, 2), (0, 1, 1, 2), (1, 2, 0, 2), (1, 2, 1, 2),
                                                          (2, 3, 0, 2), (2, 3, 1, 2), (3, 4, 0, 2), (3, 4, 1, 2)]
    assert [du.cloud_shard(r, 2, 4) for r in range(2)] == [(0, 2, 0, 1), (2, 4, 0, 1)]
    with pytest.raises(ValueError):
        du.cloud_shard(0, 6, 4)
