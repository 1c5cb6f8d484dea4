"""N>1 host logic on CPU: world_size-2 gloo.  Clouds shard over the batch with no data-path collective;
the only exchange is the SUM all-reduce of weight gradients through one flat bucket
(sph3d-gcn_b200/utils/dist_util.py).  The per-rank "gradient" here is the CPU oracle's grad_filter of the
rank's shard, so the test also checks the identity DP relies on: sum over shards == full-batch gradient."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker_query_shards(rank, world, port, out_dir):
    """two ranks share every cloud and split its query points: outputs are exact slices, grad_input sums over the group"""
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from common import features, make_cloud
    from importlib import import_module
    du = import_module("sph3d_gcn_b200.utils.dist_util")

    B, N, K, C, r = 1, 257, 16, 4, 2                       # one cloud, two ranks: cloud_shard -> shards 0 / 1 of 2
    assert du.cloud_shard(rank, world, B) == (0, 1, rank, 2)
    assert du.cloud_shard(1, 2, 4) == (2, 4, 0, 1) and du.cloud_shard(5, 8, 4) == (2, 3, 1, 2)
    xyz = make_cloud(601, B, N)
    idx, cnt, dst = O.build_sphere_neighbor(xyz, xyz, 0.25, None, K)
    filt = O.spherical_kernel(xyz, xyz, idx, cnt, dst, 0.25, [8, 2, 2])
    x, W, go = features(602, B, N, C), features(603, 33, C, r), features(604, B, N, C * r)

    class OracleConv(torch.autograd.Function):           # the CPU oracle as the row-wise op
        @staticmethod
        def forward(ctx, inp, a, b, c):
            ctx.save_for_backward(inp, a, b, c)
            return torch.from_numpy(O.depthwise_conv3d(inp.numpy(), W, a.numpy(), b.numpy(), c.numpy(), 1).astype(np.float32))

        @staticmethod
        def backward(ctx, g):
            inp, a, b, c = ctx.saved_tensors
            gi, _ = O.depthwise_conv3d_grad(inp.numpy(), W, g.numpy(), a.numpy(), b.numpy(), c.numpy())
            return torch.from_numpy(gi.astype(np.float32)), None, None, None

    xt = torch.from_numpy(x).requires_grad_(True)
    out, (m0, m1) = du.query_sharded(lambda i, a, b, c: OracleConv.apply(i, a, b, c), xt,
                                     [torch.from_numpy(idx), torch.from_numpy(cnt), torch.from_numpy(filt)], rank, world)
    out.backward(torch.from_numpy(go[:, m0:m1]))
    full = O.depthwise_conv3d(x, W, idx, cnt, filt, 1)
    gi_full, _ = O.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
    ok_rows = np.allclose(out.detach().numpy(), full[:, m0:m1], rtol=1e-6, atol=1e-6)
    ok_grad = np.allclose(xt.grad.numpy(), gi_full, rtol=1e-5, atol=1e-6)
    np.save(os.path.join(out_dir, "qs_%d.npy" % rank), np.array([ok_rows, ok_grad, m0, m1]))
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from common import features, make_cloud
    import sph3d_gcn_b200  # noqa: F401  (package import must work without a GPU)
    from importlib import import_module
    du = import_module("sph3d_gcn_b200.utils.dist_util")

    B, N, K, C, r = 6, 300, 16, 4, 2
    xyz = make_cloud(501, B, N)
    idx, cnt, dst = O.build_sphere_neighbor(xyz, xyz, 0.2, None, K)
    filt = O.spherical_kernel(xyz, xyz, idx, cnt, dst, 0.2, [8, 2, 2])
    x, W, go = features(502, B, N, C), features(503, 33, C, r), features(504, B, N, C * r)
    lo, hi = du.shard_bounds(B, rank, world)
    assert [t.shape[0] for t in du.shard_batch([torch.from_numpy(x), torch.from_numpy(go)], rank, world)] == [hi - lo] * 2
    gi, gf = O.depthwise_conv3d_grad(x[lo:hi], W, go[lo:hi], idx[lo:hi], cnt[lo:hi], filt[lo:hi])
    g_w, g_b = torch.from_numpy(gf.copy()), torch.full((5,), float(rank + 1))
    nbytes = du.allreduce_gradients([g_w, None, g_b])
    assert nbytes == (g_w.numel() + 5) * 4
    if rank == 0:
        _, full = O.depthwise_conv3d_grad(x, W, go, idx, cnt, filt)
        np.save(os.path.join(out_dir, "reduced.npy"), g_w.numpy()); np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "bias.npy"), g_b.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from importlib import import_module
    sys.path.insert(0, ROOT)
    du = import_module("sph3d_gcn_b200.utils.dist_util")
    for total, world in ((32, 8), (8, 8), (4, 8), (10, 3), (1, 2)):
        b = [du.shard_bounds(total, r, world) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == total
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    assert du.allreduce_gradients([torch.ones(3)]) == 0          # not initialised: no-op


@pytest.mark.timeout(180)
def test_gradient_allreduce_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    red, full = np.load(tmp_path / "reduced.npy"), np.load(tmp_path / "full.npy")
    scale = np.abs(full).max()
    assert np.abs(red - full).max() <= 1e-5 * scale
    assert (np.load(tmp_path / "bias.npy") == 3.0).all()


def _bucket_worker(rank, world, port, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sph3d_gcn_b200  # noqa: F401
    from importlib import import_module
    du = import_module("sph3d_gcn_b200.utils.dist_util")
    torch.manual_seed(0)                                          # same weights on every rank
    layers = [torch.nn.Linear(8, 16), torch.nn.Linear(16, 16), torch.nn.Linear(16, 4)]
    unused = torch.nn.Parameter(torch.ones(3))                    # never reached by backward: reduced by finish()
    params = [p for l in layers for p in l.parameters()] + [unused]
    buckets = du.GradBuckets(params, n_buckets=3, average=True)
    assert 2 <= len(buckets.bounds) <= 3 and buckets.flat.numel() == sum(p.numel() for p in params)
    g = torch.Generator().manual_seed(100)
    xs = torch.randn(world, 5, 8, generator=g)                    # the global batch; this rank owns slice `rank`
    for it in range(2):                                           # second iteration: zero() re-arms buckets and views
        buckets.zero()
        h = xs[rank]
        for l in layers:
            h = torch.tanh(l(h))
        h.sum().backward()
        nbytes = buckets.finish()
        assert nbytes == buckets.flat.numel() * 4
        assert all(p.grad.data_ptr() >= buckets.flat.data_ptr() for p in params)     # still views of the flat buffer
    if rank == 0:
        torch.save([p.grad.clone() for p in params], os.path.join(out_dir, "bucket_grads.pt"))
        torch.save((xs, [p.detach().clone() for p in params]), os.path.join(out_dir, "bucket_inputs.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_grad_buckets_world2(tmp_path):
    """flat gradient storage with views + hook-driven bucketed all-reduce == mean over ranks of the per-rank gradients"""
    world, port = 2, _free_port()
    mp.spawn(_bucket_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(tmp_path / "bucket_grads.pt")
    xs, weights = torch.load(tmp_path / "bucket_inputs.pt")
    ws = [w.clone().requires_grad_(True) for w in weights]
    total = 0
    for r in range(world):
        h = xs[r]
        for i in range(3):
            h = torch.tanh(h @ ws[2 * i].t() + ws[2 * i + 1])
        total = total + h.sum()
    (total / world).backward()
    for g, w in zip(got[:-1], ws[:-1]):
        assert torch.allclose(g, w.grad, rtol=1e-5, atol=1e-6)
    assert (got[-1] == 0).all()


def test_grad_buckets_single_process():
    sys.path.insert(0, ROOT)
    from importlib import import_module
    du = import_module("sph3d_gcn_b200.utils.dist_util")
    lin = torch.nn.Linear(4, 3)
    b = du.GradBuckets(list(lin.parameters()), n_buckets=8)
    b.zero()
    lin(torch.ones(2, 4)).sum().backward()
    assert b.finish() == 0 and torch.allclose(lin.bias.grad, torch.full((3,), 2.0))
    lin.weight.grad = None
    b.zero()                                                       # re-attaches the views
    lin(torch.ones(2, 4)).sum().backward()
    assert lin.weight.grad.data_ptr() >= b.flat.data_ptr() and torch.allclose(lin.weight.grad, torch.full((3, 4), 2.0))
    b.close()


def test_query_shards_of_one_cloud_gloo_world2(tmp_path):
    """B < G: ranks that share a cloud split its query rows; the only exchange is the SUM of grad_input"""
    port = _free_port()
    mp.spawn(_worker_query_shards, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "qs_0.npy"), np.load(tmp_path / "qs_1.npy")
    assert a[0] and a[1] and b[0] and b[1]
    assert (a[2], a[3], b[2], b[3]) == (0, 129, 129, 257)
