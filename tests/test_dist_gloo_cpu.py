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
