"""Streaming gather form vs scatter form of the unpool / avg-pool gradients at the Cfg-T pooling shapes (B=32, fine 10 000,
coarse 2 500, K=64, C=128) and the S3DIS ones, for every compiled CTA shape: results must agree (the two forms sum in
different orders: 1e-5 of the scale) and the times go to gpurun_out/r2_pool_stream.json.

    python profiles/check_pool_stream.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import sph3d_gcn_b200 as S


def env(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    S._lib.reload_tunables()


def timeit(fn, iters=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters, 4)


dev = torch.device("cuda", 0)
rows = []
for name, B, N, Sn, K, C in (("cfgT", 32, 10000, 2500, 64, 128), ("s3dis", 8, 8192, 2048, 64, 128)):
    g = torch.Generator().manual_seed(4321)
    xyz = torch.rand(B, N, 3, generator=g).to(dev)
    radius = bench.saturating_radius(N, K)
    idx, cnt, dst = S.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=radius, nnsample=K)
    sel = S.tf_sample.farthest_point_sample(Sn, xyz)
    bi = torch.arange(B, device=dev)[:, None]
    coarse = xyz[bi, sel.long()].contiguous()
    uidx, ucnt, udst = S.tf_nnquery.build_sphere_neighbor(coarse, xyz, radius=2 * radius, nnsample=K)
    xc = torch.randn(B, Sn, C, generator=g).to(dev)
    gof = torch.randn(B, N, C, generator=g).to(dev)
    w = ((udst + 1e-7) / (udst.sum(-1, keepdim=True) + 1e-7)).contiguous()
    pidx, pcnt = idx[bi, sel.long()].contiguous(), cnt[bi, sel.long()].contiguous()
    x = torch.randn(B, N, C, generator=g).to(dev)
    gop = torch.randn(B, Sn, C, generator=g).to(dev)
    ops = {"mean_interpolate_grad": lambda: S.tf_unpool3d.mean_interpolate_grad(xc, gof, uidx, ucnt),
           "weighted_interpolate_grad": lambda: S.tf_unpool3d.weighted_interpolate_grad(xc, gof, w, uidx, ucnt),
           "avg_pool3d_grad": lambda: S.tf_pool3d.avg_pool3d_grad(x, gop, pidx, pcnt)}
    S.tf_pool3d.GATHER_FORM_GRAD = False
    S.tf_unpool3d.GATHER_FORM_GRAD = False
    ref = {k: f() for k, f in ops.items()}
    rec = {"shape": name, "B": B, "fine": N, "coarse": Sn, "K": K, "C": C, "scatter_ms": {k: timeit(f) for k, f in ops.items()}}
    S.tf_pool3d.GATHER_FORM_GRAD = True
    S.tf_unpool3d.GATHER_FORM_GRAD = True
    for shape in (0, 328, 324, 168, 1616):          # 0 = warp-per-point gather form
        env(SPH3D_POOL_STREAM=shape)
        got = {k: f() for k, f in ops.items()}
        err = {k: float((got[k] - ref[k]).abs().max() / ref[k].abs().max()) for k in ops}
        assert all(e <= 1e-5 for e in err.values()), (name, shape, err)
        rec["gather_%s_ms" % ("warp_per_point" if shape == 0 else shape)] = {k: timeit(f) for k, f in ops.items()}
        rec["max_rel_diff_%s" % shape] = err
    env(SPH3D_POOL_STREAM=None)
    rows.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "r2_pool_stream.json"), "w"), indent=1)
