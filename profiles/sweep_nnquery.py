"""ball query: scan vs cell-grid path at the reference's config radii (unsaturated first chain step) and at the saturating radius.
    python profiles/sweep_nnquery.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
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


def timeit(fn, iters=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(11)
rows = []
for name, B, N, K, r in (("modelnet_l1", 32, 10000, 64, 0.1), ("modelnet_l2", 32, 2500, 64, 0.2), ("modelnet_l3", 32, 625, 64, 0.4),
                         ("s3dis_l1", 8, 8192, 64, 0.1), ("s3dis_l2", 8, 2048, 64, 0.2), ("s3dis_l3", 8, 768, 64, 0.4),
                         ("shapenet_l1", 16, 2048, 32, 0.08), ("cfgT_saturating", 32, 10000, 64, None),
                         ("cfg5_saturating", 4, 65536, 64, None)):
    xyz = torch.rand(B, N, 3, generator=g).to(dev)
    if name.startswith("modelnet"):            # unit-sphere normalised clouds: points spread over a ball of radius 1
        xyz = xyz * 2 - 1
    r = r or bench.saturating_radius(N, K)
    rec = {"name": name, "B": B, "N": N, "K": K, "radius": r}
    base = None
    for tag, v in (("auto", None), ("scan", 0), ("grid", 2)):
        env(SPH3D_NNQUERY_GRID=v)
        idx, cnt, dst = S.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=r, nnsample=K)
        if base is None:
            base = (idx.clone(), cnt.clone())
            rec["mean_cnt"] = float(cnt.float().mean())
        rec[tag + "_same"] = bool(torch.equal(idx, base[0]) and torch.equal(cnt, base[1]))
        rec[tag + "_ms"] = round(timeit(lambda: S.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=r, nnsample=K)), 4)
    rows.append(rec)
    print(json.dumps(rec), flush=True)
env(SPH3D_NNQUERY_GRID=None)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"rows": rows}, open(os.path.join(ROOT, "gpurun_out", "r2_nnquery.json"), "w"), indent=1)
