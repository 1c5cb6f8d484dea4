"""Strip-width / group-size sweep for channel counts that do not fill a warp with 16-byte strips
(GPU box):  python profiles/sweep_bwd_r2.py       -- results summarised in DESIGN.md section 7."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "profiles")]
import torch

import bench
import sph3d_gcn_b200 as S
from sweep_conv import time_op

SHAPES = {"c64_r1": dict(B=32, N=10000, K=64, C=64, r=1, kernel=(8, 2, 2)),
          "c64_r2": dict(B=8, N=8192, K=64, C=64, r=2, kernel=(8, 2, 2)),
          "c36_r2": dict(B=32, N=10000, K=64, C=36, r=2, kernel=(8, 2, 2)),
          "c32_r1": dict(B=32, N=10000, K=64, C=32, r=1, kernel=(8, 2, 2))}
KNOBS = ("SPH3D_BWD_G", "SPH3D_BWD_VEC", "SPH3D_FWD_VEC")
for wl, cfg in SHAPES.items():
    dev = torch.device("cuda", 0)
    host, radius, F = bench.make_inputs(cfg, 1236, dev, S)
    d = {k: v.to(dev) for k, v in host.items() if k != "xyz"}
    bwd = lambda: S.tf_conv3d.depthwise_conv3d_grad(d["x"], d["W"], d["go"], d["idx"], d["cnt"], d["filt"])
    fwd = lambda: S.tf_conv3d._forward(d["x"], d["W"], d["idx"], d["cnt"], d["filt"])
    for env in ({}, {"SPH3D_FWD_VEC": "4", "SPH3D_BWD_VEC": "4"}, {"SPH3D_BWD_G": "2"}, {"SPH3D_FWD_VEC": "1", "SPH3D_BWD_VEC": "1"}):
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            print(wl, env or "default", "fwd_ms", round(time_op(fwd), 4), "bwd_ms", round(time_op(bwd), 4), flush=True)
        except Exception as e:
            print(wl, env, "failed", e)
