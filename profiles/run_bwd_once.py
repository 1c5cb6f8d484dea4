#!/usr/bin/env python
"""Runs the Cfg-T backward (one-call form) a few times: the command ncu wraps for per-kernel times / full captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench as BM
import sph3d_gcn_b200 as S

wl = sys.argv[1] if len(sys.argv) > 1 else "cfgT"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = BM.WORKLOADS[wl]
host, radius, F = BM.make_inputs(cfg, 1234 + 2, dev, S)
d = {k: v.to(dev) for k, v in host.items() if k != "xyz"}
for _ in range(iters):
    S.tf_conv3d._forward(d["x"], d["W"], d["idx"], d["cnt"], d["filt"])
    S.tf_conv3d.depthwise_conv3d_grad(d["x"], d["W"], d["go"], d["idx"], d["cnt"], d["filt"])
torch.cuda.synchronize()
