#!/usr/bin/env python
"""Runs the fused separable layer (csrc/sepconv.cu) and the depthwise forward it contains a few times at a bench workload's
shape: the command ncu wraps.     python profiles/run_sepconv_once.py [workload] [iters] [Cout]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench as BM
import sph3d_gcn_b200 as S

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
wl = argv[0] if len(argv) > 0 else "cfgT"
iters = int(argv[1]) if len(argv) > 1 else 2
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = BM.WORKLOADS[wl]
cout = int(argv[2]) if len(argv) > 2 else 128
host, radius, F = BM.make_inputs(cfg, 1234 + 2, dev, S)
d = {k: v.to(dev) for k, v in host.items() if k != "xyz"}
W = 0.1 * torch.randn((d["W"].shape[1] * d["W"].shape[2], cout), device=dev)
img = S.tf_sepconv.pack_weights(W)
for _ in range(iters):
    S.tf_conv3d._forward(d["x"], d["W"], d["idx"], d["cnt"], d["filt"])
    S.tf_sepconv.separable_conv3d(d["x"], d["W"], W, d["idx"], d["cnt"], d["filt"], weight_image=img)
torch.cuda.synchronize()
