#!/usr/bin/env python
"""Runs the Cfg-T convolution a few times in its planned form (plans built once, outside the loop) and, with
`--one-call`, in its one-call form: the command ncu wraps for per-kernel times / full captures.

    python profiles/run_conv_once.py [workload] [iters] [--one-call]
"""
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
host, radius, F = BM.make_inputs(cfg, 1234 + 2, dev, S)
d = {k: v.to(dev) for k, v in host.items() if k != "xyz"}
C3 = S.tf_conv3d
N, K = cfg["N"], cfg["K"]
if "--one-call" in sys.argv:
    for _ in range(iters):
        C3._forward(d["x"], d["W"], d["idx"], d["cnt"], d["filt"])
        C3.depthwise_conv3d_grad(d["x"], d["W"], d["go"], d["idx"], d["cnt"], d["filt"])
else:
    fplan = C3.conv_sort(d["idx"], d["cnt"], d["filt"], F, N)
    bplan = C3.conv_transpose(d["idx"], d["cnt"], d["filt"], F, N)
    for _ in range(iters):
        C3.depthwise_conv3d_planned(d["x"], d["W"], d["cnt"], fplan, K)
        C3.depthwise_conv3d_grad_planned(d["x"], d["W"], d["go"], d["cnt"], bplan, K)
torch.cuda.synchronize()
