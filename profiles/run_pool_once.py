#!/usr/bin/env python
"""Runs the gather-form gradients of avg-pool / mean-interpolate at the Cfg-T pooling shape a few times (for ncu).
    python profiles/run_pool_once.py [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import sph3d_gcn_b200 as S

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
B, N, Sn, K, C = 32, 10000, 2500, 64, 128
g = torch.Generator().manual_seed(4321)
xyz = torch.rand(B, N, 3, generator=g).to(dev)
radius = bench.saturating_radius(N, K)
sel = S.tf_sample.farthest_point_sample(Sn, xyz)
bi = torch.arange(B, device=dev)[:, None]
coarse = xyz[bi, sel.long()].contiguous()
uidx, ucnt, udst = S.tf_nnquery.build_sphere_neighbor(coarse, xyz, radius=2 * radius, nnsample=K)
xc = torch.randn(B, Sn, C, generator=g).to(dev)
gof = torch.randn(B, N, C, generator=g).to(dev)
S.tf_unpool3d.GATHER_FORM_GRAD = True
for _ in range(iters):
    S.tf_unpool3d.mean_interpolate_grad(xc, gof, uidx, ucnt)
torch.cuda.synchronize()
