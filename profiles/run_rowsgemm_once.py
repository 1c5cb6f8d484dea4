#!/usr/bin/env python
"""Runs the three hand-written pointwise products (csrc/rowsgemm.cu, csrc/rowswgrad.cu) a few times at one layer shape: the
command ncu wraps.     python profiles/run_rowsgemm_once.py [R] [K] [N] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import sph3d_gcn_b200 as S

argv = [int(a) for a in sys.argv[1:]]
R, K, N = (argv + [320000, 128, 128])[:3] if len(argv) < 3 else argv[:3]
iters = argv[3] if len(argv) > 3 else 2
rg = S.tf_rowsgemm
x, w, g = torch.randn(R, K, device="cuda"), 0.1 * torch.randn(K, N, device="cuda"), torch.randn(R, N, device="cuda")
img, imgt = rg.pack(w), rg.pack(w, trans=True)
for _ in range(iters):
    rg.rows_gemm(x, w, image=img)
    rg.rows_gemm(g, w, trans=True, image=imgt)
    rg.rows_wgrad(x, g)
torch.cuda.synchronize()
