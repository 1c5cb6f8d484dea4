#!/usr/bin/env python
"""Backward-pass A/B on one B200: row-owned form (conv_bwd.cu) vs transposed form (conv_bwd_t.cu), with the
transposed form's phases timed separately (graph transposition, planned gradient) and its launch knobs swept.

    python profiles/bench_bwd_t.py [--workload cfgT] [--iters 20] [--out gpurun_out/bwd_t.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench as BM
import sph3d_gcn_b200 as S


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def env(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfgT")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--sweep", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = BM.WORKLOADS[args.workload]
    host, radius, F = BM.make_inputs(cfg, 1234 + 2, dev, S)
    d = {k: v.to(dev) for k, v in host.items() if k != "xyz"}
    B, N, K, C, r = cfg["B"], cfg["N"], cfg["K"], cfg["C"], cfg["r"]
    C3 = S.tf_conv3d
    res = {"workload": args.workload, "cfg": {k: cfg[k] for k in ("B", "N", "K", "C", "r")}, "F": F}

    one_call = lambda: C3.depthwise_conv3d_grad(d["x"], d["W"], d["go"], d["idx"], d["cnt"], d["filt"])
    res["fwd_ms"] = timeit(lambda: C3._forward(d["x"], d["W"], d["idx"], d["cnt"], d["filt"]), args.iters)
    env(SPH3D_BWD_ALGO=1)
    gi_a, gf_a = one_call()
    res["bwd_row_owned_ms"] = timeit(one_call, args.iters)
    env(SPH3D_BWD_ALGO=2)
    gi_b, gf_b = one_call()
    res["bwd_transposed_ms"] = timeit(one_call, args.iters)
    env(SPH3D_BWD_ALGO=None)
    res["bwd_default_ms"] = timeit(one_call, args.iters)
    res["max_abs_diff_grad_input"] = float((gi_a - gi_b).abs().max())
    res["max_abs_diff_grad_filter"] = float((gf_a - gf_b).abs().max())
    res["scale_grad_input"] = float(gi_a.abs().max()); res["scale_grad_filter"] = float(gf_a.abs().max())

    build = lambda: C3.conv_transpose(d["idx"], d["cnt"], d["filt"], F, N)
    plan = build()
    res["transpose_ms"] = timeit(build, args.iters)
    env(SPH3D_BWDT_SORT=1)
    res["transpose_canonical_ms"] = timeit(build, args.iters)
    env(SPH3D_BWDT_SORT=None)
    planned = lambda: C3.depthwise_conv3d_grad_planned(d["x"], d["W"], d["go"], d["cnt"], plan, K)
    res["planned_ms"] = timeit(planned, args.iters)
    if args.sweep:
        sw = {}
        cfgs = [("768x2", {}), ("1024x1", dict(SPH3D_BWDT_THREADS=1024)), ("640x2", dict(SPH3D_BWDT_THREADS=640)),
                ("640x3", dict(SPH3D_BWDT_THREADS=640, SPH3D_BWDT_DEPTH=3)), ("512x3", dict(SPH3D_BWDT_THREADS=512)),
                ("512x4", dict(SPH3D_BWDT_THREADS=512, SPH3D_BWDT_DEPTH=4)), ("768x2_g2", dict(SPH3D_BWDT_G=2))]
        for name, kw in cfgs:
            env(**kw)
            plan2 = build()                                   # the plan geometry follows the launch configuration
            if plan2 is not None:
                sw[name] = timeit(lambda: C3.depthwise_conv3d_grad_planned(d["x"], d["W"], d["go"], d["cnt"], plan2, K), args.iters)
            env(**{k: None for k in kw})
        res["planned_sweep_ms"] = sw
    print(json.dumps(res))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "a") as f:
            f.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
