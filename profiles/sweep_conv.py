"""Launch-shape sweep for the convolution kernels (GPU box):  python profiles/sweep_conv.py [workload]

Times conv forward / backward (CUDA events, device-resident inputs of bench.py's workload) for the
SPH3D_* launch tunables read by csrc/conv_*.cu.  Results of the sweeps are summarised in DESIGN.md.
"""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import torch

import bench
import sph3d_gcn_b200 as S


def time_op(fn, iters=5, warm=2):
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


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfgT"
    cfg = bench.WORKLOADS[wl]
    dev = torch.device("cuda", 0)
    host, radius, F = bench.make_inputs(cfg, 1236, dev, S)
    d = {k: v.to(dev) for k, v in host.items() if k != "xyz"}
    fwd = lambda: S.tf_conv3d._forward(d["x"], d["W"], d["idx"], d["cnt"], d["filt"])
    bwd = lambda: S.tf_conv3d.depthwise_conv3d_grad(d["x"], d["W"], d["go"], d["idx"], d["cnt"], d["filt"])
    grid = {
        "SPH3D_ROWS_PER_CHUNK": ["64", "128", "256"],
        "SPH3D_FWD_THREADS": ["768", "1024"],
        "SPH3D_BWD_G": ["2", "4", "8"],
        "SPH3D_BWD_THREADS": ["256", "512"],
    }
    keys = list(grid)
    base = {"SPH3D_ROWS_PER_CHUNK": "128", "SPH3D_FWD_THREADS": "1024", "SPH3D_BWD_G": "2", "SPH3D_BWD_THREADS": "512"}
    out = []
    seen = set()
    for k in keys:                       # one-factor-at-a-time around the defaults
        for v in grid[k]:
            env = dict(base); env[k] = v
            key = tuple(sorted(env.items()))
            if key in seen:
                continue
            seen.add(key)
            os.environ.update(env)
            rec = dict(env, fwd_ms=round(time_op(fwd), 4), bwd_ms=round(time_op(bwd), 4))
            out.append(rec)
            print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep_conv_%s.json" % wl), "w"), indent=1)


if __name__ == "__main__":
    main()
