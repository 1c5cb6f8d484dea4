"""Round-2 op timings on one B200: one-call vs planned convolution (forward / backward / plan builds), both forms of
the pool / unpool gradients, r = 2 shapes under both backward algorithms.

    python profiles/r2_stage.py [--out gpurun_out/r2_stage.json] [--only conv|pool|r2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import torch

import bench
import sph3d_gcn_b200 as S


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters, 4)


def env(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    S._lib.reload_tunables()


def conv_block(name, cfg, dev, res):
    C3 = S.tf_conv3d
    host, radius, F = bench.make_inputs(cfg, 1236, dev, S)
    d = {k: v.to(dev) for k, v in host.items() if k != "xyz"}
    B, N, K, C, r = (cfg[k] for k in ("B", "N", "K", "C", "r"))
    x, W, go, idx, cnt, filt = (d[k] for k in ("x", "W", "go", "idx", "cnt", "filt"))
    E = int(cnt.sum().item())
    ab_f, ab_b = bench.algorithmic_bytes(B, N, N, C, r, F, E)
    rec = {"shape": dict(cfg, kernel=list(cfg["kernel"])), "E": E, "algorithmic_MB_fwd": ab_f / 1e6, "algorithmic_MB_bwd": ab_b / 1e6}
    rec["fwd_one_call_ms"] = timeit(lambda: C3._forward(x, W, idx, cnt, filt))
    fplan = C3.conv_sort(idx, cnt, filt, F, N)
    if fplan is not None:
        rec["conv_sort_ms"] = timeit(lambda: C3.conv_sort(idx, cnt, filt, F, N))
        rec["fwd_planned_ms"] = timeit(lambda: C3.depthwise_conv3d_planned(x, W, cnt, fplan, K))
    for algo in (None, 1, 2):
        env(SPH3D_BWD_ALGO=algo)
        rec["bwd_one_call_algo_%s_ms" % (algo or "auto")] = timeit(lambda: C3.depthwise_conv3d_grad(x, W, go, idx, cnt, filt))
    env(SPH3D_BWD_ALGO=None)
    plan = C3.conv_transpose(idx, cnt, filt, F, N)
    if plan is not None:
        rec["conv_transpose_ms"] = timeit(lambda: C3.conv_transpose(idx, cnt, filt, F, N))
        rec["bwd_planned_ms"] = timeit(lambda: C3.depthwise_conv3d_grad_planned(x, W, go, cnt, plan, K))
    res[name] = rec
    print(json.dumps({name: rec}), flush=True)


def pool_block(name, B, N, Sn, K, C, dev, res):
    g = torch.Generator().manual_seed(4321)
    xyz = torch.rand(B, N, 3, generator=g).to(dev)
    radius = bench.saturating_radius(N, K)
    idx, cnt, dst = S.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=radius, nnsample=K)
    sel = S.tf_sample.farthest_point_sample(Sn, xyz)
    bi = torch.arange(B, device=dev)[:, None]
    pidx, pcnt = idx[bi, sel.long()].contiguous(), cnt[bi, sel.long()].contiguous()
    x = torch.randn(B, N, C, generator=g).to(dev)
    gop = torch.randn(B, Sn, C, generator=g).to(dev)
    coarse = xyz[bi, sel.long()].contiguous()
    uidx, ucnt, udst = S.tf_nnquery.build_sphere_neighbor(coarse, xyz, radius=2 * radius, nnsample=K)
    xc = torch.randn(B, Sn, C, generator=g).to(dev)
    gof = torch.randn(B, N, C, generator=g).to(dev)
    w = ((udst + 1e-7) / (udst.sum(-1, keepdim=True) + 1e-7)).contiguous()
    rec = {"B": B, "N": N, "S": Sn, "K": K, "C": C}
    for form in (True, False):
        S.tf_pool3d.GATHER_FORM_GRAD = form
        S.tf_unpool3d.GATHER_FORM_GRAD = form
        tag = "gather" if form else "scatter"
        rec["avg_pool3d_grad_%s_ms" % tag] = timeit(lambda: S.tf_pool3d.avg_pool3d_grad(x, gop, pidx, pcnt))
        rec["mean_interpolate_grad_%s_ms" % tag] = timeit(lambda: S.tf_unpool3d.mean_interpolate_grad(xc, gof, uidx, ucnt))
        rec["weighted_interpolate_grad_%s_ms" % tag] = timeit(lambda: S.tf_unpool3d.weighted_interpolate_grad(xc, gof, w, uidx, ucnt))
    S.tf_pool3d.GATHER_FORM_GRAD = True
    S.tf_unpool3d.GATHER_FORM_GRAD = True
    rec["avg_pool3d_ms"] = timeit(lambda: S.tf_pool3d.avg_pool3d(x, pidx, pcnt))
    rec["mean_interpolate_ms"] = timeit(lambda: S.tf_unpool3d.mean_interpolate(xc, uidx, ucnt))
    res[name] = rec
    print(json.dumps({name: rec}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_stage.json"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    res = {"device": torch.cuda.get_device_name(0)}
    W = bench.WORKLOADS
    if args.only in ("", "conv"):
        conv_block("cfgT", W["cfgT"], dev, res)
    if args.only in ("", "r2"):
        conv_block("cfgT_r2", W["cfgT_r2"], dev, res)
        conv_block("s3dis_l1", W["s3dis_l1"], dev, res)
        conv_block("s3dis_l2_c128r2_n2048", dict(B=8, N=2048, K=64, C=128, r=2, kernel=(8, 2, 2)), dev, res)
        conv_block("cfg5", dict(B=4, N=65536, K=64, C=256, r=1, kernel=(8, 2, 2)), dev, res)
    if args.only in ("", "pool"):
        pool_block("pool_cfgT", 32, 10000, 2500, 64, 128, dev, res)
        pool_block("pool_s3dis", 8, 8192, 2048, 64, 128, dev, res)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
