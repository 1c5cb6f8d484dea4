"""Per-op timing of every SURVEY.md section-8 row on one B200 (GPU box):

    python profiles/bench_ops.py [--shape cfgT|s3dis|cfg5|cfg1] [--no-ref]

For each op: CUDA-event time of this library's kernel(s), algorithmic bytes (SURVEY 8d formulas) ->
achieved GB/s and fraction of the measured HBM peak, and -- unless --no-ref -- the time of the UNMODIFIED
reference kernel (oracle/_ref) on the same device-resident inputs.  One JSON line per op; the table in
DESIGN.md / BASELINE.md is filled from this output (copied to profiles/).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np
import torch

import bench
import ref_gpu as R
import sph3d_gcn_b200 as S

SHAPES = {
    # B, N, K, C, r, S (FPS samples), radius(None = saturating), Mc (coarse cloud for unpool)
    "cfgT": dict(B=32, N=10000, K=64, C=128, r=1, S=2500, radius=None),
    "s3dis": dict(B=8, N=8192, K=64, C=64, r=2, S=2048, radius=None),
    "cfg5": dict(B=4, N=65536, K=64, C=256, r=1, S=16384, radius=None),
    "cfg1": dict(B=2, N=1024, K=20, C=3, r=2, S=256, radius=0.15),
}


def timeit(fn, iters, warm=2):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="cfgT", choices=sorted(SHAPES))
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    cfg = SHAPES[args.shape]
    B, N, K, C, r, Sn = (cfg[k] for k in ("B", "N", "K", "C", "r", "S"))
    M = N
    dev = torch.device("cuda", 0)
    peak, peak_src = bench.peaks()
    use_ref = (not args.no_ref) and R.available()
    g = torch.Generator().manual_seed(4321)
    xyz = torch.rand(B, N, 3, generator=g).to(dev)
    radius = cfg["radius"] or bench.saturating_radius(N, K)
    kernel = [8, 2, 2]
    F = 33
    u = S.sph3gcn_util
    rows = []

    def report(op, ms, abytes, ref_ms=None, extra=None):
        rec = {"shape": args.shape, "op": op, "ms": round(ms, 4), "algorithmic_MB": round(abytes / 1e6, 2),
               "GBps": round(abytes / ms / 1e6, 1), "frac_of_hbm_peak": round(abytes / ms / 1e6 / peak, 4),
               "ref_kernel_ms": None if ref_ms is None else round(ref_ms, 3),
               "speedup_vs_ref_kernel": None if ref_ms is None else round(ref_ms / ms, 1)}
        if extra:
            rec.update(extra)
        rows.append(rec)
        print(json.dumps(rec), flush=True)

    # ---- a1 nnquery -------------------------------------------------------------------------------
    idx, cnt, dst = S.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=radius, nnsample=K)
    E = int(cnt.sum().item())
    ms = timeit(lambda: S.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=radius, nnsample=K), args.iters)
    ref_ms = None
    if use_ref:
        oi, oc, od = torch.zeros_like(idx), torch.zeros_like(cnt), torch.zeros_like(dst)
        ref_ms = timeit(lambda: (oi.zero_(), oc.zero_(), od.zero_(),
                                 R.launch_raw("sphere", B, N, M, K, float(radius), xyz, xyz, oi, oc, od)), 1, warm=1)
    report("build_sphere_neighbor", ms, 4 * (3 * B * N + 3 * B * M + 2 * B * M * K + B * M), ref_ms,
           {"Gtests_per_s": round(B * M * N / ms / 1e6, 1), "mean_neighbors": E / (B * M)})

    # ---- a3 buildkernel ---------------------------------------------------------------------------
    filt = S.tf_buildkernel.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, kernel=kernel)
    ms = timeit(lambda: S.tf_buildkernel.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, kernel=kernel), args.iters)
    if use_ref:
        of = torch.zeros_like(filt)
        ref_ms = timeit(lambda: (of.zero_(), R.launch_raw("kernel", B, N, M, K, 8, 2, 2, float(radius), xyz, xyz, idx, cnt, dst, of)), 2, warm=1)
    report("spherical_kernel", ms, 4 * (3 * B * N + 3 * B * M + 3 * E + B * M), ref_ms)

    # ---- a4/a5 conv -------------------------------------------------------------------------------
    x = torch.randn(B, N, C, generator=g).to(dev)
    W = (0.1 * torch.randn(F, C, r, generator=g)).to(dev)
    go = torch.randn(B, M, C * r, generator=g).to(dev)
    ab_f, ab_b = bench.algorithmic_bytes(B, N, M, C, r, F, E)
    ms = timeit(lambda: S.tf_conv3d._forward(x, W, idx, cnt, filt), args.iters)
    if use_ref:
        out = torch.empty(B, M, C * r, device=dev)
        ref_ms = timeit(lambda: (out.zero_(), R.launch_raw("conv", B, N, M, C, r, K, idx, cnt, filt, x, W, out)), 2, warm=1)
    report("depthwise_conv3d", ms, ab_f, ref_ms, {"logical_gather_GBps": round((4.0 * E * C + 4.0 * B * M * C * r + 8.0 * E) / ms / 1e6, 1)})
    ms = timeit(lambda: S.tf_conv3d.depthwise_conv3d_grad(x, W, go, idx, cnt, filt), args.iters)
    if use_ref:
        gi, gf = torch.empty(B, N, C, device=dev), torch.empty(F, C, r, device=dev)
        ref_ms = timeit(lambda: (gi.zero_(), gf.zero_(), R.launch_raw("conv_grad", B, N, M, F, C, r, K, idx, cnt, filt, x, W, go, gi, gf)), 2, warm=1)
    report("depthwise_conv3d_grad", ms, ab_b, ref_ms)

    # ---- a6 FPS -----------------------------------------------------------------------------------
    sel = S.tf_sample.farthest_point_sample(Sn, xyz)
    ms = timeit(lambda: S.tf_sample.farthest_point_sample(Sn, xyz), max(2, args.iters // 2))
    if use_ref:
        so, tmp = torch.zeros(B, Sn, dtype=torch.int32, device=dev), torch.zeros(32, N, device=dev)
        ref_ms = timeit(lambda: R.launch_raw("fps", B, N, Sn, xyz, tmp, so), 1, warm=1)
    report("farthest_point_sample", ms, 4 * (3 * B * N + B * Sn), ref_ms, {"us_per_round": round(ms * 1e3 / Sn, 3), "rounds": Sn})

    # ---- a8/a9 pooling on the FPS rows ------------------------------------------------------------
    bi = torch.arange(B, device=dev)[:, None]
    pidx, pcnt = idx[bi, sel.long()].contiguous(), cnt[bi, sel.long()].contiguous()
    Ep = int(pcnt.sum().item())
    gop = torch.randn(B, Sn, C, generator=g).to(dev)
    po, pmi = S.tf_pool3d.max_pool3d(x, pidx, pcnt)
    ms = timeit(lambda: S.tf_pool3d.max_pool3d(x, pidx, pcnt), args.iters)
    if use_ref:
        ro, rmi = torch.empty_like(po), torch.empty_like(pmi)
        ref_ms = timeit(lambda: (ro.zero_(), rmi.zero_(), R.launch_raw("maxpool", B, N, Sn, C, K, pidx, pcnt, x, ro, rmi)), 2, warm=1)
    report("max_pool3d", ms, 4 * (B * N * C + 2 * B * Sn * C + Ep + B * Sn), ref_ms)
    ms = timeit(lambda: S.tf_pool3d.max_pool3d_grad(x, gop, pmi), args.iters)
    if use_ref:
        rg = torch.empty_like(x)
        ref_ms = timeit(lambda: (rg.zero_(), R.launch_raw("maxpool_grad", B, N, Sn, C, pmi, gop, rg)), 2, warm=1)
    report("max_pool3d_grad", ms, 4 * (2 * B * Sn * C + B * N * C), ref_ms)
    ms = timeit(lambda: S.tf_pool3d.avg_pool3d(x, pidx, pcnt), args.iters)
    if use_ref:
        ref_ms = timeit(lambda: (ro.zero_(), R.launch_raw("avgpool", B, N, Sn, C, K, pidx, pcnt, x, ro)), 2, warm=1)
    report("avg_pool3d", ms, 4 * (B * N * C + B * Sn * C + Ep + B * Sn), ref_ms)
    ms = timeit(lambda: S.tf_pool3d.avg_pool3d_grad(x, gop, pidx, pcnt), args.iters)
    if use_ref:
        ref_ms = timeit(lambda: (rg.zero_(), R.launch_raw("avgpool_grad", B, N, Sn, C, K, pidx, pcnt, gop, rg)), 2, warm=1)
    report("avg_pool3d_grad", ms, 4 * (B * Sn * C + B * N * C + Ep + B * Sn), ref_ms)

    # ---- a10/a11 unpool: coarse = FPS subset, fine = full cloud -----------------------------------
    coarse = xyz[bi, sel.long()].contiguous()
    uidx, ucnt, udst = S.tf_nnquery.build_sphere_neighbor(coarse, xyz, radius=2 * radius, nnsample=K)
    Eu = int(ucnt.sum().item())
    xc = torch.randn(B, Sn, C, generator=g).to(dev)
    gof = torch.randn(B, N, C, generator=g).to(dev)
    w = ((udst + 1e-7) / (udst.sum(-1, keepdim=True) + 1e-7)).contiguous()
    ms = timeit(lambda: S.tf_unpool3d.mean_interpolate(xc, uidx, ucnt), args.iters)
    if use_ref:
        uo = torch.empty(B, N, C, device=dev)
        ref_ms = timeit(lambda: (uo.zero_(), R.launch_raw("mean", B, N, Sn, C, K, uidx, ucnt, xc, uo)), 2, warm=1)
    report("mean_interpolate", ms, 4 * (B * Sn * C + B * N * C + Eu + B * N), ref_ms)
    ms = timeit(lambda: S.tf_unpool3d.mean_interpolate_grad(xc, gof, uidx, ucnt), args.iters)
    if use_ref:
        ug = torch.empty_like(xc)
        ref_ms = timeit(lambda: (ug.zero_(), R.launch_raw("mean_grad", B, N, Sn, C, K, uidx, ucnt, gof, ug)), 2, warm=1)
    report("mean_interpolate_grad", ms, 4 * (B * N * C + B * Sn * C + Eu + B * N), ref_ms)
    ms = timeit(lambda: S.tf_unpool3d.weighted_interpolate(xc, w, uidx, ucnt), args.iters)
    if use_ref:
        ref_ms = timeit(lambda: (uo.zero_(), R.launch_raw("weighted", B, N, Sn, C, K, uidx, ucnt, xc, w, uo)), 2, warm=1)
    report("weighted_interpolate", ms, 4 * (B * Sn * C + B * N * C + 2 * Eu + B * N), ref_ms)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"peak_GBps": peak, "peak_source": peak_src, "device": torch.cuda.get_device_name(0), "rows": rows},
              open(os.path.join(ROOT, "gpurun_out", "bench_ops_%s.json" % args.shape), "w"), indent=1)


if __name__ == "__main__":
    main()
