"""Per-layer backward of the S3DIS network (BASELINE.json configs[3]): row-owned one-call form vs the transposed form
over a shared plan, for every (points, channels, multiplier) of the encoder / decoder, on graphs built the way the model
builds them (config radii, K = 64, FPS pyramid).  Decides tf_conv3d._use_planned.

    python profiles/sweep_s3dis_layers.py [--out gpurun_out/r2_s3dis_layers.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch

import sph3d_gcn_b200 as S

C3 = S.tf_conv3d


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_s3dis_layers.json"))
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=8192)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = S.models.configs.s3dis(args.N)
    B = args.B
    g = torch.Generator().manual_seed(7)
    xyz = torch.rand(B, args.N, 3, generator=g).to(dev)
    levels = [args.N] + list(cfg.num_sample)
    # (level index, C, r) of every depthwise layer: encoder convs, then decoder convs at the same point counts
    enc_in = [64, 128, 256, 256, 512]
    layers = []
    for l in range(4):
        cin = enc_in[l]
        layers += [(l, cin, 2), (l, cfg.channels[l][0], 2)]
    layers += [(4, 512, 2), (4, 512, 2), (3, 1024, 2), (3, 256, 2), (2, 512, 2), (2, 256, 2), (1, 512, 2), (1, 128, 2)]
    rows = []
    clouds = [xyz]
    for l in range(4):
        sel = S.tf_sample.farthest_point_sample(levels[l + 1], clouds[-1])
        clouds.append(clouds[-1][torch.arange(B, device=dev)[:, None], sel.long()].contiguous())
    radii = list(cfg.radius) + [cfg.radius[-1] * 2]
    graphs = {}
    for l in range(5):
        pts = clouds[l]
        idx, cnt, dst = S.tf_nnquery.build_sphere_neighbor(pts, pts, radius=radii[l], nnsample=64)
        filt = S.tf_buildkernel.spherical_kernel(pts, pts, idx, cnt, dst, radii[l], kernel=[8, 2, 2])
        graphs[l] = (idx, cnt, filt)
    for (l, C, r) in layers:
        idx, cnt, filt = graphs[l]
        N = idx.shape[1]
        x = torch.randn(B, N, C, generator=g).to(dev)
        W = (0.1 * torch.randn(33, C, r, generator=g)).to(dev)
        go = torch.randn(B, N, C * r, generator=g).to(dev)
        rec = {"level": l, "N": N, "C": C, "r": r, "mean_cnt": float(cnt.float().mean())}
        env(SPH3D_BWD_ALGO=1)
        rec["row_owned_ms"] = timeit(lambda: C3.depthwise_conv3d_grad(x, W, go, idx, cnt, filt))
        env(SPH3D_BWD_ALGO=None)
        plan = C3.conv_transpose(idx, cnt, filt, 33, N)
        ok = plan is not None and S._lib.lib().sph3d_depthwise_conv3d_grad_planned_workspace_bytes(B, N, N, 33, C, r, 64) > 0
        if ok:
            rec["plan_ms"] = timeit(lambda: C3.conv_transpose(idx, cnt, filt, 33, N))
            rec["planned_ms"] = timeit(lambda: C3.depthwise_conv3d_grad_planned(x, W, go, cnt, plan, 64))
        rec["fwd_ms"] = timeit(lambda: C3._forward(x, W, idx, cnt, filt))
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
