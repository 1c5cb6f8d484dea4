"""Fused separable layer (csrc/sepconv.cu): correctness probes with diagnostics + timing against the three-op composition.
python profiles/check_sepconv.py [--time] -> prints, and writes profiles/r2_sepconv.json with --time."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sph3d_gcn_b200 as S  # noqa: E402

DEV = "cuda:0"
sep = S.tf_sepconv


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def probe_identity():
    rng = np.random.default_rng(0)
    N, C, cout = 128, 128, 128
    W = rng.standard_normal((C, cout)).astype(np.float32)
    idx = np.arange(N, dtype=np.int32)[None, :, None].copy()
    cnt = np.ones((1, N), np.int32)
    bins = np.zeros((1, N, 1), np.int32)
    filt = np.ones((3, C, 1), np.float32)
    eye = np.eye(128, dtype=np.float32)[None]
    out = sep.separable_conv3d(T(eye), T(filt), T(W), T(idx), T(cnt), T(bins))[0][0].cpu().numpy()
    err = np.abs(out - W).max()
    print("one-hot probe: max |out - W| = %.3e" % err)
    if err > 1e-5:
        # which row of W does each output row look like?
        d = ((out[:, None, :] - W[None, :, :]) ** 2).sum(-1)
        best = d.argmin(1)
        print("  row -> best matching W row (first 32):", best[:32].tolist())
        print("  residual of best match (first 8):", d[np.arange(128), best][:8])
        print("  out[0,:8] =", out[0, :8], "\n  W[0,:8]   =", W[0, :8])
        print("  rows all-zero:", int((np.abs(out).sum(1) == 0).sum()), " nan:", int(np.isnan(out).sum()))
    x = rng.standard_normal((1, N, C)).astype(np.float32)
    out = sep.separable_conv3d(T(x), T(filt), T(W), T(idx), T(cnt), T(bins))[0][0].cpu().numpy()
    truth = x[0].astype(np.float64) @ W.astype(np.float64)
    terms = np.abs(x[0]).astype(np.float64) @ np.abs(W).astype(np.float64)
    print("dense probe: max err / sum|terms| = %.3e (fp32 GEMM: %.3e)" % (
        (np.abs(out - truth) / terms).max(), (np.abs((x[0] @ W) - truth) / terms).max()))
    return err <= 1e-5


def cuda_ms(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def timing():
    rows = []
    s3g = S.utils.sph3gcn_util
    for name, B, N, K, C, r, cout in [("cfgT", 32, 10000, 64, 128, 1, 128), ("s3dis_l1", 8, 8192, 64, 64, 2, 64),
                                      ("s3dis_l1_o128", 8, 8192, 64, 64, 2, 128), ("s3dis_l1b_c128_r2", 8, 8192, 64, 128, 2, 128),
                                      ("modelnet_l1b", 32, 10000, 64, 64, 1, 64)]:
        g = torch.Generator().manual_seed(7)
        xyz = torch.rand((B, N, 3), generator=g).to(DEV)
        radius = float((3.0 * 2 * K / (4.0 * np.pi * N)) ** (1.0 / 3.0))
        idx, cnt, dst = S.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=radius, nnsample=K)
        bins = S.tf_buildkernel.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, kernel=[8, 2, 2])
        x = torch.randn((B, N, C), generator=g).to(DEV)
        filt = (0.1 * torch.randn((33, C, r), generator=g)).to(DEV)
        W = (0.1 * torch.randn((C * r, cout), generator=g)).to(DEV)
        bias = torch.randn((cout,), generator=g).to(DEV)
        scale = torch.rand((cout,), generator=g).to(DEV) + 0.5
        shift = torch.randn((cout,), generator=g).to(DEV)
        img = sep.pack_weights(W)

        def fused_inf():
            return sep.separable_conv3d(x, filt, W, idx, cnt, bins, bias=bias, scale=scale, shift=shift, act=1, weight_image=img)[0]

        def fused_train():
            return sep.separable_conv3d(x, filt, W, idx, cnt, bins, keep_depthwise=True, weight_image=img)[0]

        def composed_raw():
            d = S.tf_conv3d.depthwise_conv3d(x, filt, idx, cnt, bins)
            return s3g._dense(d.reshape(-1, C * r), W)

        def composed_inf():
            y = composed_raw()
            return S.utils.layer_tail.bias_act_bn(y, bias, act=1) * scale + shift

        def depthwise_only():
            return S.tf_conv3d.depthwise_conv3d(x, filt, idx, cnt, bins)

        a, b = fused_inf(), composed_inf().reshape(B, N, cout)
        rel = float((a - b).abs().max() / b.abs().max())
        row = dict(name=name, B=B, N=N, K=K, C=C, r=r, Cout=cout, max_rel_diff_vs_composition=rel,
                   depthwise_only_ms=round(cuda_ms(depthwise_only), 4),
                   fused_inference_ms=round(cuda_ms(fused_inf), 4), fused_training_ms=round(cuda_ms(fused_train), 4),
                   composed_product_ms=round(cuda_ms(composed_raw), 4), composed_inference_ms=round(cuda_ms(composed_inf), 4))
        print(row, flush=True)
        rows.append(row)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "r2_sepconv.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(out, "w"), indent=1)


if __name__ == "__main__":
    ok = probe_identity()
    if ok and "--time" in sys.argv:
        timing()
    sys.exit(0 if ok else 1)
