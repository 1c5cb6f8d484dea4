"""How much L1 does the forward gather need?  Times sph3d_depthwise_conv3d at Cfg-T with unused shared memory added to the
kernel (SPH3D_FWD_SMEM_PAD_KB: the pad shrinks the SM's L1, nothing else changes), and the fused separable kernel
(csrc/sepconv.cu) with its tile size / weight-ring depth varied.  -> gpurun_out/r2_fwd_l1.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sph3d_gcn_b200 as S  # noqa: E402

DEV = "cuda:0"


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
    return round(ev[0].elapsed_time(ev[1]) / iters, 4)


def tune(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    S._lib.reload_tunables()


def main():
    B, N, K, C, r, cout = 32, 10000, 64, 128, 1, 128
    g = torch.Generator().manual_seed(7)
    xyz = torch.rand((B, N, 3), generator=g).to(DEV)
    radius = float((3.0 * 2 * K / (4.0 * np.pi * N)) ** (1.0 / 3.0))
    idx, cnt, dst = S.tf_nnquery.build_sphere_neighbor(xyz, xyz, radius=radius, nnsample=K)
    bins = S.tf_buildkernel.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, kernel=[8, 2, 2])
    x = torch.randn((B, N, C), generator=g).to(DEV)
    filt = (0.1 * torch.randn((33, C, r), generator=g)).to(DEV)
    W = (0.1 * torch.randn((C * r, cout), generator=g)).to(DEV)
    img = S.tf_sepconv.pack_weights(W)
    out = {"device": torch.cuda.get_device_name(0), "shape": dict(B=B, N=N, K=K, C=C, r=r, Cout=cout), "forward_pad": [], "fused": []}
    for pad in (0, 16, 32, 48, 64, 80, 96, 112, 128, 144, 160, 176):
        tune(SPH3D_FWD_SMEM_PAD_KB=pad)
        ms = cuda_ms(lambda: S.tf_conv3d.depthwise_conv3d(x, filt, idx, cnt, bins))
        out["forward_pad"].append(dict(pad_kb=pad, ms=ms))
        print("forward, +%3d KB shared: %.4f ms" % (pad, ms), flush=True)
    tune(SPH3D_FWD_SMEM_PAD_KB=None)
    for tile in (64, 32):
        for stages in (6, 4, 3, 2):
            tune(SPH3D_SEPCONV_TILE=tile, SPH3D_SEPCONV_STAGES=stages)
            ms = cuda_ms(lambda: S.tf_sepconv.separable_conv3d(x, filt, W, idx, cnt, bins, weight_image=img))
            out["fused"].append(dict(tile=tile, stages=stages, ms=ms))
            print("fused, tile %d, %d weight stages: %.4f ms" % (tile, stages, ms), flush=True)
    tune(SPH3D_SEPCONV_TILE=None, SPH3D_SEPCONV_STAGES=None)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "r2_fwd_l1.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
