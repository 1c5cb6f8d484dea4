"""Generates tests/golden/ref_b200.npz: outputs of the UNMODIFIED reference CUDA kernels
(oracle/_ref/libsph3d_ref.so, compiled by oracle/Makefile from /root/reference/tf_ops/*/tf_*_gpu.cu)
executed on a B200, for small seeded inputs.  The reference has no tests and no golden vectors of
its own, so these fixtures are what pins the CPU oracle on GPU-less boxes
(tests/test_oracle_golden_cpu.py).

Run on a GPU box:   python tests/golden/make_golden.py gpurun_out/ref_b200.npz
then copy the file to tests/golden/ref_b200.npz and commit it.
Large integer outputs are stored as sha256 digests (key suffix '.sha'), small ones in full.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import ref_gpu as R                                      # noqa: E402
from common import features, make_cloud                  # noqa: E402


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main(out_path):
    assert R.available(), "needs a CUDA device and oracle/_ref/libsph3d_ref.so"
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    A = lambda t: t.cpu().numpy()
    G = {}

    # ---- g1: cfg1-like intra graph + bins + conv + pools --------------------------------------
    B, N, K, C, r = 2, 512, 16, 8, 2
    xyz = make_cloud(7001, B, N, "cube"); rad = 0.17
    idx, cnt, dst = [A(t) for t in R.build_sphere_neighbor(T(xyz), T(xyz), rad, None, K)]
    G.update(g1_xyz=xyz, g1_idx=idx, g1_cnt=cnt, g1_dst=dst, g1_radius=np.float32(rad))
    for name, kern in (("822", [8, 2, 2]), ("821", [8, 2, 1]), ("823", [8, 2, 3])):
        G["g1_filt" + name] = A(R.spherical_kernel(T(xyz), T(xyz), T(idx), T(cnt), T(dst), rad, kern))
    x, W, go = features(7002, B, N, C), (0.5 * features(7003, 33, C, r)).astype(np.float32), features(7004, B, N, C * r)
    filt = G["g1_filt822"]
    G.update(g1_x=x, g1_W=W, g1_go=go)
    G["g1_conv"] = A(R.depthwise_conv3d(T(x), T(W), T(idx), T(cnt), T(filt)))
    gi, gf = R.depthwise_conv3d_grad(T(x), T(W), T(go), T(idx), T(cnt), T(filt))
    G.update(g1_conv_gi=A(gi), g1_conv_gf=A(gf))
    sel = A(R.farthest_point_sample(128, T(xyz)))
    G["g1_fps128"] = sel
    bi = np.arange(B)[:, None]
    pidx, pcnt = idx[bi, sel], cnt[bi, sel]
    xq = np.round(x * 2).astype(np.float32)                 # ties
    mo, mi = R.max_pool3d(T(xq), T(pidx), T(pcnt))
    G.update(g1_max=A(mo), g1_maxidx=A(mi))
    G["g1_max_grad"] = A(R.max_pool3d_grad(T(xq), T(go[:, :128, :C]), mi))
    G["g1_avg"] = A(R.avg_pool3d(T(x), T(pidx), T(pcnt)))
    G["g1_avg_grad"] = A(R.avg_pool3d_grad(T(x), T(go[:, :128, :C]), T(pidx), T(pcnt)))

    # ---- g2: decoder inter graph (database = coarse FPS subset, query = fine) with retries ------
    coarse = xyz[bi, sel[:, :24]]
    i2, c2, d2 = [A(t) for t in R.build_sphere_neighbor(T(coarse), T(xyz), 0.08, None, 8)]
    G.update(g2_idx=i2, g2_cnt=c2, g2_dst=d2)
    xc = features(7005, B, 24, C)
    w = ((d2 + 1e-7) / (d2.sum(-1, keepdims=True) + 1e-7)).astype(np.float32)
    G["g2_mean"] = A(R.mean_interpolate(T(xc), T(i2), T(c2)))
    G["g2_weighted"] = A(R.weighted_interpolate(T(xc), T(w), T(i2), T(c2)))
    G["g2_mean_grad"] = A(R.mean_interpolate_grad(T(xc), T(go[:, :, :C]), T(i2), T(c2)))
    G["g2_weighted_grad"] = A(R.weighted_interpolate_grad(T(xc), T(go[:, :, :C]), T(w), T(i2), T(c2)))

    # ---- g3: radius chain with B > 32 and M > 1024, sparse database (digest only) ---------------
    xb, qb = make_cloud(7006, 34, 90, "cube"), make_cloud(7007, 34, 1300, "cube")
    i3, c3, d3 = [A(t) for t in R.build_sphere_neighbor(T(xb), T(qb), 0.12, None, 6)]
    G.update({"g3_idx.sha": sha(i3), "g3_cnt.sha": sha(c3), "g3_dst.sha": sha(d3), "g3_cnt_sum": np.int64(c3.sum())})

    # ---- g4: lattice points: ties, duplicates, neighbours exactly on the radius / on bin borders --
    xg = make_cloud(7008, 2, 600, "grid")
    i4, c4, d4 = [A(t) for t in R.build_sphere_neighbor(T(xg), T(xg), 0.25, None, 24)]
    f4 = A(R.spherical_kernel(T(xg), T(xg), T(i4), T(c4), T(d4), 0.25, [8, 2, 2]))
    G.update({"g4_idx.sha": sha(i4), "g4_cnt": c4, "g4_dst.sha": sha(d4), "g4_filt.sha": sha(f4)})
    G["g4_fps200"] = A(R.farthest_point_sample(200, T(xg)))
    ic, cc = [A(t) for t in R.build_cube_neighbor(T(xg), T(xg), 0.3, None, 12, 3)]
    G.update({"g4_cube_idx.sha": sha(ic), "g4_cube_cnt": cc})

    # ---- g5: FPS on bigger clouds (digest) + global graph of the classification head -------------
    xs = make_cloud(7009, 2, 3000, "shell")
    G["g5_fps750.sha"] = sha(A(R.farthest_point_sample(750, T(xs))))
    xg2 = make_cloud(7010, 2, 156, "shell"); qc = xg2.mean(axis=1, keepdims=True).astype(np.float32)
    i5, c5, d5 = [A(t) for t in R.build_sphere_neighbor(T(xg2), T(qc), 100.0, None, 156)]
    G.update(g5_gidx=i5, g5_gcnt=c5, g5_gdst=d5)
    G["g5_gfilt"] = A(R.spherical_kernel(T(xg2), T(qc), T(i5), T(c5), T(d5), 100.0, [8, 2, 1]))

    G["meta_device"] = np.frombuffer(torch.cuda.get_device_name(0).encode(), dtype=np.uint8)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **G)
    print("wrote", out_path, os.path.getsize(out_path), "bytes,", len(G), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_b200.npz"))
