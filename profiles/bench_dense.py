"""Pointwise products of the S3DIS / ModelNet layers: tcgen05 9xBF16 (csrc/dense_gemm.cuh) vs the cuBLAS fp32 GEMM torch calls.

    python profiles/bench_dense.py        (GPU box) -> one JSON line per shape and product, TFLOP/s of both, error vs float64
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

import torch

import sph3d_gcn_b200 as S

u = S.sph3gcn_util
SHAPES = [  # (rows, Cin*r, Cout, where)
    (65536, 256, 128, "s3dis conv1_2"), (16384, 512, 256, "s3dis conv2_2"), (6144, 512, 256, "s3dis conv3"),
    (3072, 1024, 512, "s3dis conv4_2"), (3072, 2048, 256, "s3dis deconv2_1"), (16384, 1024, 128, "s3dis deconv4_1"),
    (320000, 72, 64, "modelnet conv1_1"), (320000, 64, 64, "modelnet conv1_2"), (80000, 128, 128, "modelnet conv2_2"),
]


def timeit(fn, iters=20, warm=3):
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
    u.DENSE_CTA_PAIR = "--pair" in sys.argv          # y / gx on cta_group::2 kernels
    torch.manual_seed(0)
    tot = {"tc": 0.0, "lib": 0.0}
    for R, K, N, where in SHAPES:
        x, w, g = torch.randn(R, K, device="cuda"), torch.randn(K, N, device="cuda") * 0.1, torch.randn(R, N, device="cuda")
        flops = 2.0 * R * K * N
        ref = {"y": x.double() @ w.double(), "gx": g.double() @ w.double().t(), "gw": x.double().t() @ g.double()}
        for name, tc, lib in (("y", lambda: u._tc_gemm(0, x, w, R, N, K), lambda: x @ w),
                              ("gx", lambda: u._tc_gemm(1, g, w, R, K, N), lambda: g @ w.t()),
                              ("gw", lambda: u._weight_grad(x, g), lambda: x.t() @ g)):
            out = tc()
            if out is None:
                print(json.dumps({"where": where, "product": name, "tc": "not covered"}))
                continue
            scale = float(ref[name].abs().max())
            err_tc = float((out.double() - ref[name]).abs().max()) / scale
            err_lib = float((lib().double() - ref[name]).abs().max()) / scale
            t_tc, t_lib = timeit(tc), timeit(lib)
            tot["tc"] += t_tc
            tot["lib"] += t_lib
            print(json.dumps({"where": where, "product": name, "R": R, "K": K, "N": N, "tc_ms": round(t_tc, 4),
                              "cublas_fp32_ms": round(t_lib, 4), "tc_tflops": round(flops / t_tc / 1e9, 1),
                              "cublas_tflops": round(flops / t_lib / 1e9, 1), "speedup": round(t_lib / t_tc, 2),
                              "tc_rel_err": err_tc, "cublas_rel_err": err_lib}))
    print(json.dumps({"total_tc_ms": round(tot["tc"], 3), "total_cublas_fp32_ms": round(tot["lib"], 3)}))


if __name__ == "__main__":
    main()
