"""Hand-written tcgen05 products (csrc/rowsgemm.cu, csrc/rowswgrad.cu) against a float64 product and the cuBLAS fp32 GEMM torch calls
on the layer shapes of the three networks (profiles/r2_rowsgemm.json also holds the timings of the CUTLASS-collective
instantiation of round 1, measured with this script before that path was removed).

    python profiles/check_rowsgemm.py [--time]      (GPU box) -> one JSON line per shape and product
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

import torch

import sph3d_gcn_b200 as S

u = S.sph3gcn_util
rg = S.tf_rowsgemm
SHAPES = [  # (rows, K, N, where)
    (100, 64, 32, "tiny"), (1000, 68, 132, "ragged"), (4096, 128, 516, "three column groups"),
    (65536, 256, 128, "s3dis conv1_2"), (16384, 512, 256, "s3dis conv2_2"), (6144, 512, 256, "s3dis conv3"),
    (3072, 1024, 512, "s3dis conv4_2"), (3072, 2048, 256, "s3dis deconv2_1"), (16384, 1024, 128, "s3dis deconv4_1"),
    (320000, 72, 64, "modelnet conv1_1"), (320000, 64, 64, "modelnet conv1_2"), (80000, 128, 128, "modelnet conv2_2"),
    (32768, 128, 128, "shapenet conv1"), (16384, 256, 256, "shapenet conv2"), (320000, 128, 128, "cfgT layer"),
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
    timing = "--time" in sys.argv
    torch.manual_seed(0)
    worst = 0.0
    worst2 = 0.0
    for R, K, N, where in SHAPES:
        x, w, g = torch.randn(R, K, device="cuda"), torch.randn(K, N, device="cuda") * 0.1, torch.randn(R, N, device="cuda")
        for name, a, trans, ref in (("y", x, False, x.double() @ w.double()), ("gx", g, True, g.double() @ w.double().t())):
            out = rg.rows_gemm(a, w, trans=trans)
            torch.cuda.synchronize()
            # error relative to the sum of |terms| of every output element (the bar the fused layer uses: 1e-5)
            mag = (a.double().abs() @ (w.double().abs().t() if trans else w.double().abs()))
            err = float(((out.double() - ref).abs() / mag.clamp_min(1e-30)).max())
            worst = max(worst, err)
            out2 = rg.rows_gemm(a, w, trans=trans, terms=2)
            err2 = float(((out2.double() - ref).abs() / mag.clamp_min(1e-30)).max())
            worst2 = max(worst2, err2)
            rec = {"where": where, "product": name, "R": R, "K": a.shape[1], "N": out.shape[1], "rel_err_of_terms": err,
                   "rel_err_of_terms_2": err2, "rel_err_fp32_gemm": float((((g @ w.t() if trans else x @ w).double() - ref).abs() / mag.clamp_min(1e-30)).max())}
            if timing:
                img = rg.pack(w, trans)
                lib = (lambda: g @ w.t()) if trans else (lambda: x @ w)
                Kk, Nn = a.shape[1], out.shape[1]
                rec["rows_gemm_ms"] = round(timeit(lambda: rg.rows_gemm(a, w, trans=trans, image=img)), 4)
                rec["rows_gemm_2term_ms"] = round(timeit(lambda: rg.rows_gemm(a, w, trans=trans, image=img, terms=2)), 4)
                rec["with_pack_ms"] = round(timeit(lambda: rg.rows_gemm(a, w, trans=trans)), 4)
                rec["cublas_fp32_ms"] = round(timeit(lib), 4)
                bytes_ = 4.0 * R * (Kk + Nn)
                rec["hbm_gbs"] = round(bytes_ / rec["rows_gemm_ms"] / 1e6, 1)
                rec["tflops_fp32_equiv"] = round(2.0 * R * Kk * Nn / rec["rows_gemm_ms"] / 1e9, 1)
            print(json.dumps(rec), flush=True)
    if "--wgrad" in sys.argv:
        for R, K, N, where in SHAPES:
            x, g = torch.randn(R, K, device="cuda"), torch.randn(R, N, device="cuda")
            ref = x.double().t() @ g.double()
            mag = x.double().abs().t() @ g.double().abs()
            rec = {"where": where, "product": "gw", "R": R, "K": K, "N": N}
            for t in (3, 2):
                out = rg.rows_wgrad(x, g, terms=t)
                rec["rel_err_of_terms_%d" % t] = float(((out.double() - ref).abs() / mag.clamp_min(1e-30)).max())
            rec["rel_err_fp32_gemm"] = float((((x.t() @ g).double() - ref).abs() / mag.clamp_min(1e-30)).max())
            worst = max(worst, rec["rel_err_of_terms_3"])
            if timing:
                rec["rows_wgrad_ms"] = round(timeit(lambda: rg.rows_wgrad(x, g)), 4)
                rec["rows_wgrad_2term_ms"] = round(timeit(lambda: rg.rows_wgrad(x, g, terms=2)), 4)
                rec["cublas_fp32_ms"] = round(timeit(lambda: x.t() @ g), 4)
                u.ROWS_GEMM = False
                rec["split_k_library_ms"] = round(timeit(lambda: u._weight_grad(x, g)), 4)
                u.ROWS_GEMM = True
            print(json.dumps(rec), flush=True)
    print(json.dumps({"worst_rel_err_of_terms": worst, "worst_rel_err_of_terms_2term": worst2, "ok": worst < 1e-5}))
    assert worst < 1e-5


if __name__ == "__main__" and "--trace" not in sys.argv:
    main()


def trace(R, K, N):
    """clock stamps of CTA (0,0): where a tile's time goes (producer / issuer / epilogue)"""
    from sph3d_gcn_b200 import _lib
    x, w = torch.randn(R, K, device="cuda"), torch.randn(K, N, device="cuda") * 0.1
    img = rg.pack(w)
    buf = torch.zeros(6 * 64 * 4, dtype=torch.int64, device="cuda")
    for _ in range(3):
        rg.rows_gemm(x, w, image=img)
    _lib.lib().sph3d_rows_gemm_trace(buf.data_ptr())
    rg.rows_gemm(x, w, image=img)
    torch.cuda.synchronize()
    _lib.lib().sph3d_rows_gemm_trace(None)
    t = buf.view(6, 64, 4).cpu()
    t0 = int(t[t > 0].min())
    names = ["producer(warp0): begin / got stage / arrived", "issuer: begin wait / rows ready / issued+committed",
             ] + ["epilogue(warp%d): begin wait / accumulator ready / stored / first 64 rows loaded" % q for q in range(4)]
    for r in range(6):
        print(names[r])
        for i in range(8):
            row = [int(v) - t0 if v > 0 else -1 for v in t[r, i]]
            print("   step %2d: %s" % (i, row))


if __name__ == "__main__" and "--trace" in sys.argv:
    for shp in ((320000, 64, 64), (320000, 128, 128), (65536, 256, 128)):
        print("==== trace", shp)
        trace(*shp)
