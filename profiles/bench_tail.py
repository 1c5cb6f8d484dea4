"""Layer tail (bias -> ELU -> batch normalisation, csrc/post.cu) against the HBM roofline, next to the same chain as
separate torch nodes (what the layer library did before the op was fused):

    python profiles/bench_tail.py [--R 320000] [--C 128] [--iters 20]        (GPU box)

Traffic model (fp32, R x C matrix, T = 4*R*C bytes): the batch statistics need a full pass before the first output can
be written and the matrix (164 MB at the Cfg-T shape) does not fit the 126 MB L2, so the floor is
  forward  3T  (read x for the statistics, read x again, write out)
  backward 5T  (read x and grad_out for the two column sums, read both again, write grad_x)
One JSON line per direction: CUDA-event time, achieved GB/s on that traffic, fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

import torch

import bench
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--R", type=int, default=320000)
    ap.add_argument("--C", type=int, default=128)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    R, C = a.R, a.C
    dev = torch.device("cuda", 0)
    lt = S.utils.layer_tail
    peak, peak_src = bench.peaks()
    torch.manual_seed(0)
    x = torch.randn(R, C, device=dev, requires_grad=True)
    go = torch.randn(R, C, device=dev)
    bias = torch.zeros(C, device=dev, requires_grad=True)
    gamma = torch.ones(C, device=dev, requires_grad=True)
    beta = torch.zeros(C, device=dev, requires_grad=True)
    mm, mv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    T = 4.0 * R * C

    def fused_fwd():
        return lt.bias_act_bn(x, bias, gamma, beta, mm, mv, act=lt.ACT_ELU, training=True)

    def eager_fwd():
        y = torch.nn.functional.elu(x + bias)
        mean, var = y.mean(0), y.var(0, unbiased=False)
        with torch.no_grad():
            mm.mul_(0.99).add_(mean.detach(), alpha=0.01)
            mv.mul_(0.99).add_(var.detach(), alpha=0.01)
        return (y - mean) * torch.rsqrt(var + 1e-3) * gamma + beta

    out = {}
    for name, fwd in (("fused", fused_fwd), ("torch_nodes", eager_fwd)):
        t_f = timeit(lambda: fwd(), a.iters)
        y = fwd()

        def bwd():
            for p in (x, bias, gamma, beta):
                p.grad = None
            y.backward(go, retain_graph=True)

        t_b = timeit(bwd, a.iters)
        out[name] = (t_f, t_b)
        for direction, t, mult in (("forward", t_f, 3.0), ("backward", t_b, 5.0)):
            gbs = mult * T / (t * 1e-3) / 1e9
            print(json.dumps({"op": "bias_elu_bn_" + direction, "impl": name, "R": R, "C": C, "ms": round(t, 4),
                              "traffic_floor_bytes": mult * T, "achieved_gbs": round(gbs, 1), "peak_gbs": peak,
                              "frac": round(gbs / peak, 3), "peak_source": peak_src}))
    print(json.dumps({"speedup_forward": round(out["torch_nodes"][0] / out["fused"][0], 2),
                      "speedup_backward": round(out["torch_nodes"][1] / out["fused"][1], 2)}))


if __name__ == "__main__":
    main()
