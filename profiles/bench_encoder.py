"""Whole-network training steps for BASELINE.json configs[1..3] on synthetic clouds, through the package's model call
graphs (sph3d-gcn_b200/models/SPH3D_{modelnet,shapenet,s3dis}.py = the reference's models/SPH3D_*.py on the
`sph3gcn_util` mirror): forward, the reference's loss, backward through torch autograd.

    configs[1] ModelNet40 cls   B=32 N=10000 K=64, 3-level FPS+pool encoder + global conv + classifier
    configs[2] ShapeNet part    B=16 N=2048  K=32, encoder + unpool decoder
    configs[3] S3DIS seg        B=8  N=8192  K=64, encoder + unpool decoder, inner-point loss

    python profiles/bench_encoder.py [--model modelnet|shapenet|s3dis] [--B ..] [--N ..] [--steps 5]       (GPU box)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

import torch

import sph3d_gcn_b200 as S

s3g_util = S.sph3gcn_util
M = S.models


train_step = S.utils.train_step
make_config, DEFAULT_SHAPE = train_step.make_config, train_step.DEFAULT_SHAPE


def make_step(B, N, seed=7, model="modelnet", K=None):
    step, cfg, _ = train_step.make_step(B, N, seed, model, K)
    return step, cfg


def _dist_setup():
    """torchrun launch (one process per GPU): data-parallel over the batch, SURVEY.md 8(e)"""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world


def run(B, N, steps, warmup=2, seed=7, model="modelnet", K=None, graph=False):
    rank, world = _dist_setup()
    step, cfg = make_step(B, N, seed + rank, model, K)        # every rank owns B clouds of its own (weak scaling)
    run_step = step
    if graph:                                                 # the whole step as ONE CUDA graph (utils/graph_step.py)
        run_step = S.utils.graph_step.GraphedStep(step, s3g_util.trainable_variables, warmup=warmup)
    reduced_bytes = 0

    def full_step():
        """step + (N > 1) the one cross-rank exchange: SUM all-reduce of every weight gradient in one flat bucket"""
        nonlocal reduced_bytes
        out = run_step()
        if world > 1:
            reduced_bytes = S.utils.dist_util.allreduce_gradients([p.grad for p in s3g_util.trainable_variables()], average=True)
        return out

    for _ in range(warmup):
        pred, end, loss = full_step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        pred, end, loss = full_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:                                              # max over ranks, on the device clock
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    grads = [p.grad for p in s3g_util.trainable_variables()]
    what = {"modelnet": "SPH3D_modelnet get_model+get_loss fwd+bwd (configs[1])",
            "shapenet": "SPH3D_shapenet get_model+get_loss fwd+bwd (configs[2])",
            "s3dis": "SPH3D_s3dis get_model+get_loss fwd+bwd (configs[3])"}[model]
    feat = end['global_feat'] if model == "modelnet" else end['feats']
    return {"workload": what, "B": B, "N": N, "levels": cfg.num_sample, "K": cfg.nn_uplimit[0], "n_gpus": world,
            "B_per_gpu": B, "allreduce_bytes_per_step": reduced_bytes,
            "ms_per_step": ms, "points_per_s": world * B * N / (ms * 1e-3), "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / steps,
            "pred_shape": list(pred.shape), "feature_dim": int(feat.shape[-1]), "loss": float(loss.detach()), "n_params": len(grads),
            "all_grads_finite": bool(all(gr is not None and torch.isfinite(gr).all() for gr in grads)),
            "fused_tail": bool(s3g_util.FUSED_TAIL), "cuda_graph": bool(graph)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=0)
    ap.add_argument("--N", type=int, default=0)
    ap.add_argument("--K", type=int, default=0, help="shapenet only: neighbours per point (default 32)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--model", default="modelnet", choices=["modelnet", "shapenet", "s3dis"])
    ap.add_argument("--no-share", action="store_true", help="one graph transposition per convolution gradient (tf_conv3d.SHARE_PLANS off)")
    ap.add_argument("--no-fused-tail", action="store_true", help="bias/ELU/BN as separate torch nodes (sph3gcn_util.FUSED_TAIL off)")
    ap.add_argument("--graph", action="store_true", help="capture the step in a CUDA graph and time replays")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    S.tf_conv3d.SHARE_PLANS = not a.no_share
    s3g_util.FUSED_TAIL = not a.no_fused_tail
    B0, N0 = DEFAULT_SHAPE[a.model]
    rec = run(a.B or B0, a.N or N0, a.steps, model=a.model, K=a.K or None, graph=a.graph)
    rec["share_plans"] = bool(S.tf_conv3d.SHARE_PLANS)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(rec))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        name = "bench_%s%s%s%s.json" % ("encoder" if a.model == "modelnet" else a.model, "_noshare" if a.no_share else "",
                                        "_eagertail" if a.no_fused_tail else "", ("_graph" if a.graph else "") +
                                        ("_dp%d" % rec["n_gpus"] if rec["n_gpus"] > 1 else "") + a.tag)
        json.dump(rec, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
    if rec["n_gpus"] > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
