"""One training step of a model call graph on synthetic clouds: forward, the reference's loss, backward.

The reference's train scripts (s3dis_seg/train_s3dis.py:196-260, modelnet40_cls/train_modelnet.py:152-215) build the
graph of get_model + get_loss + the optimiser's gradients once and run it per batch; this is that step on the package's
model call graphs, for the whole-network benchmarks (BASELINE.json configs[1..3]) and tests.  Data-parallel runs give
every rank its own slice of the global batch: pass the per-rank batch size and a per-rank seed."""
import torch

from .. import models as M
from . import sph3gcn_util as s3g_util

DEFAULT_SHAPE = {"modelnet": (32, 10000), "shapenet": (16, 2048), "s3dis": (8, 8192)}


def make_config(model, N, K=None):
    if model == "modelnet":
        return M.configs.modelnet(N)
    if model == "shapenet":
        return M.configs.shapenet(N, nn_uplimit=K or 32)       # BASELINE.json configs[2]: K = 32
    if model == "s3dis":
        return M.configs.s3dis(N)
    raise ValueError("unknown model %r" % (model,))


def make_inputs(model, B, N, cfg, seed, device):
    """synthetic batch of the model's input layout -> (points, label, inner-or-None) on `device`"""
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(B, N, 3, generator=g)                    # unit cube, like a normalised cloud / a 1 m S3DIS block
    inner = None
    if model == "modelnet":
        pts, label = xyz, torch.randint(0, cfg.num_cls, (B,), generator=g)
    elif model == "shapenet":                                 # xyz + normals, 50 part classes (shapenet_seg/train_shapenet.py)
        pts = torch.cat([xyz, torch.rand(B, N, 3, generator=g)], dim=2)
        label = torch.randint(0, 50, (B, N), generator=g)
    else:                                                     # xyz + rgb (INPUT_DIM = 6, s3dis_seg/train_s3dis.py:57)
        pts = torch.cat([xyz, torch.rand(B, N, 3, generator=g)], dim=2)
        label = torch.randint(0, cfg.num_cls, (B, N), generator=g)
        inner = (torch.rand(B, N, generator=g) < 0.7).to(torch.int32)
    return pts.to(device), label.to(device), None if inner is None else inner.to(device)


def make_step(B, N, seed=7, model="modelnet", K=None, zero_grads=None):
    """-> (step function, config, inputs).  step() = clear collections, drop (or zero) the gradients, forward, the
    reference's loss, backward; returns (pred, end_points, loss).  `zero_grads`: a callable that resets the gradient
    storage in place (dist_util.GradBuckets.zero) instead of setting every .grad to None."""
    dev = torch.device("cuda", torch.cuda.current_device())
    cfg = make_config(model, N, K)
    s3g_util.reset_variables()
    pts, label, inner = make_inputs(model, B, N, cfg, seed, dev)
    hook = {"zero": zero_grads}

    def step():
        s3g_util.clear_collections()
        if hook["zero"] is not None:
            hook["zero"]()
        else:
            for p in s3g_util.trainable_variables():
                p.grad = None
        if model == "modelnet":
            pred, end = M.SPH3D_modelnet.get_model(pts, True, cfg)
            M.SPH3D_modelnet.get_loss(pred, label, end)
        elif model == "shapenet":
            pred, end = M.SPH3D_shapenet.get_model(pts, 50, True, cfg)
            M.SPH3D_shapenet.get_loss(pred, label, end)
        else:
            pred, end = M.SPH3D_s3dis.get_model(pts, True, cfg)
            M.SPH3D_s3dis.get_loss(pred, label, end, inner)
        loss = sum(s3g_util.get_collection('losses'))         # tf.add_n(tf.get_collection('losses')) in the train scripts
        loss.backward()
        return pred, end, loss

    step.set_zero_grads = lambda fn: hook.__setitem__("zero", fn)
    return step, cfg, (pts, label, inner)
