"""BASELINE.json configs[1] ("ModelNet40 cls shape: B=32 N=10000 K=64 Cin=3->128, FPS+pool 3-level encoder"):
the op sequence of models/SPH3D_modelnet.py:33-93 of the reference, written against this repository's
`sph3gcn_util` mirror exactly as the reference model calls `s3g_util.*` -- build_graph (ball query + FPS),
spherical_kernel, two separable_conv3d per level, gather_nd of the intra graph, pool3d, global max, and the
global conv over the final 156 points -- forward AND backward through torch autograd.  It is a benchmark
driver / drop-in demonstration, not a model zoo: no FC head, no loss beyond a scalar reduction.

    python profiles/bench_encoder.py [--B 32] [--N 10000] [--steps 5]       (GPU box)
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


class ModelNetConfig:                      # modelnet40_cls/modelnet_config.py of the reference
    def __init__(self, num_input):
        self.num_input = num_input
        self.mlp = 32
        self.num_sample = [num_input // 4 ** (i + 1) for i in range(10) if (num_input // 4 ** (i + 1)) > 100]
        self.radius = [0.1, 0.2, 0.4][:len(self.num_sample)]
        self.nn_uplimit = [64, 64, 64][:len(self.num_sample)]
        self.channels = [[64, 64], [64, 128], [128, 128]][:len(self.num_sample)]
        self.multiplier = [[2, 1], [1, 2], [1, 1]][:len(self.num_sample)]
        self.global_channels, self.global_multiplier = 512, 2
        self.weight_decay = 1e-5
        self.kernel = [8, 2, 2]
        self.binSize = 8 * 2 * 2 + 1
        self.pool_method, self.sample = 'max', 'FPS'
        self.use_raw, self.with_bn, self.with_bias = True, True, False


def _separable_conv3d_block(net, list_channels, bin_size, nn_index, nn_count, filt_idx, name, depth_multiplier,
                            weight_decay=None, with_bn=True, with_bias=True, is_training=None):
    for l, num_out_channels in enumerate(list_channels):                       # SPH3D_modelnet.py:22-31
        net = s3g_util.separable_conv3d(net, num_out_channels, bin_size, depth_multiplier[l], name + '_' + str(l + 1),
                                        nn_index, nn_count, filt_idx, weight_decay=weight_decay, with_bn=with_bn,
                                        with_bias=with_bias, is_training=is_training)
    return net


def encoder(points, is_training, config):
    """SPH3D_modelnet.get_model up to the global feature vector (SPH3D_modelnet.py:33-96)."""
    xyz = points
    query = xyz.mean(dim=1, keepdim=True)
    net = s3g_util.pointwise_conv3d(xyz, config.mlp, 'mlp1', weight_decay=config.weight_decay, with_bn=config.with_bn,
                                    with_bias=config.with_bias, is_training=is_training)
    global_feat = []
    for l in range(len(config.radius)):
        if config.use_raw:
            net = torch.cat([net, xyz], dim=-1)
        intra_idx, intra_cnt, intra_dst, indices = s3g_util.build_graph(xyz, config.radius[l], config.nn_uplimit[l],
                                                                        config.num_sample[l], sample_method=config.sample)
        filt_idx = s3g_util.spherical_kernel(xyz, xyz, intra_idx, intra_cnt, intra_dst, config.radius[l], kernel=config.kernel)
        net = _separable_conv3d_block(net, config.channels[l], config.binSize, intra_idx, intra_cnt, filt_idx,
                                      'conv' + str(l + 1), config.multiplier[l], weight_decay=config.weight_decay,
                                      with_bn=config.with_bn, with_bias=config.with_bias, is_training=is_training)
        if config.num_sample[l] > 1:
            xyz = s3g_util.gather_nd(xyz, indices)
            inter_idx = s3g_util.gather_nd(intra_idx, indices)
            inter_cnt = s3g_util.gather_nd(intra_cnt, indices)
            net = s3g_util.pool3d(net, inter_idx, inter_cnt, method=config.pool_method, scope='pool' + str(l + 1))
        global_feat.append(net.amax(dim=1, keepdim=True))
    nn_idx, nn_cnt, nn_dst = s3g_util.build_global_graph(xyz, query, 100.0)
    filt_idx = s3g_util.spherical_kernel(xyz, query, nn_idx, nn_cnt, nn_dst, 100.0, kernel=[8, 2, 1])
    net = s3g_util.separable_conv3d(net, config.global_channels, 17, config.global_multiplier, 'global_conv', nn_idx, nn_cnt,
                                    filt_idx, weight_decay=config.weight_decay, with_bn=config.with_bn,
                                    with_bias=config.with_bias, is_training=is_training)
    global_feat.append(net)
    return torch.cat(global_feat, dim=2).reshape(points.shape[0], -1)


class S3disConfig:                         # s3dis_seg/s3dis_config.py of the reference (scaled with num_input)
    def __init__(self, num_input):
        self.num_input, self.num_cls, self.mlp = num_input, 13, 64
        self.num_sample = [num_input // 4, num_input * 3 // 32, num_input * 3 // 64, num_input // 64]   # 2048,768,384,128 @8192
        self.radius = [0.1, 0.2, 0.4, 0.8]
        self.nn_uplimit = [64, 64, 64, 64]
        self.channels = [[128, 128], [256, 256], [256, 256], [512, 512]]
        self.multiplier = [[2, 2], [2, 2], [2, 2], [2, 2]]
        self.weight_decay = None
        self.kernel, self.binSize = [8, 2, 2], 33
        self.pool_method, self.unpool_method, self.sample = 'max', 'mean', 'FPS'
        self.with_bn, self.with_bias = True, False


def s3dis_model(points, is_training, config):
    """SPH3D_s3dis.get_model (models/SPH3D_s3dis.py:35-111): encoder + decoder with skip concats -> per-point logits.
    (The model's config-list `.reverse()` calls are done on copies.)"""
    xyz = points[:, :, 0:3]
    net = s3g_util.pointwise_conv3d(points, config.mlp, 'mlp1', weight_decay=config.weight_decay, with_bn=config.with_bn,
                                    with_bias=config.with_bias, is_training=is_training)
    xyz_layers, encoder_feats = [xyz], []
    for l in range(len(config.radius)):
        intra_idx, intra_cnt, intra_dst, indices = s3g_util.build_graph(xyz, config.radius[l], config.nn_uplimit[l],
                                                                        config.num_sample[l], sample_method=config.sample)
        filt_idx = s3g_util.spherical_kernel(xyz, xyz, intra_idx, intra_cnt, intra_dst, config.radius[l], kernel=config.kernel)
        net = _separable_conv3d_block(net, config.channels[l], config.binSize, intra_idx, intra_cnt, filt_idx,
                                      'conv' + str(l + 1), config.multiplier[l], weight_decay=config.weight_decay,
                                      with_bn=config.with_bn, with_bias=config.with_bias, is_training=is_training)
        encoder_feats.append(net)
        if config.num_sample[l] > 1:
            xyz = s3g_util.gather_nd(xyz, indices)
            xyz_layers.append(xyz)
            inter_idx = s3g_util.gather_nd(intra_idx, indices)
            inter_cnt = s3g_util.gather_nd(intra_cnt, indices)
            net = s3g_util.pool3d(net, inter_idx, inter_cnt, method=config.pool_method, scope='pool' + str(l + 1))
    radius, uplimit = config.radius[::-1], config.nn_uplimit[::-1]
    channels, multiplier = config.channels[::-1], config.multiplier[::-1]
    xyz_layers, encoder_feats = xyz_layers[::-1], encoder_feats[::-1]
    for l in range(len(radius)):
        xyz, xyz_unpool = xyz_layers[l], xyz_layers[l + 1]
        intra_idx, intra_cnt, intra_dst, inter_idx, inter_cnt, inter_dst = s3g_util.build_graph_deconv(
            xyz, xyz_unpool, radius[l], uplimit[l])
        filt_idx = s3g_util.spherical_kernel(xyz, xyz, intra_idx, intra_cnt, intra_dst, radius[l], kernel=config.kernel)
        net = _separable_conv3d_block(net, channels[l], config.binSize, intra_idx, intra_cnt, filt_idx,
                                      'deconv' + str(l + 1), multiplier[l], weight_decay=config.weight_decay,
                                      with_bn=config.with_bn, with_bias=config.with_bias, is_training=is_training)
        net = s3g_util.unpool3d(net, inter_idx, inter_cnt, inter_dst, method=config.unpool_method, scope='unpool' + str(l + 1))
        net = torch.cat((net, encoder_feats[l]), dim=2)
    return s3g_util.pointwise_conv3d(net, config.num_cls, scope='logits', with_bn=False, with_bias=config.with_bias,
                                     activation_fn=None, is_training=is_training)


def run(B, N, steps, warmup=2, seed=7, model="modelnet"):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(B, N, 3, generator=g).to(dev)            # unit cube, like a normalised ModelNet cloud
    cfg = ModelNetConfig(N) if model == "modelnet" else S3disConfig(N)
    s3g_util.reset_variables()
    if model != "modelnet":                                    # xyz + 3 extra input features (rgb in the reference)
        pts = torch.cat([pts, torch.rand(B, N, 3, generator=g).to(dev)], dim=2)

    def step():
        s3g_util.clear_collections()
        for p in s3g_util.trainable_variables():
            p.grad = None
        feat = encoder(pts, True, cfg) if model == "modelnet" else s3dis_model(pts, True, cfg).reshape(B, -1)
        loss = feat.square().mean() + sum(s3g_util.get_collection('losses'))
        loss.backward()
        return feat, loss

    for _ in range(warmup):
        feat, loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        feat, loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    grads = [p.grad for p in s3g_util.trainable_variables()]
    what = "modelnet 3-level encoder fwd+bwd (configs[1])" if model == "modelnet" else "s3dis encoder+decoder fwd+bwd (configs[3])"
    return {"workload": what, "B": B, "N": N, "levels": cfg.num_sample,
            "ms_per_step": ms, "points_per_s": B * N / (ms * 1e-3), "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / steps,
            "feature_dim": int(feat.shape[1]), "loss": float(loss), "n_params": len(grads),
            "all_grads_finite": bool(all(gr is not None and torch.isfinite(gr).all() for gr in grads))}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--N", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--model", default="modelnet", choices=["modelnet", "s3dis"])
    ap.add_argument("--no-share", action="store_true", help="one graph transposition per convolution gradient (tf_conv3d.SHARE_PLANS off)")
    a = ap.parse_args()
    S.tf_conv3d.SHARE_PLANS = not a.no_share
    rec = run(a.B, a.N, a.steps, model=a.model)
    rec["share_plans"] = bool(S.tf_conv3d.SHARE_PLANS)
    print(json.dumps(rec))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "bench_%s%s.json" % ("encoder" if a.model == "modelnet" else "s3dis", "_noshare" if a.no_share else "")), "w"), indent=1)
