"""ModelNet40 classification network -- the call graph of /root/reference/models/SPH3D_modelnet.py:33-119 on the
sph3gcn_util mirror: mlp1, three encoder levels (raw xyz re-attached before each), a global max per level, the K=N
global convolution seen from the cloud centroid, then the two-layer classifier with dropout."""
import torch
import torch.nn.functional as F

from ..utils import sph3gcn_util as s3g_util
from . import _stages


def normalize_xyz(points):
    """centre on the centroid, scale the farthest point to the unit sphere (SPH3D_modelnet.py:11-17)"""
    points = points - points.mean(dim=1, keepdim=True)
    reach = points.square().sum(dim=-1, keepdim=True).amax(dim=1, keepdim=True).sqrt()
    return points / reach


def _network(points, is_training, config):
    """points (B, N, 3) -> (logits (B, num_cls), end_points)"""
    B, N = points.shape[0], points.shape[1]
    assert N == config.num_input
    end_points = {}
    xyz = normalize_xyz(points) if config.normalize else points
    viewpoint = xyz.mean(dim=1, keepdim=True)
    layer = dict(weight_decay=config.weight_decay, with_bn=config.with_bn, with_bias=config.with_bias,
                 is_training=is_training)
    net = s3g_util.pointwise_conv3d(xyz, config.mlp, 'mlp1', **layer)
    summary = []
    s3g_util.prefetch_samples(xyz, config.num_sample, config.sample)      # the whole FPS chain starts now, on the side stream
    for level in range(len(config.radius)):
        if config.use_raw:
            net = torch.cat([net, xyz], dim=-1)
        _, coarse, net = _stages.encoder_level(xyz, net, level, config, is_training)
        if coarse is not None:
            xyz = coarse
        summary.append(net.amax(dim=1, keepdim=True))
    far = 100.0                                         # any radius >= 2 links the viewpoint to every remaining point
    idx, cnt, dst = s3g_util.build_global_graph(xyz, viewpoint, far)
    bins = s3g_util.spherical_kernel(xyz, viewpoint, idx, cnt, dst, far, kernel=[8, 2, 1])
    summary.append(s3g_util.separable_conv3d(net, config.global_channels, 17, config.global_multiplier, 'global_conv',
                                             idx, cnt, bins, **layer))
    net = torch.cat(summary, dim=2).reshape(B, -1)
    end_points['global_feat'] = net
    for width, scope in ((512, 'fc1'), (256, 'fc2')):
        net = s3g_util.fully_connected(net, width, scope=scope, **layer)
        net = F.dropout(net, 0.5, training=bool(is_training))
    net = s3g_util.fully_connected(net, config.num_cls, scope='logits', with_bn=False, with_bias=config.with_bias,
                                   activation_fn=None, is_training=is_training)
    return net, end_points


def get_model(points, is_training, config=None):
    # the samplers of build_graph run ahead on a side stream; gather_nd (the only consumer of `indices` here) joins them
    with s3g_util.async_sampling():
        return _network(points, is_training, config)


def get_loss(pred, label, end_points):
    """mean softmax cross entropy, also appended to the 'losses' collection (SPH3D_modelnet.py:112-119)"""
    classify_loss = F.cross_entropy(pred, label.long())
    s3g_util.get_variable_store().collections['losses'].append(classify_loss)
    return classify_loss
