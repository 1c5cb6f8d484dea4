"""ShapeNet part segmentation network -- the call graph of /root/reference/models/SPH3D_shapenet.py:33-123: the S3DIS
trunk fed with the raw points, plus a second point-wise layer whose output is concatenated with the mlp1 features
before the classifier."""
import torch
import torch.nn.functional as F

from ..utils import sph3gcn_util as s3g_util
from . import _stages
from .SPH3D_modelnet import normalize_xyz   # noqa: F401  (same centroid / unit-sphere normalisation, unused when config.normalize is False)


def _network(points, num_cls, is_training, config):
    end_points = {}
    xyz = points[:, :, 0:3].contiguous()
    layer = dict(weight_decay=config.weight_decay, with_bn=config.with_bn, with_bias=config.with_bias,
                 is_training=is_training)
    stem = s3g_util.pointwise_conv3d(points, config.mlp, 'mlp1', **layer)
    net = _stages.segmentation_trunk(xyz, stem, config, is_training)
    net = s3g_util.pointwise_conv3d(net, config.mlp, 'mlp2', **layer)
    net = torch.cat((net, stem), dim=2)
    end_points['feats'] = net                       # after the skip concat, as SPH3D_shapenet.py:108-111
    net = s3g_util.pointwise_conv3d(net, num_cls, scope='logits', with_bn=False, with_bias=config.with_bias,
                                    activation_fn=None, is_training=is_training)
    return net, end_points


def get_model(points, num_cls, is_training, config=None):
    # the samplers of build_graph run ahead on a side stream; gather_nd (the only consumer of `indices` here) joins them
    with s3g_util.async_sampling():
        return _network(points, num_cls, is_training, config)


def get_loss(pred, label, end_points):
    classify_loss = F.cross_entropy(pred.reshape(-1, pred.shape[-1]), label.reshape(-1).long())
    s3g_util.get_variable_store().collections['losses'].append(classify_loss)
    return classify_loss
