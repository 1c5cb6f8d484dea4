"""S3DIS semantic segmentation network -- the call graph of /root/reference/models/SPH3D_s3dis.py:35-133 on the
sph3gcn_util mirror: four encoder levels, four decoder levels with mean unpooling and skip concats, per-point logits;
the loss counts only the points of a block's inner region."""
import torch
import torch.nn.functional as F

from ..utils import sph3gcn_util as s3g_util
from . import _stages


def normalize_xyz(points):
    """centre x and y on the block's bounding box, keep the height as is (SPH3D_s3dis.py:11-19)"""
    lo, hi = points.amin(dim=1, keepdim=True), points.amax(dim=1, keepdim=True)
    centre = (hi + lo) / 2
    return torch.cat((points[:, :, 0:2] - centre[:, :, 0:2], points[:, :, 2:]), dim=2)


def _network(points, is_training, config):
    """points (B, N, >=3): xyz first; columns 6.. (if any) join the normalised xyz as input features (:36-43)"""
    end_points = {}
    xyz = points[:, :, 0:3].contiguous()
    first = normalize_xyz(xyz) if config.normalize else xyz
    net = torch.cat((first, points[:, :, 6:]), dim=2)
    net = s3g_util.pointwise_conv3d(net, config.mlp, 'mlp1', weight_decay=config.weight_decay, with_bn=config.with_bn,
                                    with_bias=config.with_bias, is_training=is_training)
    net = _stages.segmentation_trunk(xyz, net, config, is_training)
    end_points['feats'] = net
    net = s3g_util.pointwise_conv3d(net, config.num_cls, scope='logits', with_bn=False, with_bias=config.with_bias,
                                    activation_fn=None, is_training=is_training)
    return net, end_points


def get_model(points, is_training, config=None):
    # the samplers of build_graph run ahead on a side stream; gather_nd (the only consumer of `indices` here) joins them
    with s3g_util.async_sampling():
        return _network(points, is_training, config)


def get_loss(pred, label, end_points, inner_label):
    """pred (B, N, num_cls), label / inner_label (B, N)"""
    per_point = F.cross_entropy(pred.reshape(-1, pred.shape[-1]), label.reshape(-1).long(), reduction='none')
    classify_loss = _stages.masked_mean_per_cloud(per_point.reshape(label.shape), inner_label)
    s3g_util.get_variable_store().collections['losses'].append(classify_loss)
    return classify_loss
