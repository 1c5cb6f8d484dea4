"""The two stages every SPH3D network is assembled from: an encoder level (graph -> bins -> conv block -> strided
max-pool) and a decoder level (graph pair -> conv block -> unpool -> skip concat).  The reference spells these loops out
in each model file (models/SPH3D_modelnet.py:53-82, models/SPH3D_s3dis.py:54-103); the calls into the layer library are
the reference's, argument for argument."""
import torch

from ..utils import sph3gcn_util as s3g_util


def conv_block(net, graph, scope, channels, multipliers, config, is_training):
    """`_separable_conv3d_block` of the reference models (SPH3D_s3dis.py:22-32): scopes count from 1."""
    nn_index, nn_count, filt_index = graph
    for i, (cout, mult) in enumerate(zip(channels, multipliers), start=1):
        net = s3g_util.separable_conv3d(net, cout, config.binSize, mult, '%s_%d' % (scope, i), nn_index, nn_count,
                                        filt_index, weight_decay=config.weight_decay, with_bn=config.with_bn,
                                        with_bias=config.with_bias, is_training=is_training)
    return net


def encoder_level(xyz, net, level, config, is_training):
    """One resolution of the encoder.  Returns (features at this resolution, xyz of the next one or None,
    pooled features or the unpooled ones when the level does not subsample)."""
    radius, keep = config.radius[level], config.num_sample[level]
    idx, cnt, dst, picked = s3g_util.build_graph(xyz, radius, config.nn_uplimit[level], keep, sample_method=config.sample)
    bins = s3g_util.spherical_kernel(xyz, xyz, idx, cnt, dst, radius, kernel=config.kernel)
    name = 'conv%d' % (level + 1)
    net = conv_block(net, (idx, cnt, bins), name, config.channels[level], config.multiplier[level], config, is_training)
    if keep is None or keep <= 1:
        return net, None, net
    # the rows of the intra-graph that belong to the sampled points ARE the pooling graph (tf.gather_nd in the reference)
    coarse_xyz = s3g_util.gather_nd(xyz, picked)
    pooled = s3g_util.pool3d(net, s3g_util.gather_nd(idx, picked), s3g_util.gather_nd(cnt, picked),
                             method=config.pool_method, scope='pool%d' % (level + 1))
    return net, coarse_xyz, pooled


def decoder(net, xyz_pyramid, skips, config, is_training):
    """Coarse-to-fine half of the segmentation networks.  `xyz_pyramid` / `skips` are fine-to-coarse as the encoder
    produced them; the reference reverses its config lists in place (SPH3D_s3dis.py:80-85), here they are read backwards."""
    depth = len(config.radius)
    for step in range(depth):
        lvl = depth - 1 - step                           # encoder level whose radius / widths this step re-uses
        xyz, xyz_fine = xyz_pyramid[lvl + 1], xyz_pyramid[lvl]
        idx, cnt, dst, up_idx, up_cnt, up_dst = s3g_util.build_graph_deconv(xyz, xyz_fine, config.radius[lvl],
                                                                             config.nn_uplimit[lvl])
        bins = s3g_util.spherical_kernel(xyz, xyz, idx, cnt, dst, config.radius[lvl], kernel=config.kernel)
        net = conv_block(net, (idx, cnt, bins), 'deconv%d' % (step + 1), config.channels[lvl], config.multiplier[lvl],
                         config, is_training)
        net = s3g_util.unpool3d(net, up_idx, up_cnt, up_dst, method=config.unpool_method, scope='unpool%d' % (step + 1))
        net = torch.cat((net, skips[lvl]), dim=2)
    return net


def segmentation_trunk(xyz, net, config, is_training):
    """encoder + decoder; returns per-point features at the input resolution"""
    pyramid, skips = [xyz], []
    s3g_util.prefetch_samples(xyz, config.num_sample, config.sample)      # the whole FPS chain starts now, on the side stream
    for level in range(len(config.radius)):
        feats, coarse, net = encoder_level(xyz, net, level, config, is_training)
        skips.append(feats)
        if coarse is not None:
            xyz = coarse
            pyramid.append(xyz)
    return decoder(net, pyramid, skips, config, is_training)


def masked_mean_per_cloud(loss, inner_label):
    """sum over clouds of the mean loss over that cloud's inner points (0 for a cloud without any):
    the per-cloud tf.where / tf.gather_nd / tf.cond loop of SPH3D_s3dis.get_loss (:116-130) in one expression."""
    inner = (inner_label > 0).to(loss.dtype)
    count = inner.sum(dim=1)
    return ((loss * inner).sum(dim=1) / count.clamp(min=1.0)).sum()
