"""Spherical kernel bin index -- mirrors /root/reference/tf_ops/buildkernel/tf_buildkernel.py:10-34."""
import torch

from .. import _lib
from .tf_nnquery import _xyz


# Graph-build side of the convolution plans (tf_conv3d.emit_plans): with this switch on, spherical_kernel also writes
# the per-row bin-sorted edge words (forward) and, when EMIT_PLANS == "train", the transposed graph (backward) of the
# graph it has just binned, so that the convolutions over that graph start from them.
EMIT_PLANS = False


@torch.no_grad()
def spherical_kernel(database, query, nn_index, nn_count, nn_dist, radius, kernel=[8, 2, 3]):
    """Spherical-kernel bin of every edge of a ball-query graph.

    kernel = [n azimuth, p elevation, q radial] divisions; returns filt_index (B,M,K) int32 in [0, n*p*q]:
    0 for the (near-)coincident "self" neighbour, otherwise 1 + radial*p*n + elevation*n + azimuth, computed
    with the reference's mixed fp32/fp64 arithmetic on the sqrt-distances `nn_dist` and the nominal `radius`.
    Padding slots (k >= nn_count) are 0.
    """
    n, p, q = [int(v) for v in kernel]
    database = _xyz(database, "database")
    query = _xyz(query, "query")
    nn_index = _lib.cuda_tensor(nn_index, torch.int32, 3, "nn_index")
    nn_count = _lib.cuda_tensor(nn_count, torch.int32, 2, "nn_count")
    nn_dist = _lib.cuda_tensor(nn_dist, torch.float32, 3, "nn_dist")
    if not radius > 0:
        raise ValueError("Range search requires radius>0, got %r" % (radius,))
    if not (n > 2 and n % 2 == 0):
        raise ValueError("Need n_>2 and n_%%2==0, got %d" % n)
    if not (p > 0 and p % 2 == 0):
        raise ValueError("Need p_>0 and p_%%2==0, got %d" % p)
    if not q > 0:
        raise ValueError("Need q_>0, got %d" % q)
    B, N, _ = database.shape
    M, K = query.shape[1], nn_index.shape[2]
    if nn_index.shape[:2] != (B, M) or nn_count.shape != (B, M) or nn_dist.shape != nn_index.shape:
        raise ValueError("nn_index/nn_count/nn_dist shapes do not match the query")
    filt_index = torch.empty((B, M, K), dtype=torch.int32, device=database.device)
    with torch.cuda.device(database.device):
        rc = _lib.lib().sph3d_spherical_kernel(B, N, M, K, n, p, q, float(radius), _lib.ptr(database),
                                               _lib.ptr(query), _lib.ptr(nn_index), _lib.ptr(nn_count),
                                               _lib.ptr(nn_dist), _lib.ptr(filt_index), _lib.stream_ptr())
    _lib.check(rc, "spherical_kernel")
    if EMIT_PLANS:
        from . import tf_conv3d
        tf_conv3d.emit_plans(nn_index, nn_count, filt_index, n * p * q + 1, N, backward=(EMIT_PLANS == "train"))
    return filt_index
