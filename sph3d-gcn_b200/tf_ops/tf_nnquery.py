"""Neighbour queries -- mirrors /root/reference/tf_ops/nnquery/tf_nnquery.py:9-60.
No gradient (ops.NoGradient there; plain non-differentiable index outputs here)."""
import torch

from .. import _lib


def _xyz(t, name):
    t = _lib.cuda_tensor(t, torch.float32, 3, name)
    if t.shape[2] < 3:
        raise ValueError("Shape of %s points requires to be (batch, npoint, 3)" % name)
    return t[:, :, 0:3].contiguous()       # reference slices [:, :, 0:3] (tf_nnquery.py:26-27)


@torch.no_grad()
def build_sphere_neighbor(database, query, radius=0.1, dilation_rate=None, nnsample=100):
    """Ball query with the reference's semantics.

    database (B,N,>=3) and query (B,M,>=3) float32 CUDA tensors (only xyz is used); the effective radius is
    radius * dilation_rate when a dilation rate is given; at most `nnsample` neighbours per query.
    Returns nn_index (B,M,K) int32 -- the first K in-range database ids in ascending order, zero padded --,
    nn_count (B,M) int32 in [1,K] and nn_dist (B,M,K) float32 = sqrt of the Euclidean distance, zero padded.
    """
    database = _xyz(database, "database")
    query = _xyz(query, "query")
    if dilation_rate is not None:
        radius = dilation_rate * radius
    if not radius > 0:
        raise ValueError("Range search requires radius>0, got %r" % (radius,))
    if not nnsample > 0:
        raise ValueError("BuildSphereNeighbor requires nn_sample>0, got %r" % (nnsample,))
    B, N, _ = database.shape
    if query.shape[0] != B:
        raise ValueError("database and query must have the same batch size")
    M, K = query.shape[1], int(nnsample)
    dev = database.device
    nn_index = torch.empty((B, M, K), dtype=torch.int32, device=dev)
    nn_count = torch.empty((B, M), dtype=torch.int32, device=dev)
    nn_dist = torch.empty((B, M, K), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws_bytes = L.sph3d_build_sphere_neighbor_workspace_bytes(B, N, M, K)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
    with torch.cuda.device(dev):
        rc = L.sph3d_build_sphere_neighbor(B, N, M, K, float(radius), _lib.ptr(database), _lib.ptr(query),
                                           _lib.ptr(nn_index), _lib.ptr(nn_count), _lib.ptr(nn_dist),
                                           _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    _lib.check(rc, "build_sphere_neighbor")
    return nn_index, nn_count, nn_dist


@torch.no_grad()
def build_cube_neighbor(database, query, length=0.1, dilation_rate=None, nnsample=100, gridsize=3):
    """Axis-aligned cube query with a gridsize^3 bin per hit (tf_nnquery_gpu.cu:72-113).

    Per query point: the first `nnsample` database points (ascending id) whose offset lies inside the cube of edge
    `length` (times `dilation_rate`), each with the cell of the gridsize x gridsize x gridsize subdivision it falls in.
    Returns nn_index (B, M, nnsample, 2) int32 = (point id, x*g*g + y*g + z) pairs and nn_count (B, M) int32, which may
    be 0 (no radius growth here).  No model of the reference uses it; kept for API completeness.
    """
    database = _xyz(database, "database")
    query = _xyz(query, "query")
    if dilation_rate is not None:
        length = dilation_rate * length
    if not length > 0:
        raise ValueError("Cube size requires length>0, got %r" % (length,))
    if not nnsample > 0:
        raise ValueError("BuildSphereNeighbor requires nn_sample>0, got %r" % (nnsample,))
    if not gridsize > 0:
        raise ValueError("Need grid_size_>0, got %r" % (gridsize,))
    B, N, _ = database.shape
    M, K = query.shape[1], int(nnsample)
    dev = database.device
    nn_index = torch.empty((B, M, K, 2), dtype=torch.int32, device=dev)
    nn_count = torch.empty((B, M), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().sph3d_build_cube_neighbor(B, N, M, int(gridsize), K, float(length), _lib.ptr(database),
                                                  _lib.ptr(query), _lib.ptr(nn_index), _lib.ptr(nn_count),
                                                  _lib.stream_ptr())
    _lib.check(rc, "build_cube_neighbor")
    return nn_index, nn_count
