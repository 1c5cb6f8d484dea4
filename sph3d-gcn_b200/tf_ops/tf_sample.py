"""Samplers -- mirrors /root/reference/tf_ops/sampling/tf_sample.py:15-49."""
import torch

from .. import _lib
from .tf_nnquery import _xyz


@torch.no_grad()
def farthest_point_sample(neursize, database):
    """Farthest point sampling: `neursize` point ids per cloud, (B, neursize) int32.

    Starts from point 0 of each cloud of `database` (B, N, 3+) and repeatedly takes the point whose squared distance to
    the set picked so far is largest, with the reference kernel's tie rule -- bit-identical picks.
    """
    database = _xyz(database, "database")
    neursize = int(neursize)
    if not neursize > 0:
        raise ValueError("FarthestPointSample expects positive npoint")
    B, N, _ = database.shape
    L = _lib.lib()
    out = torch.empty((B, neursize), dtype=torch.int32, device=database.device)
    nbytes = L.sph3d_farthest_point_sample_workspace_bytes(B, N, neursize)
    temp = torch.empty((max(nbytes, 4) // 4,), dtype=torch.float32, device=database.device) if nbytes else None
    with torch.cuda.device(database.device):
        rc = L.sph3d_farthest_point_sample(B, N, neursize, _lib.ptr(database), _lib.ptr(temp), nbytes,
                                           _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "farthest_point_sample")
    return out


@torch.no_grad()
def inverse_density_sample(neursize, probability):
    '''Gumbel-max top-k on log(probability) (tf_sample.py:27-41).  RNG-driven: torch's generator
    replaces TensorFlow's, so only the distribution (not the draw) matches the reference.'''
    logits = torch.log(probability)
    z = -torch.log(-torch.log(torch.rand_like(logits)))
    _, neuron_index = torch.topk(logits + z, int(neursize), dim=-1)
    return neuron_index.to(torch.int32)


@torch.no_grad()
def random_sample(neursize, database):
    '''Uniform indices WITH replacement, as tf.random.uniform(minval=0, maxval=num_points) (tf_sample.py:44-49).'''
    B, N = database.shape[0], database.shape[1]
    return torch.randint(0, N, (B, int(neursize)), device=database.device, dtype=torch.int32)
