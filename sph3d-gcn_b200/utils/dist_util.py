"""Data-parallel plumbing: the reference is single-GPU; BASELINE.json asks for
clouds sharded over the batch across 1/2/4/8 B200s with an all-reduce of the weight gradients only.

Every op of the hot path is independent per cloud, so a rank simply runs the unchanged kernels on its
slice [shard_bounds(B, rank, world)] -- there is no data-path collective.  The only cross-rank step is
the SUM all-reduce of weight gradients (grad_filter of every depthwise layer + pointwise/BN
parameters), done on ONE flat fp32 bucket (a few MB at most: latency-, not bandwidth-bound on
NVLink 5 / NVSwitch, so a single fused call beats per-tensor calls).  Works with backend 'nccl' on
GPUs and 'gloo' on CPU (tests/test_dist_gloo_cpu.py, world_size 2).

Q1 caveat: the ball query's radius chain depends on the LOCAL batch index (i%32, i//32); for a
per-rank batch <= 32 the i//32 term vanishes, so sharding does not change any result.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous, balanced slice of `total` clouds for `rank`: sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world):
    """Slice every tensor of a list along dim 0 to this rank's clouds."""
    lo, hi = shard_bounds(tensors[0].shape[0], rank, world)
    return [t[lo:hi].contiguous() for t in tensors]


def flatten_grads(params):
    grads = [p.grad if getattr(p, "grad", None) is not None else torch.zeros_like(p) for p in params]
    return torch.cat([g.reshape(-1) for g in grads]) if grads else torch.zeros(0)


def allreduce_gradients(tensors, group=None, average=False):
    """In-place SUM (or mean) all-reduce of a list of gradient tensors through one flat bucket.
    Returns the number of bytes reduced.  No-op when torch.distributed is not initialised."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return flat.numel() * flat.element_size()
