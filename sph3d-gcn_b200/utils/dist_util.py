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

Fewer clouds than ranks (BASELINE configs[4]: B = 4 on 8 GPUs; SURVEY.md section 8e "B < G"): the ranks that hold the
SAME cloud split its QUERY points.  Graph construction stays whole per cloud (the radius chain depends on the query's
index inside its cloud, so a sliced ball query would not be the reference's); the convolution is exact on any slice of
the rows -- output rows [m0, m1) need only those rows of the graph -- and its backward needs ONE real exchange:
grad_input is a sum over all rows that reference a point, i.e. over the shards, so it is all-reduced inside the cloud's
rank group (`query_sharded`, NCCL over NVLink); grad_filter partial sums join the weight-gradient all-reduce as before.
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


def cloud_shard(rank, world, clouds):
    """(first cloud, one past the last cloud, query shard, query shards) of `rank`: with world <= clouds every rank takes a
    slice of whole clouds; with more ranks than clouds, world // clouds ranks share each cloud and split its query points."""
    rank, world, clouds = int(rank), int(world), int(clouds)
    if world <= clouds:
        lo, hi = shard_bounds(clouds, rank, world)
        return lo, hi, 0, 1
    if world % clouds:
        raise ValueError("more ranks than clouds: the number of ranks must be a multiple of the number of clouds")
    per = world // clouds
    return rank // per, rank // per + 1, rank % per, per


class _SumGradOverGroup(torch.autograd.Function):
    """identity in the forward pass; the gradient is SUM all-reduced over `group` (the ranks that hold other query shards
    of the same clouds)"""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def query_sharded(op, input, index_tensors, shard, nshards, group=None):
    """Rows [m0, m1) = shard_bounds(M, shard, nshards) of a row-wise graph op:  op(input, *sliced index tensors).

    `op` is any op of the hot path whose output row m depends on row m of the index tensors only (depthwise_conv3d with
    the filter bound, pooling, unpooling); `index_tensors` are (B, M, ...) tensors of the whole graph.  The returned rows
    are exact; in the backward pass the gradient w.r.t. `input` is summed over `group`.  -> (output rows, (m0, m1))"""
    m0, m1 = shard_bounds(index_tensors[0].shape[1], shard, nshards)
    sliced = [t[:, m0:m1].contiguous() for t in index_tensors]
    x = _SumGradOverGroup.apply(input, group) if nshards > 1 else input
    return op(x, *sliced), (m0, m1)


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


class GradBuckets(object):
    """Persistent flat gradient storage + bucketed, overlapped all-reduce.

    Every parameter's .grad is a VIEW into one flat fp32 buffer (no torch.cat round trip before the collective, no copy
    back after it).  Parameters are grouped into `n_buckets` contiguous buckets in REVERSE registration order -- the order
    in which a backward pass finishes them -- and a post-accumulate-grad hook starts a bucket's all-reduce on a side
    stream the moment its last gradient is written, so the collective of the decoder's gradients runs under the backward
    of the encoder.  finish() joins the side stream (and applies the 1/world average).  The hooks, the side stream and
    the NCCL calls are all stream-ordered, so a step that uses this can be captured in a CUDA graph as a whole.

        buckets = GradBuckets(params, n_buckets=4)     # after the variables exist
        buckets.zero(); loss.backward(); buckets.finish()
    With torch.distributed uninitialised (or world size 1) zero()/finish() still work and nothing is reduced.
    """

    def __init__(self, params, n_buckets=4, group=None, average=True):
        self.params = [p for p in params if p.requires_grad]
        self.group, self.average = group, average
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        order = list(reversed(self.params))                      # backward finishes the last layers first
        n_buckets = max(1, min(int(n_buckets), len(order)))
        target = (total + n_buckets - 1) // n_buckets
        self.bounds, self.bucket_of, self.pending0 = [], {}, []
        off, start, count = 0, 0, 0
        for p in order:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self.bucket_of[id(p)] = len(self.bounds)
            off += p.numel()
            count += 1
            if off - start >= target or p is order[-1]:
                self.bounds.append((start, off))
                self.pending0.append(count)
                start, count = off, 0
        self.pending = list(self.pending0)
        self.cuda = dev.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=dev) if self.cuda else None
        self.works = []
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.bytes_reduced = 0

    def zero(self):
        """reset the gradient storage in place (the views stay attached) and re-arm the buckets"""
        self.flat.zero_()
        if any(p.grad is None for p in self.params):            # a caller set .grad = None in between: re-attach the views
            self.reattach()
        self.pending = list(self.pending0)
        self.works = []
        self.bytes_reduced = 0

    def reattach(self):
        off = 0
        for p in reversed(self.params):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def _launch(self, b):
        if self.world == 1:
            return
        lo, hi = self.bounds[b]
        chunk = self.flat[lo:hi]
        self.bytes_reduced += chunk.numel() * 4
        if self.cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _hook(self, p):
        b = self.bucket_of[id(p)]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def finish(self):
        """every bucket reduced (buckets whose hooks never fired -- unused parameters -- are reduced here) and averaged;
        returns the bytes that crossed the collective"""
        for b, left in enumerate(self.pending):
            if left > 0:
                self.pending[b] = 0
                self._launch(b)
        for w in self.works:
            w.wait()
        self.works = []
        if self.cuda and self.world > 1:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)
        if self.average and self.world > 1:
            self.flat.mul_(1.0 / self.world)
        return self.bytes_reduced

    def close(self):
        for h in self.handles:
            h.remove()
        self.handles = []
