"""Depthwise spherical convolution -- mirrors /root/reference/tf_ops/convolution/tf_conv3d.py:10-32
(op + its registered gradient).  Index inputs receive no gradient, as there."""
import torch

from .. import _lib


def _check(input, filter, nn_index, nn_count, bin_index):
    input = _lib.cuda_tensor(input, torch.float32, 3, "input")
    filter = _lib.cuda_tensor(filter, torch.float32, 3, "filter")
    nn_index = _lib.cuda_tensor(nn_index, torch.int32, 3, "nn_index")
    nn_count = _lib.cuda_tensor(nn_count, torch.int32, 2, "nn_count")
    bin_index = _lib.cuda_tensor(bin_index, torch.int32, 3, "bin_index")
    if filter.shape[1] != input.shape[2]:
        raise ValueError("Input Channel size error of the filter")        # tf_conv3d.cpp:67
    if bin_index.shape != nn_index.shape or nn_count.shape != nn_index.shape[:2] or nn_index.shape[0] != input.shape[0]:
        raise ValueError("nn_index / nn_count / bin_index shapes are inconsistent")
    return input, filter, nn_index, nn_count, bin_index


def _forward(input, filter, nn_index, nn_count, bin_index, graph=None):
    B, N, C = input.shape
    F, _, r = filter.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    if (SHARE_PLANS and FORWARD_PLANS and graph is not None
            and _lib.lib().sph3d_depthwise_conv3d_planned_supported(B, N, M, F, C, r, K)):
        plan = _shared_plan("fwd", graph[0], graph[1], graph[2], F, N)
        if plan is not None:
            return depthwise_conv3d_planned(input, filter, nn_count, plan, K)
    output = torch.empty((B, M, C * r), dtype=torch.float32, device=input.device)
    with torch.cuda.device(input.device):
        rc = _lib.lib().sph3d_depthwise_conv3d(B, N, M, F, C, r, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                               _lib.ptr(bin_index), _lib.ptr(input), _lib.ptr(filter),
                                               _lib.ptr(output), _lib.stream_ptr())
    _lib.check(rc, "depthwise_conv3d")
    return output


def depthwise_conv3d_grad(input, filter, grad_output, nn_index, nn_count, bin_index):
    """conv3d_module.depthwise_conv3d_grad of the reference (tf_conv3d.py:30): -> (grad_input, grad_filter)."""
    input, filter, nn_index, nn_count, bin_index = _check(input, filter, nn_index, nn_count, bin_index)
    grad_output = _lib.cuda_tensor(grad_output, torch.float32, 3, "grad_output")
    B, N, C = input.shape
    F, _, r = filter.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    if grad_output.shape != (B, M, C * r):
        raise ValueError("grad_output must be (batch, mpoint, in_channels*multiplier)")
    L = _lib.lib()
    grad_input = torch.empty((B, N, C), dtype=torch.float32, device=input.device)
    grad_filter = torch.empty((F, C, r), dtype=torch.float32, device=input.device)
    ws_bytes = L.sph3d_depthwise_conv3d_grad_workspace_bytes(B, N, M, F, C, r, K)
    ws = torch.empty((max(ws_bytes // 4, 1),), dtype=torch.float32, device=input.device)
    with torch.cuda.device(input.device):
        rc = L.sph3d_depthwise_conv3d_grad(B, N, M, F, C, r, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                           _lib.ptr(bin_index), _lib.ptr(input), _lib.ptr(filter),
                                           _lib.ptr(grad_output), _lib.ptr(grad_input), _lib.ptr(grad_filter),
                                           _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    _lib.check(rc, "depthwise_conv3d_grad")
    return grad_input, grad_filter


def conv_sort(nn_index, nn_count, bin_index, num_bins, npoint):
    """Graph-only half of the forward pass (sph3d_conv_sort): every row's edges grouped by bin, packed as one int32
    word per edge.  Returns an opaque plan tensor for depthwise_conv3d_planned, or None where the planned form does
    not apply (num_bins > 128).  Depends only on the graph and num_bins."""
    nn_index = _lib.cuda_tensor(nn_index, torch.int32, 3, "nn_index")
    nn_count = _lib.cuda_tensor(nn_count, torch.int32, 2, "nn_count")
    bin_index = _lib.cuda_tensor(bin_index, torch.int32, 3, "bin_index")
    if bin_index.shape != nn_index.shape or nn_count.shape != nn_index.shape[:2]:
        raise ValueError("nn_index / nn_count / bin_index shapes are inconsistent")
    B, M, K = nn_index.shape
    npoint, num_bins = int(npoint), int(num_bins)
    L = _lib.lib()
    nbytes = L.sph3d_conv_sort_bytes(B, npoint, M, num_bins, K)
    if nbytes == 0:
        return None
    plan = torch.empty((nbytes // 4,), dtype=torch.int32, device=nn_index.device)
    with torch.cuda.device(nn_index.device):
        rc = L.sph3d_conv_sort(B, npoint, M, num_bins, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                               _lib.ptr(bin_index), _lib.ptr(plan), nbytes, _lib.stream_ptr())
    _lib.check(rc, "conv_sort")
    return plan


def depthwise_conv3d_planned(input, filter, nn_count, plan, nnsample):
    """depthwise_conv3d with the per-row bin grouping hoisted out (plan from conv_sort); same values bit for bit."""
    input = _lib.cuda_tensor(input, torch.float32, 3, "input")
    filter = _lib.cuda_tensor(filter, torch.float32, 3, "filter")
    nn_count = _lib.cuda_tensor(nn_count, torch.int32, 2, "nn_count")
    B, N, C = input.shape
    F, _, r = filter.shape
    M, K = nn_count.shape[1], int(nnsample)
    if filter.shape[1] != C:
        raise ValueError("Input Channel size error of the filter")
    L = _lib.lib()
    if plan is None or not L.sph3d_depthwise_conv3d_planned_supported(B, N, M, F, C, r, K):
        raise ValueError("the planned convolution does not cover this shape; use depthwise_conv3d")
    output = torch.empty((B, M, C * r), dtype=torch.float32, device=input.device)
    with torch.cuda.device(input.device):
        rc = L.sph3d_depthwise_conv3d_planned(B, N, M, F, C, r, K, _lib.ptr(nn_count), _lib.ptr(plan), plan.numel() * 4,
                                              _lib.ptr(input), _lib.ptr(filter), _lib.ptr(output), _lib.stream_ptr())
    _lib.check(rc, "depthwise_conv3d_planned")
    return output


def conv_transpose(nn_index, nn_count, bin_index, num_bins, npoint):
    """Graph-only half of the backward pass (sph3d_conv_transpose): per input point, the (output point, bin)
    pairs that reference it, grouped by bin.  Returns an opaque int32 plan tensor for
    depthwise_conv3d_grad_planned, or None where the planned form does not apply (num_bins > 72).  The plan
    depends only on the graph, num_bins and npoint (points in the input cloud), so one plan serves every
    convolution applied over that graph."""
    nn_index = _lib.cuda_tensor(nn_index, torch.int32, 3, "nn_index")
    nn_count = _lib.cuda_tensor(nn_count, torch.int32, 2, "nn_count")
    bin_index = _lib.cuda_tensor(bin_index, torch.int32, 3, "bin_index")
    if bin_index.shape != nn_index.shape or nn_count.shape != nn_index.shape[:2]:
        raise ValueError("nn_index / nn_count / bin_index shapes are inconsistent")
    B, M, K = nn_index.shape
    npoint, num_bins = int(npoint), int(num_bins)
    L = _lib.lib()
    nbytes = L.sph3d_conv_transpose_bytes(B, npoint, M, num_bins, K)
    if nbytes == 0:
        return None
    plan = torch.empty((nbytes // 4,), dtype=torch.int32, device=nn_index.device)
    with torch.cuda.device(nn_index.device):
        rc = L.sph3d_conv_transpose(B, npoint, M, num_bins, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                    _lib.ptr(bin_index), _lib.ptr(plan), nbytes, _lib.stream_ptr())
    _lib.check(rc, "conv_transpose")
    return plan


def depthwise_conv3d_grad_planned(input, filter, grad_output, nn_count, plan, nnsample):
    """depthwise_conv3d_grad with the graph transposition hoisted out (plan from conv_transpose)."""
    input = _lib.cuda_tensor(input, torch.float32, 3, "input")
    filter = _lib.cuda_tensor(filter, torch.float32, 3, "filter")
    grad_output = _lib.cuda_tensor(grad_output, torch.float32, 3, "grad_output")
    nn_count = _lib.cuda_tensor(nn_count, torch.int32, 2, "nn_count")
    B, N, C = input.shape
    F, _, r = filter.shape
    M, K = nn_count.shape[1], int(nnsample)
    if filter.shape[1] != C:
        raise ValueError("Input Channel size error of the filter")
    if grad_output.shape != (B, M, C * r):
        raise ValueError("grad_output must be (batch, mpoint, in_channels*multiplier)")
    L = _lib.lib()
    ws_bytes = L.sph3d_depthwise_conv3d_grad_planned_workspace_bytes(B, N, M, F, C, r, K)
    if ws_bytes == 0 or plan is None:
        raise ValueError("the planned gradient does not cover this shape; use depthwise_conv3d_grad")
    grad_input = torch.empty((B, N, C), dtype=torch.float32, device=input.device)
    grad_filter = torch.empty((F, C, r), dtype=torch.float32, device=input.device)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=input.device)
    with torch.cuda.device(input.device):
        rc = L.sph3d_depthwise_conv3d_grad_planned(B, N, M, F, C, r, K, _lib.ptr(nn_count), _lib.ptr(plan),
                                                   plan.numel() * 4, _lib.ptr(input), _lib.ptr(filter),
                                                   _lib.ptr(grad_output), _lib.ptr(grad_input), _lib.ptr(grad_filter),
                                                   _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    _lib.check(rc, "depthwise_conv3d_grad_planned")
    return grad_input, grad_filter


# Plan sharing.  The reference's models apply two convolutions per level over one graph (models/SPH3D_*.py call
# separable_conv3d twice with the same nn_index / nn_count / filt_index), so their forward passes can share one sorted
# edge list and their gradients one transposed graph.  With SHARE_PLANS on, a plan is built at the first pass that needs
# it -- or by tf_buildkernel.spherical_kernel itself when its EMIT_PLANS switch is on (the graph-build side) -- and kept
# as an attribute of the bin_index tensor OBJECT (it lives and dies with the graph; in-place edits of an index tensor
# change its _version and invalidate it).  With SHARE_PLANS off every call runs the one-call entry points.
SHARE_PLANS = True
# The forward's plan (per-row bin sort hoisted out of the kernel) is OFF by default: measured on B200 the in-kernel sort
# costs the forward 5 % (0.589 -> 0.559 ms at Cfg-T) while writing the words costs 0.14 ms per graph, so two
# convolutions per graph do not pay it back (profiles/r2_stage_a.json).  The entry points stay (C ABI + tests).
FORWARD_PLANS = False
_BUILDERS = {"fwd": lambda *a: conv_sort(*a), "bwd": lambda *a: conv_transpose(*a)}


def _shared_plan(kind, nn_index, nn_count, bin_index, num_bins, npoint):
    key = (kind, int(num_bins), int(npoint), tuple(nn_index.shape), nn_index.data_ptr(), nn_count.data_ptr(),
           nn_index._version, nn_count._version, bin_index._version)
    cache = getattr(bin_index, "_sph3d_plans", None)
    if cache is not None and key in cache:
        return cache[key]
    plan = _BUILDERS[kind](nn_index, nn_count, bin_index, num_bins, npoint)
    try:
        if cache is None or any(k[1:] != key[1:] for k in cache):       # a changed graph replaces the stale plans
            cache = {}
            bin_index._sph3d_plans = cache
        cache[key] = plan
    except Exception:                                    # objects that refuse attributes: no sharing, still correct
        pass
    return plan


def emit_plans(nn_index, nn_count, bin_index, num_bins, npoint, backward=True):
    """build the graph-only plans of the convolution now (graph-build side) and hang them off bin_index"""
    if FORWARD_PLANS:
        _shared_plan("fwd", nn_index, nn_count, bin_index, num_bins, npoint)
    if backward:
        _shared_plan("bwd", nn_index, nn_count, bin_index, num_bins, npoint)


def _use_planned(C, r, rows):
    """whether a shared plan pays for this layer.  r = 1: always (the one-call form transposes anyway).  r = 2: the
    transposed form gathers C*r floats per edge where the row-owned one gathers C and reduces C; measured over every layer
    of the S3DIS network (profiles/r2_s3dis_layers.json, plan shared by the level's two convolutions) it wins up to
    C*r = 512 on levels of >= 4096 rows (0.24 vs 0.39 ms at C = 64, 0.22 vs 0.32 ms at C = 256) and loses beyond
    (C = 512: 0.76 vs 0.60 ms) and on the deep, tiny levels, where the plan build is not paid back."""
    return r == 1 or (r == 2 and C * r <= 512 and rows >= 4096)


def backward_on_graph(input, filter, grad_output, graph):
    """(grad_input, grad_filter) for the graph tensor OBJECTS `graph` = (nn_index, nn_count, bin_index): the planned form
    when the shape pays and a transposed graph hangs off (or can be hung on) bin_index, else the one-call gradient."""
    g_idx, g_cnt, g_bin = graph
    grad_output = grad_output.contiguous()
    F, C, r = filter.shape
    if SHARE_PLANS and _use_planned(C, r, input.shape[0] * g_idx.shape[1]):
        L = _lib.lib()
        B, N = input.shape[0], input.shape[1]
        M, K = g_idx.shape[1], g_idx.shape[2]
        if L.sph3d_depthwise_conv3d_grad_planned_workspace_bytes(B, N, M, F, C, r, K) > 0:
            plan = _shared_plan("bwd", g_idx, g_cnt, g_bin, F, N)
            if plan is not None:
                return depthwise_conv3d_grad_planned(input, filter, grad_output, g_cnt, plan, K)
    return depthwise_conv3d_grad(input, filter, grad_output, g_idx, g_cnt, g_bin)


class _DepthwiseConv3d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, filter, nn_index, nn_count, bin_index):
        ctx.save_for_backward(input, filter, nn_index, nn_count, bin_index)
        ctx.graph = (nn_index, nn_count, bin_index)      # the tensor OBJECTS (a shared plan hangs off bin_index)
        return _forward(input, filter, nn_index, nn_count, bin_index, ctx.graph)

    @staticmethod
    def backward(ctx, grad_output):
        input, filter = ctx.saved_tensors[:2]
        gi, gf = backward_on_graph(input, filter, grad_output, ctx.graph)
        return gi, gf, None, None, None


def depthwise_conv3d(input, filter, nn_index, nn_count, bin_index):
    """Depthwise spherical convolution over a neighbour graph (differentiable in `input` and `filter`).

    input (B, N, C) float32 features of the database cloud; filter (F, C, r) float32, one C x r slab per spherical bin;
    nn_index (B, M, K) int32 neighbour ids of each of the M query points, nn_count (B, M) int32 how many of the K slots
    are valid, bin_index (B, M, K) int32 the bin of every edge (output of spherical_kernel).
    Returns (B, M, C*r) float32: channel c*r+j of a query point is the mean over its valid edges of
    input[neighbour, c] * filter[bin, c, j].
    """
    input, filter, nn_index, nn_count, bin_index = _check(input, filter, nn_index, nn_count, bin_index)
    return _DepthwiseConv3d.apply(input, filter, nn_index, nn_count, bin_index)
