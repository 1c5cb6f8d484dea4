"""Whole-network reference arm (TEST / BENCH INFRASTRUCTURE ONLY, like everything under oracle/).

The reference's model call graphs are TensorFlow-1 Python (models/SPH3D_*.py) and cannot run here; what CAN run is every
custom op of those graphs as the UNMODIFIED reference kernel (oracle/_ref, bound by ref_gpu.py).  install(pkg) swaps each
custom op the layer library calls for a torch.autograd wrapper around the reference's forward and *Grad launchers and turns
the library's fused paths off (layer tail, split-K weight gradient, tcgen05 pointwise product -> plain torch nodes), so a
step of sph3d-gcn_b200/models then executes the reference's kernels in the reference's order.  Used by
tests/test_model_vs_reference_gpu.py (parity at network scale) and bench.py --impl reference --workload *_model (timing).
"""
import torch

import ref_gpu as ref


class Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, filter, nn_index, nn_count, bin_index):
        ctx.save_for_backward(input, filter, nn_index, nn_count, bin_index)
        return ref.depthwise_conv3d(input.contiguous(), filter.contiguous(), nn_index, nn_count, bin_index)

    @staticmethod
    def backward(ctx, g):
        input, filter, nn_index, nn_count, bin_index = ctx.saved_tensors
        gi, gf = ref.depthwise_conv3d_grad(input.contiguous(), filter.contiguous(), g.contiguous(), nn_index, nn_count, bin_index)
        return gi, gf, None, None, None


class MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, nn_index, nn_count):
        out, max_index = ref.max_pool3d(input.contiguous(), nn_index, nn_count)
        ctx.save_for_backward(input, max_index)
        ctx.mark_non_differentiable(max_index)
        return out, max_index

    @staticmethod
    def backward(ctx, g, _):
        input, max_index = ctx.saved_tensors
        return ref.max_pool3d_grad(input.contiguous(), g.contiguous(), max_index), None, None


class AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, nn_index, nn_count):
        ctx.save_for_backward(input, nn_index, nn_count)
        return ref.avg_pool3d(input.contiguous(), nn_index, nn_count)

    @staticmethod
    def backward(ctx, g):
        input, nn_index, nn_count = ctx.saved_tensors
        return ref.avg_pool3d_grad(input.contiguous(), g.contiguous(), nn_index, nn_count), None, None


class MeanUnpool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, nn_index, nn_count):
        ctx.save_for_backward(input, nn_index, nn_count)
        return ref.mean_interpolate(input.contiguous(), nn_index, nn_count)

    @staticmethod
    def backward(ctx, g):
        input, nn_index, nn_count = ctx.saved_tensors
        return ref.mean_interpolate_grad(input.contiguous(), g.contiguous(), nn_index, nn_count), None, None


class WeightedUnpool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weight, nn_index, nn_count):
        ctx.save_for_backward(input, weight, nn_index, nn_count)
        return ref.weighted_interpolate(input.contiguous(), weight.contiguous(), nn_index, nn_count)

    @staticmethod
    def backward(ctx, g):
        input, weight, nn_index, nn_count = ctx.saved_tensors
        return ref.weighted_interpolate_grad(input.contiguous(), g.contiguous(), weight.contiguous(), nn_index, nn_count), None, None, None


def install(pkg, setattr_fn=None):
    """swap the custom ops of pkg.sph3gcn_util for the reference kernels.  `setattr_fn(obj, name, value)` lets a test pass
    monkeypatch.setattr (undone automatically); without it the swap is permanent for the process and the previous values
    are returned as a list of (obj, name, old) for a manual undo."""
    u = pkg.sph3gcn_util
    undo = []

    def put(obj, name, value):
        if setattr_fn is not None:
            setattr_fn(obj, name, value)
        else:
            undo.append((obj, name, getattr(obj, name)))
            setattr(obj, name, value)

    put(u, "neighbor_fn", ref.build_sphere_neighbor)
    put(u, "spherical_kernel", ref.spherical_kernel)
    put(u, "farthest_point_sample", ref.farthest_point_sample)
    put(u.tf_conv3d, "depthwise_conv3d", lambda i, f, a, b, c: Conv.apply(i, f, a, b, c))
    put(u.tf_pool3d, "max_pool3d", lambda i, a, b: MaxPool.apply(i, a, b))
    put(u.tf_pool3d, "avg_pool3d", lambda i, a, b: AvgPool.apply(i, a, b))
    put(u.tf_unpool3d, "mean_interpolate", lambda i, a, b: MeanUnpool.apply(i, a, b))
    put(u.tf_unpool3d, "weighted_interpolate", lambda i, w, a, b: WeightedUnpool.apply(i, w.detach(), a, b))
    put(u, "FUSED_TAIL", False)
    put(u, "SPLIT_K_WEIGHT_GRAD", False)
    put(u, "ROWS_GEMM", False)
    return undo
