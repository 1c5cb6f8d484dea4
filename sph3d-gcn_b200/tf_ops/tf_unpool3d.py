"""Graph unpooling -- mirrors /root/reference/tf_ops/unpooling/tf_unpool3d.py:9-28 (ops + gradients).
input is the COARSE cloud (B,M,C); nn_index (B,N,K) indexes into it; output is (B,N,C)."""
import torch

from .. import _lib


def _check(input, nn_index, nn_count, weight=None):
    input = _lib.cuda_tensor(input, torch.float32, 3, "input")
    nn_index = _lib.cuda_tensor(nn_index, torch.int32, 3, "nn_index")   # tf_unpool3d.cpp:82 rank checks
    nn_count = _lib.cuda_tensor(nn_count, torch.int32, 2, "nn_count")
    if nn_index.shape[0] != input.shape[0] or nn_count.shape != nn_index.shape[:2]:
        raise ValueError("nn_index / nn_count shapes are inconsistent with input")
    if weight is not None:
        weight = _lib.cuda_tensor(weight, torch.float32, 3, "weight")
        if weight.shape != nn_index.shape:
            raise ValueError("weight must have the shape of nn_index")
    return input, nn_index, nn_count, weight


def _dims(input, nn_index):
    B, M, C = input.shape
    return B, nn_index.shape[1], M, C, nn_index.shape[2]


# True: the gradients transpose the graph and gather (streaming form: four coarse points per warp, 32-channel columns, work
# items in order of list length; csrc/conv_bwd_t.cu); False: vector-reduction scatter (csrc/pool3d.cu).  Measured
# (profiles/r2_pool_stream.json): 1.35 vs 1.96 ms at the Cfg-T unpool shape, 0.39 vs 0.47 ms at the S3DIS one.
GATHER_FORM_GRAD = True


def _scratch(B, N, M, C, K, device):
    nbytes = _lib.lib().sph3d_interpolate_grad_workspace_bytes(B, N, M, C, K) if GATHER_FORM_GRAD else 0
    return (torch.empty((nbytes // 4,), dtype=torch.int32, device=device) if nbytes else None), nbytes


def mean_interpolate_grad(input, grad_output, nn_index, nn_count):
    input, nn_index, nn_count, _ = _check(input, nn_index, nn_count)
    grad_output = _lib.cuda_tensor(grad_output, torch.float32, 3, "grad_output")
    B, N, M, C, K = _dims(input, nn_index)
    grad_input = torch.empty((B, M, C), dtype=torch.float32, device=input.device)
    ws, ws_bytes = _scratch(B, N, M, C, K, input.device)
    with torch.cuda.device(input.device):
        rc = _lib.lib().sph3d_mean_interpolate_grad(B, N, M, C, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                                    _lib.ptr(grad_output), _lib.ptr(grad_input), _lib.ptr(ws), ws_bytes,
                                                    _lib.stream_ptr())
    _lib.check(rc, "mean_interpolate_grad")
    return grad_input


def weighted_interpolate_grad(input, grad_output, weight, nn_index, nn_count):
    input, nn_index, nn_count, weight = _check(input, nn_index, nn_count, weight)
    grad_output = _lib.cuda_tensor(grad_output, torch.float32, 3, "grad_output")
    B, N, M, C, K = _dims(input, nn_index)
    grad_input = torch.empty((B, M, C), dtype=torch.float32, device=input.device)
    ws, ws_bytes = _scratch(B, N, M, C, K, input.device)
    with torch.cuda.device(input.device):
        rc = _lib.lib().sph3d_weighted_interpolate_grad(B, N, M, C, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                                        _lib.ptr(grad_output), _lib.ptr(weight),
                                                        _lib.ptr(grad_input), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    _lib.check(rc, "weighted_interpolate_grad")
    return grad_input


class _MeanInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, nn_index, nn_count):
        B, N, M, C, K = _dims(input, nn_index)
        output = torch.empty((B, N, C), dtype=torch.float32, device=input.device)
        with torch.cuda.device(input.device):
            rc = _lib.lib().sph3d_mean_interpolate(B, N, M, C, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                                   _lib.ptr(input), _lib.ptr(output), _lib.stream_ptr())
        _lib.check(rc, "mean_interpolate")
        ctx.save_for_backward(input, nn_index, nn_count)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, nn_index, nn_count = ctx.saved_tensors
        return mean_interpolate_grad(input, grad_output.contiguous(), nn_index, nn_count), None, None


class _WeightedInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weight, nn_index, nn_count):
        B, N, M, C, K = _dims(input, nn_index)
        output = torch.empty((B, N, C), dtype=torch.float32, device=input.device)
        with torch.cuda.device(input.device):
            rc = _lib.lib().sph3d_weighted_interpolate(B, N, M, C, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                                       _lib.ptr(input), _lib.ptr(weight), _lib.ptr(output),
                                                       _lib.stream_ptr())
        _lib.check(rc, "weighted_interpolate")
        ctx.save_for_backward(input, weight, nn_index, nn_count)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, nn_index, nn_count = ctx.saved_tensors
        # no gradient to weight, as in the reference (tf_unpool3d.py:28 returns None for it)
        return weighted_interpolate_grad(input, grad_output.contiguous(), weight, nn_index, nn_count), None, None, None


def mean_interpolate(input, nn_index, nn_count):
    input, nn_index, nn_count, _ = _check(input, nn_index, nn_count)
    return _MeanInterpolate.apply(input, nn_index, nn_count)


def weighted_interpolate(input, weight, nn_index, nn_count):
    input, nn_index, nn_count, weight = _check(input, nn_index, nn_count, weight)
    return _WeightedInterpolate.apply(input, weight.detach(), nn_index, nn_count)
