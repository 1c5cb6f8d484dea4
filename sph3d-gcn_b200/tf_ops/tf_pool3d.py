"""Graph pooling -- mirrors /root/reference/tf_ops/pooling/tf_pool3d.py:9-28 (ops + gradients)."""
import torch

from .. import _lib


def _check(input, nn_index, nn_count):
    input = _lib.cuda_tensor(input, torch.float32, 3, "input")
    nn_index = _lib.cuda_tensor(nn_index, torch.int32, 3, "nn_index")   # tf_pool3d.cpp:84 rank checks
    nn_count = _lib.cuda_tensor(nn_count, torch.int32, 2, "nn_count")
    if nn_index.shape[0] != input.shape[0] or nn_count.shape != nn_index.shape[:2]:
        raise ValueError("nn_index / nn_count shapes are inconsistent with input")
    return input, nn_index, nn_count


def max_pool3d_grad(input, grad_output, max_index):
    input = _lib.cuda_tensor(input, torch.float32, 3, "input")
    grad_output = _lib.cuda_tensor(grad_output, torch.float32, 3, "grad_output")
    max_index = _lib.cuda_tensor(max_index, torch.int32, 3, "max_index")
    B, N, C = input.shape
    M = grad_output.shape[1]
    grad_input = torch.empty((B, N, C), dtype=torch.float32, device=input.device)
    with torch.cuda.device(input.device):
        rc = _lib.lib().sph3d_max_pool3d_grad(B, N, M, C, _lib.ptr(max_index), _lib.ptr(grad_output),
                                              _lib.ptr(grad_input), _lib.stream_ptr())
    _lib.check(rc, "max_pool3d_grad")
    return grad_input


# True: the gradient transposes the graph and gathers (no atomics); False: vector-reduction scatter (csrc/pool3d.cu).
# Measured (profiles/r2_pool_stream.json): avg-pool has a quarter of the edges of an unpool level and the transposition
# costs more than the reductions it saves (0.55 vs 0.49 ms at Cfg-T, 0.17 vs 0.11 ms at S3DIS): scatter stays the default.
GATHER_FORM_GRAD = False


def avg_pool3d_grad(input, grad_output, nn_index, nn_count):
    input, nn_index, nn_count = _check(input, nn_index, nn_count)
    grad_output = _lib.cuda_tensor(grad_output, torch.float32, 3, "grad_output")
    B, N, C = input.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    grad_input = torch.empty((B, N, C), dtype=torch.float32, device=input.device)
    L = _lib.lib()
    ws_bytes = L.sph3d_avg_pool3d_grad_workspace_bytes(B, N, M, C, K) if GATHER_FORM_GRAD else 0
    ws = torch.empty((ws_bytes // 4,), dtype=torch.int32, device=input.device) if ws_bytes else None
    with torch.cuda.device(input.device):
        rc = L.sph3d_avg_pool3d_grad(B, N, M, C, K, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                     _lib.ptr(grad_output), _lib.ptr(grad_input), _lib.ptr(ws), ws_bytes,
                                     _lib.stream_ptr())
    _lib.check(rc, "avg_pool3d_grad")
    return grad_input


class _MaxPool3d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, nn_index, nn_count):
        B, N, C = input.shape
        M, K = nn_index.shape[1], nn_index.shape[2]
        output = torch.empty((B, M, C), dtype=torch.float32, device=input.device)
        max_index = torch.empty((B, M, C), dtype=torch.int32, device=input.device)
        with torch.cuda.device(input.device):
            rc = _lib.lib().sph3d_max_pool3d(B, N, M, C, K, _lib.ptr(nn_index), _lib.ptr(nn_count), _lib.ptr(input),
                                             _lib.ptr(output), _lib.ptr(max_index), _lib.stream_ptr())
        _lib.check(rc, "max_pool3d")
        ctx.save_for_backward(input, max_index)
        ctx.mark_non_differentiable(max_index)
        return output, max_index

    @staticmethod
    def backward(ctx, grad_output, grad_index):
        input, max_index = ctx.saved_tensors
        return max_pool3d_grad(input, grad_output.contiguous(), max_index), None, None


class _AvgPool3d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, nn_index, nn_count):
        B, N, C = input.shape
        M, K = nn_index.shape[1], nn_index.shape[2]
        output = torch.empty((B, M, C), dtype=torch.float32, device=input.device)
        with torch.cuda.device(input.device):
            rc = _lib.lib().sph3d_avg_pool3d(B, N, M, C, K, _lib.ptr(nn_index), _lib.ptr(nn_count), _lib.ptr(input),
                                             _lib.ptr(output), _lib.stream_ptr())
        _lib.check(rc, "avg_pool3d")
        ctx.save_for_backward(input, nn_index, nn_count)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, nn_index, nn_count = ctx.saved_tensors
        return avg_pool3d_grad(input, grad_output.contiguous(), nn_index, nn_count), None, None


def max_pool3d(input, nn_index, nn_count):
    """-> (output (B,M,C), max_index (B,M,C) int32 = database point id of the maximum)."""
    input, nn_index, nn_count = _check(input, nn_index, nn_count)
    return _MaxPool3d.apply(input, nn_index, nn_count)


def avg_pool3d(input, nn_index, nn_count):
    input, nn_index, nn_count = _check(input, nn_index, nn_count)
    return _AvgPool3d.apply(input, nn_index, nn_count)
