"""bias -> activation -> batch normalisation as ONE op (csrc/post.cu, C ABI sph3d_bias_act_bn[_grad]).

The reference ends every layer with three TensorFlow graph nodes over the (B, M, C) matmul result
(/root/reference/utils/sph3gcn_util.py:147-161, :206-220, :257-271: tf.nn.bias_add, activation_fn,
tf.layers.batch_normalization(momentum=0.99, training=is_training) at :328-332).  Executed eagerly that chain is a
dozen full-tensor kernels per direction; here it is two passes per direction and the activation output is never
written to memory.  Semantics kept: ELU *before* BN, biased batch variance, eps 1e-3, moving
statistics moved with decay 0.99 while training and used for normalisation otherwise.
"""
import torch

from .. import _lib

ACT_NONE, ACT_ELU = 0, 1


def _workspace(R, C, device):
    nbytes = _lib.lib().sph3d_bias_act_bn_workspace_bytes(R, C)
    return torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=device), nbytes


class _BiasActBN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, gamma, beta, moving_mean, moving_var, act, training, eps, momentum):
        shape = x.shape
        C = shape[-1]
        x2 = x.reshape(-1, C).contiguous()
        R = x2.shape[0]
        dev = x2.device
        out = torch.empty_like(x2)
        has_bn = gamma is not None
        if has_bn:
            save_mean = torch.empty(C, dtype=torch.float32, device=dev)
            save_invstd = torch.empty(C, dtype=torch.float32, device=dev)
        else:
            save_mean = save_invstd = None
        ws, nbytes = _workspace(R, C, dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().sph3d_bias_act_bn(R, C, act, 1 if training else 0, eps, momentum, _lib.ptr(x2), _lib.ptr(bias),
                                              _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(moving_mean), _lib.ptr(moving_var),
                                              _lib.ptr(out), _lib.ptr(save_mean), _lib.ptr(save_invstd), _lib.ptr(ws),
                                              nbytes, _lib.stream_ptr())
        _lib.check(rc, "bias_act_bn")
        ctx.save_for_backward(x2, bias, gamma, save_mean, save_invstd)
        ctx.act, ctx.training, ctx.shape = act, bool(training), shape
        return out.reshape(shape)

    @staticmethod
    def backward(ctx, grad_out):
        x2, bias, gamma, save_mean, save_invstd = ctx.saved_tensors
        R, C = x2.shape
        dev = x2.device
        g = grad_out.reshape(R, C).contiguous()
        if g.dtype != torch.float32:
            g = g.float()
        grad_x = torch.empty_like(x2)
        grad_bias = torch.empty(C, dtype=torch.float32, device=dev) if bias is not None else None
        if gamma is not None:
            grad_gamma = torch.empty(C, dtype=torch.float32, device=dev)
            grad_beta = torch.empty(C, dtype=torch.float32, device=dev)
        else:
            grad_gamma = grad_beta = None
        ws, nbytes = _workspace(R, C, dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().sph3d_bias_act_bn_grad(R, C, ctx.act, 1 if ctx.training else 0, _lib.ptr(x2), _lib.ptr(bias),
                                                   _lib.ptr(gamma), _lib.ptr(save_mean), _lib.ptr(save_invstd), _lib.ptr(g),
                                                   _lib.ptr(grad_x), _lib.ptr(grad_bias), _lib.ptr(grad_gamma),
                                                   _lib.ptr(grad_beta), _lib.ptr(ws), nbytes, _lib.stream_ptr())
        _lib.check(rc, "bias_act_bn_grad")
        return grad_x.reshape(ctx.shape), grad_bias, grad_gamma, grad_beta, None, None, None, None, None, None


def _vec(t, C, name):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("%s must live on a CUDA device (sph3d-gcn_b200 has no CPU path)" % name)
    if t.dim() != 1 or t.shape[0] != C or t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError("%s must be a contiguous float32 vector of %d channels, got %s %s" % (name, C, tuple(t.shape), t.dtype))
    return t


def bias_act_bn(x, bias=None, gamma=None, beta=None, moving_mean=None, moving_var=None, act=ACT_ELU,
                training=False, eps=1e-3, momentum=0.99):
    """out = BN(act(x + bias)) over the last axis of x (any leading shape).

    bias / (gamma, beta, moving_mean, moving_var) may be None (no bias / no batch normalisation).  moving_mean and
    moving_var are updated in place when `training`.  Differentiable in x, bias, gamma, beta."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise ValueError("x must live on a CUDA device (sph3d-gcn_b200 has no CPU path)")
    if x.dtype != torch.float32 or x.dim() < 1 or x.numel() == 0:
        raise ValueError("x must be a non-empty float32 tensor, got %s %s" % (tuple(x.shape), x.dtype))
    if act not in (ACT_NONE, ACT_ELU):
        raise ValueError("act must be ACT_NONE or ACT_ELU")
    C = x.shape[-1]
    bias = _vec(bias, C, "bias")
    gamma, beta = _vec(gamma, C, "gamma"), _vec(beta, C, "beta")
    moving_mean, moving_var = _vec(moving_mean, C, "moving_mean"), _vec(moving_var, C, "moving_var")
    if (gamma is None) != (beta is None) or (gamma is not None and (moving_mean is None or moving_var is None)):
        raise ValueError("batch normalisation needs gamma, beta, moving_mean and moving_var together")
    return _BiasActBN.apply(x, bias, gamma, beta, moving_mean, moving_var, int(act), bool(training), float(eps),
                            float(momentum))
