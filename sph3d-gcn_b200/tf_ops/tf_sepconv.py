"""The separable layer as one op (csrc/sepconv.cu): depthwise spherical convolution -> pointwise product on the tensor
cores -> bias / ELU / per-channel affine.  The reference has no such op -- it is the node chain of
/root/reference/utils/sph3gcn_util.py:128-161 (tf_conv3d.depthwise_conv3d -> tf.matmul -> tf.nn.bias_add -> tf.nn.elu ->
batch normalisation) -- so this module has no counterpart under the reference's tf_ops/; `utils.sph3gcn_util.separable_conv3d`
routes through it where `supported()` says the fused kernel applies."""
import torch

from .. import _lib
from .tf_conv3d import _check

ACT_NONE, ACT_ELU = 0, 1


def supported(input, filter, nn_index, num_out_channels):
    if not (input.is_cuda and input.dtype == torch.float32 and filter.dtype == torch.float32):
        return False
    B, N, C = input.shape
    F, _, r = filter.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    return bool(_lib.lib().sph3d_separable_conv3d_supported(B, N, M, F, C, r, K, int(num_out_channels)))


def pack_weights(weights):
    """pointwise weights (C*r, Cout) fp32 -> the bf16 three-term operand image the fused kernel streams (uint8 tensor)"""
    weights = _lib.cuda_tensor(weights, torch.float32, 2, "weights")
    Kp, Cout = weights.shape
    L = _lib.lib()
    image = torch.empty((L.sph3d_sepconv_weight_image_bytes(Kp, Cout),), dtype=torch.uint8, device=weights.device)
    with torch.cuda.device(weights.device):
        rc = L.sph3d_sepconv_pack_weights(Kp, Cout, _lib.ptr(weights), _lib.ptr(image), _lib.stream_ptr())
    _lib.check(rc, "sepconv_pack_weights")
    return image


def _vec(t, n, name):
    if t is None:
        return None
    t = _lib.cuda_tensor(t, torch.float32, 1, name)
    if t.shape[0] != n:
        raise ValueError("%s must have one entry per output channel" % name)
    return t


def separable_conv3d(input, filter, weights, nn_index, nn_count, bin_index, bias=None, scale=None, shift=None,
                     act=ACT_NONE, keep_depthwise=False, weight_image=None):
    """-> (output (B, M, Cout), depthwise output (B, M, C*r) or None).
    output = act(depthwise_conv3d(input, filter, graph) @ weights + bias) * scale + shift; no autograd (the layer
    library wraps it)."""
    input, filter, nn_index, nn_count, bin_index = _check(input, filter, nn_index, nn_count, bin_index)
    weights = _lib.cuda_tensor(weights, torch.float32, 2, "weights")
    B, N, C = input.shape
    F, _, r = filter.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    Kp, Cout = weights.shape
    if Kp != C * r:
        raise ValueError("weights must be (in_channels*multiplier, out_channels)")
    L = _lib.lib()
    if not L.sph3d_separable_conv3d_supported(B, N, M, F, C, r, K, Cout):
        raise ValueError("shape not covered by the fused separable kernel (see sph3d_separable_conv3d_supported)")
    bias, scale, shift = _vec(bias, Cout, "bias"), _vec(scale, Cout, "scale"), _vec(shift, Cout, "shift")
    image = weight_image if weight_image is not None else pack_weights(weights)
    out = torch.empty((B, M, Cout), dtype=torch.float32, device=input.device)
    dw = torch.empty((B, M, Kp), dtype=torch.float32, device=input.device) if keep_depthwise else None
    with torch.cuda.device(input.device):
        rc = L.sph3d_separable_conv3d(B, N, M, F, C, r, K, Cout, _lib.ptr(nn_index), _lib.ptr(nn_count),
                                      _lib.ptr(bin_index), _lib.ptr(input), _lib.ptr(filter), _lib.ptr(image),
                                      _lib.ptr(bias), _lib.ptr(scale), _lib.ptr(shift), int(act), _lib.ptr(dw),
                                      _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "separable_conv3d")
    return out, dw
