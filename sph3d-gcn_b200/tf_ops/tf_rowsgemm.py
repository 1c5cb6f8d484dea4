"""The pointwise product over the rows of a layer (csrc/rowsgemm.cu, hand-written tcgen05): the tf.matmul of
/root/reference/utils/sph3gcn_util.py:144-146 (separable_conv3d), :203-205 (pointwise_conv3d), :254-256 (fully_connected)
and its input gradient.  The reference has no op of its own for it (it is a TensorFlow node), so this module has no
counterpart under the reference's tf_ops/; `utils.sph3gcn_util._Dense` routes y = x w and gx = g w^T through it."""
import torch

from .. import _lib


def supported(R, K, N):
    return R > 0 and K > 0 and N > 0 and K % 4 == 0 and N % 4 == 0


def pack(weights, trans=False):
    """weights fp32 -> the three-term bf16 operand image of the product x @ weights (trans=False, weights (K, N)) or
    x @ weights.T (trans=True, weights (N, K)); uint8 tensor"""
    weights = _lib.cuda_tensor(weights, torch.float32, 2, "weights")
    K, N = (weights.shape[1], weights.shape[0]) if trans else (weights.shape[0], weights.shape[1])
    L = _lib.lib()
    image = torch.empty((L.sph3d_rows_gemm_image_bytes(K, N),), dtype=torch.uint8, device=weights.device)
    with torch.cuda.device(weights.device):
        rc = L.sph3d_rows_gemm_pack(K, N, _lib.ptr(weights), 1 if trans else 0, _lib.ptr(image), _lib.stream_ptr())
    _lib.check(rc, "rows_gemm_pack")
    return image


def pack_pair(weights):
    """weights (K, N) fp32 -> (image of x @ weights, image of g @ weights.T) in one launch"""
    weights = _lib.cuda_tensor(weights, torch.float32, 2, "weights")
    K, N = weights.shape
    L = _lib.lib()
    image = torch.empty((L.sph3d_rows_gemm_image_bytes(K, N),), dtype=torch.uint8, device=weights.device)
    image_t = torch.empty((L.sph3d_rows_gemm_image_bytes(N, K),), dtype=torch.uint8, device=weights.device)
    with torch.cuda.device(weights.device):
        rc = L.sph3d_rows_gemm_pack_pair(K, N, _lib.ptr(weights), _lib.ptr(image), _lib.ptr(image_t), _lib.stream_ptr())
    _lib.check(rc, "rows_gemm_pack_pair")
    return image, image_t


# bf16 terms per operand: 3 = six cross products (2^-24 of a product), 2 = four (2^-17); see include/sph3d_b200.h
TERMS = 3


def rows_gemm(x, weights, trans=False, image=None, terms=None):
    """x (R, K) @ weights (K, N)  [trans=False]   or   x (R, K) @ weights (N, K).T  [trans=True]  -> (R, N) fp32.
    No autograd (the layer library wraps it)."""
    x = _lib.cuda_tensor(x, torch.float32, 2, "x")
    weights = _lib.cuda_tensor(weights, torch.float32, 2, "weights")
    K, N = (weights.shape[1], weights.shape[0]) if trans else (weights.shape[0], weights.shape[1])
    R = x.shape[0]
    if x.shape[1] != K:
        raise ValueError("inner dimensions of x and weights differ")
    if not supported(R, K, N) or x.data_ptr() % 16:
        raise ValueError("rows_gemm needs K and N multiples of 4 and 16-byte aligned rows")
    if image is None:
        image = pack(weights, trans)
    y = torch.empty((R, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().sph3d_rows_gemm(R, K, N, int(terms or TERMS), _lib.ptr(x), _lib.ptr(image), _lib.ptr(y), _lib.stream_ptr())
    _lib.check(rc, "rows_gemm")
    return y


def rows_wgrad(x, g, terms=None):
    """x (R, K), g (R, N) -> x.T @ g (K, N) fp32: the weight gradient of the rows product, summed over the rows in a fixed
    order (csrc/rowswgrad.cu).  No autograd."""
    x = _lib.cuda_tensor(x, torch.float32, 2, "x")
    g = _lib.cuda_tensor(g, torch.float32, 2, "g")
    R, K = x.shape
    N = g.shape[1]
    if g.shape[0] != R:
        raise ValueError("x and g must have the same number of rows")
    if not supported(R, K, N) or x.data_ptr() % 16 or g.data_ptr() % 16:
        raise ValueError("rows_wgrad needs K and N multiples of 4 and 16-byte aligned rows")
    L = _lib.lib()
    ws_bytes = L.sph3d_rows_wgrad_workspace_bytes(R, K, N)
    ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=x.device)
    gw = torch.empty((K, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.sph3d_rows_wgrad(R, K, N, int(terms or TERMS), _lib.ptr(x), _lib.ptr(g), _lib.ptr(gw), _lib.ptr(ws), ws_bytes,
                                _lib.stream_ptr())
    _lib.check(rc, "rows_wgrad")
    return gw
