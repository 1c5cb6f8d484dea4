"""ctypes binding of lib/libsph3d_b200.so (C ABI: include/sph3d_b200.h).

The reference loads one TensorFlow op library per op at import time
(`tf.load_op_library(... 'tf_conv3d_so.so')`, /root/reference/tf_ops/convolution/tf_conv3d.py:7)
and crashes at import if it is missing.  Same contract here: there is NO CPU or eager fallback.
If the shared library cannot be loaded (and cannot be built because nvcc is absent) every op
raises; tensors that are not on a CUDA device are rejected with ValueError.
"""
import ctypes
import os

import torch

from . import build as _build

_LIB = None

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/sph3d_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "sph3d_abi_version": (c_int, []),
    "sph3d_last_launch_count": (c_int, []),
    "sph3d_reload_tunables": (None, []),
    "sph3d_conv_sort_bytes": (c_size_t, [c_int] * 5),
    "sph3d_conv_sort": (c_int, [c_int] * 5 + [_P] * 4 + [c_size_t, _P]),
    "sph3d_depthwise_conv3d_planned_supported": (c_size_t, [c_int] * 7),
    "sph3d_depthwise_conv3d_planned": (c_int, [c_int] * 7 + [_P, _P, c_size_t] + [_P] * 4),
    "sph3d_build_sphere_neighbor_workspace_bytes": (c_size_t, [c_int] * 4),
    "sph3d_build_sphere_neighbor": (c_int, [c_int] * 4 + [c_float] + [_P] * 6 + [c_size_t, _P]),
    "sph3d_build_cube_neighbor": (c_int, [c_int] * 5 + [c_float] + [_P] * 5),
    "sph3d_spherical_kernel": (c_int, [c_int] * 7 + [c_float] + [_P] * 7),
    "sph3d_depthwise_conv3d": (c_int, [c_int] * 7 + [_P] * 7),
    "sph3d_depthwise_conv3d_grad_workspace_bytes": (c_size_t, [c_int] * 7),
    "sph3d_depthwise_conv3d_grad": (c_int, [c_int] * 7 + [_P] * 9 + [c_size_t, _P]),
    "sph3d_conv_transpose_bytes": (c_size_t, [c_int] * 5),
    "sph3d_conv_transpose": (c_int, [c_int] * 5 + [_P] * 4 + [c_size_t, _P]),
    "sph3d_depthwise_conv3d_grad_planned_workspace_bytes": (c_size_t, [c_int] * 7),
    "sph3d_depthwise_conv3d_grad_planned": (c_int, [c_int] * 7 + [_P, _P, c_size_t] + [_P] * 6 + [c_size_t, _P]),
    "sph3d_farthest_point_sample_workspace_bytes": (c_size_t, [c_int] * 3),
    "sph3d_farthest_point_sample": (c_int, [c_int] * 3 + [_P, _P, c_size_t, _P, _P]),
    "sph3d_max_pool3d": (c_int, [c_int] * 5 + [_P] * 6),
    "sph3d_max_pool3d_grad": (c_int, [c_int] * 4 + [_P] * 4),
    "sph3d_avg_pool3d": (c_int, [c_int] * 5 + [_P] * 5),
    "sph3d_avg_pool3d_grad_workspace_bytes": (c_size_t, [c_int] * 5),
    "sph3d_avg_pool3d_grad": (c_int, [c_int] * 5 + [_P] * 5 + [c_size_t, _P]),
    "sph3d_mean_interpolate": (c_int, [c_int] * 5 + [_P] * 5),
    "sph3d_interpolate_grad_workspace_bytes": (c_size_t, [c_int] * 5),
    "sph3d_mean_interpolate_grad": (c_int, [c_int] * 5 + [_P] * 5 + [c_size_t, _P]),
    "sph3d_weighted_interpolate": (c_int, [c_int] * 5 + [_P] * 6),
    "sph3d_weighted_interpolate_grad": (c_int, [c_int] * 5 + [_P] * 6 + [c_size_t, _P]),
    "sph3d_bias_act_bn_workspace_bytes": (c_size_t, [c_int] * 2),
    "sph3d_bias_act_bn": (c_int, [c_int] * 4 + [c_float] * 2 + [_P] * 10 + [c_size_t, _P]),
    "sph3d_bias_act_bn_grad": (c_int, [c_int] * 4 + [_P] * 11 + [c_size_t, _P]),
    "sph3d_separable_conv3d_supported": (c_int, [c_int] * 8),
    "sph3d_sepconv_weight_image_bytes": (c_size_t, [c_int] * 2),
    "sph3d_sepconv_pack_weights": (c_int, [c_int] * 2 + [_P] * 3),
    "sph3d_separable_conv3d": (c_int, [c_int] * 8 + [_P] * 9 + [c_int] + [_P] * 3),
    "sph3d_rows_gemm_image_bytes": (c_size_t, [c_int] * 2),
    "sph3d_rows_gemm_pack": (c_int, [c_int] * 2 + [_P, c_int, _P, _P]),
    "sph3d_rows_gemm_pack_pair": (c_int, [c_int] * 2 + [_P] * 4),
    "sph3d_rows_gemm": (c_int, [c_int] * 4 + [_P] * 4),
    "sph3d_rows_gemm_trace": (None, [_P]),
    "sph3d_rows_wgrad_workspace_bytes": (c_size_t, [c_int] * 3),
    "sph3d_rows_wgrad": (c_int, [c_int] * 4 + [_P] * 4 + [c_size_t, _P]),
}


def library_path():
    return _build.LIB


def lib():
    """Load (building first if the .so is missing or stale and nvcc is available)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if _build.is_stale():
        try:
            _build.build()
        except Exception as e:  # no nvcc / compile error
            if not os.path.exists(path):
                raise ImportError(
                    "sph3d-gcn_b200: %s is missing and could not be built (%s). There is no CPU "
                    "fallback: build it with `python sph3d-gcn_b200/build.py`." % (path, e))
            if os.environ.get("SPH3D_ALLOW_STALE_LIB") != "1":
                import warnings
                warnings.warn("sph3d-gcn_b200: %s is older than its sources and the rebuild failed (%s); loading the "
                              "stale binary. Set SPH3D_ALLOW_STALE_LIB=1 to silence." % (path, e), RuntimeWarning)
    handle = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)      # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if handle.sph3d_abi_version() != 5:
        raise ImportError("sph3d-gcn_b200: ABI version mismatch in %s" % path)
    _LIB = handle
    return _LIB


def reload_tunables():
    """re-read the SPH3D_* environment variables (the library reads them once, at load)"""
    lib().sph3d_reload_tunables()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return t.data_ptr() if t is not None else None


def cuda_tensor(t, dtype, ndim, name):
    """Validation the TensorFlow glue did with OP_REQUIRES (-> InvalidArgument); here ValueError."""
    if not isinstance(t, torch.Tensor):
        raise ValueError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must live on a CUDA device (sph3d-gcn_b200 has no CPU path)" % name)
    if t.dim() != ndim:
        raise ValueError("The rank of %s should be %d, got shape %s" % (name, ndim, tuple(t.shape)))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def check(code, what):
    if code != 0:
        raise RuntimeError("%s failed: cudaError %d%s" % (
            what, code, " (invalid argument)" if code == 1 else ""))
