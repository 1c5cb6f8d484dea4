"""The UNMODIFIED reference CUDA kernels as a checker (GPU box only).

TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_golden.py, bench.py's reference-kernel
timing).  oracle/_ref/libsph3d_ref.so is compiled by oracle/Makefile straight from
/root/reference/tf_ops/*/tf_*_gpu.cu; its 14 host launchers have C++ linkage, bound here by
mangled name.  Each wrapper reproduces what the TensorFlow glue (tf_ops/*/tf_*.cpp) does around
the launcher: zero-fill every output, FPS (32,n) temp buffer, dimension extraction (incl. the
N/M swap of the unpooling ops, tf_unpool3d.cpp:76-80), slicing xyz to 3 columns.  The launchers
use the legacy default stream, so we synchronise the device before and after.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsph3d_ref.so")
_LIB = None

_NAMES = {
    "conv": "_Z23depthwiseConv3dLauncheriiiiiiPKiS0_S0_PKfS2_Pf",
    "conv_grad": "_Z27depthwiseConv3dGradLauncheriiiiiiiPKiS0_S0_PKfS2_S2_PfS3_",
    "sphere": "_Z27buildSphereNeighborLauncheriiiifPKfS0_PiS1_Pf",
    "cube": "_Z25buildCubeNeighborLauncheriiiiifPKfS0_PiS1_",
    "kernel": "_Z23sphericalKernelLauncheriiiiiiifPKfS0_PKiS2_S0_Pi",
    "fps": "_Z27farthestPointSampleLauncheriiiPKfPfPi",
    "maxpool": "_Z17maxPool3dLauncheriiiiiPKiS0_PKfPfPi",
    "maxpool_grad": "_Z21maxPool3dGradLauncheriiiiPKiPKfPf",
    "avgpool": "_Z17avgPool3dLauncheriiiiiPKiS0_PKfPf",
    "avgpool_grad": "_Z21avgPool3dGradLauncheriiiiiPKiS0_PKfPf",
    "mean": "_Z23meanInterpolateLauncheriiiiiPKiS0_PKfPf",
    "mean_grad": "_Z27meanInterpolateGradLauncheriiiiiPKiS0_PKfPf",
    "weighted": "_Z27weightedInterpolateLauncheriiiiiPKiS0_PKfS2_Pf",
    "weighted_grad": "_Z31weightedInterpolateGradLauncheriiiiiPKiS0_PKfS2_Pf",
}
_I, _F, _P = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
_ARGS = {
    "conv": [_I] * 6 + [_P] * 6, "conv_grad": [_I] * 7 + [_P] * 8,
    "sphere": [_I] * 4 + [_F] + [_P] * 5, "cube": [_I] * 5 + [_F] + [_P] * 4,
    "kernel": [_I] * 7 + [_F] + [_P] * 6, "fps": [_I] * 3 + [_P] * 3,
    "maxpool": [_I] * 5 + [_P] * 5, "maxpool_grad": [_I] * 4 + [_P] * 3,
    "avgpool": [_I] * 5 + [_P] * 4, "avgpool_grad": [_I] * 5 + [_P] * 4,
    "mean": [_I] * 5 + [_P] * 4, "mean_grad": [_I] * 5 + [_P] * 4,
    "weighted": [_I] * 5 + [_P] * 5, "weighted_grad": [_I] * 5 + [_P] * 5,
}


def available():
    return os.path.exists(_SO) and torch.cuda.is_available()


def _fn(key):
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(_SO)
    f = getattr(_LIB, _NAMES[key])
    f.restype = None
    f.argtypes = _ARGS[key]
    return f


def _call(key, *args):
    torch.cuda.synchronize()
    _fn(key)(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args])
    torch.cuda.synchronize()


def launch_raw(key, *args):
    """Enqueue a launcher on the legacy default stream WITHOUT synchronising (for timing loops)."""
    _fn(key)(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args])


def _f(t):
    return t.to(torch.float32).contiguous()


def _i(t):
    return t.to(torch.int32).contiguous()


def _z(shape, dtype, like):
    return torch.zeros(shape, dtype=dtype, device=like.device)


def build_sphere_neighbor(database, query, radius=0.1, dilation_rate=None, nnsample=100):
    database, query = _f(database[:, :, 0:3]), _f(query[:, :, 0:3])
    if dilation_rate is not None:
        radius = dilation_rate * radius
    B, N, _ = database.shape
    M, K = query.shape[1], int(nnsample)
    idx, cnt, dst = _z((B, M, K), torch.int32, database), _z((B, M), torch.int32, database), _z((B, M, K), torch.float32, database)
    _call("sphere", B, N, M, K, float(radius), database, query, idx, cnt, dst)
    return idx, cnt, dst


def build_cube_neighbor(database, query, length=0.1, dilation_rate=None, nnsample=100, gridsize=3):
    database, query = _f(database[:, :, 0:3]), _f(query[:, :, 0:3])
    if dilation_rate is not None:
        length = dilation_rate * length
    B, N, _ = database.shape
    M, K = query.shape[1], int(nnsample)
    idx, cnt = _z((B, M, K, 2), torch.int32, database), _z((B, M), torch.int32, database)
    _call("cube", B, N, M, int(gridsize), K, float(length), database, query, idx, cnt)
    return idx, cnt


def spherical_kernel(database, query, nn_index, nn_count, nn_dist, radius, kernel=[8, 2, 3]):
    n, p, q = kernel
    database, query = _f(database[:, :, 0:3]), _f(query[:, :, 0:3])
    nn_index, nn_count, nn_dist = _i(nn_index), _i(nn_count), _f(nn_dist)
    B, N, _ = database.shape
    M, K = query.shape[1], nn_index.shape[2]
    filt = _z((B, M, K), torch.int32, database)
    _call("kernel", B, N, M, K, int(n), int(p), int(q), float(radius), database, query, nn_index, nn_count, nn_dist, filt)
    return filt


def depthwise_conv3d(input, filter, nn_index, nn_count, bin_index):
    input, filter = _f(input), _f(filter)
    nn_index, nn_count, bin_index = _i(nn_index), _i(nn_count), _i(bin_index)
    B, N, C = input.shape
    r = filter.shape[2]
    M, K = nn_index.shape[1], nn_index.shape[2]
    out = _z((B, M, C * r), torch.float32, input)
    _call("conv", B, N, M, C, r, K, nn_index, nn_count, bin_index, input, filter, out)
    return out


def depthwise_conv3d_grad(input, filter, grad_output, nn_index, nn_count, bin_index):
    input, filter, grad_output = _f(input), _f(filter), _f(grad_output)
    nn_index, nn_count, bin_index = _i(nn_index), _i(nn_count), _i(bin_index)
    B, N, C = input.shape
    F, _, r = filter.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    gi, gf = _z((B, N, C), torch.float32, input), _z((F, C, r), torch.float32, input)
    _call("conv_grad", B, N, M, F, C, r, K, nn_index, nn_count, bin_index, input, filter, grad_output, gi, gf)
    return gi, gf


def farthest_point_sample(neursize, database):
    database = _f(database[:, :, 0:3])
    B, N, _ = database.shape
    out = _z((B, int(neursize)), torch.int32, database)
    temp = _z((32, N), torch.float32, database)
    _call("fps", B, N, int(neursize), database, temp, out)
    return out


def max_pool3d(input, nn_index, nn_count):
    input, nn_index, nn_count = _f(input), _i(nn_index), _i(nn_count)
    B, N, C = input.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    out, mi = _z((B, M, C), torch.float32, input), _z((B, M, C), torch.int32, input)
    _call("maxpool", B, N, M, C, K, nn_index, nn_count, input, out, mi)
    return out, mi


def max_pool3d_grad(input, grad_output, max_index):
    input, grad_output, max_index = _f(input), _f(grad_output), _i(max_index)
    B, N, C = input.shape
    M = grad_output.shape[1]
    gi = _z((B, N, C), torch.float32, input)
    _call("maxpool_grad", B, N, M, C, max_index, grad_output, gi)
    return gi


def avg_pool3d(input, nn_index, nn_count):
    input, nn_index, nn_count = _f(input), _i(nn_index), _i(nn_count)
    B, N, C = input.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    out = _z((B, M, C), torch.float32, input)
    _call("avgpool", B, N, M, C, K, nn_index, nn_count, input, out)
    return out


def avg_pool3d_grad(input, grad_output, nn_index, nn_count):
    input, grad_output, nn_index, nn_count = _f(input), _f(grad_output), _i(nn_index), _i(nn_count)
    B, N, C = input.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    gi = _z((B, N, C), torch.float32, input)
    _call("avgpool_grad", B, N, M, C, K, nn_index, nn_count, grad_output, gi)
    return gi


def mean_interpolate(input, nn_index, nn_count):
    input, nn_index, nn_count = _f(input), _i(nn_index), _i(nn_count)
    B, M, C = input.shape
    N, K = nn_index.shape[1], nn_index.shape[2]
    out = _z((B, N, C), torch.float32, input)
    _call("mean", B, N, M, C, K, nn_index, nn_count, input, out)
    return out


def mean_interpolate_grad(input, grad_output, nn_index, nn_count):
    input, grad_output, nn_index, nn_count = _f(input), _f(grad_output), _i(nn_index), _i(nn_count)
    B, M, C = input.shape
    N, K = nn_index.shape[1], nn_index.shape[2]
    gi = _z((B, M, C), torch.float32, input)
    _call("mean_grad", B, N, M, C, K, nn_index, nn_count, grad_output, gi)
    return gi


def weighted_interpolate(input, weight, nn_index, nn_count):
    input, weight, nn_index, nn_count = _f(input), _f(weight), _i(nn_index), _i(nn_count)
    B, M, C = input.shape
    N, K = nn_index.shape[1], nn_index.shape[2]
    out = _z((B, N, C), torch.float32, input)
    _call("weighted", B, N, M, C, K, nn_index, nn_count, input, weight, out)
    return out


def weighted_interpolate_grad(input, grad_output, weight, nn_index, nn_count):
    input, grad_output, weight = _f(input), _f(grad_output), _f(weight)
    nn_index, nn_count = _i(nn_index), _i(nn_count)
    B, M, C = input.shape
    N, K = nn_index.shape[1], nn_index.shape[2]
    gi = _z((B, M, C), torch.float32, input)
    _call("weighted_grad", B, N, M, C, K, nn_index, nn_count, grad_output, weight, gi)
    return gi
