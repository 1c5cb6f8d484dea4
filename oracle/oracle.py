"""numpy front-end of the CPU oracle (oracle/sph3d_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under sph3d-gcn_b200/ imports this module.

Function names, argument order and return arity follow the reference's Python op wrappers
(/root/reference/tf_ops/*/tf_*.py), with numpy arrays in place of TF tensors.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = ctypes.POINTER(ctypes.c_int)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile liboracle.so (and, when /root/reference is present, oracle/_ref)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sph3d_oracle.c")
    stale = (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []),
                       check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE, "ref"], check=False, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_atan2f.restype = ctypes.c_float
        _LIB.oracle_atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pf(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(_i32p)


def atan2f(y, x):
    return float(lib().oracle_atan2f(ctypes.c_float(y), ctypes.c_float(x)))


# ---- a1 / a2 : tf_nnquery.py:9-60 -----------------------------------------------------------
def build_sphere_neighbor(database, query, radius=0.1, dilation_rate=None, nnsample=100):
    database = _f(np.asarray(database)[:, :, 0:3])
    query = _f(np.asarray(query)[:, :, 0:3])
    if dilation_rate is not None:
        radius = dilation_rate * radius
    B, N, _ = database.shape
    M = query.shape[1]
    idx = np.empty((B, M, nnsample), np.int32)
    cnt = np.empty((B, M), np.int32)
    dst = np.empty((B, M, nnsample), np.float32)
    lib().oracle_build_sphere_neighbor(B, N, M, int(nnsample), ctypes.c_float(radius),
                                       _pf(database), _pf(query), _pi(idx), _pi(cnt), _pf(dst))
    return idx, cnt, dst


def build_cube_neighbor(database, query, length=0.1, dilation_rate=None, nnsample=100, gridsize=3):
    database = _f(np.asarray(database)[:, :, 0:3])
    query = _f(np.asarray(query)[:, :, 0:3])
    if dilation_rate is not None:
        length = dilation_rate * length
    B, N, _ = database.shape
    M = query.shape[1]
    idx = np.empty((B, M, nnsample, 2), np.int32)
    cnt = np.empty((B, M), np.int32)
    lib().oracle_build_cube_neighbor(B, N, M, int(nnsample), ctypes.c_float(length), int(gridsize),
                                     _pf(database), _pf(query), _pi(idx), _pi(cnt))
    return idx, cnt


# ---- a3 : tf_buildkernel.py:10-34 -----------------------------------------------------------
def spherical_kernel(database, query, nn_index, nn_count, nn_dist, radius, kernel=[8, 2, 3]):
    n, p, q = kernel
    database = _f(np.asarray(database)[:, :, 0:3])
    query = _f(np.asarray(query)[:, :, 0:3])
    nn_index, nn_count, nn_dist = _i(nn_index), _i(nn_count), _f(nn_dist)
    B, N, _ = database.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    filt = np.empty((B, M, K), np.int32)
    lib().oracle_spherical_kernel(B, N, M, K, int(n), int(p), int(q), ctypes.c_float(radius),
                                  _pf(database), _pf(query), _pi(nn_index), _pi(nn_count),
                                  _pf(nn_dist), _pi(filt))
    return filt


# ---- a4 / a5 : tf_conv3d.py:10-32 -----------------------------------------------------------
def depthwise_conv3d(input, filter, nn_index, nn_count, bin_index, mode=1):
    """mode 0 = fp32 in the reference's evaluation order, mode 1 = fp64 accumulate (truth)."""
    input, filter = _f(input), _f(filter)
    nn_index, nn_count, bin_index = _i(nn_index), _i(nn_count), _i(bin_index)
    B, N, C = input.shape
    F, C2, r = filter.shape
    assert C2 == C
    M, K = nn_index.shape[1], nn_index.shape[2]
    out = np.empty((B, M, C * r), np.float32)
    lib().oracle_depthwise_conv3d(B, N, M, C, r, K, int(mode), _pi(nn_index), _pi(nn_count),
                                  _pi(bin_index), _pf(input), _pf(filter), _pf(out))
    return out


def depthwise_conv3d_grad(input, filter, grad_output, nn_index, nn_count, bin_index):
    input, filter, grad_output = _f(input), _f(filter), _f(grad_output)
    nn_index, nn_count, bin_index = _i(nn_index), _i(nn_count), _i(bin_index)
    B, N, C = input.shape
    F, _, r = filter.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    gi = np.empty((B, N, C), np.float32)
    gf = np.empty((F, C, r), np.float32)
    lib().oracle_depthwise_conv3d_grad(B, N, M, F, C, r, K, _pi(nn_index), _pi(nn_count),
                                       _pi(bin_index), _pf(input), _pf(filter), _pf(grad_output),
                                       _pf(gi), _pf(gf))
    return gi, gf


# ---- a6 : tf_sample.py:15-24 ----------------------------------------------------------------
def farthest_point_sample(neursize, database):
    database = _f(np.asarray(database)[:, :, 0:3])
    B, N, _ = database.shape
    out = np.zeros((B, neursize), np.int32)
    lib().oracle_farthest_point_sample(B, N, int(neursize), _pf(database), _pi(out))
    return out


# ---- a8 / a9 : tf_pool3d.py -----------------------------------------------------------------
def max_pool3d(input, nn_index, nn_count):
    input, nn_index, nn_count = _f(input), _i(nn_index), _i(nn_count)
    B, N, C = input.shape
    M, K = nn_index.shape[1], nn_index.shape[2]
    out = np.empty((B, M, C), np.float32)
    mi = np.empty((B, M, C), np.int32)
    lib().oracle_max_pool3d(B, N, M, C, K, _pi(nn_index), _pi(nn_count), _pf(input), _pf(out), _pi(mi))
    return out, mi


def max_pool3d_grad(input, grad_output, max_index):
    input, grad_output, max_index = _f(input), _f(grad_output), _i(max_index)
    B, N, C = input.shape
    M = grad_output.shape[1]
    gi = np.empty((B, N, C), np.float32)
    lib().oracle_max_pool3d_grad(B, N, M, C, _pi(max_index), _pf(grad_output), _pf(gi))
    return gi


def _gather_reduce(input, nn_index, nn_count, weight, mode):
    input, nn_index, nn_count = _f(input), _i(nn_index), _i(nn_count)
    weight = _f(weight) if weight is not None else None
    B, src, C = input.shape
    rows, K = nn_index.shape[1], nn_index.shape[2]
    out = np.empty((B, rows, C), np.float32)
    lib().oracle_gather_reduce(B, src, rows, C, K, int(mode), _pi(nn_index), _pi(nn_count),
                               _pf(weight), _pf(input), _pf(out))
    return out


def _scatter_grad(input, grad_output, nn_index, nn_count, weight):
    input, grad_output = _f(input), _f(grad_output)
    nn_index, nn_count = _i(nn_index), _i(nn_count)
    weight = _f(weight) if weight is not None else None
    B, src, C = input.shape
    rows, K = nn_index.shape[1], nn_index.shape[2]
    gi = np.empty((B, src, C), np.float32)
    lib().oracle_scatter_grad(B, src, rows, C, K, _pi(nn_index), _pi(nn_count), _pf(weight),
                              _pf(grad_output), _pf(gi))
    return gi


def avg_pool3d(input, nn_index, nn_count, mode=1):
    return _gather_reduce(input, nn_index, nn_count, None, mode)


def avg_pool3d_grad(input, grad_output, nn_index, nn_count):
    return _scatter_grad(input, grad_output, nn_index, nn_count, None)


# ---- a10 / a11 : tf_unpool3d.py -------------------------------------------------------------
def mean_interpolate(input, nn_index, nn_count, mode=1):
    return _gather_reduce(input, nn_index, nn_count, None, mode)


def mean_interpolate_grad(input, grad_output, nn_index, nn_count):
    return _scatter_grad(input, grad_output, nn_index, nn_count, None)


def weighted_interpolate(input, weight, nn_index, nn_count, mode=1):
    return _gather_reduce(input, nn_index, nn_count, weight, mode)


def weighted_interpolate_grad(input, grad_output, weight, nn_index, nn_count):
    return _scatter_grad(input, grad_output, nn_index, nn_count, weight)
