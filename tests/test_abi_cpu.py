"""The C-ABI library loads on a GPU-less box and exports exactly what include/sph3d_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "sph3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sph3d_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_all_reference_launchers():
    syms = _header_symbols()
    for want in ["sph3d_build_sphere_neighbor", "sph3d_build_cube_neighbor", "sph3d_spherical_kernel",
                 "sph3d_depthwise_conv3d", "sph3d_depthwise_conv3d_grad", "sph3d_farthest_point_sample",
                 "sph3d_max_pool3d", "sph3d_max_pool3d_grad", "sph3d_avg_pool3d", "sph3d_avg_pool3d_grad",
                 "sph3d_mean_interpolate", "sph3d_mean_interpolate_grad", "sph3d_weighted_interpolate",
                 "sph3d_weighted_interpolate_grad"]:      # the 14 host launchers of tf_ops/*/tf_*_gpu.cu
        assert want in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.library_path())
    for s in _header_symbols():
        assert hasattr(lib, s), "missing export " + s
    assert set(pkg._lib.SIGNATURES) == set(_header_symbols())
    assert pkg._lib.lib().sph3d_abi_version() == 5


def test_no_cpu_fallback(pkg):
    import torch
    x = torch.zeros(1, 8, 3)
    with pytest.raises(ValueError, match="CUDA"):
        pkg.tf_nnquery.build_sphere_neighbor(x, x, radius=0.1, nnsample=4)
    with pytest.raises(ValueError, match="CUDA"):
        pkg.tf_sample.farthest_point_sample(4, x)
    with pytest.raises(ValueError, match="CUDA"):
        pkg.tf_conv3d.depthwise_conv3d(x, torch.zeros(3, 3, 1), torch.zeros(1, 8, 2, dtype=torch.int32),
                                       torch.zeros(1, 8, dtype=torch.int32), torch.zeros(1, 8, 2, dtype=torch.int32))


def test_product_never_imports_oracle():
    """oracle/ is a checker: nothing under sph3d-gcn_b200/ may reference it."""
    base = os.path.join(ROOT, "sph3d-gcn_b200")
    for d, _, files in os.walk(base):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "liboracle" not in txt and "import oracle" not in txt and "ref_gpu" not in txt, f


def test_workspace_queries_need_no_gpu(pkg):
    L = pkg._lib.lib()
    assert L.sph3d_farthest_point_sample_workspace_bytes(4, 8192, 2048) == 0          # on-chip plan
    assert L.sph3d_farthest_point_sample_workspace_bytes(2, 200000, 10) == 2 * 200000 * 4
    assert L.sph3d_depthwise_conv3d_grad_workspace_bytes(32, 10000, 10000, 33, 128, 1, 64) > 0
    assert L.sph3d_depthwise_conv3d_grad_workspace_bytes(0, 1, 1, 1, 1, 1, 1) == 0
    # layer products: three bf16 terms, units of 128 output columns x 64 k = 16 KB; one partial block of gw per row slab
    assert L.sph3d_rows_gemm_image_bytes(256, 128) == 4 * 3 * 1 * 16384
    assert L.sph3d_rows_gemm_image_bytes(68, 132) == 2 * 3 * 2 * 16384
    assert L.sph3d_rows_gemm_image_bytes(0, 128) == 0
    ws = L.sph3d_rows_wgrad_workspace_bytes(65536, 256, 128)
    assert ws > 0 and ws % (256 * 128 * 4) == 0
    assert L.sph3d_rows_wgrad_workspace_bytes(100, 4, 4) == 4 * 4 * 4            # a single slab
