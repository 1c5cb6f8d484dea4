"""The cell-grid accelerator of the ball query (csrc/nnquery.cu) must be invisible: forced on for every
cloud size (SPH3D_NNQUERY_GRID=2) it has to reproduce the oracle bit for bit on all the adversarial cases of
test_parity_gpu.py -- radius chain with B>32 / M>1024, retries on a sparse database, lattice points exactly on
the radius, queries outside the database's bounding box, K > N."""
import os

import numpy as np
import pytest
import torch

from common import assert_equal, make_cloud, saturating_radius
from test_parity_gpu import SPHERE_CASES

pytestmark = pytest.mark.gpu
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
A = lambda t: t.detach().cpu().numpy()


@pytest.fixture()
def force_grid(pkg):
    old = os.environ.get("SPH3D_NNQUERY_GRID")
    os.environ["SPH3D_NNQUERY_GRID"] = "2"
    pkg._lib.reload_tunables()            # the library reads its tunables once, at load
    yield
    if old is None:
        os.environ.pop("SPH3D_NNQUERY_GRID", None)
    else:
        os.environ["SPH3D_NNQUERY_GRID"] = old
    pkg._lib.reload_tunables()


@pytest.mark.parametrize("case", SPHERE_CASES, ids=[c[0] for c in SPHERE_CASES])
def test_grid_path_matches_oracle(case, pkg, oracle, force_grid):
    name, B, N, M, K, radius, kind = case
    xyz = make_cloud(11, B, N, kind)
    q = xyz if M is None else make_cloud(12, B, M, kind)
    radius = radius or saturating_radius(N, K)
    assert pkg._lib.lib().sph3d_build_sphere_neighbor_workspace_bytes(B, N, q.shape[1], K) > 0 or N < 32
    oi, oc, od = oracle.build_sphere_neighbor(xyz, q, radius, None, K)
    gi, gc, gd = pkg.tf_nnquery.build_sphere_neighbor(T(xyz), T(q), radius=radius, nnsample=K)
    assert_equal(A(gc), oc, name + " nn_count"); assert_equal(A(gi), oi, name + " nn_index"); assert_equal(A(gd), od, name + " nn_dist")


def test_grid_queries_outside_the_box_and_flat_clouds(pkg, oracle, force_grid):
    """decoder-style query set far outside the database box; a planar database (one cell layer)."""
    rng = np.random.default_rng(5)
    db = make_cloud(21, 2, 3000, "cube") * np.float32(0.5)                  # box [0, 0.5]^3
    q = (rng.random((2, 1500, 3), dtype=np.float32) * np.float32(1.4) - np.float32(0.45)).astype(np.float32)
    for r, K in ((0.08, 32), (0.2, 16)):
        oi, oc, od = oracle.build_sphere_neighbor(db, q, r, None, K)
        gi, gc, gd = pkg.tf_nnquery.build_sphere_neighbor(T(db), T(q), radius=r, nnsample=K)
        assert_equal(A(gc), oc); assert_equal(A(gi), oi); assert_equal(A(gd), od)
    flat = make_cloud(22, 1, 4000, "cube"); flat[..., 2] = np.float32(0.25)
    oi, oc, od = oracle.build_sphere_neighbor(flat, flat, 0.05, None, 24)
    gi, gc, gd = pkg.tf_nnquery.build_sphere_neighbor(T(flat), T(flat), radius=0.05, nnsample=24)
    assert_equal(A(gc), oc); assert_equal(A(gi), oi); assert_equal(A(gd), od)


def test_auto_mode_large_cloud(pkg, oracle):
    """default policy: the grid switches itself on for N >= 98304 (beyond every BASELINE shape; profiles/r2_nnquery.json)."""
    N, K, r = 100000, 32, 0.03
    xyz = make_cloud(23, 1, N, "cube")
    assert pkg._lib.lib().sph3d_build_sphere_neighbor_workspace_bytes(1, N, N, K) > 0
    assert pkg._lib.lib().sph3d_build_sphere_neighbor_workspace_bytes(4, 65536, 65536, 64) == 0       # cfg5 scans
    oi, oc, od = oracle.build_sphere_neighbor(xyz, xyz, r, None, K)
    gi, gc, gd = pkg.tf_nnquery.build_sphere_neighbor(T(xyz), T(xyz), radius=r, nnsample=K)
    assert_equal(A(gc), oc); assert_equal(A(gi), oi); assert_equal(A(gd), od)
