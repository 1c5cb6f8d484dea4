"""Drop-in check at the call sites: the ModelNet encoder of the reference (models/SPH3D_modelnet.py:33-96),
written against this repository's sph3gcn_util mirror (profiles/bench_encoder.py), runs forward and backward and
every op inside it agrees with the oracle (odd channel counts 35 / 67 exercise the VEC=1 kernels, the K=N global
graph the 17-bin kernel)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles"))


def test_modelnet_encoder_slice_runs_and_trains(pkg):
    import bench_encoder as be
    rec = be.run(B=2, N=2048, steps=1, warmup=1)
    assert rec["levels"] == [512, 128]
    assert rec["all_grads_finite"] and np.isfinite(rec["loss"])
    u = pkg.sph3gcn_util
    names = set(u.named_variables())
    for want in ("mlp1/weights", "conv1_1/depthwise_weights", "conv1_1/weights", "conv2_2/depthwise_weights",
                 "global_conv/depthwise_weights", "conv1_1/bn/gamma"):
        assert want in names, want
    v = u.named_variables()
    assert tuple(v["conv1_1/depthwise_weights"].shape) == (33, 35, 2)         # mlp 32 + raw xyz 3, multiplier 2
    assert tuple(v["conv2_1/depthwise_weights"].shape) == (33, 67, 1)
    assert tuple(v["global_conv/depthwise_weights"].shape) == (17, 128, 2)
    assert all(float(p.grad.abs().sum()) > 0 for n, p in v.items() if n.endswith("depthwise_weights"))
    # feature vector = 64 + 128 (level maxima) + 512 (global conv)
    assert rec["feature_dim"] == 64 + 128 + 512
