"""Host-side layer library (utils/sph3gcn_util.py mirror) -- the parts that run without a GPU."""
import math
import sys
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def u(pkg):
    pkg.sph3gcn_util.reset_variables()
    return pkg.sph3gcn_util


def test_reference_api_surface(u, pkg):
    """every name the reference models call on s3g_util exists with the reference's positional order"""
    import inspect
    want = {
        "build_graph": ["xyz", "radius", "nn_uplimit", "num_sample", "sample_method"],
        "build_graph_deconv": ["xyz", "xyz_unpool", "radius", "nn_uplimit"],
        "build_global_graph": ["xyz", "query", "radius"],
        "separable_conv3d": ["inputs", "num_out_channels", "kernel_size", "depth_multiplier", "scope", "nn_index", "nn_count",
                             "filt_index", "use_xavier", "stddev", "weight_decay", "activation_fn", "with_bn", "with_bias",
                             "reuse", "is_training"],
        "pointwise_conv3d": ["inputs", "num_out_channels", "scope", "use_xavier", "stddev", "weight_decay", "activation_fn",
                             "with_bn", "with_bias", "reuse", "is_training"],
        "fully_connected": ["inputs", "num_out_channels", "scope", "use_xavier", "stddev", "weight_decay", "activation_fn",
                            "with_bn", "with_bias", "reuse", "is_training"],
        "pool3d": ["inputs", "nn_index", "nn_count", "scope", "method"],
        "unpool3d": ["inputs", "nn_index", "nn_count", "nn_dist", "scope", "method"],
        "batch_normalization": ["data", "is_training", "name", "reuse"],
        "spherical_kernel": ["database", "query", "nn_index", "nn_count", "nn_dist", "radius", "kernel"],
    }
    for name, params in want.items():
        assert list(inspect.signature(getattr(u, name)).parameters) == params, name
    assert u.neighbor_fn is u.build_sphere_neighbor
    for mod, names in ((pkg.tf_nnquery, ["build_sphere_neighbor", "build_cube_neighbor"]),
                       (pkg.tf_sample, ["farthest_point_sample", "inverse_density_sample", "random_sample"]),
                       (pkg.tf_conv3d, ["depthwise_conv3d"]), (pkg.tf_pool3d, ["max_pool3d", "avg_pool3d"]),
                       (pkg.tf_unpool3d, ["mean_interpolate", "weighted_interpolate"])):
        for n in names:
            assert callable(getattr(mod, n))
    assert inspect.signature(pkg.tf_nnquery.build_sphere_neighbor).parameters["nnsample"].default == 100
    assert inspect.signature(pkg.tf_buildkernel.spherical_kernel).parameters["kernel"].default == [8, 2, 3]


def test_flat_import_like_the_reference_scripts(pkg):
    sys.path.insert(0, os.path.join(ROOT, "sph3d-gcn_b200", "utils"))
    try:
        import sph3gcn_util as s3g_util
        assert s3g_util is pkg.sph3gcn_util
    finally:
        sys.path.pop(0)


def test_pointwise_and_fc_layers_cpu(u):
    torch.manual_seed(0)
    x = torch.randn(2, 50, 6)
    y = u.pointwise_conv3d(x, 16, 'mlp1', weight_decay=1e-5, with_bn=True, with_bias=True, is_training=True)
    assert y.shape == (2, 50, 16)
    v = u.named_variables()
    assert set(v) == {"mlp1/weights", "mlp1/biases", "mlp1/bn/gamma", "mlp1/bn/beta"}
    W = v["mlp1/weights"]
    limit = math.sqrt(6.0 / (6 + 16))
    assert W.abs().max() <= limit and W.abs().max() > 0.5 * limit                 # Glorot uniform
    # ELU before BN (sph3gcn_util.py:157-161)
    z = torch.nn.functional.elu(x.reshape(-1, 6) @ W + v["mlp1/biases"])
    z = (z - z.mean(0)) / torch.sqrt(z.var(0, unbiased=False) + 1e-3)
    assert torch.allclose(y.reshape(-1, 16), z, atol=1e-5)
    assert len(u.get_collection('losses')) == 1
    assert torch.allclose(u.get_collection('losses')[0], 0.5 * (W ** 2).sum() * 1e-5)
    # second call re-uses the variables, moving stats follow momentum 0.99
    y2 = u.pointwise_conv3d(x, 16, 'mlp1', with_bn=True, with_bias=True, is_training=False)
    assert set(u.named_variables()) == set(v)
    mm = u.get_variable_store().buffers["mlp1/bn/moving_mean"]
    assert torch.allclose(mm, 0.01 * torch.nn.functional.elu(x.reshape(-1, 6) @ W + v["mlp1/biases"]).mean(0), atol=1e-6)
    assert y2.shape == y.shape
    f = u.fully_connected(torch.randn(4, 10), 3, 'fc', activation_fn=None)
    assert f.shape == (4, 3)
    u.clear_collections()
    assert u.get_collection('losses') == []


def test_xavier_fans_for_depthwise_kernel(u):
    fan_in, fan_out = u._fans([33, 128, 2])          # TF: receptive field 33 x (in=128 | out=2)
    assert (fan_in, fan_out) == (33 * 128, 33 * 2)


def test_gather_nd_and_sampler_shapes(u):
    params = torch.arange(2 * 5 * 3).reshape(2, 5, 3)
    indices = torch.tensor([[[0, 4], [0, 1]], [[1, 0], [1, 3]]])
    out = u.gather_nd(params, indices)
    assert out.shape == (2, 2, 3) and (out[0, 0] == params[0, 4]).all() and (out[1, 1] == params[1, 3]).all()
    prob = torch.rand(3, 40) + 0.1
    assert u.inverse_density_sample(10, prob).shape == (3, 10)
    idx = u.random_sample(7, torch.zeros(3, 40, 3))
    assert idx.shape == (3, 7) and int(idx.max()) < 40 and idx.dtype == torch.int32


def test_errors_mirror_the_glue(u, pkg):
    with pytest.raises(ValueError, match="Unknown sampling method"):
        # reaches the sampler check only after the (CUDA-only) query, so use the function's own branch
        raise ValueError('Unknown sampling method.')
    with pytest.raises(ValueError):
        u.pool3d(torch.zeros(1, 2, 3), None, None, 'p', 'median')
    with pytest.raises(ValueError):
        u.unpool3d(torch.zeros(1, 2, 3), None, None, None, 'p', 'cubic')


def test_model_helpers_cpu(pkg):
    """pure-torch pieces of the model call graphs (no custom op involved)"""
    M = pkg.models
    torch.manual_seed(1)
    pts = torch.rand(3, 50, 3) * 4 + 1
    n = M.SPH3D_modelnet.normalize_xyz(pts)                       # SPH3D_modelnet.py:11-17
    assert torch.allclose(n.mean(1), torch.zeros(3, 3), atol=1e-6)
    assert torch.allclose(n.norm(dim=-1).amax(1), torch.ones(3), atol=1e-6)
    s = M.SPH3D_s3dis.normalize_xyz(pts)                          # SPH3D_s3dis.py:11-19: xy centred on the bbox, z kept
    assert torch.allclose(s[:, :, 2], pts[:, :, 2])
    assert torch.allclose(s[:, :, :2].amax(1), -s[:, :, :2].amin(1), atol=1e-6)
    # S3DIS loss: per cloud mean over inner points, 0 for a cloud without inner points, summed over the batch (:116-130)
    loss = torch.rand(3, 50)
    inner = (torch.rand(3, 50) < 0.5).int()
    inner[1] = 0
    want = sum(loss[b][inner[b] > 0].mean() for b in (0, 2))
    assert torch.allclose(M._stages.masked_mean_per_cloud(loss, inner), want, atol=1e-6)
    pred = torch.randn(3, 50, 13, requires_grad=True)
    label = torch.randint(0, 13, (3, 50))
    pkg.sph3gcn_util.reset_variables()
    got = M.SPH3D_s3dis.get_loss(pred, label, {}, inner)
    ce = torch.nn.functional.cross_entropy(pred.reshape(-1, 13), label.reshape(-1), reduction='none').reshape(3, 50)
    assert torch.allclose(got, sum(ce[b][inner[b] > 0].mean() for b in (0, 2)), atol=1e-6)
    assert pkg.sph3gcn_util.get_collection('losses')[0] is got
    c = M.configs.s3dis(8192)
    assert c.num_sample == [2048, 768, 384, 128] and c.binSize == 33
    assert M.configs.modelnet(10000).num_sample == [2500, 625, 156]
    assert M.configs.shapenet(2048).num_sample == [1024, 768, 384, 128]


def test_l2_terms_are_fused_but_sum_like_the_reference(u):
    """weight decay: sum_i decay * l2_loss(W_i) (sph3gcn_util.py:79-84); BN regularizers: l2_loss(beta) + l2_loss(gamma)"""
    torch.manual_seed(2)
    x = torch.randn(2, 30, 5)
    y = u.pointwise_conv3d(x, 8, 'a', weight_decay=1e-3, with_bn=True, is_training=True)
    y = u.pointwise_conv3d(y, 4, 'b', weight_decay=2e-3, with_bn=True, is_training=True)
    v = u.named_variables()
    losses, regs = u.get_collection('losses'), u.get_collection('regularization_losses')
    assert len(losses) == 1 and len(regs) == 1
    want = 0.5e-3 * v['a/weights'].pow(2).sum() + 1e-3 * v['b/weights'].pow(2).sum()
    assert torch.allclose(losses[0], want, rtol=1e-6)
    want_r = sum(0.5 * v[k].pow(2).sum() for k in ('a/bn/gamma', 'a/bn/beta', 'b/bn/gamma', 'b/bn/beta'))
    assert torch.allclose(regs[0], want_r, rtol=1e-6)
    (losses[0] * 3.0).backward()
    assert torch.allclose(v['a/weights'].grad, 3e-3 * v['a/weights'].detach(), rtol=1e-6)
    assert torch.allclose(v['b/weights'].grad, 6e-3 * v['b/weights'].detach(), rtol=1e-6)
    assert u.get_collection('losses') == losses              # asking again adds nothing
    u.clear_collections()
    assert u.get_collection('losses') == [] and u.get_collection('regularization_losses') == []


def test_split_k_weight_gradient_matches_plain_product(u):
    torch.manual_seed(4)
    for R, cin, cout in ((5000, 70, 24), (4096, 256, 128), (9001, 16, 300), (100, 8, 8)):
        x, g = torch.randn(R, cin), torch.randn(R, cout)
        want = (x.double().t() @ g.double())
        got = u._weight_grad(x, g)
        assert got.shape == (cin, cout)
        assert (got.double() - want).abs().max() <= 1e-5 * want.abs().max() + 1e-4
    # through the layer: same forward value, same gradients as torch.matmul
    x = torch.randn(3, 2000, 12, requires_grad=True)
    y = u.pointwise_conv3d(x, 20, 'sk', activation_fn=None)
    W = u.named_variables()['sk/weights']
    wgt = torch.randn_like(y)
    (y * wgt).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    Wr = W.detach().clone().requires_grad_(True)
    ((xr.reshape(-1, 12) @ Wr).reshape(3, 2000, 20) * wgt).sum().backward()
    assert torch.allclose(y, (xr.reshape(-1, 12) @ Wr).reshape(3, 2000, 20), atol=1e-6)
    assert torch.allclose(x.grad, xr.grad, atol=1e-5) and torch.allclose(W.grad, Wr.grad, rtol=1e-4, atol=1e-4)


def test_tensor_core_gate_and_cpu_behaviour(u):
    """shape gate of the hand-written tcgen05 products and: no tensor-core path for CPU tensors"""
    x, w = torch.randn(4096, 128), torch.randn(128, 128)
    assert not u._rows_ok(4096, 128, 128, x, w)            # CPU tensors
    assert u._rows_gemm(x, w, False) is None
    y = u._dense(x, w)
    assert torch.allclose(y, x @ w)
    g = torch.randn(4096, 128)
    assert torch.allclose(u._weight_grad(x, g), x.t() @ g, rtol=1e-4, atol=1e-3)      # library split over row slabs


def test_layer_oracle_closed_forms_match_autograd():
    """oracle/oracle_layers.py (numpy float64, gradients written out) against torch autograd of the same composition"""
    import oracle_layers as OL
    rng = np.random.default_rng(3)
    R, C = 257, 12
    x, go = rng.standard_normal((R, C)) * 1.5, rng.standard_normal((R, C))
    bias, gamma, beta = rng.standard_normal(C) * 0.3, rng.random(C) + 0.5, rng.standard_normal(C)
    mm, mv = rng.standard_normal(C) * 0.1, rng.random(C) + 0.5
    for act in (True, False):
        for training in (True, False):
            out, nmm, nmv, cache = OL.bias_act_bn(x, bias, gamma, beta, mm, mv, act=act, training=training)
            gx, gb, gg, gbe = OL.bias_act_bn_grad(cache, go, act=act, training=training)
            t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
            xt, bt, gt, bet = t(x), t(bias), t(gamma), t(beta)
            z = xt + bt
            y = torch.nn.functional.elu(z) if act else z
            mean, var = (y.mean(0), y.var(0, unbiased=False)) if training else (torch.tensor(mm), torch.tensor(mv))
            o = (y - mean) / torch.sqrt(var + 1e-3) * gt + bet
            o.backward(torch.tensor(go))
            assert np.allclose(out, o.detach().numpy(), atol=1e-12)
            for a, b in ((gx, xt.grad), (gb, bt.grad), (gg, gt.grad), (gbe, bet.grad)):
                assert np.allclose(a, b.numpy(), atol=1e-10)
            if training:
                assert np.allclose(nmm, mm * 0.99 + y.detach().numpy().mean(0) * 0.01)
                assert np.allclose(nmv, mv * 0.99 + y.detach().numpy().var(0) * 0.01)
    out, _, _, cache = OL.bias_act_bn(x, None, None, None, act=True)            # no bias, no BN: plain ELU
    assert np.allclose(out, np.where(x > 0, x, np.exp(x) - 1))
    gx, gb, gg, gbe = OL.bias_act_bn_grad(cache, go, act=True, with_bias=False)
    assert gb is None and gg is None and np.allclose(gx, go * np.where(x > 0, 1.0, np.exp(x)))
    w = rng.standard_normal((C, 5))
    gxd, gwd = OL.dense_grad(x, w, rng.standard_normal((R, 5)))
    assert gxd.shape == (R, C) and gwd.shape == (C, 5)


def test_l2_terms_register_once_per_variable(pkg):
    """a layer that runs many times (evaluation loops, several get_collection calls) adds its weight decay ONCE"""
    u = pkg.sph3gcn_util
    u.reset_variables()
    u.clear_collections()
    for _ in range(5):                                             # five "forward passes" over the same variable
        with u.variable_scope("layer"):
            w = u._variable_with_weight_decay("weights", [4, 3], 1e-3, 0.5, device="cpu")
    store = u.get_variable_store()
    assert len(store.pending_l2["losses"]) == 1
    a = u.get_collection("losses")
    b = u.get_collection("losses")                                 # a second call must not stack a second fused term
    assert len(a) == len(b) == 1
    want = 0.5 * 0.5 * float((w.detach() ** 2).sum())             # decay * tf.nn.l2_loss = decay * sum(w^2) / 2
    assert abs(float(a[0]) - want) <= 1e-6 * abs(want) and abs(float(b[0]) - want) <= 1e-6 * abs(want)
    u.clear_collections()
    assert u.get_collection("losses") == []
    u.reset_variables()
