"""Drop-in check at the call sites: the reference's model call graphs (models/SPH3D_modelnet.py, SPH3D_s3dis.py,
SPH3D_shapenet.py), written against this repository's sph3gcn_util mirror (sph3d-gcn_b200/models, driven by
profiles/bench_encoder.py), run forward and backward and
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


def test_separable_conv_odd_channels_matches_oracle(pkg, oracle):
    """C=35 (32 + raw xyz): the layer zero-pads channels to 36 to stay on the vector kernels; results and
    gradients must equal the un-padded oracle, and the variables keep the reference's shapes."""
    from common import assert_close, features, make_cloud, saturating_radius
    u = pkg.sph3gcn_util
    u.reset_variables()
    B, N, K, C, r, Cout = 2, 700, 32, 35, 2, 24
    xyz = make_cloud(151, B, N, "cube")
    rad = saturating_radius(N, K)
    idx, cnt, dst = oracle.build_sphere_neighbor(xyz, xyz, rad, None, K)
    filt = oracle.spherical_kernel(xyz, xyz, idx, cnt, dst, rad, [8, 2, 2])
    x = features(152, B, N, C)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
    xt = t(x).requires_grad_(True)
    y = u.separable_conv3d(xt, Cout, 33, r, 'odd', t(idx), t(cnt), t(filt), activation_fn=None)
    v = u.named_variables()
    Wd, Wp = v['odd/depthwise_weights'], v['odd/weights']
    assert tuple(Wd.shape) == (33, C, r) and tuple(Wp.shape) == (C * r, Cout)
    dw = oracle.depthwise_conv3d(x, Wd.detach().cpu().numpy(), idx, cnt, filt, 1)
    want = dw.reshape(-1, C * r).astype(np.float64) @ Wp.detach().cpu().numpy().astype(np.float64)
    assert_close(y.detach().cpu().numpy().reshape(-1, Cout), want, 2e-5, "separable_conv3d C=35")
    go = features(153, B, N, Cout)
    y.backward(t(go))
    g_dw = (go.reshape(-1, Cout).astype(np.float64) @ Wp.detach().cpu().numpy().astype(np.float64).T).reshape(B, N, C * r).astype(np.float32)
    gi, gf = oracle.depthwise_conv3d_grad(x, Wd.detach().cpu().numpy(), g_dw, idx, cnt, filt)
    assert_close(xt.grad.cpu().numpy(), gi, 2e-5, "grad_input through the padded layer")
    assert_close(Wd.grad.cpu().numpy(), gf, 2e-5, "grad depthwise_weights through the padded layer")


def test_s3dis_model_slice_runs_and_trains(pkg):
    """segmentation family (models/SPH3D_s3dis.py:35-133): build_graph_deconv, inter-graph ball queries from the fine
    to the coarse cloud (retries), mean unpooling and skip concats, inner-point loss, forward + backward."""
    import bench_encoder as be
    B, N = 2, 1024
    rec = be.run(B=B, N=N, steps=1, warmup=1, model="s3dis")
    assert rec["levels"] == [256, 96, 48, 16]
    assert rec["all_grads_finite"] and np.isfinite(rec["loss"])
    assert rec["pred_shape"] == [B, N, 13]                                      # per-point logits
    assert rec["feature_dim"] == 128 + 128                                      # unpooled deconv4_2 (+) conv1_2 skip
    names = set(pkg.sph3gcn_util.named_variables())
    for want in ("conv4_2/depthwise_weights", "deconv1_1/depthwise_weights", "deconv4_2/weights", "logits/weights"):
        assert want in names, want
    v = pkg.sph3gcn_util.named_variables()
    assert tuple(v["mlp1/weights"].shape) == (3, 64)                             # normalised xyz only: points[:, :, 6:] is empty
    assert tuple(v["deconv2_1/depthwise_weights"].shape) == (33, 512 + 512, 2)      # unpooled 512 (+) skip 512
    assert all(float(p.grad.abs().sum()) > 0 for n, p in v.items() if n.endswith("depthwise_weights"))


def test_shapenet_model_slice_runs_and_trains(pkg):
    """models/SPH3D_shapenet.py:33-123 at K=32 (BASELINE.json configs[2]): raw 6-column input, mlp2 (+) mlp1 before the logits"""
    import bench_encoder as be
    B, N = 2, 1024
    rec = be.run(B=B, N=N, steps=1, warmup=1, model="shapenet")
    assert rec["levels"] == [512, 384, 192, 64] and rec["K"] == 32
    assert rec["all_grads_finite"] and np.isfinite(rec["loss"])
    assert rec["pred_shape"] == [B, N, 50]
    v = pkg.sph3gcn_util.named_variables()
    assert tuple(v["mlp1/weights"].shape) == (6, 64) and tuple(v["mlp2/weights"].shape) == (256, 64)
    assert tuple(v["logits/weights"].shape) == (64 + 64, 50)


def test_modelnet_model_slice_runs_and_trains(pkg):
    import bench_encoder as be
    rec = be.run(B=2, N=2048, steps=1, warmup=1)
    assert rec["levels"] == [512, 128]
    assert rec["all_grads_finite"] and np.isfinite(rec["loss"])
    u = pkg.sph3gcn_util
    names = set(u.named_variables())
    for want in ("mlp1/weights", "conv1_1/depthwise_weights", "conv1_1/weights", "conv2_2/depthwise_weights",
                 "global_conv/depthwise_weights", "conv1_1/bn/gamma", "fc1/weights", "fc2/bn/beta", "logits/weights"):
        assert want in names, want
    v = u.named_variables()
    assert tuple(v["conv1_1/depthwise_weights"].shape) == (33, 35, 2)         # mlp 32 + raw xyz 3, multiplier 2
    assert tuple(v["conv2_1/depthwise_weights"].shape) == (33, 67, 1)
    assert tuple(v["global_conv/depthwise_weights"].shape) == (17, 128, 2)
    assert all(float(p.grad.abs().sum()) > 0 for n, p in v.items() if n.endswith("depthwise_weights"))
    # feature vector = 64 + 128 (level maxima) + 512 (global conv); classifier 704 -> 512 -> 256 -> 40
    assert rec["feature_dim"] == 64 + 128 + 512 and rec["pred_shape"] == [2, 40]
    assert tuple(v["fc1/weights"].shape) == (704, 512)


def test_graphed_step_equals_eager_step(pkg):
    """utils/graph_step.GraphedStep: the whole S3DIS training step (ball queries, FPS on its side stream, bins,
    convolutions, unpooling, inner-point loss, backward) captured in one CUDA graph; replays on NEW input data must
    reproduce the eager step (float atomics in the gradients: tolerance)."""
    from common import assert_close
    u, M = pkg.sph3gcn_util, pkg.models
    dev = torch.device("cuda", 0)
    B, N = 2, 1024
    cfg = M.configs.s3dis(N)
    g = torch.Generator().manual_seed(21)
    pts = torch.rand(B, N, 6, generator=g).to(dev)
    label = torch.randint(0, 13, (B, N), generator=g).to(dev)
    inner = (torch.rand(B, N, generator=g) < 0.6).int().to(dev)
    u.reset_variables()

    def step():
        u.clear_collections()
        for p in u.trainable_variables():
            p.grad = None
        pred, end = M.SPH3D_s3dis.get_model(pts, True, cfg)
        loss = M.SPH3D_s3dis.get_loss(pred, label, end, inner)
        loss.backward()
        return pred, loss

    gstep = pkg.utils.graph_step.GraphedStep(step, u.trainable_variables, warmup=2)
    for seed in (22, 23):
        pts.copy_(torch.rand(B, N, 6, generator=torch.Generator().manual_seed(seed)))
        pred_g, loss_g = gstep()
        torch.cuda.synchronize()
        got = [pred_g.detach().cpu().numpy().copy(), float(loss_g)] + \
              [p.grad.detach().cpu().numpy().copy() for p in u.trainable_variables()]
        pred_e, loss_e = step()                                   # eager, same data, same weights
        want = [pred_e.detach().cpu().numpy(), float(loss_e)] + [p.grad.cpu().numpy() for p in u.trainable_variables()]
        assert abs(got[1] - want[1]) <= 1e-5 * abs(want[1])
        assert_close(got[0], want[0], 1e-4, "logits, graph replay vs eager")
        names = list(u.named_variables())
        floor = 1e-6 * max(float(np.abs(w).max()) for w in want[2:])
        for nme, a, w in zip(names, got[2:], want[2:]):
            # two runs of the same fp32 network: float-atomic summation order differs, and column sums that cancel
            # (BN beta / gamma gradients of ~1e-7) amplify that noise -- a consistency bound, not a parity bound
            err = np.abs(a.astype(np.float64) - w)
            tol = 2e-2 * float(np.abs(w).max()) + floor
            assert err.max() <= tol, "grad of %s, graph replay vs eager: max err %.3e > %.3e" % (nme, err.max(), tol)
    assert gstep.replays == 2


def test_async_sampling_joins_before_indices_are_read(pkg, oracle):
    """with async_sampling(): build_graph returns while FPS still runs on the side stream; gather_nd must join it.
    The gathered rows are checked against the oracle's FPS for a cloud large enough that FPS outlasts the query."""
    from common import assert_equal, make_cloud
    u = pkg.sph3gcn_util
    B, N, S = 4, 8192, 2048
    xyz_np = make_cloud(171, B, N, "cube")
    xyz = torch.from_numpy(xyz_np).to("cuda:0")
    want = oracle.farthest_point_sample(S, xyz_np)                               # (B, S) ids
    rows = np.take_along_axis(xyz_np, want[:, :, None].astype(np.int64).repeat(3, axis=2), axis=1)
    for _ in range(3):
        with u.async_sampling():
            idx, cnt, dst, indices = u.build_graph(xyz, 0.05, 16, S, sample_method='FPS')
            assert getattr(indices, "_sph3d_ready", None) is not None            # not joined yet
            picked = u.gather_nd(xyz, indices)
            assert indices._sph3d_ready is None
        assert_equal(picked.cpu().numpy(), rows, "xyz rows selected by async FPS")
        assert_equal(indices[..., 1].cpu().numpy(), want, "FPS ids")
    with u.async_sampling():                                                     # never consumed: joined when the block ends
        _, _, _, indices = u.build_graph(xyz, 0.05, 16, S, sample_method='FPS')
    assert not u._PENDING_SAMPLES
    assert_equal(indices[..., 1].cpu().numpy(), want, "FPS ids, joined at block exit")
    _, _, _, indices = u.build_graph(xyz, 0.05, 16, S, sample_method='FPS')      # outside the block: joined on return
    assert getattr(indices, "_sph3d_ready", None) is None
    assert_equal(indices[..., 1].cpu().numpy(), want, "FPS ids, synchronous")


def test_sampling_pyramid_prefetch_changes_nothing_but_the_schedule(pkg, monkeypatch):
    """prefetch_samples runs the whole FPS chain ahead on the side stream; build_graph / gather_nd then hand out ITS picks
    and coarse clouds.  Forward results must be bit-identical to the level-by-level schedule."""
    u, M = pkg.sph3gcn_util, pkg.models
    B, N = 2, 2048
    cfg = M.configs.s3dis(N)
    g = torch.Generator().manual_seed(5)
    pts = torch.rand(B, N, 6, generator=g).to("cuda:0")
    u.reset_variables()
    with torch.no_grad():
        u.clear_collections()
        a, _ = M.SPH3D_s3dis.get_model(pts, False, cfg)
        torch.cuda.synchronize()
        assert not u._PREFETCHED and not u._PENDING_SAMPLES                # everything joined and released
        seen = []
        real = u.prefetch_samples
        monkeypatch.setattr(u, "prefetch_samples", lambda *args, **kw: seen.append(args[1]))
        u.clear_collections()
        b, _ = M.SPH3D_s3dis.get_model(pts, False, cfg)
        torch.cuda.synchronize()
    assert seen == [cfg.num_sample]
    assert torch.equal(a, b)
    # the chain itself: picks per level equal FPS run level by level
    monkeypatch.setattr(u, "prefetch_samples", real)
    xyz = pts[:, :, :3].contiguous()
    with u.async_sampling():
        u.prefetch_samples(xyz, cfg.num_sample, 'FPS')
        level = xyz
        for s in cfg.num_sample:
            _, _, _, ind = u.build_graph(level, 0.2, 16, s, sample_method='FPS')
            want = pkg.tf_sample.farthest_point_sample(s, level)
            coarse = u.gather_nd(level, ind)
            assert torch.equal(ind[..., 1], want)
            assert torch.equal(coarse, level[torch.arange(B, device=level.device)[:, None], want.long()])
            level = coarse
