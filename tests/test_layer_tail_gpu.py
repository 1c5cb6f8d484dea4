"""Layer tail (bias -> ELU -> batch normalisation, csrc/post.cu) against a float64 composition of the same nodes
(/root/reference/utils/sph3gcn_util.py:147-161 + :328-332 semantics: ELU before BN, biased variance, eps 1e-3,
momentum 0.99).  fp32 tolerance 1e-5 of the tensor's scale (north_star), written in assert_close calls below."""
import itertools

import numpy as np
import pytest
import torch

from common import assert_close

pytestmark = pytest.mark.gpu
EPS, MOM = 1e-3, 0.99


def _reference(x, bias, gamma, beta, mm, mv, act, training):
    """float64 autograd composition -> (out, new moving mean, new moving var)"""
    z = x if bias is None else x + bias
    y = torch.nn.functional.elu(z) if act else z
    if gamma is None:
        return y, None, None
    if training:
        mean = y.mean(0)
        var = y.var(0, unbiased=False)
        nmm = mm * MOM + mean.detach() * (1 - MOM)
        nmv = mv * MOM + var.detach() * (1 - MOM)
    else:
        mean, var, nmm, nmv = mm, mv, mm, mv
    return (y - mean) / torch.sqrt(var + EPS) * gamma + beta, nmm, nmv


def _run_case(pkg, R, C, with_bias, with_bn, act, training, seed, shift=0.0, lead=None):
    lt = pkg.utils.layer_tail
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, C, generator=g) * 1.7 + shift
    go = torch.randn(R, C, generator=g)
    bias = torch.randn(C, generator=g) * 0.5 if with_bias else None
    gamma = torch.rand(C, generator=g) + 0.5 if with_bn else None
    beta = torch.randn(C, generator=g) if with_bn else None
    mm = torch.randn(C, generator=g) * 0.1 if with_bn else None
    mv = torch.rand(C, generator=g) + 0.5 if with_bn else None
    dev = "cuda:0"
    d = lambda t, rg=False: None if t is None else t.to(dev).clone().requires_grad_(rg)
    xd, bd, gd, bed, mmd, mvd = d(x, True), d(bias, True), d(gamma, True), d(beta, True), d(mm), d(mv)
    xin = xd if lead is None else xd.reshape(*lead, C)
    out = lt.bias_act_bn(xin, bd, gd, bed, mmd, mvd, act=lt.ACT_ELU if act else lt.ACT_NONE, training=training,
                         eps=EPS, momentum=MOM)
    assert out.shape == xin.shape
    out.reshape(R, C).backward(go.to(dev))
    f = lambda t, rg=False: None if t is None else t.double().clone().requires_grad_(rg)
    xr, br, gr, ber = f(x, True), f(bias, True), f(gamma, True), f(beta, True)
    want, nmm, nmv = _reference(xr, br, gr, ber, f(mm), f(mv), act, training)
    want.backward(go.double())
    tag = "R=%d C=%d bias=%d bn=%d act=%d train=%d" % (R, C, with_bias, with_bn, act, training)
    assert_close(out.detach().cpu().numpy().reshape(R, C), want.detach().numpy(), 1e-5, "out " + tag)
    assert_close(xd.grad.cpu().numpy(), xr.grad.numpy(), 1e-5, "grad_x " + tag)
    if with_bias:
        # a column sum of grad_x: the tolerance is relative to the size of the summands (under training-mode BN
        # without activation the exact sum is 0 and only rounding noise is left)
        summands = float(xr.grad.abs().sum(0).max())
        err = np.abs(bd.grad.cpu().numpy().astype(np.float64) - br.grad.numpy())
        assert err.max() <= 1e-5 * summands, "grad_bias %s: max err %.3e (summands %.3e)" % (tag, err.max(), summands)
    if with_bn:
        assert_close(gd.grad.cpu().numpy(), gr.grad.numpy(), 1e-5, "grad_gamma " + tag)
        assert_close(bed.grad.cpu().numpy(), ber.grad.numpy(), 1e-5, "grad_beta " + tag)
        assert_close(mmd.cpu().numpy(), nmm.numpy(), 1e-5, "moving_mean " + tag)
        assert_close(mvd.cpu().numpy(), nmv.numpy(), 1e-5, "moving_var " + tag)


@pytest.mark.parametrize("C", [13, 32, 36, 64, 128, 131, 1024])
def test_layer_tail_channel_counts(pkg, C):
    """every strip width: 13/131 scalar, 36 and 64 two channels per lane, 128/1024 16-byte strips, 8 chunks at 1024"""
    _run_case(pkg, 3001, C, True, True, True, True, seed=C)


@pytest.mark.parametrize("with_bias,with_bn,act,training",
                         [c for c in itertools.product([False, True], repeat=4) if c[0] or c[1] or c[2]])
def test_layer_tail_flag_combinations(pkg, with_bias, with_bn, act, training):
    _run_case(pkg, 777, 96, with_bias, with_bn, act, training, seed=17)


@pytest.mark.parametrize("R", [1, 5, 8, 9, 4736, 50000])
def test_layer_tail_row_counts(pkg, R):
    """fewer rows than warps in a CTA, exact multiples, more rows than row-walkers"""
    _run_case(pkg, R, 64, True, True, True, True, seed=R)


def test_layer_tail_large_mean_is_stable(pkg):
    """mean >> std: E[y^2]-mean^2 in fp32 would lose the variance; the shifted / Chan form must not"""
    _run_case(pkg, 20000, 128, False, True, False, True, seed=5, shift=300.0)


def test_layer_tail_leading_shape_and_determinism(pkg):
    _run_case(pkg, 6 * 500, 128, True, True, True, True, seed=9, lead=(6, 500))
    lt = pkg.utils.layer_tail
    torch.manual_seed(3)
    x = torch.randn(8, 4096, 128, device="cuda:0")
    gam, bet = torch.ones(128, device="cuda:0"), torch.zeros(128, device="cuda:0")
    outs = []
    for _ in range(2):
        mm, mv = torch.zeros(128, device="cuda:0"), torch.ones(128, device="cuda:0")
        xg = x.clone().requires_grad_(True)
        gg = gam.clone().requires_grad_(True)
        o = lt.bias_act_bn(xg, None, gg, bet, mm, mv, training=True)
        o.square().sum().backward()
        outs.append((o.detach().clone(), xg.grad.clone(), gg.grad.clone(), mm.clone(), mv.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)                         # fixed-order folds: bit-reproducible


def test_layer_library_fused_tail_equals_composition(pkg):
    """pointwise_conv3d with FUSED_TAIL on and off: same outputs, gradients and moving statistics"""
    u = pkg.sph3gcn_util
    torch.manual_seed(11)
    x = torch.randn(4, 900, 48, device="cuda:0")
    wgt = torch.randn(4, 900, 64, device="cuda:0")
    res = []
    for fused in (True, False):
        u.reset_variables()
        u.FUSED_TAIL = fused
        try:
            torch.manual_seed(12)
            xg = x.clone().requires_grad_(True)
            y = u.pointwise_conv3d(xg, 64, 'pw', with_bn=True, with_bias=True, is_training=True)
            with torch.no_grad():
                u.named_variables()['pw/biases'].add_(0.3)
            y = u.pointwise_conv3d(xg, 64, 'pw', with_bn=True, with_bias=True, is_training=True)
            (y * wgt).sum().backward()                  # per-element weights: a per-channel weight has zero gradient through BN
            v = u.named_variables()
            b = u.get_variable_store().buffers
            res.append([y.detach().cpu().numpy(), xg.grad.cpu().numpy()] +
                       [v[k].grad.cpu().numpy() for k in ('pw/weights', 'pw/biases', 'pw/bn/gamma', 'pw/bn/beta')] +
                       [b['pw/bn/moving_mean'].cpu().numpy(), b['pw/bn/moving_variance'].cpu().numpy()])
        finally:
            u.FUSED_TAIL = True
    for i, (a, w) in enumerate(zip(*res)):
        assert_close(a, w, 2e-5, "fused vs composition, item %d" % i)


def test_layer_tail_rejects_cpu_and_bad_shapes(pkg):
    lt = pkg.utils.layer_tail
    with pytest.raises(ValueError, match="CUDA"):
        lt.bias_act_bn(torch.zeros(4, 8))
    x = torch.zeros(4, 8, device="cuda:0")
    with pytest.raises(ValueError):
        lt.bias_act_bn(x, torch.zeros(7, device="cuda:0"))
    with pytest.raises(ValueError):
        lt.bias_act_bn(x, None, torch.ones(8, device="cuda:0"))          # gamma without beta / moving statistics


@pytest.mark.parametrize("act,training", [(True, True), (True, False), (False, True)])
def test_layer_tail_against_numpy_oracle(pkg, act, training):
    """sph3d_bias_act_bn[_grad] against oracle/oracle_layers.py (float64 numpy, gradients in closed form)"""
    import oracle_layers as OL
    lt = pkg.utils.layer_tail
    rng = np.random.default_rng(41)
    R, C = 5000, 128
    x = (rng.standard_normal((R, C)) * 1.3).astype(np.float32)
    go = rng.standard_normal((R, C)).astype(np.float32)
    bias = (rng.standard_normal(C) * 0.4).astype(np.float32)
    gamma, beta = (rng.random(C) + 0.5).astype(np.float32), rng.standard_normal(C).astype(np.float32)
    mm, mv = (rng.standard_normal(C) * 0.1).astype(np.float32), (rng.random(C) + 0.5).astype(np.float32)
    want, nmm, nmv, cache = OL.bias_act_bn(x, bias, gamma, beta, mm, mv, act=act, training=training)
    gx, gb, gg, gbe = OL.bias_act_bn_grad(cache, go, act=act, training=training)
    d = lambda a, rg=False: torch.from_numpy(a.copy()).to("cuda:0").requires_grad_(rg)
    xd, bd, gd, bed, mmd, mvd = d(x, True), d(bias, True), d(gamma, True), d(beta, True), d(mm), d(mv)
    out = lt.bias_act_bn(xd, bd, gd, bed, mmd, mvd, act=lt.ACT_ELU if act else lt.ACT_NONE, training=training)
    out.backward(d(go))
    tag = " act=%d train=%d" % (act, training)
    assert_close(out.detach().cpu().numpy(), want, 1e-5, "out" + tag)
    assert_close(xd.grad.cpu().numpy(), gx, 1e-5, "grad_x" + tag)
    assert_close(gd.grad.cpu().numpy(), gg, 1e-5, "grad_gamma" + tag)
    assert_close(bed.grad.cpu().numpy(), gbe, 1e-5, "grad_beta" + tag)
    summands = float(np.abs(gx).sum(0).max())
    assert np.abs(bd.grad.cpu().numpy() - gb).max() <= 1e-5 * summands
    assert_close(mmd.cpu().numpy(), nmm, 1e-5, "moving_mean" + tag)
    assert_close(mvd.cpu().numpy(), nmv, 1e-5, "moving_var" + tag)
