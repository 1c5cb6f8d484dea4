"""The three products of a layer's pointwise product (y = x w, gx = g w^T, gw = x^T g) as the layer library routes them:
hand-written tcgen05 kernels (csrc/rowsgemm.cu, csrc/rowswgrad.cu) at fp32 accuracy where the shape allows, the library
GEMM elsewhere, against a float64 product.  Tolerance: 1e-5 of the result's scale (north_star), i.e. what the fp32 SIMT GEMM
of the reference's tf.matmul delivers; a single-pass TF32 or BF16 product would miss it by two orders of magnitude.
(The kernels themselves are tested against the elementwise bound in tests/test_rowsgemm_gpu.py.)"""
import numpy as np
import pytest
import torch

from common import assert_close, assert_close_terms

pytestmark = pytest.mark.gpu


def _ops(pkg):
    return pkg.sph3gcn_util


@pytest.mark.parametrize("R,K,N", [(4096, 128, 128), (65536, 256, 128), (3072, 2048, 256), (20000, 72, 64), (2052, 36, 512)])
def test_forward_and_input_gradient_products(pkg, R, K, N):
    u = _ops(pkg)
    g = torch.Generator().manual_seed(R + K + N)
    x = (torch.randn(R, K, generator=g) * 2 + 0.5).cuda()
    w = (torch.randn(K, N, generator=g) * 0.3).cuda()
    go = torch.randn(R, N, generator=g).cuda()
    y = u._rows_gemm(x, w, False)
    assert y is not None, "the tensor-core path refused an aligned shape"
    assert_close(y.cpu().numpy(), (x.double() @ w.double()).cpu().numpy(), 1e-5, "y = x w (%d,%d,%d)" % (R, K, N))
    gx = u._rows_gemm(go, w, True)
    assert gx is not None
    assert_close(gx.cpu().numpy(), (go.double() @ w.double().t()).cpu().numpy(), 1e-5, "gx = g w^T (%d,%d,%d)" % (R, K, N))


@pytest.mark.parametrize("terms", [3, 2])
@pytest.mark.parametrize("R,K,N", [(65536, 256, 128), (6144, 512, 256), (50001, 72, 64), (3072, 2048, 256), (1000, 68, 132),
                                   (100, 4, 4), (40000, 128, 516)])
def test_weight_gradient_kernel(pkg, R, K, N, terms):
    """csrc/rowswgrad.cu against float64, elementwise bound; ragged rows (last stage, last slab), K / N tails, one and
    many blocks of gw"""
    rg = pkg.tf_rowsgemm
    g = torch.Generator().manual_seed(R + 3 * K + N)
    x = (torch.randn(R, K, generator=g) + 0.25).cuda()
    go = torch.randn(R, N, generator=g).cuda()
    gw = rg.rows_wgrad(x, go, terms=terms)
    xd, gd = x.double(), go.double()
    assert_close_terms(gw.cpu().numpy(), (xd.t() @ gd).cpu().numpy(), (xd.abs().t() @ gd.abs()).cpu().numpy(), 1e-5,
                       "gw = x^T g (%d,%d,%d) terms=%d" % (R, K, N, terms))
    assert torch.equal(gw, rg.rows_wgrad(x, go, terms=terms))       # slabs summed in a fixed order


@pytest.mark.parametrize("R,K,N", [(65536, 256, 128), (6144, 512, 256), (50001, 72, 64), (50001, 3, 32), (8192, 64, 13)])
def test_weight_gradient_split_k(pkg, R, K, N):
    u = _ops(pkg)
    g = torch.Generator().manual_seed(R)
    x = (torch.randn(R, K, generator=g) + 0.25).cuda()
    go = torch.randn(R, N, generator=g).cuda()
    gw = u._weight_grad(x, go)
    assert_close(gw.cpu().numpy(), (x.double().t() @ go.double()).cpu().numpy(), 1e-5, "gw = x^T g (%d,%d,%d)" % (R, K, N))
    again = u._weight_grad(x, go)
    assert torch.equal(gw, again)                                   # slabs summed in a fixed order


def test_unaligned_shapes_fall_back_to_the_library_gemm(pkg):
    u = _ops(pkg)
    x = torch.randn(4096, 3, device="cuda")
    w = torch.randn(3, 32, device="cuda")
    assert u._rows_gemm(x, w, False) is None                          # K = 3: no 16-byte rows
    assert u._rows_gemm(torch.randn(4096, 64, device="cuda"), torch.randn(64, 13, device="cuda"), False) is None
    y = u._dense(x, w)
    assert_close(y.cpu().numpy(), (x.double() @ w.double()).cpu().numpy(), 1e-5, "fallback product")


def test_layer_with_and_without_tensor_cores(pkg):
    """pointwise_conv3d end to end: same outputs and gradients with the tensor-core products on and off"""
    u = _ops(pkg)
    torch.manual_seed(5)
    x = torch.randn(8, 2048, 64, device="cuda")
    wgt = torch.randn(8, 2048, 128, device="cuda")
    res = []
    for on in (True, False):
        u.reset_variables()
        u.ROWS_GEMM = on
        try:
            torch.manual_seed(6)
            xg = x.clone().requires_grad_(True)
            y = u.pointwise_conv3d(xg, 128, 'tc', with_bn=True, is_training=True)
            (y * wgt).sum().backward()
            v = u.named_variables()
            res.append([y.detach().cpu().numpy(), xg.grad.cpu().numpy(), v['tc/weights'].grad.cpu().numpy()])
        finally:
            u.ROWS_GEMM = True
    for i, (a, b) in enumerate(zip(*res)):
        assert_close(a, b, 2e-5, "tensor cores on vs off, item %d" % i)


def test_layer_products_against_numpy_oracle(pkg):
    """_Dense (forward + both gradients through autograd) against oracle/oracle_layers.py"""
    import oracle_layers as OL
    u = _ops(pkg)
    rng = np.random.default_rng(8)
    R, K, N = 16384, 256, 128
    x, w, go = (rng.standard_normal(s).astype(np.float32) for s in ((R, K), (K, N), (R, N)))
    xd = torch.from_numpy(x).cuda().requires_grad_(True)
    wd = torch.from_numpy(w).cuda().requires_grad_(True)
    y = u._dense(xd, wd)
    y.backward(torch.from_numpy(go).cuda())
    gx, gw = OL.dense_grad(x, w, go)
    assert_close(y.detach().cpu().numpy(), OL.dense(x, w), 1e-5, "y")
    assert_close(xd.grad.cpu().numpy(), gx, 1e-5, "grad_x")
    assert_close(wd.grad.cpu().numpy(), gw, 1e-5, "grad_w")


def test_weight_gradient_on_the_second_stream_changes_nothing(pkg):
    """OVERLAP_WEIGHT_GRAD: gw is enqueued on a side stream next to gx; results are bit-identical to the serial order"""
    u = _ops(pkg)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(6144, 256, generator=g).cuda()
    w = (torch.randn(256, 128, generator=g) * 0.2).cuda()
    go = torch.randn(6144, 128, generator=g).cuda()
    res = []
    for on in (True, False):
        u.OVERLAP_WEIGHT_GRAD = on
        try:
            for _ in range(3):                                      # repeated: a missing join would show up as a race
                xd, wd = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
                y = u._dense(xd, wd)
                (y * 2.0).backward(go)
            torch.cuda.synchronize()
            res.append((xd.grad.clone(), wd.grad.clone()))
        finally:
            u.OVERLAP_WEIGHT_GRAD = True
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
