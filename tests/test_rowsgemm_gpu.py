"""The hand-written tcgen05 rows product (csrc/rowsgemm.cu, C ABI sph3d_rows_gemm): y = x w and gx = g w^T of a layer's
pointwise product (tf.matmul over the B*M rows, /root/reference/utils/sph3gcn_util.py:144-146, :203-205, :254-256) against
a float64 product.  Tolerance: every element within 1e-5 of the sum of the magnitudes of its own terms (the elementwise
backward-error bound of common.assert_close_terms) -- what an fp32 GEMM delivers; a single bf16 or tf32 pass misses it by
two to three orders of magnitude."""
import numpy as np
import pytest
import torch

from common import assert_close, assert_close_terms

pytestmark = pytest.mark.gpu

# (rows, K, N): one tile, ragged rows / channels, K tail inside a 64-chunk, three column groups, a resident and a streamed
# weight image, one block per CTA (few row tiles), the headline layer, more row tiles than SMs
SHAPES = [(100, 64, 32), (1000, 68, 132), (4096, 128, 516), (65536, 256, 128), (6144, 512, 256), (3072, 2048, 256),
          (20000, 72, 64), (2052, 36, 512), (40001, 128, 128), (129, 4, 4)]


def _inputs(R, K, N):
    g = torch.Generator().manual_seed(R + 7 * K + 13 * N)
    x = (torch.randn(R, K, generator=g) * 2 + 0.5).cuda()
    w = (torch.randn(K, N, generator=g) * 0.3).cuda()
    go = torch.randn(R, N, generator=g).cuda()
    return x, w, go


@pytest.mark.parametrize("terms", [3, 2])
@pytest.mark.parametrize("R,K,N", SHAPES)
def test_rows_product_and_its_input_gradient(pkg, R, K, N, terms):
    rg = pkg.tf_rowsgemm
    x, w, go = _inputs(R, K, N)
    y = rg.rows_gemm(x, w, terms=terms)
    gx = rg.rows_gemm(go, w, trans=True, terms=terms)
    xd, wd, gd = x.double(), w.double(), go.double()
    assert_close_terms(y.cpu().numpy(), (xd @ wd).cpu().numpy(), (xd.abs() @ wd.abs()).cpu().numpy(), 1e-5,
                       "y = x w (%d,%d,%d) terms=%d" % (R, K, N, terms))
    assert_close_terms(gx.cpu().numpy(), (gd @ wd.t()).cpu().numpy(), (gd.abs() @ wd.abs().t()).cpu().numpy(), 1e-5,
                       "gx = g w^T (%d,%d,%d) terms=%d" % (R, K, N, terms))


def test_packed_image_is_reusable_and_rows_beyond_the_last_tile_are_untouched(pkg):
    rg = pkg.tf_rowsgemm
    x, w, _ = _inputs(1000, 68, 132)
    img = rg.pack(w)
    a = rg.rows_gemm(x, w, image=img)
    b = rg.rows_gemm(x, w, image=img)
    assert torch.equal(a, b), "two launches over one image differ"
    # the kernel writes rows < R only: a guard band behind the output stays as it was
    R, K, N = 1000, 68, 132
    buf = torch.full((R + 64, N), 7.0, device="cuda")
    L = pkg._lib.lib()
    rc = L.sph3d_rows_gemm(R, K, N, 3, x.data_ptr(), img.data_ptr(), buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0
    assert torch.equal(buf[:R], a) and bool((buf[R:] == 7.0).all())


def test_pair_pack_equals_two_packs(pkg):
    rg = pkg.tf_rowsgemm
    for K, N in ((68, 132), (256, 128), (4, 516)):
        w = torch.randn(K, N, device="cuda")
        a, b = rg.pack_pair(w)
        assert torch.equal(a, rg.pack(w)) and torch.equal(b, rg.pack(w, trans=True))


def test_refuses_what_it_cannot_do(pkg):
    rg = pkg.tf_rowsgemm
    x = torch.randn(64, 6, device="cuda")
    w = torch.randn(6, 8, device="cuda")
    with pytest.raises(ValueError):
        rg.rows_gemm(x, w)                                # K not a multiple of 4
    L = pkg._lib.lib()
    y = torch.empty(64, 8, device="cuda")
    img = torch.empty(L.sph3d_rows_gemm_image_bytes(8, 8), dtype=torch.uint8, device="cuda")
    assert L.sph3d_rows_gemm(64, 8, 8, 4, x.data_ptr(), img.data_ptr(), y.data_ptr(), 0) == 1     # terms must be 2 or 3
    assert L.sph3d_rows_gemm(64, 6, 8, 3, x.data_ptr(), img.data_ptr(), y.data_ptr(), 0) == 1


def test_layer_product_routes_through_the_rows_product(pkg, monkeypatch):
    """_Dense (forward + both gradients) with the hand-written product on and off agrees to fp32 accuracy, and the kernel
    really runs when it is on"""
    u = pkg.sph3gcn_util
    x, w, go = _inputs(8192, 128, 256)
    outs = []
    for on in (True, False):
        monkeypatch.setattr(u, "ROWS_GEMM", on)
        xd, wd = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        calls = []
        real = pkg.tf_rowsgemm.rows_gemm
        monkeypatch.setattr(pkg.tf_rowsgemm, "rows_gemm", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
        y = u._dense(xd, wd)
        y.backward(go)
        monkeypatch.setattr(pkg.tf_rowsgemm, "rows_gemm", real)
        assert (len(calls) == 2) == on
        outs.append((y.detach().cpu().numpy(), xd.grad.cpu().numpy(), wd.grad.cpu().numpy()))
    for a, b, name in zip(outs[0], outs[1], ("y", "grad_x", "grad_w")):
        assert_close(a, b, 2e-5, "rows product on vs off: " + name)
