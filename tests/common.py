"""Seeded synthetic inputs shared by the CPU and GPU tests."""
import numpy as np


def make_cloud(seed, B, N, kind="cube"):
    """xyz (B,N,3) float32: 'cube' = uniform in the unit cube, 'shell' = unit sphere shell with 1% radial
    noise, 'grid' = points on a coarse lattice (many exact ties / duplicates / on-radius pairs)."""
    rng = np.random.default_rng(seed)
    if kind == "cube":
        xyz = rng.random((B, N, 3), dtype=np.float32)
    elif kind == "shell":
        v = rng.standard_normal((B, N, 3)).astype(np.float32)
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        xyz = (v * (1 + 0.01 * rng.standard_normal((B, N, 1)).astype(np.float32))).astype(np.float32)
    elif kind == "grid":
        xyz = (rng.integers(0, 9, size=(B, N, 3)).astype(np.float32) * np.float32(0.125))
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(xyz, dtype=np.float32)


def saturating_radius(N, K):
    """radius at which the mean in-range count in the unit cube is ~2K."""
    return float((3.0 * 2 * K / (4.0 * np.pi * N)) ** (1.0 / 3.0))


def features(seed, *shape):
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)


def assert_close(new, truth, rtol=1e-5, what=""):
    """|new - truth| <= rtol * max|truth| elementwise + rtol relative (fp32 tolerance of north_star)."""
    new = np.asarray(new, dtype=np.float64)
    truth = np.asarray(truth, dtype=np.float64)
    assert new.shape == truth.shape, (what, new.shape, truth.shape)
    scale = float(np.max(np.abs(truth))) if truth.size else 0.0
    err = np.abs(new - truth)
    tol = rtol * np.abs(truth) + rtol * scale + 1e-30
    bad = err > tol
    assert not bad.any(), "%s: %d/%d elements off, max err %.3e (scale %.3e)" % (
        what, int(bad.sum()), bad.size, float(err.max()), scale)


def assert_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind == "f":
        a, b = a.view(np.uint32), b.view(np.uint32)
    neq = a != b
    assert not neq.any(), "%s: %d/%d elements differ (first at %s)" % (
        what, int(neq.sum()), neq.size, tuple(np.argwhere(neq)[0]))


def assert_close_terms(new, truth, abs_terms, rtol=1e-5, what=""):
    """ELEMENTWISE bound for a sum of products: |new - truth| <= rtol * (sum of the magnitudes of that element's own
    summands) -- the backward-error form of "1e-5 relative fp32" (an fp32 sum of n terms is only accurate relative to
    sum |terms|, never relative to a result that cancels).  abs_terms is the same op evaluated on |inputs|.
    Returns (max error / abs_terms, max plain relative error over elements above 1e-3 of the scale) for reporting."""
    new = np.asarray(new, dtype=np.float64)
    truth = np.asarray(truth, dtype=np.float64)
    abs_terms = np.asarray(abs_terms, dtype=np.float64)
    assert new.shape == truth.shape == abs_terms.shape, (what, new.shape, truth.shape, abs_terms.shape)
    err = np.abs(new - truth)
    bad = err > rtol * abs_terms + 1e-30
    worst = float(np.max(err / np.maximum(abs_terms, 1e-30))) if err.size else 0.0
    scale = float(np.max(np.abs(truth))) if truth.size else 0.0
    big = np.abs(truth) > 1e-3 * scale
    rel = float(np.max(err[big] / np.abs(truth[big]))) if big.any() else 0.0
    assert not bad.any(), "%s: %d/%d elements beyond %g x sum|terms| (worst %.3e; max plain relative error %.3e)" % (
        what, int(bad.sum()), bad.size, rtol, worst, rel)
    return worst, rel
