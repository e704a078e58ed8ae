"""Pinning of the projection / binning checker (VERDICT r1 "Next" 5).

oracle.project(backend="reference") runs the reference's OWN `stereograph_project` / `angular_project`
kernels (.cl:488-538), compiled from the reference source (oracle/_ref/libref_proj.so; the OpenCL vector
literals are rewritten on the fly, oracle/proj_shim.cpp).  The NumPy restatement (backend="port") is
pinned to it here: bit-exact where only +, *, /, sqrt are involved (stereographic; x, y, power), and to
1e-6 of the radius where float transcendental functions are (angular: glibc's acosf/atan2f/sinf/cosf vs
NumPy's; OpenCL itself only promises 4-5 ulp for them)."""

import numpy as np
import pytest

import util
from oracle import oracle

needs_ref = pytest.mark.skipif(not oracle.available("reference_proj"), reason="oracle/_ref/libref_proj.so not built")


def _points(n, seed, radius=1000.0):
    rng = np.random.RandomState(seed)
    d = rng.normal(size=(n, 3))
    d[:, 2] = np.abs(d[:, 2])
    d /= np.linalg.norm(d, axis=1)[:, None]
    pos = np.zeros((n, 4), np.float32)
    pos[:, :3] = (radius * d).astype(np.float32)
    pos[0, :3] = (0.0, 0.0, radius)  # the pole: r = 0 takes the A = 1 branch of .cl:531
    pwr = rng.rand(n).astype(np.float32)
    return pos, pwr


def _ulps(a, b):
    ai, bi = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-2 ** 31) - ai, ai)
    bi = np.where(bi < 0, np.int64(-2 ** 31) - bi, bi)
    return np.abs(ai - bi)


@needs_ref
@pytest.mark.parametrize("rot", [None, "z0.7"])
def test_stereographic_port_is_bit_exact(rot):
    pos, pwr = _points(50_000, 1)
    r = None if rot is None else np.array([[np.cos(0.7), -np.sin(0.7), 0], [np.sin(0.7), np.cos(0.7), 0], [0, 0, 1]], np.float32)
    a = oracle.project("stereographic", pos, pwr, r, backend="reference")
    b = oracle.project("stereographic", pos, pwr, r, backend="port")
    for x, y, what in zip(a, b, ("x", "y", "pwr")):
        assert np.array_equal(util.bits(x), util.bits(y)), what


@needs_ref
def test_angular_port_within_libm_tolerance():
    pos, pwr = _points(50_000, 2)
    a = oracle.project("angular", pos, pwr, backend="reference")
    b = oracle.project("angular", pos, pwr, backend="port")
    r = np.hypot(a[0].astype(np.float64), a[1].astype(np.float64)).astype(np.float32)
    for x, y, what in zip(a, b, ("x", "y", "pwr")):
        # x = r cos(phi), y = r sin(phi): compare at the scale of r (a coordinate near 0 has no relative meaning)
        if what == "pwr":
            assert _ulps(x, y).max() <= 4, what
        else:
            # phi = atan2(v, u) is good to ~2 ulp of up to pi, which moves r cos(phi) by up to ~7 ulp OF r: 1e-6 r
            assert np.max(np.abs(x.astype(np.float64) - y) / np.maximum(r, 1e-30)) <= 1e-6, what
    assert a[0][0] == 0.0 and a[1][0] == 0.0 and a[2][0] == pwr[0]  # the pole


def test_bin2d_is_histogram2d_with_the_reference_weights():
    pos, pwr = _points(5_000, 3)
    x, y, w = oracle.project("angular", pos, pwr, backend="port")
    lim = ((-1.0, 1.0), (-1.0, 1.0))
    H, xe, ye = oracle.bin2d(x, y, w, lim, 40, "angular")
    dxdy = (2.0 / 40) ** 2
    inside = (np.abs(x) <= 1) & (np.abs(y) <= 1)
    assert abs(H.sum() - np.float64(w[inside]).sum() / dxdy) <= 1e-9 * H.sum()
    Hxy, _, _ = oracle.bin2d(pos[:, 0], pos[:, 1], pwr, ((-1000, 1000), (-1000, 1000)), 16, "xy")
    assert abs(Hxy.sum() - np.float64(pwr).sum()) <= 1e-5 * Hxy.sum()
