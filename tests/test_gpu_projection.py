"""GPU projection + binning against the NumPy restatement of the reference's
stereograph_project / angular_project kernels (cl:488-538) and np.histogram2d calls (it:416,529,560).

x, y and the corrected power come from float32 transcendental functions whose last bits OpenCL
leaves to the implementation, so they are compared at 1e-6 relative (north_star tolerance); a ray
whose projected coordinate lies within that distance of a bin edge may land in the neighbouring
bin, therefore per-bin power is compared after allowing those few rays to move (their count is
bounded and reported), while the total binned power must agree to 1e-9."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tracer():
    from lightpycl_b200 import iterative_tracer as it, scenes

    ls, meshes, params = scenes.directivity_lens(ray_count=60_000)
    tr = it.CL_Tracer()
    tr.iterative_tracer(ls, meshes, **params)
    return tr


@pytest.fixture(scope="module")
def traced():
    return _tracer()


@pytest.mark.parametrize("kind,limits,points", [
    ("xy", ((-1000, 1000), (-1000, 1000)), 64),
    ("stereographic", ((-1, 1), (-1, 1)), 50),
    ("angular", ((-np.pi / 2.0, np.pi / 2.0), (-np.pi / 2.0, np.pi / 2.0)), 30),
    ("angular", ((-1, 1), (-1, 1)), 500),
])
def test_binned_data_matches_numpy(traced, kind, limits, points):
    from lightpycl_b200 import projection

    tr = traced
    pos, pwr = tr.get_measured_rays()
    pwr = np.asarray(pwr).reshape(-1)
    assert len(pwr) > 50_000
    method = {"xy": tr.get_binned_data, "stereographic": tr.get_binned_data_stereographic,
              "angular": tr.get_binned_data_angular}[kind]
    H, xe, ye = method(limits=limits, points=points)
    x, y, w = projection.project_numpy(kind, pos, pwr)
    Hn, xn, yn = projection.bin_numpy(x, y, w, limits, points, kind)
    assert H.shape == Hn.shape and np.array_equal(xe, xn) and np.array_equal(ye, yn)
    total, total_n = H.sum(), Hn.sum()
    assert abs(total - total_n) <= 1e-6 * abs(total_n)
    diff = np.abs(H - Hn)
    wmax = float(np.max(np.abs(np.float64(w)))) / (1.0 if kind == "xy" else
                                                    ((limits[0][1] - limits[0][0]) / points) ** 2)
    moved = diff.sum() / (2 * wmax) if wmax else 0.0  # each edge-flipped ray changes two bins by <= wmax
    bad = np.argwhere(diff > 1e-6 * np.maximum(np.abs(Hn), 1e-300) + 0)
    # bins that differ must be explained by a handful of edge rays
    assert moved <= 1e-4 * len(pwr) + 2, "too much power moved between bins: ~%g rays" % moved
    assert len(bad) <= 4 * (int(moved) + 1)
    if kind == "xy":  # no transcendental functions involved: exact bin membership
        assert len(bad) == 0


def test_projection_kernels_match_numpy(traced):
    import torch

    from lightpycl_b200 import _cabi, projection

    tr = traced
    lib = _cabi.load()
    rec = tr.measured_device()
    k = rec.shape[0]
    pos, pwr = tr.get_measured_rays()
    rot = np.array([[np.cos(0.7), -np.sin(0.7), 0], [np.sin(0.7), np.cos(0.7), 0], [0, 0, 1]], dtype=np.float32)
    for kind in ("stereographic", "angular"):
        for r in (None, rot):
            x, y, w = (torch.empty(k, dtype=torch.float32, device=rec.device) for _ in range(3))
            rc = lib.lpcl_project(rec.data_ptr(), k, r.ctypes.data if r is not None else None, projection.KINDS[kind],
                                  x.data_ptr(), y.data_ptr(), w.data_ptr(), None)
            _cabi.check(rc, "lpcl_project")
            xn, yn, wn = projection.project_numpy(kind, pos, np.asarray(pwr).reshape(-1), r)
            ok = np.isfinite(xn) & np.isfinite(yn)
            scale = np.maximum(np.abs(xn[ok]), np.abs(yn[ok])) + 1e-3
            assert np.max(np.abs(x.cpu().numpy()[ok] - xn[ok]) / scale) < 2e-6
            assert np.max(np.abs(y.cpu().numpy()[ok] - yn[ok]) / scale) < 2e-6
            assert np.allclose(w.cpu().numpy()[ok], wn[ok], rtol=2e-6, atol=0)


def test_replicated_sources_sum(traced):
    H1, _, _ = traced.get_binned_data_angular(limits=((-1.6, 1.6), (-1.6, 1.6)), points=40)
    H4, _, _ = traced.get_binned_data_replicated(limits=((-1.6, 1.6), (-1.6, 1.6)), points=40, axis="z", sources=4)
    assert abs(H4.sum() - 4 * H1.sum()) <= 1e-6 * abs(H4.sum())  # rotation about z keeps every ray inside the disc
