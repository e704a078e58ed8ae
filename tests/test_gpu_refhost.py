"""The CUDA path against fixtures produced by the REFERENCE'S OWN host code (tests/golden/refhost_*.npz: the reference's
iterative_tracer.py loop, get_measured_rays, projection / replication functions and beam metrics, run by path under
oracle/ref_host.py around the reference kernels; generator tests/golden/make_golden_host.py).  The scenes are rebuilt
here with this package's generators (pinned to the reference's by tests/test_refhost_golden.py) and traced on the GPU:

  results        every iteration's (origin, dest, power, measured), in the reference's order, bit for bit
  measured rays  (pos, pwr) bit for bit, shape quirks included
  hist_data      get_binned_data exact; stereographic / angular / replicated: total to 1e-6, per bin to 1e-6 except bins
                 touched by a ray within 2e-6 of a bin edge (float transcendental functions; counted and reported)
  beam metrics   the numbers the reference prints"""

import hashlib
import os

import numpy as np
import pytest

import refhost_cases as cases
import util

pytestmark = pytest.mark.gpu


def _load(name):
    return np.load(os.path.join(util.GOLDEN_DIR, name), allow_pickle=False)


def _sha1(a):
    return hashlib.sha1(np.ascontiguousarray(np.asarray(a)).tobytes()).hexdigest()


def _trace(name):
    from lightpycl_b200 import iterative_tracer as it, scenes

    case = cases.trace_cases()[name]
    ls, meshes, params = getattr(scenes, case["scene"])(**case["kw"])
    tr = it.CL_Tracer()
    res = tr.iterative_tracer(ls, meshes, **params)
    return case, tr, res


@pytest.mark.parametrize("name", sorted(cases.trace_cases()))
def test_gpu_trace_matches_reference_host_run(name):
    case, tr, res = _trace(name)
    z = _load("refhost_trace_%s.npz" % name)
    assert len(res) == int(z["n_iter"]), ([len(r[0]) for r in res], z["rays"].tolist())
    for k, arrs in enumerate(res):
        for key, a in zip(("origin", "dest", "power", "meas"), arrs):
            a = np.asarray(a)
            assert tuple(z["it%d_%s_shape" % (k, key)]) == a.shape, (k, key, a.shape)
            if case.get("digest_only"):
                if key in ("power", "meas"):  # the 4-vectors' w lane is don't-care: digests only cover the scalars
                    assert _sha1(a) == str(z["it%d_%s_sha1" % (k, key)]), (k, key)
                continue
            want = z["it%d_%s" % (k, key)]
            if key in ("origin", "dest"):
                a, want = a[:, :3], want[:, :3]
            assert np.array_equal(util.bits(np.ascontiguousarray(a)), util.bits(np.ascontiguousarray(want))), (k, key)
    pos, pwr = tr.get_measured_rays()
    n = 0 if pos is None else len(pos)
    assert n == int(z["measured_count"])
    if n and not case.get("digest_only"):
        assert np.asarray(pwr).shape == z["measured_pwr"].shape
        assert np.array_equal(util.bits(np.ascontiguousarray(np.asarray(pos)[:, :3])), util.bits(np.ascontiguousarray(z["measured_pos"][:, :3])))
        assert np.array_equal(util.bits(np.asarray(pwr)), util.bits(z["measured_pwr"]))


def _edge_bins(x, y, limits, points, rel):
    """Bins a ray within rel bin widths of a bin edge could fall into: its own and the neighbour across that edge."""
    touched = np.zeros((points, points), bool)
    wx, wy = (limits[0][1] - limits[0][0]) / points, (limits[1][1] - limits[1][0]) / points
    tx, ty = (np.float64(x) - limits[0][0]) / wx, (np.float64(y) - limits[1][0]) / wy
    near = (np.abs(tx - np.round(tx)) <= rel) | (np.abs(ty - np.round(ty)) <= rel)
    for a, b in zip(tx[near], ty[near]):
        for da in (-1, 0, 1):
            for db in (-1, 0, 1):
                i, j = int(np.floor(a)) + da, int(np.floor(b)) + db
                if 0 <= i < points and 0 <= j < points:
                    touched[i, j] = True
    return touched, int(near.sum())


@pytest.mark.parametrize("name", [n for n, c in cases.trace_cases().items() if c.get("post")])
def test_gpu_post_processing_matches_reference_host_run(name, capsys):
    from oracle import oracle

    case, tr, _ = _trace(name)
    z = _load("refhost_trace_%s.npz" % name)
    pts, lim = case["points"], ((-1, 1), (-1, 1))
    notes = []
    pos, pwr = z["measured_pos"], np.asarray(z["measured_pwr"]).reshape(-1)
    H, xe, ye = tr.get_binned_data(limits=case["limits_xy"], points=pts)
    assert np.array_equal(np.stack([xe, ye]), z["hist_xy_edges"])
    assert np.allclose(H, z["hist_xy"], rtol=1e-12, atol=0)  # same rays into the same bins; float64 sums in another order
    for method, key, kind in ((tr.get_binned_data_stereographic, "hist_stereo", "stereographic"),
                              (tr.get_binned_data_angular, "hist_angular", "angular")):
        H, xe, ye = method(limits=lim, points=pts)
        want = z[key]
        assert np.array_equal(np.stack([xe, ye]), z[key + "_edges"])
        assert abs(H.sum() - want.sum()) <= 1e-6 * want.sum(), kind
        x, y, _ = oracle.project(kind, pos, pwr)
        touched, n_near = _edge_bins(x, y, lim, pts, 2e-6 * 1.6 / (2.0 / pts))
        bad = np.abs(H - want) > 1e-6 * np.abs(want).max()
        notes.append("%s %s: %d measured rays, %d within 2e-6 of a bin edge, %d of %d bins differ by more than 1e-6 (all touched by such a ray: %s)"
                     % (name, kind, len(pwr), n_near, int(bad.sum()), pts * pts, bool(np.all(touched[bad]))))
        assert np.all(touched[bad]), kind
    H, xe, ye = tr.get_binned_data_replicated(limits=lim, points=pts, axis="z", sources=case["replicas"])
    want = z["hist_replicated"]
    assert np.array_equal(np.stack([xe, ye]), z["hist_replicated_edges"])
    assert abs(H.sum() - want.sum()) <= 1e-6 * want.sum()
    assert np.sum(np.abs(H - want) > 1e-6 * np.abs(want).max()) <= 4 * case["replicas"]
    total, angle, throughput = tr.get_beam_width_half_power()
    want = z["half_power"]
    n = (len(want) - 1) // 2
    assert abs(total - want[0]) <= 1e-9 * abs(want[0])
    assert np.allclose(np.atleast_1d(angle) / np.pi * 180.0, want[1:1 + n], rtol=2e-8, atol=0)
    assert np.allclose(np.atleast_1d(throughput) * 100.0, want[1 + n:], rtol=2e-8, atol=0)
    total, angle, throughput = tr.get_beam_HWHM(points=case["hwhm_points"])
    want = z["hwhm"]
    n = len(np.atleast_1d(angle))
    assert abs(total - want[0]) <= 1e-9 * abs(want[0])
    assert np.allclose(np.atleast_1d(angle) / np.pi * 180.0, want[1:1 + n], rtol=2e-8, atol=0)
    assert np.allclose(np.atleast_1d(throughput) * 100.0, want[1 + n:], rtol=2e-8, atol=0)
    capsys.readouterr()  # the beam metrics print the reference's report lines
    with capsys.disabled():
        print("\n" + "\n".join(notes))
