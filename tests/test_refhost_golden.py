"""Host side pinned to the reference's OWN Python code.

tests/golden/refhost_*.npz were produced by running /root/reference/{light_source, geo_optical_elements,
iterative_tracer}.py themselves (oracle/ref_host.py: Python-2 source executed by path under a Python-2-semantics shim,
pyopencl replaced by a stand-in whose kernels are the reference .cl compiled by g++; generator:
tests/golden/make_golden_host.py).  Here this package's mirrors - and the oracle's restated host loop - run the same
calls with the same seeds (tests/refhost_cases.py) and must reproduce those outputs:

  light sources        origin / direction / power arrays bit for bit, (N,1) vs (N,) power shapes included
  mesh generators      tribuf() vertex buffers bit for bit (sha1 over all triangles + head/tail samples), materials
  trace loop           oracle.trace == the reference's iterative_tracer loop, every iteration's
                       (origin, dest, power, measured) bit for bit, ray partitions (> 50,000 rays) included
  projections          oracle.project + bin2d == get_binned_data / _stereographic / _angular / replicated `hist_data`
  beam metrics         get_beam_width_half_power / get_beam_HWHM numbers == the numbers the reference prints

No GPU: the CUDA path is compared with the same fixtures in tests/test_gpu_refhost.py."""

import hashlib
import os

import numpy as np
import pytest

import refhost_cases as cases
import util
from lightpycl_b200 import geo_optical_elements as goe, iterative_tracer as it, light_source as lsm, scenes
from oracle import oracle


def _load(name):
    return np.load(os.path.join(util.GOLDEN_DIR, name), allow_pickle=False)


def _sha1(a):
    return hashlib.sha1(np.ascontiguousarray(np.asarray(a)).tobytes()).hexdigest()


def _same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(util.bits(a), util.bits(b))


@pytest.mark.parametrize("name", sorted(cases.source_cases(lsm)))
def test_light_source_matches_reference(name):
    z = _load("refhost_sources.npz")
    s = cases.source_cases(lsm)[name]()
    assert int(s.ray_count) == int(z[name + "/ray_count"])
    for key, mine in (("origin", s.rays_origin), ("dir", s.rays_dir), ("power", s.rays_power)):
        want = z[name + "/" + key]
        mine = np.asarray(mine)
        assert mine.shape == want.shape, (key, mine.shape, want.shape)
        assert mine.dtype == want.dtype, (key, mine.dtype, want.dtype)
        assert np.array_equal(util.bits(mine), util.bits(want)), key


@pytest.mark.parametrize("name", sorted(cases.geometry_cases(goe)))
def test_mesh_generator_matches_reference(name):
    z = _load("refhost_geometry.npz")
    m = cases.geometry_cases(goe)[name]()
    v = [np.asarray(b, dtype=np.float32) for b in m.tribuf()]
    assert len(v[0]) == int(z[name + "/count"])
    assert np.array_equal(np.stack([b[:4] for b in v]), z[name + "/head"])
    assert np.array_equal(np.stack([b[-4:] for b in v]), z[name + "/tail"])
    assert [_sha1(b) for b in v] == list(z[name + "/sha1"])
    mat = m.getMaterialBuf()
    assert np.array_equal(np.array([mat["type"], mat["IOR"], mat["R"], mat["dissipation"]], dtype=np.float64), z[name + "/material"])


def _check_results(results, z, digest_only):
    assert len(results) == int(z["n_iter"]), ([len(r[0]) for r in results], z["rays"])
    for k, arrs in enumerate(results):
        for key, a in zip(("origin", "dest", "power", "meas"), arrs):
            a = np.asarray(a)
            assert tuple(z["it%d_%s_shape" % (k, key)]) == a.shape, (k, key)
            if key in ("origin", "dest"):
                # the w lane of a float4 is don't-care (cl: float3 buffers are 16 bytes wide, lane 3 is never written
                # by the kernels): compare what the reference's arrays hold in x, y, z
                pass
            if digest_only:
                assert _sha1(a) == str(z["it%d_%s_sha1" % (k, key)]), (k, key)
            else:
                want = z["it%d_%s" % (k, key)]
                assert a.dtype == want.dtype and np.array_equal(util.bits(a), util.bits(want)), (k, key)


@pytest.mark.parametrize("name", sorted(cases.trace_cases()))
def test_oracle_host_loop_matches_reference_host_loop(name):
    """oracle.trace (the restated loop, it:77-393, around the reference kernels) against the reference's own loop."""
    if not oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    case = cases.trace_cases()[name]
    z = _load("refhost_trace_%s.npz" % name)
    ls, meshes, params = getattr(scenes, case["scene"])(**case["kw"])
    out = oracle.trace(ls, meshes, backend="reference", power_sum="reference", keep_debug=False, **params)
    _check_results(out["results"], z, case.get("digest_only", False))
    pos, pwr = oracle.measured_rays(out["results"])
    assert (0 if pos is None else len(pos)) == int(z["measured_count"])
    if int(z["measured_count"]) and not case.get("digest_only"):
        assert _same_bits(pos, z["measured_pos"]) and _same_bits(pwr, z["measured_pwr"])


POST = [n for n, c in cases.trace_cases().items() if c.get("post")]


@pytest.mark.parametrize("name", POST)
def test_projection_checker_matches_reference_pipeline(name):
    """oracle.project + oracle.bin2d (the checker of the GPU projection kernels) against the hist_data the reference's
    get_binned_data* / replicate_lightsources_and_plot produce from the same measured rays."""
    case = cases.trace_cases()[name]
    z = _load("refhost_trace_%s.npz" % name)
    pos, pwr = z["measured_pos"], np.asarray(z["measured_pwr"]).reshape(-1)
    pts, lim = case["points"], ((-1, 1), (-1, 1))
    backend = "reference" if oracle.available("reference_proj") else "port"
    exact = backend == "reference"
    for kind, key, limits in (("xy", "hist_xy", case["limits_xy"]), ("stereographic", "hist_stereo", lim), ("angular", "hist_angular", lim)):
        x, y, w = (pos[:, 0], pos[:, 1], pwr) if kind == "xy" else oracle.project(kind, pos, pwr, backend=backend)
        H, xe, ye = oracle.bin2d(x, y, w, limits, pts, kind)
        assert np.array_equal(np.stack([xe, ye]), z[key + "_edges"])
        if exact or kind != "angular":
            assert np.array_equal(H, z[key]), kind
        else:
            assert abs(H.sum() - z[key].sum()) <= 1e-6 * z[key].sum()
    # replicated sources (it:564-628): rotation k 2 pi / sources about z, angular projection, concatenate, one histogram
    total = None
    xs, ys, ws = [], [], []
    for k in range(case["replicas"]):
        a = k * 2.0 * np.pi / case["replicas"]
        R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]).astype(np.float32)
        x, y, w = oracle.project("angular", pos, pwr, R, backend=backend)
        xs.append(x), ys.append(y), ws.append(w)
    H, xe, ye = oracle.bin2d(np.concatenate(xs), np.concatenate(ys), np.concatenate(ws), lim, pts, "angular")
    if exact:
        assert np.array_equal(H, z["hist_replicated"])
    else:
        assert abs(H.sum() - z["hist_replicated"].sum()) <= 1e-6 * z["hist_replicated"].sum()


class _Measured(it.CL_Tracer):
    """The beam-metric methods only need get_measured_rays(): a tracer shell without a device."""

    def __init__(self, pos, pwr):  # noqa: D401 - deliberately does not call CL_Tracer.__init__ (needs a GPU)
        self._pos, self._pwr = pos, pwr

    def get_measured_rays(self, gather=True):
        return self._pos, self._pwr

    def __del__(self):
        pass


@pytest.mark.parametrize("name", POST)
def test_beam_metrics_match_reference_printout(name, capsys):
    case = cases.trace_cases()[name]
    z = _load("refhost_trace_%s.npz" % name)
    tr = _Measured(z["measured_pos"], z["measured_pwr"])
    total, angle, throughput = tr.get_beam_width_half_power()
    want = z["half_power"]  # [total, angle(s) in degrees ..., throughput(s) in percent ...]
    n = (len(want) - 1) // 2
    assert len(np.atleast_1d(angle)) == n
    assert abs(total - want[0]) <= 1e-9 * abs(want[0])
    # the reference PRINTS its arrays with NumPy's default 8 significant digits
    assert np.allclose(np.atleast_1d(angle) / np.pi * 180.0, want[1:1 + n], rtol=2e-8, atol=0)
    assert np.allclose(np.atleast_1d(throughput) * 100.0, want[1 + n:], rtol=2e-8, atol=0)
    total, angle, throughput = tr.get_beam_HWHM(points=case["hwhm_points"])
    want = z["hwhm"]
    n = len(np.atleast_1d(angle))
    assert abs(total - want[0]) <= 1e-9 * abs(want[0])
    assert np.allclose(np.atleast_1d(angle) / np.pi * 180.0, want[1:1 + n], rtol=2e-8, atol=0)
    assert np.allclose(np.atleast_1d(throughput) * 100.0, want[1 + n:], rtol=2e-8, atol=0)
    capsys.readouterr()
