"""The rest of the CL_Tracer surface a LightPyCL script touches after the trace (reference
it:395-779): lazy `results`, pickling, DXF export, beam metrics - exercised on a small scene."""

import glob
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def traced():
    from lightpycl_b200 import iterative_tracer as it, scenes

    ls, meshes, params = scenes.parabolic_mirror(ray_count=4000, profile_pts=60, ang_pts=37)
    tr = it.CL_Tracer(platform_name="NVIDIA", device_name="460")  # reference-style arguments are accepted
    res = tr.iterative_tracer(light_source=ls, meshes=meshes, **params)
    assert res is tr.results
    return tr, ls, meshes, params


def test_results_behave_like_the_reference_list(traced):
    tr, ls, meshes, params = traced
    res = tr.results
    assert len(res) == len(res.ray_counts) >= 2
    proc = sum(len(r[3]) for r in res)  # the loop every reference example runs (ex_lens:86-90)
    assert proc == sum(res.ray_counts)
    o, d, p, m = res[0]
    assert o.shape == (4000, 4) and d.shape == (4000, 4) and p.shape == (4000, 1) and m.dtype == np.int32
    assert set(np.unique(m)) <= {-1, 0, 1}
    assert res[-1][0].shape[0] == res.ray_counts[-1] and res[0:1][0][2].shape == (4000, 1)
    assert int(tr.tri_count) == sum(len(x.triangles) for x in meshes) and len(tr.geometry) == 3


def test_pickle_roundtrip_and_dxf(traced, tmp_path, monkeypatch):
    from lightpycl_b200 import iterative_tracer as it

    tr, ls, meshes, params = traced
    monkeypatch.chdir(tmp_path)
    fname = tr.pickle_results()
    assert fname and os.path.exists(fname)
    other = it.CL_Tracer()
    other.load_pickle_results(fname)
    assert len(other.results) == len(tr.results)
    for a, b in zip(other.results, tr.results):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    tr.save_traced_scene(str(tmp_path / "scene.dxf"))
    txt = open(tmp_path / "scene.dxf").read()
    assert txt.count("3DFACE") == sum(tr.results.ray_counts) + int(tr.tri_count) and txt.rstrip().endswith("EOF")
    meshes[2].write_dxf(str(tmp_path / "plate.dxf"))
    assert open(tmp_path / "plate.dxf").read().count("3DFACE") == 12
    ls[0].save_dxf(str(tmp_path / "rays.dxf"))


def test_beam_metrics(traced, capsys):
    tr, ls, meshes, params = traced
    total, angle, frac = tr.get_beam_width_half_power(points=100, pole=[0, 0, 1, 0])
    pos, pwr = tr.get_measured_rays()
    assert abs(total - float(np.sum(np.asarray(pwr, np.float64)))) <= 1e-9 * total
    assert 0.0 <= float(angle[0]) <= np.pi / 2 and 0.4 < float(frac[0]) < 0.6
    total2, hwhm, frac2 = tr.get_beam_HWHM(points=100, pole=[0, 0, 1, 0])
    assert abs(total2 - total) <= 1e-9 * total and 0.0 <= float(hwhm[0]) <= np.pi / 2
    x, inten = tr.get_elevation_histogram(points=90, pole=[0, 0, 1, 0])
    assert x.shape == (90,) and np.all(np.isfinite(inten))
    out = capsys.readouterr().out
    assert "Beam halfpower angle" in out and "Beam HWHM angle" in out
    # a collimated beam from a parabola: almost all measured power within a few degrees of the pole
    elev = np.arccos(pos[:, 2] / np.linalg.norm(pos[:, :3], axis=1))
    w = np.asarray(pwr, np.float64).reshape(-1)
    assert w[elev < np.deg2rad(5)].sum() > 0.9 * w.sum()


@pytest.mark.parametrize("name,kw", [("c4_human_eye", dict(ray_count=500, sources=2)),
                                     ("c5p_lenslet_array_66_meshes", dict(ray_count=300, sources=4, grid=5, ang_pts=9)),
                                     ("c1_nested_cubes", dict(ray_count=200))])
def test_device_side_triangle_gather_equals_host_flatten(name, kw, monkeypatch):
    """The default set-up uploads the INDEXED meshes and gathers v0/v1/v2 on the GPU (index_meshes); with
    LPCL_DEVICE_GATHER=0 the host gathers (flatten_meshes, it:121-151).  Same device buffers, same trace, same
    `geometry` attribute - few meshes (one upload per mesh) and many (27 meshes: one concatenated upload)."""
    from lightpycl_b200 import iterative_tracer as it, scenes

    ls, meshes, params = scenes.CONFIGS[name](**kw)
    params = dict(params, trace_iterations=min(int(params["trace_iterations"]), 5))
    tr_dev = it.CL_Tracer()
    assert tr_dev.device_gather
    res_dev = tr_dev.iterative_tracer(ls, meshes, **params)
    monkeypatch.setenv("LPCL_DEVICE_GATHER", "0")
    tr_host = it.CL_Tracer()
    assert not tr_host.device_gather
    res_host = tr_host.iterative_tracer(ls, meshes, **params)
    for k in ("v0", "v1", "v2"):
        a, b = tr_dev._scene_tensors[k].cpu().numpy(), tr_host._scene_tensors[k].cpu().numpy()
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), k
    util.assert_results_equal(list(res_dev), list(res_host), name)
    assert tr_dev._geometry is None  # not needed for tracing: built on first access
    for a, b in zip(tr_dev.geometry, tr_host.geometry):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_search_variant_hint_keeps_results(monkeypatch):
    """lpcl_scene_set_tc_warps picks the epilogue variant of the tensor-core search (0 = deferred flags + second pass,
    the default; 8 / 16 = a vote per chunk test with fat / thin warps); CL_Tracer.SEARCH_VARIANT passes it on.  The
    results do not depend on it."""
    from lightpycl_b200 import iterative_tracer as it, scenes

    ls, meshes, params = scenes.directivity_lens(ray_count=30_000)
    tr = it.CL_Tracer()
    res1 = [tuple(np.array(a) for a in r) for r in tr.iterative_tracer(ls, meshes, **params)]
    assert set(tr.stats["tc_warps"]) == {0}, tr.stats["tc_warps"]
    assert tr._lib.lpcl_search_launch_count(tr._scene) == 3  # scratch initialisation, sweep, second pass
    for variant in (8, 16):
        monkeypatch.setattr(it.CL_Tracer, "SEARCH_VARIANT", variant)
        tr2 = it.CL_Tracer()
        res2 = tr2.iterative_tracer(ls, meshes, **params)
        assert set(tr2.stats["tc_warps"]) == {variant}
        assert tr2._lib.lpcl_search_launch_count(tr2._scene) == 2  # no second pass
        util.assert_results_equal(list(res2), res1, "epilogue variant %d" % variant)
    rc = tr2._lib.lpcl_scene_set_tc_warps(tr2._scene, 12)
    assert rc != 0 and b"warps" in tr2._lib.lpcl_last_error()


def test_streamed_upload_of_generation_0_keeps_results(monkeypatch):
    """Large pageable ray sets are uploaded slice by slice under generation 0's search (LPCL_STREAM_UPLOAD, default on):
    same `results`, same measured rays, and a second run() on the now-resident rays gives them again."""
    from lightpycl_b200 import iterative_tracer as it, scenes

    ls, meshes, params = scenes.directivity_lens(ray_count=300_000)
    params = dict(params, trace_iterations=4)
    tr = it.CL_Tracer()
    assert tr.stream_upload and 300_000 >= tr.STREAM_MIN_RAYS
    res = [tuple(np.array(a) for a in r) for r in tr.iterative_tracer(ls, meshes, **params)]
    assert tr.stats["tiles"][0] == 2 and tr._job["resident"]
    pos, pwr = (np.array(a) for a in tr.get_measured_rays())
    again = tr.run(**params)  # rays resident: one launch for generation 0
    assert tr.stats["tiles"][0] == 1
    util.assert_results_equal(list(again), res, "second run on resident rays")
    monkeypatch.setenv("LPCL_STREAM_UPLOAD", "0")
    tr0 = it.CL_Tracer()
    assert not tr0.stream_upload
    res0 = tr0.iterative_tracer(ls, meshes, **params)
    assert tr0.stats["tiles"][0] == 1
    util.assert_results_equal(res, list(res0), "streamed vs one-piece upload")
    pos0, pwr0 = tr0.get_measured_rays()
    assert np.array_equal(util.bits(pos), util.bits(np.asarray(pos0))) and np.array_equal(util.bits(pwr), util.bits(np.asarray(pwr0)))
