"""End-to-end parity of the production path (CL_Tracer -> C ABI -> fused sm_100a kernels) against
the oracle: `results` bit-identical per iteration (same ray ORDER as the reference), hit
mesh/triangle/n1/n2 identical, measured (pos, pwr) identical; plus size-independent properties at
larger ray counts."""

import glob
import os

import numpy as np
import pytest

import util
from oracle import oracle

pytestmark = pytest.mark.gpu


def _trace_gpu(ls, meshes, params, **kw):
    from lightpycl_b200 import iterative_tracer as it

    tr = it.CL_Tracer(platform_name="NVIDIA", device_name="B200", record_hits=True, **kw)
    tr.iterative_tracer(light_source=ls, meshes=meshes, **params)
    return tr


def _check(name, ls, meshes, params, **kw):
    ref = oracle.trace(ls, meshes, backend="reference" if oracle.available("reference") else "port",
                       power_sum="f64", **params)
    tr = _trace_gpu(ls, meshes, params, **kw)
    util.assert_results_equal(list(tr.results), ref["results"], name)
    assert tr.results.ray_counts == [len(r[0]) for r in ref["results"]]
    for k, dbg in enumerate(ref["debug"]):
        hit = tr.hit_records[k].cpu().numpy()
        for col, key in enumerate(("hit_mesh", "hit_tri", "n1", "n2")):
            assert np.array_equal(hit[:, col], dbg[key]), (name, k, key)
    pos, pwr = tr.get_measured_rays()
    rpos, rpwr = oracle.measured_rays(ref["results"])
    if rpwr is not None and np.asarray(rpwr).size:
        assert np.asarray(pwr).shape == np.asarray(rpwr).shape
        assert np.array_equal(util.bits(pos[:, :3]), util.bits(rpos[:, :3]))
        assert np.array_equal(util.bits(np.asarray(pwr)), util.bits(np.asarray(rpwr)))
        total, rtotal = np.sum(np.asarray(pwr, np.float64)), np.sum(np.asarray(rpwr, np.float64))
        assert abs(total - rtotal) <= 1e-6 * abs(rtotal)
        assert abs(sum(tr.stats["measured"]) - np.asarray(rpwr).size) == 0
    else:
        assert len(np.asarray(pwr).reshape(-1)) == 0
    assert tr.stats["live"] == ref["live"]
    return tr, ref


FIXTURES = sorted(f for f in glob.glob(os.path.join(util.GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(f).startswith("refhost_"))  # refhost_*: tests/test_refhost_golden.py


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_trace_matches_golden_fixture(path):
    """GPU against the committed outputs of the reference kernels (same frozen inputs)."""
    z = np.load(path)
    ls, meshes = util.arrays_to_scene(z)
    params = dict(trace_iterations=int(z["param_trace_iterations"]),
                  trace_until_dissipated=float(z["param_trace_until_dissipated"]),
                  max_ray_len=np.float32(z["param_max_ray_len"]), ior_env=np.float32(z["param_ior_env"]))
    tr = _trace_gpu(ls, meshes, params)
    n_iter = int(z["n_iter"])
    want = [(z["it%d_origin" % k], z["it%d_dest" % k], z["it%d_power" % k], z["it%d_meas" % k]) for k in range(n_iter)]
    util.assert_results_equal(list(tr.results), want, os.path.basename(path))
    for k in range(n_iter):
        hit = tr.hit_records[k].cpu().numpy()
        for col, key in enumerate(("hit_mesh", "hit_tri", "n1", "n2")):
            assert np.array_equal(hit[:, col], z["it%d_%s" % (k, key)]), (k, key)
    pos, pwr = tr.get_measured_rays()
    assert np.array_equal(util.bits(pos[:, :3]), util.bits(z["measured_pos"][:, :3]))
    assert np.array_equal(util.bits(np.asarray(pwr).reshape(-1)), util.bits(z["measured_pwr"]))


@pytest.mark.parametrize("name,kw", [
    ("c1_nested_cubes", dict(ray_count=3000)),
    ("c2_directivity_lens", dict(ray_count=4000)),
    ("c3_parabolic_mirror", dict(ray_count=3000)),
    ("c4_human_eye", dict(ray_count=200)),
    ("dissipative_cube", dict(ray_count=3000)),
    ("c5_lenslet_array", dict(ray_count=150, sources=4, grid=2, ang_pts=17)),
    ("c5_lenslet_array", dict(ray_count=100, sources=4, grid=2, ang_pts=9, per_lenslet_mesh=True)),
])
def test_trace_matches_oracle_on_configs(name, kw):
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.CONFIGS[name](**kw)
    _check(name, ls, meshes, params)


def test_trace_with_lineage_and_without_results():
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.directivity_lens(ray_count=3000)
    tr, ref = _check("lens+lineage", ls, meshes, params, lineage=True)
    # lineage order == reference order: every generation's lineage must be strictly increasing
    for gen, _dm, _ps in tr.results._dev:
        lin = gen.logical(gen.lineage).cpu().numpy()
        assert np.all(np.diff(lin) > 0)
    lean = _trace_gpu(ls, meshes, params, keep_results=False)
    a, b = tr.get_measured_rays(), lean.get_measured_rays()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and len(lean.results) == 0
    assert lean.results.ray_counts == tr.results.ray_counts


def test_determinism_and_shard_union_at_scale():
    """Size-independent properties at BASELINE's full C2 size, 1M rays x 19,006 triangles (far too big
    for the CPU oracle in a test): (1) two runs are bit-identical; (2) linearity - tracing the two halves of the rays
    separately yields the same measured records as the full trace (rays are independent);
    (3) power bookkeeping: measured + live + lost <= input, measured power reproduced by stats."""
    from lightpycl_b200 import scenes

    n_full = 1_000_000  # BASELINE's full C2 size: 1M rays x 19,006 triangles, depth 10
    ls, meshes, params = scenes.directivity_lens(ray_count=n_full)
    a = _trace_gpu(ls, meshes, params, lineage=True)
    b = _trace_gpu(ls, meshes, params, lineage=True)
    pa, pb = a.get_measured_rays(), b.get_measured_rays()
    assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
    assert a.results.ray_counts == b.results.ray_counts

    src = ls[0]
    halves = []
    for lo, hi in ((0, n_full // 2), (n_full // 2, n_full)):
        part = util.ArraySource(src.rays_origin[lo:hi], src.rays_dir[lo:hi], src.rays_power[lo:hi])
        # same number of iterations as the full run so that the same generations are compared
        p = dict(params, trace_iterations=len(a.results), trace_until_dissipated=1.0)
        halves.append(_trace_gpu([part], meshes, p).get_measured_rays())
    full = _trace_gpu(ls, meshes, dict(params, trace_iterations=len(a.results), trace_until_dissipated=1.0))
    fpos, fpwr = full.get_measured_rays()
    upos = np.concatenate([h[0] for h in halves])
    upwr = np.concatenate([np.asarray(h[1]).reshape(-1) for h in halves])
    key = lambda pos, pwr: np.sort(np.ascontiguousarray(np.concatenate((pos[:, :3], pwr.reshape(-1, 1)), axis=1))
                                   .view([("", np.float32)] * 4).reshape(-1), axis=0)
    assert np.array_equal(key(fpos, np.asarray(fpwr).reshape(-1)), key(upos, upwr))
    measured = float(np.sum(np.asarray(pa[1], np.float64)))
    assert 0.3 * a.input_power < measured <= a.input_power * (1 + 1e-5)


@pytest.mark.parametrize("name,kw", [
    ("c5_lenslet_array", dict(ray_count=60, sources=4, grid=3, ang_pts=9, per_lenslet_mesh=True)),  # 11 meshes > 8
    ("c5_lenslet_array", dict(ray_count=40, sources=1, grid=4, ang_pts=5, per_lenslet_mesh=True)),  # 18 meshes
])
def test_many_meshes_take_the_generic_resolve_path(name, kw):
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.CONFIGS[name](**kw)
    assert len(meshes) > 8
    _check(name, ls, meshes, dict(params, trace_iterations=6))


def test_edge_cases_single_ray_all_miss_and_tiny_scene():
    from lightpycl_b200 import geo_optical_elements as goe, light_source as lsm

    oe = goe.optical_elements()
    plate = oe.cube(center=(0, 0, 10, 0), size=[2, 2, 1, 0])
    plate.setMaterial(mat_type="measure")
    params = dict(trace_iterations=5, trace_until_dissipated=0.99, max_ray_len=np.float32(100.0), ior_env=np.float32(1.0))
    # one ray straight at the plate
    one = util.ArraySource(np.array([[0.1, 0.2, 0, 0]], np.float32), np.array([[0, 0, 1, 0]], np.float32),
                           np.array([2.5], np.float32))
    tr, ref = _check("one ray", [one], [plate], params)
    assert tr.results.ray_counts == [1] and tr.stats["measured"] == [1]
    # every ray misses everything: one generation, all terminated, nothing measured
    np.random.seed(2)
    away = lsm.light_source(center=np.array([0, 0, 0, 0], np.float32), direction=(0, 0, -1), ray_count=777, power=3.0)
    tr, ref = _check("all miss", [away], [plate], params)
    assert tr.results.ray_counts == [777] and tr.stats["terminated"] == [777] and tr.stats["live"] == [0]
    pos, pwr = tr.get_measured_rays()
    assert pos.shape == (0, 4) and np.asarray(pwr).size == 0
    # rays exactly along a face of a cube (parallel, DEN == 0) and through an edge shared by two triangles
    box = oe.cube(center=(0, 0, 5, 0), size=[2, 2, 2, 0])
    box.setMaterial(mat_type="refractive", IOR=1.5)
    o = np.array([[1.0, 0.0, 0, 0], [0.0, 0.0, 0, 0], [0.5, 0.5, 0, 0], [-1.0, -1.0, 0, 0]], np.float32)
    d = np.tile(np.array([[0, 0, 1, 0]], np.float32), (4, 1))
    graze = util.ArraySource(o, d, np.ones(4, np.float32))
    _check("grazing", [graze], [plate, box], params)


def test_single_iteration_keeps_the_sources_power_shape():
    """Reference quirk (it:404-409): with a single iteration get_measured_rays returns
    the power in the source's shape, (K,1) for random_rays; later iterations flatten it."""
    from lightpycl_b200 import geo_optical_elements as goe, light_source as lsm

    np.random.seed(3)
    src = lsm.light_source(power=1000.0, ray_count=2000)
    dome = goe.optical_elements().hemisphere(center=[0, 0, 0, 0], radius=500.0)
    dome.setMaterial(mat_type="measure")
    params = dict(trace_iterations=4, trace_until_dissipated=0.99, max_ray_len=np.float32(1e3), ior_env=np.float32(1.0))
    tr, ref = _check("bare source", [src], [dome], params)
    pos, pwr = tr.get_measured_rays()
    assert len(tr.results) == 1 and np.asarray(pwr).ndim == 2 and np.asarray(pwr).shape[1] == 1
    assert abs(float(np.sum(np.asarray(pwr, np.float64))) - 1000.0) < 1.0  # energy conservation (ex_sph:105-109)


def test_full_size_mirror_power_bookkeeping():
    """C3 at full size (1M rays x 100,552 triangles): every source ray hits the mirror, its child is
    either measured on the dome or absorbed by the terminator plate, so
    measured power = 0.98 x (power of the rays that miss the plate) and the counters add up."""
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.parabolic_mirror(ray_count=1_000_000)
    tr = _trace_gpu(ls, meshes, params, keep_results=False)
    assert tr.stats["rays"][0] == 1_000_000 and len(tr.stats["rays"]) == 2
    gen1 = tr.stats["rays"][1]
    assert 0.99999 * 1_000_000 <= gen1 <= 1_000_000          # (a handful of rays slip through facet seams)
    assert tr.stats["measured"][1] + tr.stats["terminated"][1] == gen1 and tr.stats["live"][1] == 0
    pos, pwr = tr.get_measured_rays()
    w = np.asarray(pwr, np.float64).reshape(-1)
    assert len(w) == sum(tr.stats["measured"])
    # collimated beam: measured end points lie on the r = 500 dome, within the mirror's 10-unit radius plus the
    # spread of the flat facets (113 wedges => normals off by up to 1.6 deg => ~3 deg of divergence over 500 units)
    assert np.all(np.abs(np.linalg.norm(pos[:, :3], axis=1) - 500.0) < 0.5)
    assert np.all(np.hypot(pos[:, 0], pos[:, 1]) < 45.0)
    blocked = tr.stats["terminated"][1] / gen1               # fraction of the rays the 4x4 plate shadows
    assert 0.03 < blocked < 0.15                              # >= 16 / (pi 10^2) = 5.1 %: the beam is denser near the axis
    # mirror loss 2 % (reflectivity 0.98 set by attribute); the plate's share from the ideal parabola
    assert abs(w.sum() / tr.input_power - 0.98 * (1 - _blocked_power(tr, ls))) < 3e-2  # facets blur the plate's shadow


def _blocked_power(tr, ls):
    # power fraction of the source rays whose reflected ray passes through the plate |x|,|y| <= 2 at z ~ 100:
    # for a parabola with the source at the focus the reflected ray keeps the (x, y) of the mirror hit point
    d = ls[0].rays_dir[:, :3].astype(np.float64)
    p = ls[0].rays_power.reshape(-1).astype(np.float64)
    f = 5.0
    t = 2.0 * f / (1.0 - d[:, 2])                             # |focus -> paraboloid| along a unit direction (axis +z, opening +z)
    x, y = t * d[:, 0], t * d[:, 1]
    return p[(np.abs(x) <= 2.0) & (np.abs(y) <= 2.0)].sum() / p.sum()


def test_empty_source():
    """Zero rays (an 'empty input'): one empty iteration, then the power test ends the trace."""
    from lightpycl_b200 import geo_optical_elements as goe

    plate = goe.optical_elements().cube(center=(0, 0, 10, 0), size=[2, 2, 1, 0])
    plate.setMaterial(mat_type="measure")
    empty = util.ArraySource(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), np.zeros((0,), np.float32))
    tr = _trace_gpu([empty], [plate], dict(trace_iterations=3, max_ray_len=np.float32(100.0)))
    assert tr.results.ray_counts == [0] and len(tr.results) == 1
    o, d, p, m = tr.results[0]
    assert o.shape == (0, 4) and d.shape == (0, 4) and p.shape == (0,) and m.shape == (0,)
    pos, pwr = tr.get_measured_rays()
    assert pos.shape == (0, 4) and np.asarray(pwr).size == 0


def test_ray_tiles_give_bit_identical_results(monkeypatch):
    """HBM-bounded ray tiling (the reference's <= 50,000-ray partitions, it:218-273): with a tiny
    scratch budget every generation of C2 at 100k rays is traced in several ray tiles that reuse one
    scratch buffer; `results`, hit records and measured rays must equal the untiled run bit for bit."""
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.directivity_lens(ray_count=100_000)
    whole = _trace_gpu(ls, meshes, params, lineage=True)
    assert max(whole.stats["tiles"]) == 1
    tiled = _trace_gpu(ls, meshes, params, lineage=True, ray_tile_bytes=1)  # -> the 65,536-ray minimum tile
    assert min(t for t, n in zip(tiled.stats["tiles"], tiled.stats["rays"]) if n > 65536) >= 2
    assert tiled.results.ray_counts == whole.results.ray_counts and tiled.stats["live"] == whole.stats["live"]
    util.assert_results_equal(list(tiled.results), list(whole.results), "tiled vs whole")
    for a, b in zip(tiled.hit_records, whole.hit_records):
        assert bool((a == b).all())
    for (ga, _d, _p), (gb, _d2, _p2) in zip(tiled.results._dev, whole.results._dev):
        assert bool((ga.logical(ga.lineage) == gb.logical(gb.lineage)).all())
        import torch

        assert bool((ga.logical(ga.dir_prev).view(torch.int32) == gb.logical(gb.dir_prev).view(torch.int32)).all())
    pa, pb = tiled.get_measured_rays(), whole.get_measured_rays()
    assert np.array_equal(util.bits(pa[0]), util.bits(pb[0])) and np.array_equal(util.bits(np.asarray(pa[1])),
                                                                                  util.bits(np.asarray(pb[1])))
    # the budget can also come from the environment, and without retained results the end points are per-tile scratch
    monkeypatch.setenv("LPCL_RAY_TILE_BYTES", "4000000")
    lean = _trace_gpu(ls, meshes, params, keep_results=False)
    assert max(lean.stats["tiles"]) >= 2
    pc = lean.get_measured_rays()
    assert np.array_equal(util.bits(pc[0]), util.bits(pb[0])) and np.array_equal(util.bits(np.asarray(pc[1])),
                                                                                  util.bits(np.asarray(pb[1])))


def test_deep_trace_lineage_epochs():
    """40 generations (the reference allows 100 by default, it:77): past the 31 branch bits of a 64-bit
    lineage id the tracer starts a new lineage epoch; results stay bit-identical to the oracle and
    every generation's ids stay strictly increasing (= the reference order) and equal the oracle's."""
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.glass_cube_cavity(ray_count=500, trace_iterations=40)
    tr, ref = _check("cavity", ls, meshes, params, lineage=True)
    assert len(tr.results) == 40 and min(tr.stats["measured"][1:]) > 0
    want = oracle.trace(ls, meshes, power_sum="f64", lineage0=np.arange(500), **params)["lineage"]
    for k, (gen, _dm, _ps) in enumerate(tr.results._dev):
        lin = gen.logical(gen.lineage).cpu().numpy()
        assert np.all(np.diff(lin) > 0), k
        assert np.array_equal(lin, want[k]), k


def test_held_results_survive_later_traces():
    """arrays returned by get_measured_rays / tr.geometry must never be overwritten by a later
    call that recycles host buffers."""
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.directivity_lens(ray_count=5000)
    tr = _trace_gpu(ls, meshes, params)
    pos, pwr = tr.get_measured_rays()
    geo0 = tr.geometry[0]
    xyz, pw_copy, geo_copy = pos[:, :3], np.array(pwr, copy=True), np.array(geo0, copy=True)
    xyz_copy = np.array(xyz, copy=True)
    del pos
    ls2, meshes2, params2 = scenes.parabolic_mirror(ray_count=5000, profile_pts=40, ang_pts=17)
    for _ in range(5):
        tr.iterative_tracer(ls2, meshes2, **params2)
        tr.get_measured_rays()
    assert np.array_equal(xyz, xyz_copy) and np.array_equal(pwr, pw_copy) and np.array_equal(geo0, geo_copy)


@pytest.mark.parametrize("name,kw", [
    ("c1_nested_cubes", dict(ray_count=10)),       # the reference example verbatim: 10 rays x 48 triangles
    ("c1_nested_cubes", dict(ray_count=700)),      # grows past the one-CTA limit: handed over to the general loop
    ("glass_box_in_box", dict(ray_count=40, trace_iterations=40)),  # 31 generations in one launch, then the general loop
    ("dissipative_cube", dict(ray_count=5)),
])
def test_one_launch_trace_of_small_scenes(name, kw, monkeypatch):
    """Small scenes run every generation in ONE kernel launch (csrc/lpcl_small.cuh).  Bit-identical to the oracle,
    to the general loop (LPCL_SMALL_PATH=0), and across the hand-over from one to the other."""
    from lightpycl_b200 import geo_optical_elements as goe, iterative_tracer as it, scenes

    if name == "glass_box_in_box":  # glass_cube_cavity with a 12-triangle measure box instead of the 10,366-triangle sphere
        ls, meshes, params = scenes.glass_cube_cavity(**kw)
        box = goe.optical_elements().cube(center=(0, 0, 0, 0), size=[200, 200, 200, 0])
        box.setMaterial(mat_type="measure")
        meshes = [box, meshes[1]]
    else:
        ls, meshes, params = scenes.CONFIGS[name](**kw)
    if name == "dissipative_cube":  # 10,378 triangles: only a handful of rays fit the pair limit
        assert sum(s.ray_count for s in ls) * 10378 <= it.CL_Tracer.SMALL_PAIR_LIMIT
    tr, ref = _check(name + " one launch", ls, meshes, params, lineage=True)
    assert tr.stats.get("one_launch_generations", 0) >= 1
    if name == "c1_nested_cubes" and kw["ray_count"] == 10:
        assert tr.stats["one_launch_generations"] == len(tr.results) == 4 and tr.kernel_launches == 2
    if kw.get("ray_count") == 700 or name == "glass_box_in_box":
        assert tr.stats["one_launch_generations"] < len(tr.results)  # the general loop finished the trace
    monkeypatch.setenv("LPCL_SMALL_PATH", "0")
    general = _trace_gpu(ls, meshes, params, lineage=True)
    assert "one_launch_generations" not in general.stats
    util.assert_results_equal(list(tr.results), list(general.results), name)
    for a, b in zip(tr.hit_records, general.hit_records):
        assert bool((a == b).all())
    for (ga, _d, _p), (gb, _d2, _p2) in zip(tr.results._dev, general.results._dev):
        assert bool((ga.logical(ga.lineage) == gb.logical(gb.lineage)).all())
    for k in ("rays", "live", "measured", "terminated"):
        assert tr.stats[k] == general.stats[k], k
    pa, pb = tr.get_measured_rays(), general.get_measured_rays()
    assert np.array_equal(util.bits(pa[0]), util.bits(pb[0]))
    assert np.array_equal(util.bits(np.asarray(pa[1]).reshape(-1)), util.bits(np.asarray(pb[1]).reshape(-1)))
    monkeypatch.delenv("LPCL_SMALL_PATH")
    monkeypatch.setattr(it.CL_Tracer, "SMALL_PAIR_LIMIT", 48 * 12)  # hand over after the first generation(s)
    if name == "c1_nested_cubes" and kw["ray_count"] == 10:
        early = _trace_gpu(ls, meshes, params, lineage=True)
        assert 1 <= early.stats["one_launch_generations"] < 4
        util.assert_results_equal(list(early.results), list(general.results), name + " early hand-over")


@pytest.mark.parametrize("per_ray", [0, 1])
def test_record_budget_is_a_matter_of_speed_only(per_ray):
    """The deferred-flag search hands flagged (ray, chunk) records to its second pass through a buffer in the workspace,
    LPCL_REC_PER_RAY records per ray (read once per process, hence the subprocess); records that do not fit are evaluated
    inside the search kernel.  With no buffer at all (0) and with one that overflows (1; C2 flags ~4 chunks per ray) the
    measured rays are the ones the default budget gives."""
    import hashlib
    import subprocess
    import sys

    code = ("import hashlib, numpy as np\n"
            "from lightpycl_b200 import iterative_tracer as it, scenes\n"
            "ls, meshes, params = scenes.directivity_lens(ray_count=20000)\n"
            "tr = it.CL_Tracer()\n"
            "tr.iterative_tracer(ls, meshes, **params)\n"
            "pos, pwr = tr.get_measured_rays()\n"
            "print('SIG', hashlib.sha1(np.ascontiguousarray(pos).tobytes() + np.ascontiguousarray(pwr).tobytes()).hexdigest(),"
            " sum(tr.stats['candidates']))\n")
    sigs = []
    for env_value in (None, str(per_ray)):
        env = {k: v for k, v in os.environ.items() if k != "LPCL_REC_PER_RAY"}
        if env_value is not None:
            env["LPCL_REC_PER_RAY"] = env_value
        out = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        line = [l for l in out.stdout.splitlines() if l.startswith("SIG ")][-1].split()
        sigs.append(line[1])
        assert int(line[2]) > 20000  # more flagged chunks than a budget of one per ray holds
    assert sigs[0] == sigs[1]
