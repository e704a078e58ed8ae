"""Known-answer tests that follow from the reference kernel formulas:
the reference ships no tests, so these pin the physics of the oracle (and, via the parity tests,
of the CUDA path)."""

import ctypes as C

import numpy as np

import util
from oracle import oracle


def _split(d, normal, n1, n2, power=1.0):
    lib = oracle._lib("port")
    f3 = lambda v: np.ascontiguousarray(v, dtype=np.float32)
    at, d, nrm = f3([0, 0, 0]), f3(d), f3(normal)
    rd, td = np.zeros(3, np.float32), np.zeros(3, np.float32)
    rp, tp, rm, tm = C.c_float(), C.c_float(), C.c_int(), C.c_int()
    fp = C.POINTER(C.c_float)
    rc = lib.orc_fresnel_split(at.ctypes.data_as(fp), d.ctypes.data_as(fp), C.c_float(power), nrm.ctypes.data_as(fp),
                               C.c_float(n1), C.c_float(n2), rd.ctypes.data_as(fp), C.byref(rp), C.byref(rm),
                               td.ctypes.data_as(fp), C.byref(tp), C.byref(tm))
    return rc, rd, rp.value, rm.value, td, tp.value, tm.value


def test_normal_incidence_fresnel():
    # R = ((n1-n2)/(n1+n2))^2  (cl:313-316)
    for n1, n2 in ((1.0, 1.5), (1.5, 1.0), (1.0, 2.5), (1.3374, 1.42)):
        rc, rd, rp, rm, td, tp, tm = _split([0, 0, 1], [0, 0, -1], n1, n2)
        want = ((n1 - n2) / (n1 + n2)) ** 2
        assert rc == 0 and rm == 0 and tm == 0
        assert abs(rp - want) < 1e-6 and abs(tp - (1 - want)) < 1e-6
        assert np.allclose(rd, [0, 0, -1], atol=1e-7) and np.allclose(td, [0, 0, 1], atol=1e-6)


def test_normal_flip_is_symmetric():
    a = _split([0.3, 0, 0.9539392], [0, 0, -1], 1.0, 1.5)
    b = _split([0.3, 0, 0.9539392], [0, 0, 1], 1.0, 1.5)  # cl:304-307 flips the normal to face the ray
    assert a[2] == b[2] and np.array_equal(a[1], b[1]) and np.array_equal(a[4], b[4])


def test_total_internal_reflection():
    # beyond the critical angle asin(1/1.5) = 41.8 deg all power is reflected, refract child dead (cl:329-341)
    ang = np.deg2rad(60.0)
    rc, rd, rp, rm, td, tp, tm = _split([np.sin(ang), 0, np.cos(ang)], [0, 0, -1], 1.5, 1.0, power=2.0)
    assert rc == 0 and rp == 2.0 and rm == 0 and tm == -1 and tp == 0.0 and np.all(td == 0)
    assert np.allclose(rd, [np.sin(ang), 0, -np.cos(ang)], atol=1e-6)


def test_snell_and_energy():
    ang = np.deg2rad(35.0)
    rc, rd, rp, rm, td, tp, tm = _split([np.sin(ang), 0, np.cos(ang)], [0, 0, -1], 1.0, 1.5)
    assert abs(np.linalg.norm(td) - 1) < 1e-6 and abs(np.linalg.norm(rd) - 1) < 1e-6
    assert abs(td[0] / np.linalg.norm(td) - np.sin(ang) / 1.5) < 1e-6  # n1 sin(t1) = n2 sin(t2)
    assert abs(rp + tp - 1.0) < 1e-7


def test_negative_ior_gives_negative_refracted_power():
    # Rs, Rp > 1 for n2 < 0 (example_nested_cubes_refraction.py uses IOR -2)
    rc, rd, rp, rm, td, tp, tm = _split([0.2, 0, 0.9797959], [0, 0, -1], 1.0, -2.0)
    assert rp > 1.0 and tp < 0.0 and rm == 0 and tm == 0


def test_ray_triangle_basics():
    V0, V1, V2 = [0, 0, 5], [1, 0, 5], [0, 1, 5]
    assert oracle.ray_triangle([0.25, 0.25, 0], [0, 0, 1], V0, V1, V2)[:2] == (1, 5.0)
    assert oracle.ray_triangle([0.25, 0.25, 0], [0, 0, -1], V0, V1, V2)[:2] == (1, -5.0)  # no sign test on t (cl:98)
    assert oracle.ray_triangle([0.25, 0.25, 0], [0, 0, 1], V0, V2, V1)[0] == 1  # no back-face culling
    assert oracle.ray_triangle([0.75, 0.75, 0], [0, 0, 1], V0, V1, V2)[0] == 0
    assert oracle.ray_triangle([0.25, 0.25, 0], [1, 0, 0], V0, V1, V2)[0] == 0  # parallel: |DEN| < 1e-6
    # Q7: the DEN cut-off is absolute - a 1e-3 x 1e-3 triangle (DEN = 1e-6 * 0.5) is invisible
    s = 7e-4
    assert oracle.ray_triangle([1e-4, 1e-4, 0], [0, 0, 1], [0, 0, 5], [s, 0, 5], [0, s, 5])[0] == 0


def test_beer_lambert_and_measure_flags():
    z = np.load(util.GOLDEN_DIR + "/mini_lossy.npz")
    ls, meshes = util.arrays_to_scene(z)
    out = oracle.trace(ls, meshes, trace_iterations=3, max_ray_len=np.float32(2e3))
    # iteration 1 holds the rays that entered the lossy block (mesh 1, alpha = 0.08): P' = P exp(-alpha L)
    o, d, p, m = out["results"][1]
    dbg = out["debug"][1]
    inside = np.where(dbg["n1"] == 1)[0]
    assert len(inside) > 0
    parents0 = out["results"][0]
    # children of iteration 0 are [reflected..., refracted...]; recompute their pre-loss power independently
    L = np.linalg.norm((d[inside, :3] - o[inside, :3]).astype(np.float64), axis=1)
    assert np.all(L > 0)
    # power after the step must be strictly below the Fresnel-transmitted power by exp(-0.08 L)
    alive0 = np.concatenate([np.zeros(0)])  # (structure check only; exact value checked below)
    ch = oracle.reflect_refract(out["scene"], parents0[0], parents0[1], out["debug"][0]["dir"],
                                np.ascontiguousarray(parents0[2].reshape(-1).copy()),
                                np.zeros(len(parents0[0]), np.int32), out["debug"][0]["n1"], out["debug"][0]["n2"],
                                out["debug"][0]["hit_mesh"], out["debug"][0]["hit_tri"], np.float32(1.0),
                                np.float32(2e3))
    pre = np.concatenate((ch["refl_power"], ch["refr_power"]))[np.concatenate((ch["refl_meas"], ch["refr_meas"])) == 0]
    want = pre[inside].astype(np.float64) * np.exp(-np.float64(np.float32(0.08)) * L)
    assert np.allclose(p[inside], want, rtol=2e-6)
    # flags: measured rays ended on mesh 0 (type 3), terminated ones hit nothing
    for (o, d, p, m), dbg in zip(out["results"], out["debug"]):
        assert np.all(dbg["hit_mesh"][m == 1] == 0)
        assert np.all(dbg["hit_mesh"][m == -1] == -1)
        assert np.all(dbg["hit_mesh"][(m == 0)] >= 1)


def test_parabola_collimates_and_conserves_energy():
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.parabolic_mirror(ray_count=300, profile_pts=60, ang_pts=37)
    out = oracle.trace(ls, meshes, **params)
    o, d, p, m = out["results"][1]  # reflected rays: children of the mirror hits
    dirs = (d[:, :3] - o[:, :3]).astype(np.float64)
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    came_from_mirror = out["debug"][1]["prev_mesh"] == 1
    assert came_from_mirror.sum() > 100
    # a source at the focus is collimated along +z up to the facet angle of the coarse mesh
    assert np.all(dirs[came_from_mirror, 2] > 0.995)
    # bare-source energy conservation (reference example_directivity_sphericalMeasSurf.py:105-109)
    from lightpycl_b200 import geo_optical_elements as goe, light_source as lsm
    np.random.seed(3)
    src = lsm.light_source(power=1000.0, ray_count=500)
    dome = goe.optical_elements().hemisphere(center=[0, 0, 0, 0], radius=500.0)
    dome.setMaterial(mat_type="measure")
    out = oracle.trace([src], [dome], trace_iterations=4, max_ray_len=np.float32(1e3))
    pos, pwr = oracle.measured_rays(out["results"])
    # Q8: a sliver of the revolved dome is double-covered, those rays count as "exiting" -> also measured
    assert abs(float(np.sum(pwr, dtype=np.float64)) - 1000.0) < 1.0


def test_nested_cubes_window_logic():
    """SURVEY.md worked example: a +z ray inside m1 sees m1-top, m0-top (exiting) and m2-bottom
    (entering) inside one EPSILON window -> direct m1 -> m2 transition, hit mesh reported as m1."""
    from lightpycl_b200 import scenes

    _, meshes, params = scenes.nested_cubes()
    sc = oracle.Scene(meshes)
    origin = np.array([[-20, 0, 22.0, 0]], dtype=np.float32)
    direc = np.array([[0, 0, 1, 0]], dtype=np.float32)
    nt, h, ni = oracle.intersect(sc, origin, direc, params["max_ray_len"])
    pp = oracle.postproc(sc, origin, direc, np.array([1], np.int32), nt, h, ni, params["max_ray_len"])
    assert pp["hit_mesh"][0] == 1 and pp["n1"][0] == 1 and pp["n2"][0] == 2
    assert abs(pp["dest"][0, 2] - 25.0) < 1e-3
