"""Pin the oracle: (1) the C restatement (trace_oracle.c) against every golden fixture generated
from the reference kernels; (2) where oracle/_ref exists (the unmodified reference .cl compiled by
path), the restatement against it live, kernel by kernel and over whole traces."""

import glob
import os

import numpy as np
import pytest

import util
from oracle import oracle

FIXTURES = sorted(f for f in glob.glob(os.path.join(util.GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(f).startswith("refhost_"))  # refhost_*: tests/test_refhost_golden.py


def _params(z):
    return dict(trace_iterations=int(z["param_trace_iterations"]),
                trace_until_dissipated=float(z["param_trace_until_dissipated"]),
                max_ray_len=np.float32(z["param_max_ray_len"]), ior_env=np.float32(z["param_ior_env"]))


def test_fixtures_exist():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
@pytest.mark.parametrize("backend", ["port", "reference"])
def test_oracle_matches_golden(path, backend):
    if not oracle.available(backend) and backend == "reference":
        pytest.skip("oracle/_ref not present on this box")
    z = np.load(path)
    ls, meshes = util.arrays_to_scene(z)
    out = oracle.trace(ls, meshes, backend=backend, **_params(z))
    n_iter = int(z["n_iter"])
    want = [(z["it%d_origin" % k], z["it%d_dest" % k], z["it%d_power" % k], z["it%d_meas" % k]) for k in range(n_iter)]
    util.assert_results_equal(out["results"], want, os.path.basename(path))
    for k in range(n_iter):
        for key in ("hit_mesh", "hit_tri", "n1", "n2", "hits", "near_tri"):
            assert np.array_equal(out["debug"][k][key], z["it%d_%s" % (k, key)]), (k, key)
        assert np.array_equal(util.bits(out["debug"][k]["near_t"]), util.bits(z["it%d_near_t" % k]))
    pos, pwr = oracle.measured_rays(out["results"])
    if len(z["measured_pwr"]):
        assert np.array_equal(util.bits(pos[:, :3]), util.bits(z["measured_pos"][:, :3]))
        assert np.array_equal(util.bits(np.asarray(pwr).reshape(-1)), util.bits(z["measured_pwr"]))


@pytest.mark.skipif(not oracle.available("reference"), reason="oracle/_ref not present on this box")
@pytest.mark.parametrize("name", ["c1", "c2", "c3", "dissipative"])
def test_port_matches_reference_live(name):
    from lightpycl_b200 import scenes

    builder, kw = {"c1": (scenes.nested_cubes, dict(ray_count=500)),
                   "c2": (scenes.directivity_lens, dict(ray_count=300)),
                   "c3": (scenes.parabolic_mirror, dict(ray_count=200)),
                   "dissipative": (scenes.dissipative_cube, dict(ray_count=500))}[name]
    ls, meshes, params = builder(**kw)
    a = oracle.trace(ls, meshes, backend="reference", **params)
    b = oracle.trace(ls, meshes, backend="port", **params)
    util.assert_results_equal(b["results"], a["results"], name)
    for da, db in zip(a["debug"], b["debug"]):
        for key in ("hit_mesh", "hit_tri", "n1", "n2", "hits", "near_tri", "entering"):
            assert np.array_equal(da[key], db[key]), key


@pytest.mark.skipif(not oracle.available("reference"), reason="oracle/_ref not present on this box")
def test_ray_triangle_probe_random():
    rng = np.random.RandomState(5)
    n_hit = 0
    for _ in range(20000):
        V = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
        O = rng.uniform(-3, 3, 3).astype(np.float32)
        # aim near the triangle so that hits, edge cases and misses all occur
        w = rng.dirichlet([1, 1, 1]) * rng.choice([1.0, 1.0, 1.02])
        D = (w @ V - O).astype(np.float32)
        a = oracle.ray_triangle(O, D, V[0], V[1], V[2], "reference")
        b = oracle.ray_triangle(O, D, V[0], V[1], V[2], "port")
        assert a[0] == b[0]
        if a[0]:
            n_hit += 1
            assert np.float32(a[1]).tobytes() == np.float32(b[1]).tobytes()
    assert 2000 < n_hit < 19999


def test_power_sum_rules_agree_on_fixtures():
    """The GPU path breaks on a float64 power sum, the reference on a sorted sequential float32
    sum (it:372).  They must take the same decisions on every fixture."""
    for path in FIXTURES:
        z = np.load(path)
        ls, meshes = util.arrays_to_scene(z)
        a = oracle.trace(ls, meshes, backend="port", power_sum="reference", **_params(z))
        b = oracle.trace(ls, meshes, backend="port", power_sum="f64", **_params(z))
        assert a["live"] == b["live"]
        if os.path.basename(path).startswith("c1_"):
            continue  # negative IOR: child powers of +-1e5 cancel, the two sums differ in the 3rd digit
        for x, y in zip(a["power_left"], b["power_left"]):
            assert abs(x - y) <= 1e-5 * max(abs(x), 1e-30) + 1e-9


def test_c1_both_power_sum_rules_give_the_same_generations():
    """the GPU path ends the trace on a float64 sum of the surviving power, the
    reference on a sorted sequential float32 sum (it:372).  On C1 (negative IOR: child powers of
    +-1e5 cancel, so the two sums differ in the third digit) the BREAK DECISIONS must still agree:
    identical `live` lists, at 300 and at 1,000 rays, with the reference kernels where present.
    Tolerance statement: the rule is a threshold test on
    power_left < 0.01 * input_power; the two sums differ by <= 1e-2 relative on C1 and <= 1e-5
    elsewhere, so a decision could only flip for a generation whose surviving power lies within that
    band of the threshold - none of the configurations has one (closest approach asserted below)."""
    from lightpycl_b200 import scenes

    backend = "reference" if oracle.available("reference") else "port"
    for n in (300, 1000):
        ls, meshes, params = scenes.nested_cubes(ray_count=n)
        a = oracle.trace(ls, meshes, backend=backend, power_sum="reference", **params)
        b = oracle.trace(ls, meshes, backend=backend, power_sum="f64", **params)
        assert a["live"] == b["live"] and len(a["results"]) == len(b["results"])
        thr = 0.01 * float(np.asarray(a["input_power"]).reshape(-1)[0])
        for x, y in zip(a["power_left"], b["power_left"]):
            assert (x < thr) == (y < thr)
            assert min(abs(x - thr), abs(y - thr)) > 4.0 * abs(x - y)  # no generation near the threshold
