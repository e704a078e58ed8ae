"""The fixtures tests/golden/refhost_*.npz must be what the reference's own Python produces TODAY in the authoring
container: this re-runs a slice of tests/golden/make_golden_host.py (oracle/ref_host.py executing
/root/reference/*.py by path) and compares with the committed files.  Skipped where /root/reference does not exist
(the GPU box): there the fixtures are the only trace of the reference."""

import os

import numpy as np
import pytest

import refhost_cases as cases
import util
from oracle import ref_host

pytestmark = pytest.mark.skipif(not (ref_host.available() and ref_host.kernels_available()),
                                reason="needs /root/reference and oracle/_ref (authoring container)")


def _load(name):
    return np.load(os.path.join(util.GOLDEN_DIR, name), allow_pickle=False)


def test_reference_light_sources_reproduce_the_fixture():
    lsm = ref_host.load("light_source")
    z = _load("refhost_sources.npz")
    with ref_host.captured_stdout():
        made = {name: make() for name, make in cases.source_cases(lsm).items()}
    for name, s in made.items():
        assert np.array_equal(np.asarray(s.rays_dir), z[name + "/dir"]), name
        assert np.array_equal(np.asarray(s.rays_origin), z[name + "/origin"]), name
        assert np.array_equal(np.asarray(s.rays_power), z[name + "/power"]), name


def test_reference_trace_loop_reproduces_the_fixture():
    """The reference's CL_Tracer (its own __init__, its own loop) on its own kernels, nested cubes at 300 rays."""
    z = _load("refhost_trace_c1.npz")
    with ref_host.reference_scene_modules() as scenes:
        ls, meshes, params = scenes.nested_cubes(ray_count=300)
    tr = ref_host.reference_tracer()
    assert type(tr).__module__.startswith("reference_") and tr.prg.source.startswith("/*")  # the .cl text it read itself
    with ref_host.captured_stdout() as out:
        res = tr.iterative_tracer(ls, meshes, **params)
    assert "ITERATION" in out.getvalue()  # the reference's own progress prints
    assert len(res) == int(z["n_iter"])
    for k, arrs in enumerate(res):
        for key, a in zip(("origin", "dest", "power", "meas"), arrs):
            assert np.array_equal(util.bits(np.asarray(a)), util.bits(z["it%d_%s" % (k, key)])), (k, key)


def test_python2_division_semantics_of_the_shim():
    assert ref_host._py2_div(7, 2) == 3 and ref_host._py2_div(-7, 2) == -4
    assert ref_host._py2_div(7.0, 2) == 3.5 and ref_host._py2_div(np.int32(7), np.int64(2)) == 3
    assert ref_host._py2_div(np.float32(1), 4) == 0.25
    src = ref_host._rewrite_prints('\tprint "a # not a comment", x # comment\n\tprint\n')
    assert 'print("a # not a comment", x)  # comment' in src and "print()" in src
