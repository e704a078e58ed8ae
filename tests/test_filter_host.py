"""Property test of the conservative search filter on the CPU: filter_candidate() (lane-for-lane
the arithmetic of the packed CUDA loop, lpcl_math.cuh) must accept every (ray, triangle) pair that
the reference-ordered exact test accepts - over scene scales, triangle sizes from 1e-3 to 1 of the
nominal size, far-away origins, non-unit and reversed directions, edge/corner grazing hits and
degenerate triangles.  The filter may only cost time (false candidates), never drop a hit."""

import ctypes as C
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "filter_harness.so")


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO,
                           os.path.join(HERE, "filter_harness.cpp")])
    lib = C.CDLL(SO)
    for fn in (lib.filter_property, lib.filter_property_tc):
        fn.restype = C.c_long
        fn.argtypes = [C.c_long, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_int,
                       C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_double)]
    return lib


@pytest.mark.parametrize("which", ["fp32", "tensor_core_model"])
@pytest.mark.parametrize("scene_radius,tri_size,origin_radius,unit_dirs", [
    (10.0, 1.0, 10.0, 1),       # compact scene
    (1000.0, 40.0, 1000.0, 1),  # C2-like: dome triangles of tens of units, r = 1000
    (1000.0, 1.0, 50.0, 1),     # small triangles far from the origin
    (5.0, 0.05, 500.0, 1),      # origins 100x outside the scene
    (100.0, 10.0, 100.0, 0),    # non-unit directions, 1e-2 .. 1e2
    (0.01, 0.001, 0.01, 1),     # tiny scene
    (1e5, 1e3, 1e5, 1),         # huge scene
])
def test_filter_never_rejects_a_reference_hit(harness, which, scene_radius, tri_size, origin_radius, unit_dirs):
    """fp32: filter_candidate(), the packed FFMA2 loop.  tensor_core_model: filter_candidate_tc(), the split-TF32
    GEMM of lpcl_search_tc.cuh with its three outputs pushed to every corner of the budgeted MMA rounding error."""
    n_exact, n_cand, worst = C.c_long(), C.c_long(), C.c_double()
    n = 2_000_000 if which == "fp32" else 500_000
    fn = harness.filter_property if which == "fp32" else harness.filter_property_tc
    bad = fn(n, 12345, scene_radius, tri_size, origin_radius, unit_dirs, C.byref(n_exact), C.byref(n_cand), C.byref(worst))
    # (the tiny scene mostly fails the reference's ABSOLUTE |DEN| < 1e-6 cut-off)
    assert n_exact.value > (100 if tri_size < 0.01 else n // 20), "test generator no longer produces hits"
    assert bad == 0, "%d reference hits rejected by the filter (of %d)" % (bad, n_exact.value)
    assert n_cand.value >= n_exact.value
