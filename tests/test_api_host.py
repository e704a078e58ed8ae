"""Host-side API shell: generators, light sources, flattening, C-ABI symbol table.  No GPU."""

import ctypes
import os
import re

import numpy as np
import pytest

from lightpycl_b200 import _cabi, dist, geo_optical_elements as goe, light_source as lsm, scenes
from lightpycl_b200 import iterative_tracer as it
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_triangle_counts_match_survey():
    oe = goe.optical_elements()
    assert len(oe.cube((0, 0, 0, 0), [1, 1, 1, 0]).triangles) == 12
    assert len(oe.hemisphere([0, 0, 0, 0], 10.0).triangles) == 2 * 71 * 73
    assert len(oe.sphere([0, 0, 0, 0], 10.0).triangles) == 2 * 71 * 73
    assert len(oe.parabolic_mirror().triangles) == 2 * 199 * 73
    assert len(oe.lens_spherical_biconcave((0, 0, 0), 60.0, 6000.0, 50.0, 2.5).triangles) == 2 * 120 * 36
    assert len(oe.spherical_lens_nofoc(7.8, 6.5, 0.0, 0.55, 5.0).triangles) == 2 * 99 * 72
    for name, want in (("c2_directivity_lens", 19006), ("c3_parabolic_mirror", 100552), ("c4_human_eye", 67390)):
        _, meshes, _ = scenes.CONFIGS[name](ray_count=4)
        assert sum(len(m.triangles) for m in meshes) == want


def test_revolve_curve_matches_loop_definition():
    """Vectorised revolve_curve == the per-vertex construction of reference geo:285-300."""
    curve = [[0.0, 0.0], [1.0, 2.0], [3.0, 2.5], [4.0, 0.5]]
    ang, ang_pts = np.pi, 5
    mesh = goe.optical_elements().revolve_curve(curve, axis="x", ang=ang, ang_pts=ang_pts)
    d_ang = ang / (ang_pts - 1.0)
    angs = np.linspace(0.0, ang - d_ang, ang_pts)
    verts, tris = [], []
    for k, (p1, p2) in enumerate(zip(angs, angs + d_ang)):
        for i in range(len(curve) - 1):
            for phi, c in ((p1, curve[i]), (p1, curve[i + 1]), (p2, curve[i]), (p2, curve[i + 1])):
                verts.append(np.array([c[0], np.cos(phi) * c[1], np.sin(phi) * c[1], 0], dtype=np.float32))
            base = i * 4 + 4 * k * (len(curve) - 1)
            tris += [[base, base + 1, base + 2], [base + 2, base + 3, base + 1]]
    assert np.array_equal(np.asarray(mesh.vertices), np.array(verts))
    assert np.array_equal(np.asarray(mesh.triangles), np.array(tris))


def test_material_quirks():
    m = goe.optical_elements().cube((0, 0, 0, 0), [1, 1, 1, 0])
    m.setMaterial(mat_type="mirror", reflectivity=0.5)  # Q2: reflectivity is NOT stored
    assert m.getMaterialBuf() == {"type": 1, "IOR": 1.0, "R": 1.0, "dissipation": 0.0}
    m.setMaterial(mat_type="refractive", IOR=1.5, dissipation=0.1)
    assert m.getMaterialBuf() == {"type": 0, "IOR": 1.5, "R": 1.0, "dissipation": 0.1}
    with pytest.raises(NameError):
        m.setMaterial(mat_type="glass")


def test_light_source_is_seed_reproducible_and_rng_order():
    np.random.seed(42)
    a = lsm.light_source(ray_count=100, power=10.0)
    np.random.seed(42)
    u = np.random.rand(100, 1)
    v = np.random.rand(100, 1)
    el, az = np.arccos(u), 2.0 * np.pi * v
    assert a.rays_power.shape == (100, 1) and a.rays_dir.dtype == np.float32
    # direction (0,0,1): elevation0 = 0, azimuth0 = +pi/2; row vector x Rz(pi/2): (x, y, z) -> (y, -x, z)
    assert np.allclose(a.rays_dir[:, 2], np.cos(el)[:, 0], atol=1e-6)
    assert np.allclose(a.rays_dir[:, 1], -(np.sin(el) * np.cos(az))[:, 0], atol=1e-6)
    assert np.allclose(a.rays_dir[:, 0], (np.sin(el) * np.sin(az))[:, 0], atol=1e-6)
    assert abs(float(a.rays_power.sum()) - 10.0) < 1e-4
    np.random.seed(42)
    b = lsm.light_source(ray_count=100, power=10.0)
    assert np.array_equal(a.rays_dir, b.rays_dir) and np.array_equal(a.rays_power, b.rays_power)
    b.random_collimated_rays(diameter=2.0)  # consumes two more draws, square footprint, (N,) power
    assert b.rays_power.shape == (100,) and np.all(np.abs(b.rays_origin[:, :2]) <= 1.0)
    assert np.allclose(b.rays_dir[:, :3], [0, 0, 1], atol=1e-7)


def test_flatten_matches_oracle_flatten():
    ls, meshes, _ = scenes.human_eye(ray_count=50)
    o, d, p = it.flatten_light_sources(ls)
    oo, od, op, oin = oracle.flatten_sources(ls)
    assert np.array_equal(o, oo) and np.array_equal(d, od) and np.array_equal(p, op)
    assert it.sorted_sequential_sum_f32(p) == oin
    flat = it.flatten_meshes(meshes)
    sc = oracle.Scene(meshes)
    for k, ref in (("v0", sc.v0), ("v1", sc.v1), ("v2", sc.v2), ("tri_mesh", sc.tri_mesh), ("mat_type", sc.mat_type),
                   ("ior", sc.ior), ("refl", sc.refl), ("diss", sc.diss)):
        assert np.array_equal(flat[k], ref), k
    ls1, _, _ = scenes.nested_cubes(ray_count=7)  # (N,1) power: np.sort along the last axis is a no-op
    _, _, p1 = it.flatten_light_sources(ls1)
    seq = np.float32(0)
    for x in p1.reshape(-1):
        seq = np.float32(seq + x)
    assert it.sorted_sequential_sum_f32(p1) == seq


def test_cabi_exports_every_declared_symbol():
    """The shared library must load without a GPU and export exactly what include/*.h declares."""
    header = open(os.path.join(ROOT, "include", "lightpycl_b200.h")).read()
    declared = set(re.findall(r"\b(lpcl_[a-z_0-9]+)\s*\(", header))
    declared -= {"lpcl_step_counters", "lpcl_generation"}
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    lib = _cabi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.lpcl_abi_version() == _cabi.ABI_VERSION
    assert ctypes.sizeof(_cabi.StepCounters) == 64 and ctypes.sizeof(_cabi.Generation) == 56
    # argument errors are reported, not crashed on (no GPU needed: validation happens first)
    assert lib.lpcl_scene_create(None, None, None, None, None, 0, None, None, None, None, 0, None) == -1
    assert b"NULL" in lib.lpcl_last_error()


def test_tracer_refuses_to_run_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        it.CL_Tracer()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000003):
        for ws in (1, 2, 3, 8):
            b = [dist.shard_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_flatten_fast_path_equals_tribuf_path():
    """flatten_meshes gathers indexed meshes (stock GeoObject.tribuf) straight into float32 buffers;
    duck-typed meshes go through tribuf().  Both must give the same bits (cast commutes with gather),
    also after rotate/translate and for float64 cube vertices (geo:154-162)."""
    from lightpycl_b200 import geo_optical_elements as goe
    from lightpycl_b200 import scenes
    from lightpycl_b200.iterative_tracer import flatten_meshes

    _, meshes, _ = scenes.nested_cubes(ray_count=4)
    _, more, _ = scenes.directivity_lens(ray_count=4)
    meshes = list(meshes) + list(more)
    meshes[1].rotate(axis="x", angle=0.3, pivot=(0, 0, 0, 0))

    class Duck:  # same data, but only the reference's duck-typed interface (it:130,137)
        def __init__(self, m):
            self._m = m

        def getMaterialBuf(self):
            return self._m.getMaterialBuf()

        def tribuf(self):
            return goe.GeoObject.tribuf(self._m)

    fast, slow = flatten_meshes(meshes), flatten_meshes([Duck(m) for m in meshes])
    for k in ("v0", "v1", "v2"):
        assert fast[k].dtype == np.float32 and fast[k].flags["C_CONTIGUOUS"]
        assert np.array_equal(fast[k].view(np.uint32), slow[k].view(np.uint32)), k
    for k in ("tri_mesh", "mat_type", "ior", "refl", "diss"):
        assert np.array_equal(fast[k], slow[k]), k


def test_measured_result_buffers_are_reused_only_when_dropped():
    """get_measured_rays hands out arrays backed by a small pool (first-touch page faults of fresh host
    memory cost more than the copy): a backing buffer may be reused only when the caller holds no array
    derived from it any more - results already handed out must never be overwritten.  Ownership is
    tracked by an explicit lease per hand-out (not by reference counts)."""
    import gc

    from lightpycl_b200.iterative_tracer import _HostPool

    pool = _HostPool()
    key = ("pos", 4, np.dtype(np.float32).str)
    a = pool.array(key, (1000, 4), np.float32)
    a[:] = 1.0
    a2 = pool.array(key, (900, 4), np.float32)  # a still referenced: different memory
    assert not np.shares_memory(a, a2)
    addr = a.ctypes.data
    del a
    gc.collect()
    a3 = pool.array(key, (800, 4), np.float32)  # first array dropped: its memory is reused
    assert a3.ctypes.data == addr and len(pool._entries[key]) == 2
    # every kind of derived array protects the buffer: a slice, a column view, a reshape, a dtype view
    for derive in (lambda x: x[:10], lambda x: x[:, :3], lambda x: x.reshape(-1), lambda x: x.view(np.uint32)):
        held = pool.array(key, (700, 4), np.float32)
        held[:] = 7.0
        keep = derive(held)
        del held
        gc.collect()
        others = [pool.array(key, (700, 4), np.float32) for _ in range(3)]
        for o in others:
            o[:] = -1.0
            assert not np.shares_memory(keep, o)
        assert np.all(keep.view(np.float32) == 7.0)
        del keep, others, o
        gc.collect()
    big = pool.array(key, (5000, 4), np.float32)  # larger than any pooled buffer: a new one
    assert big.shape == (5000, 4) and big.dtype == np.float32
    assert pool.array(key, (0, 4), np.float32).shape == (0, 4)


def test_a_held_result_survives_later_calls():
    """The user-level version of the above: hold the first (pos, pwr) across more hand-outs."""
    from lightpycl_b200.iterative_tracer import CL_Tracer

    class Holder:  # _host_result_arrays only needs the pool
        _host_array = CL_Tracer._host_array

    from lightpycl_b200.iterative_tracer import _HostPool

    t = Holder()
    t._host_pool = _HostPool()
    pos, pwr = CL_Tracer._host_result_arrays(t, 1000)
    pos[:], pwr[:] = 3.0, 4.0
    xyz = pos[:, :3]
    del pos
    for _ in range(6):
        p2, w2 = CL_Tracer._host_result_arrays(t, 1000)
        p2[:], w2[:] = 0.0, 0.0
    assert np.all(xyz == 3.0) and np.all(pwr == 4.0)


def test_ray_tile_size_follows_the_library_workspace_sizing():
    """CL_Tracer._tile_rays: rays per tile = ray_tile_bytes / (marginal workspace bytes per ray as the LIBRARY sizes them -
    12 B per mesh of search results + the record budget of the deferred-flag search - + the per-tile end points), in
    multiples of 4096, at least 65,536; a generation that fits is not tiled.  No GPU: the library is a stand-in."""
    from lightpycl_b200 import iterative_tracer as it

    class FakeLib:
        calls = 0

        def lpcl_step_workspace_bytes(self, scene, n):
            FakeLib.calls += 1
            return 3 * 12 * n + 256 + 8 * 32 * n + 4096  # key + count for 3 meshes, 32 records of 8 bytes per ray

    tr = object.__new__(it.CL_Tracer)
    tr._lib, tr._scene, tr.mesh_count, tr._scratch_per_ray = FakeLib(), None, 3, None
    tr.ray_tile_bytes = 1 << 30
    per_ray = 36 + 256 + 17
    want = ((1 << 30) // per_ray) // 4096 * 4096
    assert tr._tile_rays(10 * want) == want
    assert tr._tile_rays(want) == want and tr._tile_rays(1000) == 1000
    assert FakeLib.calls == 2  # sized once per scene
    tr.ray_tile_bytes = 1
    assert tr._tile_rays(1_000_000) == 65536
