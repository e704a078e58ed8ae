"""Shared helpers for the test-suite (test infrastructure)."""

from __future__ import annotations

import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class ArraySource:
    """Duck-typed light source (the tracer reads exactly these four attributes, reference it:102-105)."""

    def __init__(self, origin, direc, power):
        self.rays_origin = np.asarray(origin, dtype=np.float32)
        self.rays_dir = np.asarray(direc, dtype=np.float32)
        self.rays_power = np.asarray(power, dtype=np.float32)
        self.ray_count = len(self.rays_origin)


class ArrayMesh:
    """Duck-typed mesh (the tracer calls getMaterialBuf() and tribuf(), reference it:130,137)."""

    def __init__(self, v0, v1, v2, mat_type, ior, refl, diss):
        self.v = (np.asarray(v0, np.float32), np.asarray(v1, np.float32), np.asarray(v2, np.float32))
        self.mat = {"type": int(mat_type), "IOR": float(ior), "R": float(refl), "dissipation": float(diss)}

    def getMaterialBuf(self):
        return dict(self.mat)

    def tribuf(self):
        return self.v


def scene_to_arrays(ls, meshes):
    """Freeze a (light sources, meshes) pair into plain arrays for a golden fixture."""
    from oracle import oracle

    origin, direc, power, _ = oracle.flatten_sources(ls)
    sc = oracle.Scene(meshes)
    return dict(origin=origin, dir=direc, power=power, v0=sc.v0, v1=sc.v1, v2=sc.v2, tri_mesh=sc.tri_mesh,
                mat_type=sc.mat_type, ior=sc.ior, refl=sc.refl, diss=sc.diss)


def arrays_to_scene(z):
    """Inverse of scene_to_arrays -> ([source], [meshes])."""
    src = ArraySource(z["origin"], z["dir"], z["power"])
    meshes = []
    tm = z["tri_mesh"]
    for j in range(len(z["mat_type"])):
        sel = tm == j
        meshes.append(ArrayMesh(z["v0"][sel], z["v1"][sel], z["v2"][sel], z["mat_type"][j], z["ior"][j], z["refl"][j],
                                z["diss"][j]))
    return [src], meshes


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_results_equal(got, want, what=""):
    """Bit-exact comparison of two `results` lists (x, y, z lanes of the 4-vectors; w is don't-care)."""
    assert len(got) == len(want), "%s: %d iterations vs %d" % (what, len(got), len(want))
    for k, (a, b) in enumerate(zip(got, want)):
        for x, y, name in zip(a, b, ("origin", "dest", "power", "measured")):
            x, y = np.asarray(x), np.asarray(y)
            assert x.shape == y.shape, "%s iter %d %s: shape %s vs %s" % (what, k, name, x.shape, y.shape)
            if name in ("origin", "dest"):
                x, y = x[:, :3], y[:, :3]
            if not np.array_equal(bits(x), bits(y)):
                bad = np.argwhere(bits(x) != bits(y))
                raise AssertionError("%s iter %d %s differs at %d places, first %s: %r vs %r"
                                     % (what, k, name, len(bad), bad[0], x[tuple(bad[0])], y[tuple(bad[0])]))


def mini_scenes():
    """Small versions of the config families (hundreds to ~2k triangles) used for golden fixtures
    and fast CPU tests.  Built through the package's generators."""
    from lightpycl_b200 import geo_optical_elements as goe
    from lightpycl_b200 import light_source as lsm
    from lightpycl_b200 import scenes

    oe = goe.optical_elements()
    out = {}

    def dome(radius, n=12):
        phi = np.linspace(0.0, np.pi, n)
        m = oe.revolve_curve([[a, b] for a, b in zip(np.cos(phi) * radius, np.sin(phi) * radius)], axis="x",
                             ang=np.pi, ang_pts=n + 1)
        m.setMaterial(mat_type="measure")
        return m

    def src(seed, n, direction=(0, 0, 1), power=1000.0, center=(0, 0, 0, 0)):
        np.random.seed(seed)
        return lsm.light_source(center=np.array(center, dtype=np.float32), direction=direction,
                                directivity=lambda x, y: np.cos(y), power=power, ray_count=n)

    # nested cubes: the reference example verbatim (10 rays) and a denser one
    out["c1_cubes_10"] = scenes.nested_cubes(ray_count=10)
    out["c1_cubes_300"] = scenes.nested_cubes(ray_count=300)

    # lens inside a measure dome (C2 family)
    lens = oe.revolve_curve(oe.lens_spherical_2r((0, 0, 0), 60.0, 6000.0, 50.0, 1, 2.5), axis="x", ang=np.pi, ang_pts=8)
    lens.setMaterial(mat_type="refractive", IOR=2.5)
    lens.rotate(axis="y", angle=-np.pi / 2.0, pivot=(0, 0, 0, 0))
    out["mini_lens"] = ([src(7, 400)], [dome(1000.0), lens],
                        dict(trace_iterations=10, trace_until_dissipated=0.99, max_ray_len=np.float32(2e3),
                             ior_env=np.float32(1.0)))

    # parabolic mirror + terminator plate + dome (C3 family)
    yn = np.linspace(0.0, 10.0, 24)
    xn = yn ** 2 / 20.0 - 5.0
    mirror = oe.revolve_curve([[a, b] for a, b in zip(xn, yn)], axis="x", ang=2.0 * np.pi, ang_pts=17)
    mirror.setMaterial(mat_type="mirror")
    mirror.reflectivity = 0.98
    mirror.rotate(axis="y", angle=-np.pi / 2, pivot=(0, 0, 0, 0))
    plate = oe.cube(center=(0, 0, 100, 0), size=[4, 4, 1, 0])
    plate.setMaterial(mat_type="terminator")
    out["mini_mirror"] = ([src(8, 400, direction=(0, 0, -1), power=1.0)], [dome(500.0), mirror, plate],
                          dict(trace_iterations=10, trace_until_dissipated=0.99, max_ray_len=np.float32(1e3),
                               ior_env=np.float32(1.0)))

    # lossy block (Beer-Lambert) + an anisotropic (type 4) block, two sources of (N,1) power
    block = oe.cube(center=(0, 0, 20, 0), size=[10, 10, 10, 0])
    block.setMaterial(mat_type="refractive", IOR=1.3, dissipation=0.08)
    aniso = oe.cube(center=(30, 0, 20, 0), size=[10, 10, 10, 0])
    aniso.setMaterial(mat_type="refractive_anisotropic")
    out["mini_lossy"] = ([src(9, 200), src(10, 150, direction=(0.3, 0, 1))], [dome(1000.0), block, aniso],
                         dict(trace_iterations=8, trace_until_dissipated=0.99, max_ray_len=np.float32(2e3),
                              ior_env=np.float32(1.0)))

    # collimated beam (power shape (N,)) through two touching blocks offset by a relative 1e-6 (C4 family logic)
    np.random.seed(11)
    beam = lsm.light_source(center=np.array([0, 0, -10, 0], dtype=np.float32), direction=(0, 0.01, 1),
                            directivity=lambda x, y: 1.0 + 0.0 * np.cos(y), power=1000.0, ray_count=300)
    beam.random_collimated_rays(diameter=5.0)
    b1 = oe.cube(center=(0, 0, 5, 0), size=[20, 20, 10, 0])
    b1.setMaterial(mat_type="refractive", IOR=1.3771)
    b2 = oe.cube(center=(0, 0, 15 * (1.0 + 1e-6), 0), size=[20, 20, 10, 0])
    b2.setMaterial(mat_type="refractive", IOR=1.42)
    screen = oe.cube(center=(0, 0, 30, 0), size=[40, 40, 1, 0])
    screen.setMaterial(mat_type="measure")
    out["mini_stack"] = ([beam], [screen, b1, b2],
                         dict(trace_iterations=12, trace_until_dissipated=0.99, max_ray_len=np.float32(4e1),
                              ior_env=np.float32(1.0)))
    return out
