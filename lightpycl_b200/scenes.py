"""The benchmark / parity configurations C1-C5 (BASELINE.json `configs`).

Every scene is built through this package's ``geo_optical_elements`` / ``light_source``
(the py3 mirrors of the reference modules), so the CPU oracle and the CUDA path consume
identical float32 arrays.  Rays are reproducible: ``np.random.seed(1234 + source_index)`` is
called immediately before each source's constructor.

Each builder returns ``(light_sources, meshes, params)`` where ``params`` holds the keyword
arguments for ``CL_Tracer.iterative_tracer``.
"""

from __future__ import annotations

import numpy as np

from . import geo_optical_elements as goe
from . import light_source as ls_mod

SEED0 = 1234


def _cosine(x, y):
    return np.cos(y)


def _flat(x, y):
    return 1.0 + 0.0 * np.cos(y)


def _source(index, **kw):
    np.random.seed(SEED0 + index)
    return ls_mod.light_source(**kw)


def nested_cubes(ray_count=10, trace_iterations=16):
    """C1 - reference example_nested_cubes_refraction.py:45-85 verbatim in parameters: four
    refractive boxes (IOR 1.5, -2, -2, -2), two of them offset by a relative 1e-6 so that
    their faces fall inside one EPSILON window of the slab's top face; no measure surface."""
    src = _source(0, center=np.array([0, 0, 0, 0], dtype=np.float32), direction=(0, 0, 1), directivity=_cosine,
                  power=1000.0, ray_count=ray_count)
    oe = goe.optical_elements()
    meshes = []
    m = oe.cube(center=(0, 0, 20, 0), size=[100, 100, 10, 0])
    m.setMaterial(mat_type="refractive", IOR=1.5)
    meshes.append(m)
    m = oe.cube(center=(-20, 0, 22.5 * (1.0 - 1e-6), 0), size=[30, 30, 5, 0])
    m.setMaterial(mat_type="refractive", IOR=-2.0)
    meshes.append(m)
    m = oe.cube(center=(-20, 0, 27.5 * (1.0 + 1e-6), 0), size=[30, 30, 5, 0])
    m.setMaterial(mat_type="refractive", IOR=-2.0)
    meshes.append(m)
    m = oe.cube(center=(20, 0, 20, 0), size=[30, 30, 5, 0])
    m.setMaterial(mat_type="refractive", IOR=-2.0)
    meshes.append(m)
    params = dict(trace_iterations=trace_iterations, trace_until_dissipated=0.99, max_ray_len=np.float32(2.0e2),
                  ior_env=np.float32(1.0))
    return [src], meshes, params


def directivity_lens(ray_count=1_000_000, trace_iterations=10):
    """C2 - reference example_directivity_lens.py:48-68: measure hemisphere r=1000 (10,366
    triangles) + biconcave lens IOR 2.5 (8,640 triangles); BASELINE: 1M rays, depth 10."""
    src = _source(0, center=np.array([0, 0, 0, 0], dtype=np.float32), direction=(0, 0, 1), directivity=_cosine,
                  power=1000.0, ray_count=ray_count)
    oe = goe.optical_elements()
    dome = oe.hemisphere(center=[0, 0, 0, 0], radius=1000.0)
    dome.setMaterial(mat_type="measure")
    lens = oe.lens_spherical_biconcave(focus=(0, 0, 0), r1=60.0, r2=6000.0, diameter=50.0, IOR=2.5)
    lens.rotate(axis="y", angle=-np.pi / 2.0, pivot=(0, 0, 0, 0))
    params = dict(trace_iterations=trace_iterations, trace_until_dissipated=0.99, max_ray_len=np.float32(2.0e3),
                  ior_env=np.float32(1.0))
    return [src], [dome, lens], params


def parabolic_mirror(ray_count=1_000_000, trace_iterations=10, profile_pts=400, ang_pts=113, reflectivity=0.98):
    """C3 - reference example_directivity_parabolic_mirror.py:41-77 scaled to BASELINE's
    "1M rays, ~100k triangles": source at the focus pointing -z; measure hemisphere r=500
    (10,366); mirror = the geo:233-236 parabola (f=5, d=20) revolved with 400 profile points x
    113 wedges = 90,174 triangles, reflectivity set BY ATTRIBUTE; plus a 12-triangle
    terminator plate in the collimated beam (feed-blockage stand-in).  Total 100,552 triangles."""
    src = _source(0, center=np.array([0, 0, 0, 0], dtype=np.float32), direction=(0, 0, -1), directivity=_cosine,
                  power=1.0, ray_count=ray_count)
    oe = goe.optical_elements()
    dome = oe.hemisphere(center=[0, 0, 0, 0], radius=500.0)
    dome.setMaterial(mat_type="measure")
    focal_length, diameter = 5.0, 20.0
    yn = np.linspace(0.0, diameter / 2.0, profile_pts)
    xn = yn ** 2 / (4.0 * focal_length) - focal_length
    mirror = oe.revolve_curve([[a, b] for (a, b) in zip(xn, yn)], axis="x", ang=2.0 * np.pi, ang_pts=ang_pts)
    mirror.setMaterial(mat_type="mirror")
    mirror.reflectivity = reflectivity
    mirror.rotate(axis="y", angle=-np.pi / 2, pivot=(0, 0, 0, 0))
    plate = oe.cube(center=(0, 0, 100, 0), size=[4, 4, 1, 0])
    plate.setMaterial(mat_type="terminator")
    params = dict(trace_iterations=trace_iterations, trace_until_dissipated=0.99, max_ray_len=np.float32(1.0e3),
                  ior_env=np.float32(1.0))
    return [src], [dome, mirror, plate], params


def human_eye(ray_count=1_000_000, sources=4, trace_iterations=16):
    """C4 - reference example_human_eye.py:62-103 geometry (retina measure surface + cornea,
    lens, aqueous and vitreous bodies whose interfaces are offset by a relative 1e-6, far
    inside EPSILON = 4e-5); BASELINE: 4 collimated sources x 1M rays."""
    dirs = [(0, 0.01, 1), (0, -0.01, 1), (0.01, 0, 1), (-0.01, 0, 1)]
    lights = []
    for k in range(sources):
        s = _source(k, center=np.array([0, 0, -10, 0], dtype=np.float32), direction=dirs[k % 4], directivity=_flat,
                    power=1000.0, ray_count=ray_count)
        s.random_collimated_rays(diameter=5.0)
        lights.append(s)
    oe = goe.optical_elements()
    r_cornea, r_lens = 5.0, 5.0
    r_ac, d_ac, r_pc, d_pc = 7.8, 0.0, 6.5, 0.55
    r_al, d_al, r_pl, d_pl = 10.2, 3.6, -6.0, 7.6
    r_r, d_r = -12.1, 24.2
    retina = oe.hemisphere(center=[0, 0, d_r / 2.0, 0], radius=-r_r * (1.0 - 1e-3))
    retina.setMaterial(mat_type="measure")
    cornea = oe.spherical_lens_nofoc(r1=r_ac, r2=r_pc, x1=d_ac, x2=d_pc, d=r_cornea)
    cornea.setMaterial(mat_type="refractive", IOR=1.3771)
    lens = oe.spherical_lens_nofoc(r1=r_al, r2=r_pl, x1=d_al, x2=d_pl, d=r_lens)
    lens.setMaterial(mat_type="refractive", IOR=1.4200)
    aqueous = oe.spherical_lens_nofoc(r1=r_pc, r2=r_al, x1=d_pc * (1.0 + 1e-6), x2=d_al * (1.0 - 1e-6), d=r_cornea,
                                      d2=r_lens)
    aqueous.setMaterial(mat_type="refractive", IOR=1.3374)
    vitreous = oe.spherical_lens_nofoc(r1=r_pl, r2=r_r * (1.0 + 1e-6), x1=d_pl * (1.0 + 1e-6), x2=d_r, d=r_lens,
                                       d2=r_lens, sign2_arcsin=-1.0)
    vitreous.setMaterial(mat_type="refractive", IOR=1.336)
    params = dict(trace_iterations=trace_iterations, trace_until_dissipated=0.99, max_ray_len=np.float32(4.0e1),
                  ior_env=np.float32(1.0))
    return lights, [retina, cornea, lens, aqueous, vitreous], params


def lenslet_profile(radius=5.0, curvature_radius=10.0, pts=60):
    """Closed symmetric two-arc (biconvex) profile, 2*pts+1 points, optical axis = x."""
    phi = np.arcsin(radius / curvature_radius)
    cx = curvature_radius * np.cos(phi)
    a1 = np.linspace(np.pi - phi, np.pi + phi, pts)  # front surface, bulging to -x
    a2 = np.linspace(-phi, phi, pts)  # back surface, bulging to +x
    x = np.concatenate((cx + curvature_radius * np.cos(a1), -cx + curvature_radius * np.cos(a2)))
    y = np.concatenate((curvature_radius * np.sin(a1), curvature_radius * np.sin(a2)))
    x = np.append(x, x[0])
    y = np.append(y, y[0])
    return [[a, b] for (a, b) in zip(x, y)]


def lenslet_array(ray_count=1_000_000, sources=16, grid=8, trace_iterations=12, per_lenslet_mesh=False,
                  ang_pts=65):
    """C5 - synthetic: measure hemisphere r=1000 (10,366) + an (grid x grid) array of biconvex
    lenslets (each 2*120*ang_pts triangles; 8x8x15,600 = 998,400), IOR 1.5, radius 5, pitch 12,
    centred at z=50, merged into ONE refractive mesh with GeoObject.append (geo:134) + a terminator
    slab under the sources (12).  16 cosine sources x 1M rays on a 4x4 grid at z=0.  All triangle
    areas are >> 1e-6.  ``per_lenslet_mesh=True`` is variant C5' (mesh_count=66)."""
    side = int(round(np.sqrt(sources)))
    lights = []
    for k in range(sources):
        cx = ((k % side) - (side - 1) / 2.0) * 12.0
        cy = ((k // side) - (side - 1) / 2.0) * 12.0
        lights.append(_source(k, center=np.array([cx, cy, 0, 0], dtype=np.float32), direction=(0, 0, 1),
                              directivity=_cosine, power=1000.0, ray_count=ray_count))
    oe = goe.optical_elements()
    dome = oe.hemisphere(center=[0, 0, 0, 0], radius=1000.0)
    dome.setMaterial(mat_type="measure")
    proto = oe.revolve_curve(lenslet_profile(), axis="x", ang=np.pi, ang_pts=ang_pts)
    proto.rotate(axis="y", angle=-np.pi / 2.0, pivot=(0, 0, 0, 0))
    pv = np.asarray(proto.vertices, dtype=np.float32)
    pt = np.asarray(proto.triangles, dtype=np.int64)
    lenslets = []
    for gy in range(grid):
        for gx in range(grid):
            off = np.array([(gx - (grid - 1) / 2.0) * 12.0, (gy - (grid - 1) / 2.0) * 12.0, 50.0, 0.0])
            lenslets.append((pv + off).astype(np.float32))
    if per_lenslet_mesh:
        lens_meshes = []
        for v in lenslets:
            m = goe.GeoObject(v, pt.copy())
            m.setMaterial(mat_type="refractive", IOR=1.5)
            lens_meshes.append(m)
    else:
        m = goe.GeoObject(lenslets[0], pt.copy())
        for v in lenslets[1:]:
            m.append(v, pt)
        m.setMaterial(mat_type="refractive", IOR=1.5)
        lens_meshes = [m]
    slab = oe.cube(center=(0, 0, -5, 0), size=[400, 400, 1, 0])
    slab.setMaterial(mat_type="terminator")
    params = dict(trace_iterations=trace_iterations, trace_until_dissipated=0.99, max_ray_len=np.float32(2.0e3),
                  ior_env=np.float32(1.0))
    return lights, [dome] + lens_meshes + [slab], params


def dissipative_cube(ray_count=10_000, trace_iterations=16, ior=1.0, dissipation=1.0):
    """Reference example_directivity_dissipative_cube.py:48-67: a lossy block (IOR 1.0,
    dissipation 1.0) inside a measure dome; exercises the Beer-Lambert branch cl:385-394 and the
    zero-power reflect children of an n1 == n2 interface."""
    src = _source(0, center=np.array([0, 0, 0, 0], dtype=np.float32), direction=(0, 0, 1), directivity=_cosine,
                  power=1000.0, ray_count=ray_count)
    oe = goe.optical_elements()
    dome = oe.hemisphere(center=[0, 0, 0, 0], radius=1000.0)
    dome.setMaterial(mat_type="measure")
    block = oe.cube(center=(0, 0, 20, 0), size=[10, 10, 10, 0])
    block.setMaterial(mat_type="refractive", IOR=ior, dissipation=dissipation)
    params = dict(trace_iterations=trace_iterations, trace_until_dissipated=0.99, max_ray_len=np.float32(2.0e3),
                  ior_env=np.float32(1.0))
    return [src], [dome, block], params


def glass_cube_cavity(ray_count=1000, trace_iterations=40):
    """A source INSIDE a glass cube (IOR 1.5, edge 10) inside a measure sphere (r = 100): at every bounce
    the internally reflected ray stays in the cube (totally, beyond 41.8 degrees) and the refracted one
    leaves and is measured on the sphere one iteration later.  The ray count stays bounded while the
    trace runs to any depth with measured rays in every generation - the deep-trace case behind the
    reference's default of 100 iterations (it:77) and the lineage epochs of dist.renumber_lineage.
    ``trace_until_dissipated = 1.0``: the power test never fires."""
    src = _source(0, center=np.array([0.5, -0.25, 0.125, 0], dtype=np.float32), direction=(0, 0, 1), directivity=_cosine,
                  power=1000.0, ray_count=ray_count)
    oe = goe.optical_elements()
    shell = oe.sphere(center=[0, 0, 0, 0], radius=100.0)
    shell.setMaterial(mat_type="measure")
    cube = oe.cube(center=(0, 0, 0, 0), size=[10, 10, 10, 0])
    cube.setMaterial(mat_type="refractive", IOR=1.5)
    params = dict(trace_iterations=trace_iterations, trace_until_dissipated=1.0, max_ray_len=np.float32(1.0e3),
                  ior_env=np.float32(1.0))
    return [src], [shell, cube], params


def lenslet_array_per_mesh(ray_count=1_000_000, **kw):
    """C5' - the lenslet array with every lenslet its own mesh (mesh_count = 66): stresses the
    per-(ray, mesh) search scratch and the generic (M > 8) media-resolution path."""
    return lenslet_array(ray_count=ray_count, per_lenslet_mesh=True, **kw)


CONFIGS = {
    "c5p_lenslet_array_66_meshes": lenslet_array_per_mesh,
    "c1_nested_cubes": nested_cubes,
    "c2_directivity_lens": directivity_lens,
    "c3_parabolic_mirror": parabolic_mirror,
    "c4_human_eye": human_eye,
    "c5_lenslet_array": lenslet_array,
    "dissipative_cube": dissipative_cube,
    "glass_cube_cavity": glass_cube_cavity,
}
