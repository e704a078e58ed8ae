"""Case definitions shared by tests/golden/make_golden_host.py (run with the REFERENCE's modules) and
tests/test_refhost_golden.py / tests/test_gpu_refhost.py (run with this package's modules): the same calls, the same
seeds - only the module objects differ."""

import numpy as np


def _flat(x, y):
    return 1.0 + 0.0 * np.cos(y)


def _lobe(x, y):
    return np.cos(y) ** 2 * (1.0 + 0.25 * np.cos(x))


def source_cases(lsm):
    """name -> zero-argument builder of a light_source.  Every builder seeds NumPy's global RNG itself."""

    def seeded(seed, post=None, **kw):
        def make():
            np.random.seed(seed)
            s = lsm.light_source(**kw)
            if post is not None:
                post(s)
            return s

        return make

    c = np.array([1.5, -2.0, 0.25, 0], dtype=np.float32)
    return {
        "default": seeded(11),
        "z_cosine_1000": seeded(12, power=1000.0, ray_count=257),
        "minus_z": seeded(13, direction=(0, 0, -1), power=2.0, ray_count=100),
        "plus_x": seeded(14, direction=(1, 0, 0), ray_count=64),
        "minus_y_offset": seeded(15, center=c, direction=(0, -1, 0), power=3.5, ray_count=64),
        "oblique": seeded(16, center=c, direction=(0.3, -0.2, 0.5), ray_count=128),
        "oblique_unnormalised": seeded(17, direction=(-3.0, 4.0, -12.0), ray_count=128),
        "flat_directivity": seeded(18, directivity=_flat, power=10.0, ray_count=200),
        "lobed_directivity": seeded(19, directivity=_lobe, power=10.0, ray_count=200),
        "collimated": seeded(20, direction=(0, 0.01, 1), directivity=_flat, power=1000.0, ray_count=150,
                             post=lambda s: s.random_collimated_rays(diameter=5.0)),
        "collimated_offset": seeded(21, center=c, direction=(1, 1, 0), ray_count=90,
                                    post=lambda s: s.random_collimated_rays(diameter=0.5)),
        "grid": seeded(22, ray_count=50, post=lambda s: s.grid_rays()),
        "grid_oblique": seeded(23, center=c, direction=(0.3, -0.2, 0.5), power=7.0, ray_count=144, post=lambda s: s.grid_rays()),
        "rotated_z": seeded(24, ray_count=80, post=lambda s: s.rotate_rays(axis="z", pivot=[0, 0, 0, 0], ang=np.pi / 3.0)),
        "rotated_x_pivot": seeded(25, center=c, ray_count=80,
                                  post=lambda s: s.rotate_rays(axis="x", pivot=[1.0, 2.0, 3.0, 0], ang=-0.7)),
        "rotated_y": seeded(26, ray_count=80, post=lambda s: s.rotate_rays(axis="y", ang=2.1)),
    }


def geometry_cases(goe):
    """name -> zero-argument builder of a GeoObject."""
    oe = goe.optical_elements()
    curve = [[0.0, 0.0], [1.0, 2.0], [3.0, 2.5], [4.0, 0.5]]

    def with_(m, f):
        f(m)
        return m

    def lens2r(sign):
        def make():
            return oe.revolve_curve(oe.lens_spherical_2r((0.5, 0.0), 60.0, 600.0, 25.0, sign, 1.5), axis="x", ang=np.pi, ang_pts=9)

        return make

    def curve_lens(direction, sign):
        def make():
            xy, f, d = oe.curve_lens_spherical((0, 0), 40.0, 20.0, direction, sign, 1.6)
            return oe.revolve_curve(xy, axis="x", ang=2 * np.pi, ang_pts=7)

        return make

    def merged():
        a = oe.cube(center=(0, 0, 0, 0), size=[1, 2, 3, 0])
        b = oe.cube(center=(5, 0, 0, 0), size=[1, 1, 1, 0])
        a.append(b.vertices, b.triangles)
        return a

    def biconcave_curve(axis):
        def make():
            out = oe.curve_lens_spherical_biconcave((0, 0), 60.0, 6000.0, 3.0, 50.0, axis, 2.5)
            xy = out[0] if isinstance(out, tuple) else out
            return oe.revolve_curve(xy, axis="x", ang=np.pi, ang_pts=5)

        return make

    return {
        "cube": lambda: oe.cube(center=(0, 0, 20, 0), size=[100, 100, 10, 0]),
        "cube_offset": lambda: oe.cube(center=(-20, 0, 22.5 * (1.0 - 1e-6), 0), size=[30, 30, 5, 0]),
        "sphere": lambda: oe.sphere(center=[1.0, -2.0, 3.0, 0], radius=100.0),
        "hemisphere_1000": lambda: oe.hemisphere(center=[0, 0, 0, 0], radius=1000.0),
        "hemisphere_retina": lambda: oe.hemisphere(center=[0, 0, 12.1, 0], radius=12.1 * (1.0 - 1e-3)),
        "parabolic_mirror": lambda: oe.parabolic_mirror(focus=(0, 0, 0), focal_length=5.0, diameter=20.0, reflectivity=0.98),
        "parabolic_mirror_small": lambda: oe.parabolic_mirror(focus=(1, 2, 3), focal_length=2.0, diameter=6.0),
        "topless_cylinder": lambda: oe.topless_cylinder(center=(0, 0, 0), diameter=20.0, height=10.0),
        "revolve_x_half": lambda: oe.revolve_curve(curve, axis="x", ang=np.pi, ang_pts=5),
        "revolve_y_full": lambda: oe.revolve_curve(curve, axis="y", ang=2 * np.pi, ang_pts=36),
        "revolve_z_partial": lambda: oe.revolve_curve(curve, axis="z", ang=1.0, ang_pts=4),
        "extrude_open": lambda: oe.extrude_by_vector(curve, [0.5, -0.25, 3.0], capped=False),
        "lens_biconcave": lambda: oe.lens_spherical_biconcave((0, 0, 0), 60.0, 6000.0, 50.0, 2.5),
        "lens_2r_plus": lens2r(1),
        "lens_2r_minus": lens2r(-1),
        "curve_lens_fwd_plus": curve_lens(1, 1),
        "curve_lens_back_minus": curve_lens(-1, -1),
        "curve_lens_biconcave_x": biconcave_curve("x"),
        "cornea": lambda: oe.spherical_lens_nofoc(r1=7.8, r2=6.5, x1=0.0, x2=0.55, d=5.0),
        "eye_lens": lambda: oe.spherical_lens_nofoc(r1=10.2, r2=-6.0, x1=3.6, x2=7.6, d=5.0),
        "nofoc_d2_signs": lambda: oe.spherical_lens_nofoc(r1=9.0, r2=7.0, x1=0.0, x2=1.0, d=4.0, d2=3.0, sign1_arcsin=-1.0,
                                                        sign2_arcsin=1.0),
        "translated": lambda: with_(oe.cube(center=(0, 0, 0, 0), size=[1, 2, 3, 0]), lambda m: m.translate(np.array([0.1, -0.2, 0.3, 0], dtype=np.float32))),
        "rotated_y": lambda: with_(oe.lens_spherical_biconcave((0, 0, 0), 60.0, 6000.0, 50.0, 2.5),
                                   lambda m: m.rotate(axis="y", angle=-np.pi / 2.0, pivot=(0, 0, 0, 0))),
        "rotated_x_pivot": lambda: with_(oe.cube(center=(0, 0, 0, 0), size=[1, 2, 3, 0]), lambda m: m.rotate(axis="x", angle=0.3, pivot=(1, 1, 1, 0))),
        "rotated_z": lambda: with_(oe.cube(center=(3, 0, 0, 0), size=[1, 2, 3, 0]), lambda m: m.rotate(axis="z", angle=-1.1)),
        "appended": merged,
        "material_mirror": lambda: with_(oe.cube(center=(0, 0, 0, 0), size=[1, 1, 1, 0]), lambda m: m.setMaterial(mat_type="mirror", reflectivity=0.5)),
        "material_lossy": lambda: with_(oe.cube(center=(0, 0, 0, 0), size=[1, 1, 1, 0]),
                                        lambda m: m.setMaterial(mat_type="refractive", IOR=1.5, dissipation=0.1)),
        "material_measure": lambda: with_(oe.sphere(center=[0, 0, 0, 0], radius=2.0), lambda m: m.setMaterial(mat_type="measure")),
        "material_terminator": lambda: with_(oe.cube(center=(0, 0, 0, 0), size=[1, 1, 1, 0]), lambda m: m.setMaterial(mat_type="terminator")),
    }


def trace_cases():
    """name -> which lightpycl_b200.scenes builder, its arguments, and what to record after the trace."""
    post = dict(post=True, points=40, limits_xy=((-1000, 1000), (-1000, 1000)), replicas=5, hwhm_points=60)
    return {
        "c1": dict(scene="nested_cubes", kw=dict(ray_count=300)),
        "c1_partitions": dict(scene="nested_cubes", kw=dict(ray_count=60_000, trace_iterations=3), digest_only=True),
        "c2": dict(scene="directivity_lens", kw=dict(ray_count=1500), **post),
        "c3": dict(scene="parabolic_mirror", kw=dict(ray_count=1500, profile_pts=40, ang_pts=17),
                   **dict(post, limits_xy=((-500, 500), (-500, 500)))),
        "c4": dict(scene="human_eye", kw=dict(ray_count=120, sources=2, trace_iterations=7)),
        "lossy": dict(scene="dissipative_cube", kw=dict(ray_count=600, ior=1.3, dissipation=0.05)),
        "cavity": dict(scene="glass_cube_cavity", kw=dict(ray_count=40, trace_iterations=40)),
    }
