"""CPU oracle driver for the LightPyCL trace step.  TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``lightpycl_b200``) never imports it and has no CPU fallback.

Two interchangeable kernel backends, same call signatures:

``"reference"``  oracle/_ref/libref_strict.so - the UNMODIFIED reference kernel file
                 (kernel_reflect_refract_intersect.cl) compiled by path behind cl_shim.h,
                 strict IEEE, no contraction.  Built in the authoring container by
                 ``make -C oracle ref``; the .so travels to the GPU box, the source does not.
``"port"``       oracle/_build/liboracle_port.so - trace_oracle.c, our plain-C restatement.
                 Pinned bit-for-bit to ``"reference"`` by tests/test_oracle_pinning.py.
(``"reference_fast"`` / ``"port_fast"`` are the -O3 timing builds of the same sources.)

``project()`` / ``bin2d()`` check the measured-ray post-processing:
``"reference"`` = oracle/_ref/libref_proj.so, the reference's own ``stereograph_project`` /
``angular_project`` kernels (.cl:488-538) compiled from the reference source (its OpenCL vector
literals rewritten on the fly, see proj_shim.cpp); ``"port"`` = the NumPy restatement below,
pinned to it by tests/test_oracle_projection.py.  ``bin2d`` is ``np.histogram2d`` exactly as the
reference calls it (it:416,529,560).

On top of the kernels, ``trace()`` restates the reference's host loop
(iterative_tracer.py:77-393): source/mesh flattening (it:99-151), the three launches per
generation (it:288/303/318; the <=50,000-ray chunking of it:218-273 is omitted because rays
are independent), results retention (it:355), child pruning and ordering (it:366-373) and
the two termination tests (it:372-391).  ``measured_rays()`` restates it:395-411.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {
    "reference_proj": os.path.join(HERE, "_ref", "libref_proj.so"),
    "reference": os.path.join(HERE, "_ref", "libref_strict.so"),
    "reference_fast": os.path.join(HERE, "_ref", "libref_fast.so"),
    "port": os.path.join(HERE, "_build", "liboracle_port.so"),
    "port_fast": os.path.join(HERE, "_build", "liboracle_port_fast.so"),
}
_loaded = {}

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build(quiet=True):
    """Compile the checkers (``make -C oracle all``): the port always, the shim-compiled
    reference only where /root/reference exists.  Building the checker is not using it."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def available(backend):
    return os.path.exists(_LIBS[backend])


def _lib(backend):
    if backend not in _loaded:
        path = _LIBS[backend]
        if not os.path.exists(path):
            if backend.startswith("port"):
                build()
            if not os.path.exists(path):
                raise FileNotFoundError("oracle backend %r not built: %s" % (backend, path))
        _loaded[backend] = C.CDLL(path)
    return _loaded[backend]


def set_threads(backend, n=None):
    """Use n (default: all) host threads in the OpenMP loops of `backend`; returns the count in use."""
    lib = _lib(backend)
    pre = "ref" if backend.startswith("reference") else "orc"
    getattr(lib, pre + "_set_threads")(C.c_int(int(n or os.cpu_count() or 1)))
    return int(getattr(lib, pre + "_max_threads")())


def _f(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_fp)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


class Scene:
    """Flattened scene exactly as the reference host loop builds it (it:121-151)."""

    def __init__(self, meshes):
        self.mesh_count = len(meshes)
        self.mat_type = np.zeros(self.mesh_count, dtype=np.int32)
        self.ior = np.zeros(self.mesh_count, dtype=np.float32)
        self.refl = np.zeros(self.mesh_count, dtype=np.float32)
        self.diss = np.zeros(self.mesh_count, dtype=np.float32)
        v0s, v1s, v2s, ids = [], [], [], []
        for j, mesh in enumerate(meshes):
            mat = mesh.getMaterialBuf()
            self.mat_type[j] = np.int32(mat.get("type"))
            self.ior[j] = np.float32(mat.get("IOR"))
            self.refl[j] = np.float32(mat.get("R"))
            self.diss[j] = np.float32(mat.get("dissipation"))
            tb = mesh.tribuf()
            v0s.append(np.array(tb[0], dtype=np.float32).reshape(-1, 4))
            v1s.append(np.array(tb[1], dtype=np.float32).reshape(-1, 4))
            v2s.append(np.array(tb[2], dtype=np.float32).reshape(-1, 4))
            ids.append(np.zeros(len(tb[0]), dtype=np.int32) + j)
        self.v0 = np.ascontiguousarray(np.concatenate(v0s, axis=0), dtype=np.float32)
        self.v1 = np.ascontiguousarray(np.concatenate(v1s, axis=0), dtype=np.float32)
        self.v2 = np.ascontiguousarray(np.concatenate(v2s, axis=0), dtype=np.float32)
        self.tri_mesh = np.ascontiguousarray(np.concatenate(ids, axis=0), dtype=np.int32)
        self.tri_count = len(self.v0)


def flatten_sources(light_source):
    """it:99-118 -> (origin (N,4) f32, dir (N,4) f32, power f32 [(N,1) or (N,)], input_power)."""
    origin = direc = power = None
    for k, light in enumerate(light_source):
        if k == 0:
            origin = np.float32(light.rays_origin)
            direc = np.float32(light.rays_dir)
            power = np.float32(light.rays_power)
        else:
            origin = np.append(origin, light.rays_origin, axis=0).astype(np.float32)
            direc = np.append(direc, light.rays_dir, axis=0).astype(np.float32)
            power = np.append(power, light.rays_power, axis=0).astype(np.float32)
    direc = np.ascontiguousarray(np.asarray(direc), dtype=np.float32)
    # a single grid_rays source carries surplus origin rows (ls:46-49); the loop slices [0, ray_count) (it:276-283)
    origin = np.ascontiguousarray(np.asarray(origin)[: len(direc)], dtype=np.float32)
    power = np.asarray(power, dtype=np.float32)
    return origin, direc, power, seq_sum_sorted_f32(power)


def seq_sum_sorted_f32(p):
    """``sum(np.sort(p))`` of the reference (it:115,372): np.sort along the LAST axis (a no-op
    for the (N,1) power of random_rays), then Python's left-to-right float32 accumulation."""
    s = np.sort(np.asarray(p, dtype=np.float32))
    flat = s.reshape(-1)
    if flat.size == 0:
        return np.float32(0.0)
    return np.float32(np.cumsum(flat, dtype=np.float32)[-1])  # cumsum accumulates sequentially


# ----------------------------------------------------------------------------- kernel launches
def intersect(scene, origin, direc, max_ray_len, backend="port"):
    """reference `intersect` (cl:243) -> (near_t (n,M) f32, hits (n,M) i32, near_tri (n,M) i32).
    near_t is pre-filled with max_ray_len (it:267); hits/near_tri start at 0 like the
    reference's zero-initialised buffers (it:236-237)."""
    n, M = len(origin), scene.mesh_count
    near_t = np.zeros((n, M), dtype=np.float32) + np.float32(max_ray_len)
    hits = np.zeros((n, M), dtype=np.int32)
    near_tri = np.zeros((n, M), dtype=np.int32)
    lib = _lib(backend)
    if backend.startswith("reference"):
        dest = np.zeros((n, 4), dtype=np.float32)
        dummy = np.zeros(max(n, 1), dtype=np.int32)
        lib.ref_launch_intersect(C.c_int(n), _f(origin), _f(direc), _f(dest), _i(dummy), _i(dummy), _i(dummy),
                                 _f(scene.v0), _f(scene.v1), _f(scene.v2), _i(scene.tri_mesh), _f(near_t),
                                 _i(hits), _i(near_tri), C.c_int(M), C.c_int(scene.tri_count), C.c_int(n),
                                 C.c_float(max_ray_len))
    else:
        lib.orc_intersect(C.c_int(n), _f(origin), _f(direc), _f(scene.v0), _f(scene.v1), _f(scene.v2),
                          _i(scene.tri_mesh), _f(near_t), _i(hits), _i(near_tri), C.c_int(M),
                          C.c_int(scene.tri_count), C.c_float(max_ray_len))
    return near_t, hits, near_tri


def postproc(scene, origin, direc, prev_mesh, near_t, hits, near_tri, max_ray_len, backend="port"):
    """reference `intersect_postproc` (cl:105) -> dict(dest, n1, n2, entering, hit_mesh, hit_tri)."""
    n, M = len(origin), scene.mesh_count
    dest = np.zeros((n, 4), dtype=np.float32)
    n1 = np.zeros(n, dtype=np.int32)
    n2 = np.zeros(n, dtype=np.int32)
    entering = np.zeros(n, dtype=np.int32)
    hit_mesh = np.zeros(n, dtype=np.int32)
    hit_tri = np.zeros(n, dtype=np.int32)
    prev_mesh = np.ascontiguousarray(prev_mesh, dtype=np.int32)
    lib = _lib(backend)
    if backend.startswith("reference"):
        lib.ref_launch_intersect_postproc(C.c_int(n), _f(origin), _f(direc), _f(dest), _i(prev_mesh), _i(n1), _i(n2),
                                          _i(entering), _i(hit_mesh), _i(hit_tri), _f(scene.v0), _f(scene.v1),
                                          _f(scene.v2), _i(scene.tri_mesh), _i(scene.mat_type), _f(near_t),
                                          _i(hits), _i(near_tri), C.c_int(M), C.c_int(n), C.c_float(max_ray_len))
    else:
        lib.orc_postproc(C.c_int(n), _f(origin), _f(direc), _f(dest), _i(prev_mesh), _i(n1), _i(n2), _i(entering),
                         _i(hit_mesh), _i(hit_tri), _i(scene.mat_type), _f(near_t), _i(hits), _i(near_tri),
                         C.c_int(M), C.c_float(max_ray_len))
    return dict(dest=dest, n1=n1, n2=n2, entering=entering, hit_mesh=hit_mesh, hit_tri=hit_tri)


def reflect_refract(scene, origin, dest, direc, power, measured, n1, n2, hit_mesh, hit_tri, ior_env, max_ray_len,
                    backend="port"):
    """reference `reflect_refract_rays` (cl:346).  ``power`` and ``measured`` are updated IN
    PLACE (cl:392,467,470).  Returns the two uncompacted child sets."""
    n, M = len(origin), scene.mesh_count
    ch = {}
    for k in ("refl", "refr"):
        ch[k + "_origin"] = np.zeros((n, 4), dtype=np.float32)
        ch[k + "_dir"] = np.zeros((n, 4), dtype=np.float32)
        ch[k + "_power"] = np.zeros(n, dtype=np.float32)
        ch[k + "_meas"] = np.zeros(n, dtype=np.int32)
    lib = _lib(backend)
    if backend.startswith("reference"):
        dummy = np.zeros(max(n, 1), dtype=np.int32)
        lib.ref_launch_reflect_refract_rays(
            C.c_int(n), _f(origin), _f(dest), _f(direc), _f(power), _i(measured), _i(dummy), _i(n1), _i(n2),
            _f(ch["refl_origin"]), _f(ch["refl_dir"]), _f(ch["refl_power"]), _i(ch["refl_meas"]),
            _f(ch["refr_origin"]), _f(ch["refr_dir"]), _f(ch["refr_power"]), _i(ch["refr_meas"]),
            _i(hit_mesh), _i(hit_tri), _f(scene.v0), _f(scene.v1), _f(scene.v2), _i(scene.tri_mesh),
            _i(scene.mat_type), _f(scene.ior), _f(scene.refl), _f(scene.diss), C.c_float(ior_env), C.c_int(M),
            C.c_int(n), C.c_float(max_ray_len))
    else:
        lib.orc_reflect_refract(
            C.c_int(n), _f(origin), _f(dest), _f(direc), _f(power), _i(measured), _i(n1), _i(n2),
            _f(ch["refl_origin"]), _f(ch["refl_dir"]), _f(ch["refl_power"]), _i(ch["refl_meas"]),
            _f(ch["refr_origin"]), _f(ch["refr_dir"]), _f(ch["refr_power"]), _i(ch["refr_meas"]),
            _i(hit_mesh), _i(hit_tri), _f(scene.v0), _f(scene.v1), _f(scene.v2), _i(scene.mat_type), _f(scene.ior),
            _f(scene.refl), _f(scene.diss), C.c_float(ior_env))
    return ch


def ray_triangle(O, D, V0, V1, V2, backend="port"):
    """Single ray x triangle probe (cl:50-101) -> (hit, t, u, v)."""
    a = [np.ascontiguousarray(x, dtype=np.float32) for x in (O, D, V0, V1, V2)]
    t, u, v = C.c_float(0), C.c_float(0), C.c_float(0)
    lib = _lib(backend)
    fn = lib.ref_intersect_triangle if backend.startswith("reference") else lib.orc_ray_triangle
    hit = fn(*[_f(x) for x in a], C.byref(t), C.byref(u), C.byref(v))
    return int(hit), t.value, u.value, v.value


# ----------------------------------------------------------------------------- host loop
def trace(light_source, meshes, trace_iterations=100, trace_until_dissipated=0.99, max_ray_len=np.float32(1e3),
          ior_env=np.float32(1.0), backend="port", keep_debug=True, power_sum="reference", lineage0=None):
    """Restatement of CL_Tracer.iterative_tracer (it:77-393).

    Returns a dict with
      results   list per iteration of (rays_origin (n,4), rays_dest (n,4), rays_pow, rays_meas i32)
                exactly like the reference's ``self.results`` (iteration-0 power keeps the source's
                shape, e.g. (N,1))
      debug     list per iteration of dict(hit_mesh, hit_tri, n1, n2, entering, near_t, hits, near_tri,
                dir, prev_mesh) when keep_debug
      live      list per iteration of surviving child count; power_left: list of the termination sums
      input_power, tri_count, scene
    ``power_sum``: "reference" = sorted sequential float32 sum (it:372); "f64" = float64 sum, the
    rule the GPU path uses.  The break test itself is evaluated in float64 for both.
    ``lineage0``: optional int64 id per source ray; the result then carries ``lineage`` (one int64
    array per iteration, aligned with ``results``): a child keeps its parent's id, the refracted
    child of iteration k additionally gets bit 32 + k - the id the CUDA path carries, which lets a
    test match the rays of a sub-sampled oracle run with those of a full-size GPU run.  Every 31
    iterations (the branch bits of a 64-bit id) a new epoch starts: ids become 0..n-1 in generation
    order and the branch bits start over, as lightpycl_b200.dist.renumber_lineage does.
    """
    max_ray_len = np.float32(max_ray_len)
    ior_env = np.float32(ior_env)
    origin, direc, power, input_power = flatten_sources(light_source)
    n = len(origin)
    rays_pow = np.array(power, dtype=np.float32)
    rays_meas = np.zeros(n, dtype=np.int32)
    prev_mesh = np.zeros(n, dtype=np.int32) - 2
    scene = Scene(meshes)

    results, debug, live, power_left, lineages = [], [], [], [], []
    lin = None if lineage0 is None else np.asarray(lineage0, dtype=np.int64).copy()
    epoch_start = 0
    for k_iter in range(int(trace_iterations)):
        if lin is not None and k_iter - epoch_start >= 31:
            lin, epoch_start = np.arange(len(lin), dtype=np.int64), k_iter
        n = len(origin)
        pw = np.ascontiguousarray(rays_pow.reshape(-1), dtype=np.float32)  # kernels see a flat float buffer
        ms = np.ascontiguousarray(rays_meas, dtype=np.int32)
        near_t, hits, near_tri = intersect(scene, origin, direc, max_ray_len, backend)
        pp = postproc(scene, origin, direc, prev_mesh, near_t, hits, near_tri, max_ray_len, backend)
        ch = reflect_refract(scene, origin, pp["dest"], direc, pw, ms, pp["n1"], pp["n2"], pp["hit_mesh"],
                             pp["hit_tri"], ior_env, max_ray_len, backend)
        results.append((origin, pp["dest"], pw.reshape(rays_pow.shape), ms))
        if lin is not None:
            lineages.append(lin)
        if keep_debug:
            d = dict(pp)
            d.update(near_t=near_t, hits=hits, near_tri=near_tri, dir=direc, prev_mesh=prev_mesh)
            debug.append(d)

        # it:366-373: children = reflected then refracted, parent order kept, survivors have measured == 0
        child_meas = np.concatenate((ch["refl_meas"], ch["refr_meas"]), axis=0)
        idx = np.where(child_meas == 0)[0]
        origin = np.ascontiguousarray(np.concatenate((ch["refl_origin"], ch["refr_origin"]), axis=0)[idx])
        direc = np.ascontiguousarray(np.concatenate((ch["refl_dir"], ch["refr_dir"]), axis=0)[idx])
        rays_pow = np.ascontiguousarray(np.concatenate((ch["refl_power"], ch["refr_power"]), axis=0)[idx])
        rays_meas = np.ascontiguousarray(child_meas[idx].astype(np.int32))
        prev_mesh = np.ascontiguousarray(np.concatenate((pp["hit_mesh"], pp["hit_mesh"]), axis=0)[idx])
        if lin is not None:
            branch = np.int64(1) << np.int64(32 + k_iter - epoch_start)
            lin = np.concatenate((lin, lin | branch), axis=0)[idx]

        if power_sum == "reference":
            p_scene = float(seq_sum_sorted_f32(rays_pow))
        else:
            p_scene = float(np.sum(rays_pow.astype(np.float64)))
        live.append(len(origin))
        power_left.append(p_scene)
        if p_scene < (1.0 - float(trace_until_dissipated)) * float(np.asarray(input_power).reshape(-1)[0]):
            break  # it:383-387
        if len(origin) == 0:
            break  # it:389-391
    return dict(results=results, debug=debug, live=live, power_left=power_left, input_power=input_power,
                tri_count=scene.tri_count, scene=scene, lineage=lineages)


def measured_rays(results):
    """it:395-411: end points and powers of every ray flagged measured (>= .9), iteration-major."""
    pos = pwr = None
    for k, (_o, dest, p, meas) in enumerate(results):
        idx = np.where(meas >= 0.9)[0]
        d, q = dest[idx], p[idx]
        if k == 0:
            pos, pwr = d, q
        else:
            pos = np.concatenate((pos, d), axis=0)
            pwr = np.concatenate((pwr.flatten(), q.flatten()), axis=0)
    return pos, pwr


# ----------------------------------------------------------------------------- measured-ray post-processing
_F = np.float32
PROJECTIONS = ("xy", "stereographic", "angular")


def _project_reference(kind, pos, pwr, rot):
    """The reference kernels themselves (libref_proj.so), launched like it:515 / it:546: pivot (0,0,0,0),
    R = the 4x4 float32 matrix whose upper-left 3x3 is `rot` (identity by default, it:514,545)."""
    lib = _lib("reference_proj")
    n = len(pwr)
    vecs = np.ascontiguousarray(np.asarray(pos, dtype=_F).reshape(-1, 4))
    r4 = np.zeros((4, 4), dtype=_F)
    r4[:3, :3] = np.eye(3, dtype=_F) if rot is None else np.asarray(rot, dtype=_F)[:3, :3]
    pivot = np.zeros((1, 4), dtype=_F)
    x, y, w = (np.zeros(n, dtype=_F) for _ in range(3))
    fn = lib.ref_launch_stereograph_project if kind == "stereographic" else lib.ref_launch_angular_project
    fn(C.c_int(n), _f(vecs), _f(pwr), _f(r4), _f(pivot), _f(x), _f(y), _f(w))
    return x, y, w


def _project_port(kind, pos, pwr, rot):
    """NumPy restatement of .cl:488-538: float32 in the kernels' operation order, double where OpenCL C
    promotes (the literals 1.0 / 4.0 of .cl:501)."""
    p = np.asarray(pos, dtype=_F)[:, :3]
    r = np.eye(3, dtype=_F) if rot is None else np.asarray(rot, dtype=_F)[:3, :3]
    # dot(vec, rot_mtx[k]) = (x*r0 + y*r1) + z*r2, each product and sum rounded to float32 (pivot = 0)
    u, v, w = [(p[:, 0] * r[k, 0] + p[:, 1] * r[k, 1]) + p[:, 2] * r[k, 2] for k in range(3)]
    l = np.sqrt((u * u + v * v) + w * w, dtype=_F)
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == "stereographic":  # .cl:499-505
            xt = u / (l + w)
            yt = v / (l + w)
            s = (1.0 + (xt * xt).astype(np.float64)) + (yt * yt).astype(np.float64)
            a = (4.0 / (s * s)).astype(_F)
            return xt.astype(_F), yt.astype(_F), (pwr / a).astype(_F)
        # angular, .cl:519-537: vecr = (u,v,w)/l; cosT = dot((0,0,1), vecr) = (0*x + 0*y) + 1*z
        cos_t = (_F(0) * (u / l) + _F(0) * (v / l)) + _F(1) * (w / l)
        phi = np.arctan2(v, u, dtype=_F)
        rr = np.arccos(cos_t, dtype=_F)
        xt = rr * np.cos(phi, dtype=_F)
        yt = rr * np.sin(phi, dtype=_F)
        a = np.ones_like(rr)
        big = rr > _F(0.000001)
        a[big] = np.sin(rr[big], dtype=_F) / rr[big]
        return xt.astype(_F), yt.astype(_F), (pwr / a).astype(_F)


def project(kind, pos, pwr, rot=None, backend=None):
    """-> (x, y, corrected power), float32, like the reference kernels' x / y / pwrs_cor buffers
    (kind "xy": the plain x, y, power that get_binned_data bins, it:413-418)."""
    pwr = np.ascontiguousarray(np.asarray(pwr, dtype=_F).reshape(-1))
    if kind == "xy":
        p = np.asarray(pos, dtype=_F)
        return p[:, 0].copy(), p[:, 1].copy(), pwr
    if kind not in PROJECTIONS:
        raise ValueError("unknown projection %r" % (kind,))
    if backend is None:
        backend = "reference" if available("reference_proj") else "port"
    return (_project_reference if backend == "reference" else _project_port)(kind, pos, pwr, rot)


def bin2d(x, y, w, limits, points, kind):
    """np.histogram2d exactly as the reference calls it: raw weights for get_binned_data (it:416),
    float64(pwr) / (dx * dy) for the two projections (it:525-529, 556-560)."""
    if kind == "xy":
        weights = np.asarray(w)
    else:
        dx = np.float64(limits[0][1] - limits[0][0]) / np.float64(points)
        dy = np.float64(limits[1][1] - limits[1][0]) / np.float64(points)
        weights = np.float64(w) / (dx * dy)
    return np.histogram2d(x=np.asarray(x).flatten(), y=np.asarray(y).flatten(), bins=points, range=limits,
                          weights=np.asarray(weights).flatten())
