"""Seam-level parity on the GPU: each C-ABI entry point with the reference kernels' buffer
semantics (lpcl_intersect / lpcl_intersect_postproc / lpcl_reflect_refract_rays) is fed the
exact per-iteration inputs of an oracle trace and must reproduce the oracle's outputs
bit-for-bit - the same seams as the reference's three launches (iterative_tracer.py:288/303/318)."""

import ctypes as C
import glob
import os

import numpy as np
import pytest

import util
from oracle import oracle

pytestmark = pytest.mark.gpu

# the checker: the reference kernels themselves (oracle/_ref, compiled from the reference .cl) where present,
# else the C restatement pinned to them on the CPU (tests/test_oracle_pinning.py)
BACKEND = "reference" if oracle.available("reference") else "port"


class GpuSeams:
    def __init__(self, scene):
        import torch

        from lightpycl_b200 import _cabi

        self.torch, self.lib, self.cabi = torch, _cabi.load(), _cabi
        self.dev = torch.device("cuda:0")
        self.M = scene.mesh_count
        self.v = [torch.from_numpy(a).to(self.dev) for a in (scene.v0, scene.v1, scene.v2)]
        self.handle = C.c_void_p()
        rc = self.lib.lpcl_scene_create(C.byref(self.handle), self.v[0].data_ptr(), self.v[1].data_ptr(),
                                        self.v[2].data_ptr(), scene.tri_mesh.ctypes.data, scene.tri_count,
                                        scene.mat_type.ctypes.data, scene.ior.ctypes.data, scene.refl.ctypes.data,
                                        scene.diss.ctypes.data, scene.mesh_count, None)
        _cabi.check(rc, "lpcl_scene_create")

    def close(self):
        self.lib.lpcl_scene_destroy(self.handle)

    def t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def intersect(self, origin, direc, max_ray_len):
        torch = self.torch
        n = len(origin)
        o, d = self.t(origin), self.t(direc)
        near_t = torch.empty((n, self.M), dtype=torch.float32, device=self.dev)
        hits = torch.empty((n, self.M), dtype=torch.int32, device=self.dev)
        near_tri = torch.empty((n, self.M), dtype=torch.int32, device=self.dev)
        scratch = torch.empty(self.lib.lpcl_scratch_bytes(self.handle, n), dtype=torch.uint8, device=self.dev)
        rc = self.lib.lpcl_intersect(self.handle, o.data_ptr(), d.data_ptr(), n, C.c_float(max_ray_len),
                                     near_t.data_ptr(), hits.data_ptr(), near_tri.data_ptr(), scratch.data_ptr(), None)
        self.cabi.check(rc, "lpcl_intersect")
        torch.cuda.synchronize()
        return near_t.cpu().numpy(), hits.cpu().numpy(), near_tri.cpu().numpy()

    def postproc(self, origin, direc, prev, near_t, hits, near_tri, max_ray_len):
        torch = self.torch
        n = len(origin)
        i32 = lambda: torch.full((n,), -7, dtype=torch.int32, device=self.dev)
        dest = torch.zeros((n, 4), dtype=torch.float32, device=self.dev)
        n1, n2, ent, hm, ht = i32(), i32(), torch.zeros(n, dtype=torch.int32, device=self.dev), i32(), i32()
        args = [self.t(origin), self.t(direc), dest, self.t(prev.astype(np.int32)), n1, n2, ent, hm, ht,
                self.t(near_t), self.t(hits), self.t(near_tri)]
        rc = self.lib.lpcl_intersect_postproc(self.handle, *[a.data_ptr() for a in args], n, C.c_float(max_ray_len), None)
        self.cabi.check(rc, "lpcl_intersect_postproc")
        torch.cuda.synchronize()
        return dict(dest=dest.cpu().numpy(), n1=n1.cpu().numpy(), n2=n2.cpu().numpy(), entering=ent.cpu().numpy(),
                    hit_mesh=hm.cpu().numpy(), hit_tri=ht.cpu().numpy())

    def reflect_refract(self, origin, dest, direc, power, measured, n1, n2, hit_mesh, hit_tri, ior_env):
        torch = self.torch
        n = len(origin)
        f4 = lambda: torch.zeros((n, 4), dtype=torch.float32, device=self.dev)
        f1 = lambda: torch.zeros(n, dtype=torch.float32, device=self.dev)
        i1 = lambda: torch.zeros(n, dtype=torch.int32, device=self.dev)
        pw, ms = self.t(power.reshape(-1)), self.t(measured)
        out = [f4(), f4(), f1(), i1(), f4(), f4(), f1(), i1()]
        args = [self.t(origin), self.t(dest), self.t(direc), pw, ms, self.t(n1), self.t(n2)] + out + \
               [self.t(hit_mesh), self.t(hit_tri)]
        rc = self.lib.lpcl_reflect_refract_rays(self.handle, *[a.data_ptr() for a in args], n, C.c_float(ior_env), None)
        self.cabi.check(rc, "lpcl_reflect_refract_rays")
        torch.cuda.synchronize()
        keys = ("refl_origin", "refl_dir", "refl_power", "refl_meas", "refr_origin", "refr_dir", "refr_power", "refr_meas")
        res = {k: v.cpu().numpy() for k, v in zip(keys, out)}
        res["power"], res["measured"] = pw.cpu().numpy(), ms.cpu().numpy()
        return res


def _cases():
    from lightpycl_b200 import scenes

    cases = {name: val for name, val in util.mini_scenes().items()}
    cases["c2_lens_1500"] = scenes.directivity_lens(ray_count=1500)
    cases["c3_mirror_1500"] = scenes.parabolic_mirror(ray_count=1500)
    cases["c4_eye_4x150"] = scenes.human_eye(ray_count=150)
    return cases


CASES = None


def _get(name):
    global CASES
    if CASES is None:
        CASES = _cases()
    return CASES[name]


NAMES = ["c1_cubes_10", "c1_cubes_300", "mini_lens", "mini_mirror", "mini_lossy", "mini_stack", "c2_lens_1500",
         "c3_mirror_1500", "c4_eye_4x150"]


@pytest.mark.parametrize("name", NAMES)
def test_three_seams_match_oracle(name):
    ls, meshes, params = _get(name)
    ref = oracle.trace(ls, meshes, backend=BACKEND, **params)
    g = GpuSeams(ref["scene"])
    try:
        mrl = np.float32(params["max_ray_len"])
        for k, (res, dbg) in enumerate(zip(ref["results"], ref["debug"])):
            origin, dest, power_after, meas_after = res
            direc, prev = dbg["dir"], dbg["prev_mesh"]
            # 1. search
            near_t, hits, near_tri = g.intersect(origin, direc, mrl)
            assert np.array_equal(hits, dbg["hits"]), (name, k, "hits")
            assert np.array_equal(near_tri, dbg["near_tri"]), (name, k, "near_tri")
            assert np.array_equal(util.bits(near_t), util.bits(dbg["near_t"])), (name, k, "near_t")
            # 2. media resolution
            pp = g.postproc(origin, direc, prev, dbg["near_t"], dbg["hits"], dbg["near_tri"], mrl)
            for key in ("n1", "n2", "hit_mesh", "hit_tri"):
                assert np.array_equal(pp[key], dbg[key]), (name, k, key)
            hit = dbg["hit_mesh"] >= 0
            assert np.array_equal(pp["entering"][hit], dbg["entering"][hit])
            assert np.array_equal(util.bits(pp["dest"][:, :3]), util.bits(dest[:, :3])), (name, k, "dest")
            # 3. Fresnel / materials: feed the parents' PRE-step power and flags
            if k == 0:
                p_in = np.ascontiguousarray(oracle.flatten_sources(ls)[2].reshape(-1))
            else:
                p_in = p_next
            m_in = np.zeros(len(origin), np.int32)
            want = oracle.reflect_refract(ref["scene"], origin, dest, direc, p_in.copy(), m_in.copy(), dbg["n1"],
                                          dbg["n2"], dbg["hit_mesh"], dbg["hit_tri"], np.float32(params["ior_env"]), mrl,
                                          BACKEND)
            got = g.reflect_refract(origin, dest, direc, p_in.copy(), m_in.copy(), dbg["n1"], dbg["n2"],
                                    dbg["hit_mesh"], dbg["hit_tri"], np.float32(params["ior_env"]))
            for key in ("refl_power", "refr_power", "refl_meas", "refr_meas"):
                assert np.array_equal(util.bits(got[key]), util.bits(want[key])), (name, k, key)
            for key in ("refl_origin", "refl_dir", "refr_origin", "refr_dir"):
                assert np.array_equal(util.bits(got[key][:, :3]), util.bits(want[key][:, :3])), (name, k, key)
            assert np.array_equal(util.bits(got["power"]), util.bits(power_after.reshape(-1))), (name, k, "parent power")
            assert np.array_equal(got["measured"], meas_after), (name, k, "parent measured")
            keep = np.concatenate((want["refl_meas"], want["refr_meas"])) == 0
            p_next = np.ascontiguousarray(np.concatenate((want["refl_power"], want["refr_power"]))[keep])
    finally:
        g.close()


SEARCH_ENV = ("LPCL_SEARCH_IMPL", "LPCL_SEARCH_R", "LPCL_SEARCH_RT", "LPCL_SEARCH_CHUNKS", "LPCL_TC_WARPS")


def _set_search(impl, R, chunks):
    """impl 'simt': R rays per thread; impl 'tc' (deferred flags + second pass, the default) / 'tc8' / 'tc16' (a vote
    per chunk test with fat / thin warps, LPCL_TC_WARPS=8 / 16): R ray tiles (128 rays) per CTA; 0 = heuristic."""
    for k in SEARCH_ENV:
        os.environ.pop(k, None)
    os.environ["LPCL_SEARCH_IMPL"] = "tc" if impl.startswith("tc") else impl
    if impl in ("tc8", "tc16"):
        os.environ["LPCL_TC_WARPS"] = impl[2:]
    if R:
        os.environ["LPCL_SEARCH_RT" if impl.startswith("tc") else "LPCL_SEARCH_R"] = str(R)
    if chunks:
        os.environ["LPCL_SEARCH_CHUNKS"] = str(chunks)


def _clear_search():
    for k in SEARCH_ENV:
        os.environ.pop(k, None)


@pytest.mark.parametrize("impl,R,chunks", [("simt", 1, 1), ("simt", 2, 1), ("simt", 4, 1), ("simt", 1, 7), ("simt", 2, 3),
                                           ("simt", 4, 64), ("tc", 1, 1), ("tc", 2, 1), ("tc", 4, 3), ("tc", 8, 1),
                                           ("tc", 1, 64), ("tc", 0, 0), ("tc8", 1, 1), ("tc8", 2, 1), ("tc8", 4, 3),
                                           ("tc8", 8, 1), ("tc8", 0, 0), ("tc16", 2, 1), ("tc16", 4, 3), ("tc16", 8, 1),
                                           ("tc16", 0, 0)])
def test_search_variants_agree(impl, R, chunks):
    """Both search kernels (packed-FP32 filter, tensor-core filter), their rays-per-CTA variants and
    triangle-range splits (atomic merge) give results identical to the oracle's."""
    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.directivity_lens(ray_count=5000)
    sc = oracle.Scene(meshes)
    origin, direc, _, _ = oracle.flatten_sources(ls)
    want = oracle.intersect(sc, origin, direc, params["max_ray_len"], BACKEND)
    g = GpuSeams(sc)
    _set_search(impl, R, chunks)
    try:
        got = g.intersect(origin, direc, np.float32(params["max_ray_len"]))
    finally:
        _clear_search()
        g.close()
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert np.array_equal(util.bits(got[0]), util.bits(want[0]))


def test_error_conventions():
    from lightpycl_b200 import _cabi, scenes

    import torch

    lib = _cabi.load()
    _, meshes, _ = scenes.nested_cubes()
    sc = oracle.Scene(meshes)
    v = [torch.from_numpy(a).cuda() for a in (sc.v0, sc.v1, sc.v2)]
    h = C.c_void_p()
    bad_ids = sc.tri_mesh.copy()
    bad_ids[bad_ids == 2] = 3  # skips an id -> the reference kernel would silently misbehave (cl:262-265)
    rc = lib.lpcl_scene_create(C.byref(h), v[0].data_ptr(), v[1].data_ptr(), v[2].data_ptr(), bad_ids.ctypes.data,
                               sc.tri_count, sc.mat_type.ctypes.data, sc.ior.ctypes.data, sc.refl.ctypes.data,
                               sc.diss.ctypes.data, sc.mesh_count, None)
    assert rc == -3 and b"mesh" in lib.lpcl_last_error()
    rc = lib.lpcl_scene_create(C.byref(h), v[0].data_ptr() + 4, v[1].data_ptr(), v[2].data_ptr(),
                               sc.tri_mesh.ctypes.data, sc.tri_count, sc.mat_type.ctypes.data, sc.ior.ctypes.data,
                               sc.refl.ctypes.data, sc.diss.ctypes.data, sc.mesh_count, None)
    assert rc == -4


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_resolve_and_scatter_seams_on_random_inputs(seed):
    """Stress the media-resolution window logic (cl:187-229) and the material handling (cl:346-474)
    far beyond what real scenes produce: random per-mesh scratch with exact ties, values inside /
    just outside the EPSILON window, random crossing parities, every material code (incl. 4 =
    anisotropic), negative and equal IORs, dissipative media, pre-set measured flags."""
    rng = np.random.RandomState(100 + seed)
    from lightpycl_b200 import scenes

    M = [3, 7, 12][seed]
    n = 20000
    # a scene with M one-triangle-ish meshes (geometry only matters for the normal of the hit triangle)
    class Blob:
        def __init__(self, j):
            self.j = j
            self.v = rng.uniform(-5, 5, (3, 4, 4)).astype(np.float32)
            self.v[:, :, 3] = 0
            self.mat = {"type": int(rng.choice([0, 0, 0, 1, 2, 3, 4])), "IOR": float(rng.choice([1.0, 1.5, 2.5, -2.0, 1.3374])),
                        "R": float(rng.uniform(0.5, 1.0)), "dissipation": float(rng.choice([0.0, 0.0, 0.07, 1e-7]))}

        def getMaterialBuf(self):
            return dict(self.mat)

        def tribuf(self):
            return (self.v[0], self.v[1], self.v[2])

    meshes = [Blob(j) for j in range(M)]
    sc = oracle.Scene(meshes)
    max_ray_len = np.float32(200.0)
    eps = np.float32(1e-6 * 200.0)
    origin = np.zeros((n, 4), np.float32)
    origin[:, :3] = rng.uniform(-10, 10, (n, 3))
    direc = np.zeros((n, 4), np.float32)
    d = rng.normal(size=(n, 3))
    direc[:, :3] = d / np.linalg.norm(d, axis=1)[:, None]
    # per-mesh nearest distances: a base value per ray plus offsets drawn from {0, +-eps/2, +-eps, +-2eps, big}, or "no hit"
    base = rng.uniform(1.0, 150.0, (n, 1)).astype(np.float32)
    offs = rng.choice([0.0, 0.5, -0.5, 1.0, -1.0, 2.0, 0.999, 1.001, 40.0, 1e4], size=(n, M)).astype(np.float32)
    near_t = (base + offs * eps).astype(np.float32)
    miss = rng.rand(n, M) < 0.35
    near_t[miss] = max_ray_len
    near_t[near_t >= max_ray_len] = max_ray_len
    hits = rng.randint(0, 5, (n, M)).astype(np.int32)
    hits[miss] = rng.randint(0, 2, miss.sum())  # crossings beyond max_ray_len still count (Q5)
    near_tri = np.full((n, M), -1, np.int32)
    for j in range(M):
        lo = int(np.searchsorted(sc.tri_mesh, j))
        hi = int(np.searchsorted(sc.tri_mesh, j, side="right"))
        near_tri[:, j] = rng.randint(lo, hi, n)
    near_tri[near_t >= max_ray_len] = -1
    prev = rng.randint(-2, M, n).astype(np.int32)

    want = oracle.postproc(sc, origin, direc, prev, near_t, hits, near_tri, max_ray_len, BACKEND)
    g = GpuSeams(sc)
    try:
        got = g.postproc(origin, direc, prev, near_t, hits, near_tri, max_ray_len)
        for key in ("n1", "n2", "hit_mesh", "hit_tri"):
            assert np.array_equal(got[key], want[key]), key
        hit = want["hit_mesh"] >= 0
        assert np.array_equal(got["entering"][hit], want["entering"][hit])
        assert np.array_equal(util.bits(got["dest"][:, :3]), util.bits(want["dest"][:, :3]))
        assert (want["n1"] >= 0).sum() > n // 10 and (want["n2"] >= 0).sum() > n // 10

        power = rng.uniform(0.0, 3.0, n).astype(np.float32)
        measured = rng.choice([0, 0, 0, 0, 1, -1], size=n).astype(np.int32)
        args = (origin, want["dest"], direc)
        w = oracle.reflect_refract(sc, *args, power.copy(), measured.copy(), want["n1"], want["n2"], want["hit_mesh"],
                                   want["hit_tri"], np.float32(1.0), max_ray_len, BACKEND)
        # the oracle call updates its power/measured arguments in place: redo with kept copies
        p_ref, m_ref = power.copy(), measured.copy()
        w = oracle.reflect_refract(sc, *args, p_ref, m_ref, want["n1"], want["n2"], want["hit_mesh"], want["hit_tri"],
                                   np.float32(1.0), max_ray_len, BACKEND)
        got = g.reflect_refract(origin, want["dest"], direc, power.copy(), measured.copy(), want["n1"], want["n2"],
                                want["hit_mesh"], want["hit_tri"], np.float32(1.0))
        ok = np.isfinite(w["refl_dir"]).all(axis=1) & np.isfinite(w["refr_dir"]).all(axis=1) & \
            np.isfinite(w["refl_power"]) & np.isfinite(w["refr_power"])  # NaN outputs are unspecified (cl:342)
        assert ok.mean() > 0.95
        def explain(key, a, b):
            bad = np.argwhere(util.bits(a) != util.bits(b))
            i = int(np.where(ok)[0][bad[0][0]])
            hm = int(want["hit_mesh"][i])
            return "%s: %d mismatches; ray %d got %r want %r | hit_mesh %d mat %s n1 %d n2 %d power_in %r measured_in %d " \
                   "power_out got %r want %r refl got %r want %r" % (
                       key, len(bad), i, got[key][i], w[key][i], hm, meshes[hm].mat if hm >= 0 else None, want["n1"][i],
                       want["n2"][i], power[i], measured[i], got["power"][i], p_ref[i], got["refl_power"][i], w["refl_power"][i])

        for key in ("refl_power", "refr_power", "refl_meas", "refr_meas"):
            assert np.array_equal(util.bits(got[key][ok]), util.bits(w[key][ok])), explain(key, got[key][ok], w[key][ok])
        for key in ("refl_origin", "refl_dir", "refr_origin", "refr_dir"):
            assert np.array_equal(util.bits(got[key][ok][:, :3]), util.bits(w[key][ok][:, :3])), \
                explain(key, got[key][ok][:, :3], w[key][ok][:, :3])
        assert np.array_equal(util.bits(got["power"][ok]), util.bits(p_ref[ok]))
        assert np.array_equal(got["measured"], m_ref)
    finally:
        g.close()


@pytest.mark.parametrize("seed,scale", [(0, 1.0), (1, 300.0), (2, 0.05), (3, 3.0e4), (4, 2.0e-3)])
def test_search_on_adversarial_triangle_soup(seed, scale):
    """The search (filter + exact re-evaluation + atomic merge) against the oracle on a random
    triangle soup with rays aimed AT edges, vertices and centroids, origins lying on other
    triangles (t ~ EPSILON), rays parallel to faces, reversed and non-unit directions, duplicated
    triangles (exact t ties between triangle indices and between meshes) and degenerate triangles."""
    rng = np.random.RandomState(7 + seed)
    T, M, n = 4000, 5, 12000
    v0 = rng.uniform(-10, 10, (T, 3)) * scale
    e1 = rng.normal(size=(T, 3)) * rng.choice([0.02, 0.3, 2.0], size=(T, 1)) * scale
    e2 = rng.normal(size=(T, 3)) * rng.choice([0.02, 0.3, 2.0], size=(T, 1)) * scale
    v1, v2 = v0 + e1, v0 + e2
    dup = rng.choice(T, 200, replace=False)           # exact duplicates => exact t ties
    src = rng.choice(T, 200)
    v0[dup], v1[dup], v2[dup] = v0[src], v1[src], v2[src]
    deg = rng.choice(T, 50, replace=False)
    v2[deg] = v1[deg]                                 # zero-area triangles
    tri_mesh = np.sort(rng.randint(0, M, T)).astype(np.int32)
    tri_mesh[:M] = np.arange(M)                       # every mesh non-empty, ascending
    tri_mesh = np.sort(tri_mesh)

    class Soup:
        pass

    sc = Soup()
    pad = lambda a: np.ascontiguousarray(np.concatenate((a, np.zeros((len(a), 1))), axis=1), dtype=np.float32)
    sc.v0, sc.v1, sc.v2, sc.tri_mesh = pad(v0), pad(v1), pad(v2), tri_mesh
    sc.tri_count, sc.mesh_count = T, M
    sc.mat_type = np.zeros(M, np.int32)
    sc.ior = np.ones(M, np.float32)
    sc.refl = np.ones(M, np.float32)
    sc.diss = np.zeros(M, np.float32)

    tgt = rng.randint(0, T, n)
    kind = rng.randint(0, 6, n)
    a, b, c = sc.v0[tgt, :3].astype(np.float64), sc.v1[tgt, :3].astype(np.float64), sc.v2[tgt, :3].astype(np.float64)
    u = rng.rand(n, 1)
    point = np.where(kind[:, None] == 0, a + u * (b - a),                       # on edge v0-v1
             np.where(kind[:, None] == 1, a + u * (c - a),                      # on edge v0-v2
             np.where(kind[:, None] == 2, b + u * (c - b),                      # on edge v1-v2
             np.where(kind[:, None] == 3, a, (a + b + c) / 3.0))))              # vertex / centroid
    origin3 = rng.uniform(-12, 12, (n, 3)) * scale
    on_tri = rng.rand(n) < 0.25                                                  # start ON another triangle
    other = rng.randint(0, T, n)
    w = rng.dirichlet([1, 1, 1], n)
    origin3[on_tri] = (w[:, :1] * sc.v0[other, :3] + w[:, 1:2] * sc.v1[other, :3] + w[:, 2:] * sc.v2[other, :3])[on_tri]
    d3 = point - origin3
    par = kind == 5                                                              # parallel to the target's plane
    d3[par] = (b - a)[par] + 0.3 * (c - a)[par]
    norm = np.linalg.norm(d3, axis=1, keepdims=True)
    norm[norm == 0] = 1.0
    d3 = d3 / norm * rng.choice([1.0, 1.0, 1.0, -1.0, 0.01, 37.0], size=(n, 1))
    origin = pad(origin3)
    direc = pad(d3)
    max_ray_len = np.float32(60.0 * scale)

    want = oracle.intersect(sc, origin, direc, max_ray_len, BACKEND)
    g = GpuSeams(sc)
    try:
        for impl, R, chunks in (("simt", 0, 0), ("simt", 4, 5), ("simt", 1, 3), ("tc", 0, 0), ("tc", 1, 3), ("tc", 4, 5)):
            _set_search(impl, R, chunks)
            try:
                got = g.intersect(origin, direc, max_ray_len)
            finally:
                _clear_search()
            assert np.array_equal(got[1], want[1]), "crossing counts differ (%s R=%d)" % (impl, R)
            assert np.array_equal(got[2], want[2]), "nearest triangle differs (%s R=%d)" % (impl, R)
            assert np.array_equal(util.bits(got[0]), util.bits(want[0])), "nearest t differs (%s R=%d)" % (impl, R)
    finally:
        g.close()
    # the soup really is hit, often more than once (tiny scenes lose hits to the reference's ABSOLUTE |DEN| < 1e-6 cut-off, Q7)
    assert (want[1] > 0).mean() > (0.3 if scale >= 0.05 else 0.1) and (want[1] > 1).sum() > 100


@pytest.mark.parametrize("impl,chunks", [("tc", 0), ("tc", 1), ("tc8", 0), ("tc16", 0), ("simt", 0)])
def test_search_with_thousands_of_coincident_triangles(impl, chunks):
    """Every ray crosses 3,000 copies of the same two triangles (plus a far plate in a second mesh):
    3,000 exact-t ties per ray (the first triangle index must win, cl:278), crossing counts in the
    thousands, and a candidate density that overflows the tensor-core kernels' per-warp queues, so
    their overflow paths run too: the in-place evaluation of the per-chunk-vote kernels; for the default
    deferred-flag kernel the mid-sweep flush of a full queue (chunks = 1: one CTA sweeps all 94 tiles, 6,000
    records per warp) and the in-kernel evaluation of the records that do not fit the global record buffer
    (262,500 records against a capacity of 32 per ray)."""
    rng = np.random.RandomState(11)
    copies = 3000
    quad = np.array([[[-1, -1, 5], [1, -1, 5], [1, 1, 5]], [[-1, -1, 5], [1, 1, 5], [-1, 1, 5]]], dtype=np.float64)
    stack = np.tile(quad, (copies, 1, 1))                       # mesh 0: 6,000 triangles, pairwise coincident
    plate = quad * np.array([3.0, 3.0, 1.0])
    plate[:, :, 2] = 9.0                                        # mesh 1: two larger triangles further out
    tris = np.concatenate((stack, plate))
    T, M, n = len(tris), 2, 700

    class Soup:
        pass

    sc = Soup()
    pad = lambda a: np.ascontiguousarray(np.concatenate((a, np.zeros((len(a), 1))), axis=1), dtype=np.float32)
    sc.v0, sc.v1, sc.v2 = pad(tris[:, 0]), pad(tris[:, 1]), pad(tris[:, 2])
    sc.tri_mesh = np.concatenate((np.zeros(len(stack), np.int32), np.ones(len(plate), np.int32)))
    sc.tri_count, sc.mesh_count = T, M
    sc.mat_type, sc.ior = np.zeros(M, np.int32), np.ones(M, np.float32)
    sc.refl, sc.diss = np.ones(M, np.float32), np.zeros(M, np.float32)
    origin = pad(rng.uniform(-0.2, 0.2, (n, 3)))
    target = np.concatenate((rng.uniform(-0.9, 0.9, (n, 2)), np.full((n, 1), 5.0)), axis=1)
    d3 = target - origin[:, :3]
    direc = pad(d3 / np.linalg.norm(d3, axis=1, keepdims=True))
    max_ray_len = np.float32(100.0)

    want = oracle.intersect(sc, origin, direc, max_ray_len, BACKEND)
    assert (want[1][:, 0] == copies).all() and (want[1][:, 1] == 1).all()     # every ray: 3,000 + 1 crossings
    assert (want[2][:, 0] < 2).all()                                             # ties: the first copy wins
    g = GpuSeams(sc)
    _set_search(impl, 0, chunks)
    try:
        got = g.intersect(origin, direc, max_ray_len)
    finally:
        _clear_search()
        g.close()
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert np.array_equal(util.bits(got[0]), util.bits(want[0]))
