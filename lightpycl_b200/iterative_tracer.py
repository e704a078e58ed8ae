"""The trace loop: drop-in for the reference's ``iterative_tracer.CL_Tracer`` on B200.

Reference: iterative_tracer.py:32-411 ("it:").  Same class name, constructor and method
signatures and the same result structure; the body is new.  Where the reference uploads rays,
launches three OpenCL kernels, downloads 12 arrays and prunes with NumPy on every iteration
(it:261-375), this loop keeps every ray generation in HBM as torch tensors and drives the
sm_100a kernels of ``liblightpycl_b200.so`` through the C ABI (``_cabi``): one search kernel
plus one fused resolve/Fresnel/compaction kernel per generation, and a 64-byte counter
read-back per iteration for the reference's two termination tests (it:372-391).

Multi-GPU (one process per GPU, ``torch.distributed``): rays are sharded contiguously over the
ranks, the scene is replicated, children stay on the GPU that produced them.  The only
collectives are a per-iteration all-reduce of (live rays, live power) and the final gather of
measured rays; see ``dist.py``.

There is no CPU fallback: constructing a ``CL_Tracer`` without the CUDA extension or without a
CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
import os
import pickle
import threading
import weakref
from time import strftime, time

import numpy as np

from . import _cabi
from . import dist as _dist

__all__ = ["CL_Tracer", "TraceResults"]


# ------------------------------------------------------------------------------------------ host-side flattening
def flatten_light_sources(light_source, lo=None, hi=None):
    """it:99-118.  Returns (origin (n,4) f32, dir (n,4) f32, power f32 in the sources' shape) of the rays
    [lo, hi) of the concatenated sources (default: all of them).  Only the sources overlapping the window
    are touched, and a window inside ONE source is returned as views: a rank of a ray-sharded run does not
    concatenate the other ranks' rays."""
    origins, dirs, powers, counts = [], [], [], []
    for light in light_source:
        o = np.asarray(light.rays_origin, dtype=np.float32)
        d = np.asarray(light.rays_dir, dtype=np.float32)
        p = np.asarray(light.rays_power, dtype=np.float32)
        if len(o) < int(light.ray_count) or len(d) != int(light.ray_count):
            raise ValueError("light source arrays do not match ray_count")
        # grid_rays leaves surplus origin rows behind (ls:46-49: sized before ray_count is rounded down to a square);
        # the reference loop only ever reads rows [0, ray_count) (it:276-283)
        o = o[: int(light.ray_count)]
        origins.append(o.reshape(-1, 4))
        dirs.append(d.reshape(-1, 4))
        powers.append(p)
        counts.append(len(o))
    if not origins:
        raise ValueError("light_source must be a non-empty list of sources")
    shapes = {p.ndim for p in powers}
    if len(shapes) != 1:  # np.append(..., axis=0) of (N,1) and (N,) raises in the reference too (it:111)
        raise ValueError("all light sources must give rays_power with the same number of dimensions")
    n_all = int(sum(counts))
    lo = 0 if lo is None else int(lo)
    hi = n_all if hi is None else int(hi)
    parts, start = [], 0
    for k, c in enumerate(counts):  # the pieces of [lo, hi) source by source
        a, b = max(lo, start), min(hi, start + c)
        if b > a:
            parts.append((k, a - start, b - start))
        start += c
    if not parts:
        empty_pow = powers[0][:0]
        return np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), np.ascontiguousarray(empty_pow)
    pick = lambda arrs: [arrs[k][a:b] for k, a, b in parts]  # noqa: E731
    if len(parts) == 1:  # one source: no copy (the arrays are only read)
        return tuple(np.ascontiguousarray(x[0]) for x in (pick(origins), pick(dirs), pick(powers)))
    return tuple(np.ascontiguousarray(np.concatenate(x, axis=0)) for x in (pick(origins), pick(dirs), pick(powers)))


def source_powers(light_source):
    """The per-source power arrays (float32, the sources' shapes) without concatenating them."""
    return [np.asarray(light.rays_power, dtype=np.float32) for light in light_source]


def sorted_sequential_sum_f32(p):
    """``sum(np.sort(p))`` as the reference evaluates it (it:115): sort along the last axis, then
    a left-to-right float32 accumulation (np.cumsum accumulates sequentially)."""
    if isinstance(p, (list, tuple)):  # per-source arrays (source_powers): same additions, no concatenation
        chunks = [np.asarray(c, dtype=np.float32) for c in p]
        if chunks and all(c.ndim > 1 and c.shape[-1] == 1 for c in chunks):  # np.sort along the last axis is a no-op
            total = np.float32(0.0)
            for c in chunks:
                if c.size:  # continue the left-to-right float32 accumulation across the chunk boundary
                    total = np.float32(np.cumsum(np.concatenate((np.array([total], np.float32), c.reshape(-1))),
                                                 dtype=np.float32)[-1])
            return total
        p = np.concatenate(chunks, axis=0) if chunks else np.zeros(0, np.float32)
    p = np.asarray(p, dtype=np.float32)
    # np.sort works along the LAST axis: a no-op for the (N,1) power of random_rays
    flat = (p if (p.ndim > 1 and p.shape[-1] == 1) else np.sort(p)).reshape(-1)
    if flat.size == 0:
        return np.float32(0.0)
    return np.float32(np.cumsum(flat, dtype=np.float32)[-1])


from .geo_optical_elements import GeoObject as _GeoObject  # noqa: E402

_INDEXED_TRIBUF = _GeoObject.tribuf  # meshes whose tribuf() is the stock gather take the fast path below


def flatten_meshes(meshes, alloc=None):
    """it:121-151.  Returns dict(v0, v1, v2 (T,4) f32, tri_mesh (T,) i32, mat_type i32, ior, refl, diss f32).
    ``alloc(name, rows, cols, dtype)`` may supply the output arrays (the tracer recycles host buffers)."""
    m = len(meshes)
    if m == 0:
        raise ValueError("meshes must be a non-empty list")
    mat_type = np.zeros(m, dtype=np.int32)
    ior = np.zeros(m, dtype=np.float32)
    refl = np.zeros(m, dtype=np.float32)
    diss = np.zeros(m, dtype=np.float32)
    parts = []  # (vertices f32 (V,4), triangles (T,3) int) for indexed meshes, else (v0, v1, v2) f32 arrays
    total = 0
    for j, mesh in enumerate(meshes):
        mat = mesh.getMaterialBuf()
        mat_type[j] = np.int32(mat.get("type"))
        ior[j] = np.float32(mat.get("IOR"))
        refl[j] = np.float32(mat.get("R"))
        diss[j] = np.float32(mat.get("dissipation"))
        verts, tris = getattr(mesh, "vertices", None), getattr(mesh, "triangles", None)
        if verts is not None and tris is not None and type(mesh).tribuf is _INDEXED_TRIBUF:
            # same values as tribuf() -> float32 (it:141-149): the cast commutes with the gather, and
            # gathering float32 rows straight into the output skips three float64 temporaries per mesh
            v = np.ascontiguousarray(np.asarray(verts, dtype=np.float32).reshape(-1, 4))
            t = np.asarray(tris).reshape(-1, 3)
            part, count = (v, t), len(t)
        else:
            tb = mesh.tribuf()
            part = tuple(np.asarray(tb[k], dtype=np.float32).reshape(-1, 4) for k in range(3))
            count = len(part[0])
        if count == 0:
            raise ValueError("mesh %d has no triangles (the reference kernel requires non-empty meshes)" % j)
        parts.append((part, count))
        total += count
    if alloc is None:
        alloc = lambda name, rows, cols, dtype: np.empty((rows, cols) if cols else rows, dtype=dtype)  # noqa: E731
    v0 = alloc("v0", total, 4, np.float32)
    v1 = alloc("v1", total, 4, np.float32)
    v2 = alloc("v2", total, 4, np.float32)
    tri_mesh = alloc("tri_mesh", total, 0, np.int32)
    off = 0
    for j, (part, count) in enumerate(parts):
        sl = slice(off, off + count)
        if len(part) == 2:
            v, t = part
            for k, out in enumerate((v0, v1, v2)):
                np.take(v, t[:, k], axis=0, out=out[sl])
        else:
            v0[sl], v1[sl], v2[sl] = part
        tri_mesh[sl] = j
        off += count
    return dict(v0=v0, v1=v1, v2=v2, tri_mesh=tri_mesh, mat_type=mat_type, ior=ior, refl=refl, diss=diss)


def index_meshes(meshes):
    """The scene in INDEXED form for the device-side gather: dict(parts = [(verts (V,4) f32, tris (T,3) int)] per mesh,
    tri_mesh, mat_type, ior, refl, diss), or None when a mesh does not use the stock ``tribuf()`` (then
    ``flatten_meshes`` gathers on the host).  ``verts[tris[:, k]]`` equals flatten_meshes' ``v_k`` bit for bit (the float32
    cast commutes with the gather, it:141-149); gathering 3 x T rows on the GPU instead of the host saves ~2 ms per
    100k triangles of every trace.  No copies here: the parts are views of the meshes' own arrays where they already
    are float32 / integer ndarrays."""
    m = len(meshes)
    if m == 0:
        raise ValueError("meshes must be a non-empty list")
    mat_type = np.zeros(m, dtype=np.int32)
    ior = np.zeros(m, dtype=np.float32)
    refl = np.zeros(m, dtype=np.float32)
    diss = np.zeros(m, dtype=np.float32)
    parts, total = [], 0
    for j, mesh in enumerate(meshes):
        verts, tris = getattr(mesh, "vertices", None), getattr(mesh, "triangles", None)
        if verts is None or tris is None or type(mesh).tribuf is not _INDEXED_TRIBUF:
            return None
        mat = mesh.getMaterialBuf()
        mat_type[j] = np.int32(mat.get("type"))
        ior[j] = np.float32(mat.get("IOR"))
        refl[j] = np.float32(mat.get("R"))
        diss[j] = np.float32(mat.get("dissipation"))
        v = np.ascontiguousarray(np.asarray(verts, dtype=np.float32).reshape(-1, 4))
        t = np.ascontiguousarray(np.asarray(tris).reshape(-1, 3))
        if len(t) == 0:
            raise ValueError("mesh %d has no triangles (the reference kernel requires non-empty meshes)" % j)
        if t.dtype.kind not in "iu" or t.min() < 0 or t.max() >= len(v):  # np.take on the host would wrap / raise: leave odd meshes to it
            return None
        if t.dtype not in (np.int64, np.int32):
            t = t.astype(np.int64)
        parts.append((v, t))
        total += len(t)
    tri_mesh = np.empty(total, dtype=np.int32)
    off = 0
    for j, (_, t) in enumerate(parts):
        tri_mesh[off: off + len(t)] = j
        off += len(t)
    return dict(parts=parts, tri_mesh=tri_mesh, mat_type=mat_type, ior=ior, refl=refl, diss=diss)


# ------------------------------------------------------------------------------------------ device-side state
class _Generation:
    """One ray generation in HBM (packed layout of include/lightpycl_b200.h)."""

    __slots__ = ("org_pow", "dir_prev", "lineage", "n", "split", "base")

    def __init__(self, org_pow, dir_prev, lineage, n, split, base):
        self.org_pow, self.dir_prev, self.lineage = org_pow, dir_prev, lineage
        self.n, self.split, self.base = int(n), int(split), int(base)

    def c_struct(self, lo=0, hi=None):
        """The generation - or the ray tile [lo, hi) of it - as the C ABI's lpcl_generation.  A tile
        is itself a two-segment generation over the same buffers: pointers advanced to logical ray
        ``lo``, the segment split and base moved accordingly (a tile may straddle the segments)."""
        hi = self.n if hi is None else hi
        cap = self.org_pow.shape[0]
        if lo >= self.split:  # entirely inside the second segment
            first, split, base = self.base + (lo - self.split), hi - lo, hi - lo
        else:
            first, split, base = lo, min(self.split, hi) - lo, self.base - lo
        g = _cabi.Generation()
        g.d_org_pow = self.org_pow.data_ptr() + 16 * first
        g.d_dir_prev = self.dir_prev.data_ptr() + 16 * first
        g.d_lineage = self.lineage.data_ptr() + 8 * first if self.lineage is not None else None
        g.n, g.seg_split, g.seg_base = hi - lo, split, base
        g.capacity = cap - first
        return g

    def child_struct(self, n_parents, refl_done=0, refr_done=0):
        """This (still empty) generation as the children argument of one step over a tile of an
        ``n_parents``-ray parent generation, after earlier tiles produced ``refl_done`` reflected and
        ``refr_done`` refracted survivors: reflected survivors continue at physical index refl_done,
        refracted ones at n_parents + refr_done."""
        g = _cabi.Generation()
        g.d_org_pow = self.org_pow.data_ptr() + 16 * refl_done
        g.d_dir_prev = self.dir_prev.data_ptr() + 16 * refl_done
        g.d_lineage = self.lineage.data_ptr() + 8 * refl_done if self.lineage is not None else None
        g.n, g.seg_split = 0, 0
        g.seg_base = n_parents + refr_done - refl_done
        g.capacity = self.org_pow.shape[0] - refl_done
        return g

    def logical(self, t):
        """Logical-order view (copy only when the generation has two segments)."""
        import torch

        if self.split == self.n:
            return t[: self.n]
        return torch.cat((t[: self.split], t[self.base: self.base + (self.n - self.split)]), dim=0)



class _HostPool:
    """Recycles the host memory behind arrays the tracer hands to its caller (flattened vertices,
    measured (pos, pwr)).  First-touch page faults on newly mapped host memory cost more than the
    copies themselves (5 ms vs 1 ms for the 18 MB of measured records), so a buffer is handed out
    again once the caller has dropped EVERY array that refers to it - and only then.

    Ownership is tracked explicitly, not inferred from reference counts: each hand-out wraps the
    pooled memory in a fresh lease object (a ctypes array over the buffer) and the returned
    ndarray - and every view, slice or reshape derived from it - keeps that lease alive through
    NumPy's ``.base`` chain.  ``weakref.finalize`` on the lease marks the buffer free when the last
    of them is gone."""

    def __init__(self, max_buffers_per_key=4, pinned=False):
        self._entries = {}
        self._max = max_buffers_per_key
        self._pinned = pinned  # page-locked buffers (the tracer's pool): host<->device copies run at the DMA rate

    class _Entry:
        __slots__ = ("mem", "busy", "owner")

        def __init__(self, mem, owner=None):
            self.mem, self.busy, self.owner = mem, False, owner  # owner: the pinned torch tensor behind mem, if any

        def release(self):
            self.busy = False

    def array(self, key, shape, dtype):
        dtype = np.dtype(dtype)
        count = int(np.prod(shape, dtype=np.int64))
        nbytes = count * dtype.itemsize
        if nbytes == 0:
            return np.empty(shape, dtype=dtype)
        entries = self._entries.setdefault(key, [])
        entry = next((e for e in entries if not e.busy and e.mem.nbytes >= nbytes), None)
        if entry is None:
            cap = max(int(nbytes * 1.25), 16384)
            if self._pinned:
                import torch

                owner = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
                entry = self._Entry(owner.numpy(), owner)
            else:
                entry = self._Entry(np.empty(cap, dtype=np.uint8))
            if len(entries) >= self._max:  # forget the oldest idle buffer (busy ones stay with their holders)
                idle = [e for e in entries if not e.busy]
                if idle:
                    entries.remove(idle[0])
            if len(entries) < self._max:
                entries.append(entry)
        entry.busy = True
        lease = (C.c_char * nbytes).from_buffer(entry.mem)
        weakref.finalize(lease, entry.release)
        return np.frombuffer(lease, dtype=dtype, count=count).reshape(shape)


class TraceResults:
    """``self.results`` of the reference (it:355,393): a list with one tuple per iteration,
    ``(rays_origin (n,4) f32, rays_dest (n,4) f32, rays_pow f32, rays_meas (n,) i32)``.

    The data stays in HBM; an entry is downloaded and converted to NumPy the first time it is
    indexed or iterated.  ``ray_counts`` gives ``len(res[3])`` per iteration without any
    download.  Iteration 0's power keeps the light sources' shape, e.g. (N,1).
    """

    def __init__(self):
        self._dev = []  # (generation, dest_meas tensor, power shape or None)
        self._host = {}
        self.ray_counts = []

    def _append(self, gen, dest_meas, pow_shape):
        self._dev.append((gen, dest_meas, pow_shape))
        self.ray_counts.append(gen.n)

    def __len__(self):
        return len(self._dev)

    def _materialise(self, k):
        if k not in self._host:
            gen, dest_meas, pow_shape = self._dev[k]
            op = gen.logical(gen.org_pow).cpu().numpy()
            dm = dest_meas.cpu().numpy()
            origin = np.ascontiguousarray(op.copy())
            power = np.ascontiguousarray(origin[:, 3].copy())
            origin[:, 3] = 0.0
            dest = np.ascontiguousarray(dm.copy())
            meas = np.ascontiguousarray(dest[:, 3].copy()).view(np.int32)
            dest[:, 3] = 0.0
            if pow_shape is not None:
                power = power.reshape(pow_shape)
            self._host[k] = (origin, dest, power, meas)
        return self._host[k]

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._materialise(i) for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        if not 0 <= k < len(self):
            raise IndexError(k)
        return self._materialise(k)

    def __iter__(self):
        for k in range(len(self)):
            yield self._materialise(k)


class CL_Tracer:
    """B200 trace engine with the reference's ``CL_Tracer`` interface (it:32-411).

    ``platform_name`` / ``device_name`` are accepted for source compatibility (it:36); the CUDA
    device is ``device`` if given, else ``cuda:$LOCAL_RANK`` (or the current device).

    Extra keyword arguments (all optional, defaults keep the reference's behaviour):
      keep_results   retain every generation for ``self.results`` (default True, like it:355);
                     False keeps only the measured-ray buffer (for runs that would not fit).
      record_hits    also retain per-ray (hit mesh, hit triangle, n1, n2) in ``self.hit_records``.
      lineage        carry a 64-bit lineage id per ray (auto-enabled when world_size > 1).
      process_group  torch.distributed group to shard over (default: the world group if
                     torch.distributed is initialised).
      shard_input    True: every rank passes the SAME sources and traces its contiguous slice;
                     False: every rank passes its OWN rays (used by bench.py weak scaling).
      ray_tile_bytes device-memory budget of the per-generation scratch (search results, 12 B per
                     ray and mesh).  A generation that needs more is traced in RAY TILES that reuse
                     one scratch buffer, like the reference's <= 50,000-ray partitions (it:218-273);
                     results are bit-identical.  Default: $LPCL_RAY_TILE_BYTES or 8 GiB.
      profile        record CUDA events around the search and the resolve kernel of every
                     generation -> ``self.kernel_times`` (list of dict(rays, search_ms, resolve_ms)).
    """

    results = None
    geometry = None

    def __init__(self, platform_name="NVIDIA", device_name="770", debug=False, device=None, keep_results=True,
                 record_hits=False, lineage=None, process_group=None, shard_input=True, verbose=False,
                 profile=False, ray_tile_bytes=None):
        import torch

        self.debug = debug
        self.verbose = verbose or debug
        self._torch = torch
        self._lib = _cabi.load()  # raises if the CUDA extension is missing: no fallback
        if not torch.cuda.is_available():
            raise RuntimeError("lightpycl_b200.CL_Tracer needs a CUDA device (B200, sm_100a); none is visible")
        if device is None:
            device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())))
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.platform_name, self.device_name = platform_name, device_name
        if self.verbose:
            print("Using CUDA device: ", torch.cuda.get_device_name(self.device))
        self.keep_results = keep_results
        self.record_hits = record_hits
        self._group = _dist.resolve_group(process_group)
        self.world_size, self.rank = _dist.world(self._group)
        if self.world_size > 1 and lineage is not None and not lineage:
            raise ValueError("lineage=False is not possible with world_size > 1: the measured rays of the ranks are "
                             "merged into the reference order by lineage")
        self._lineage = bool(lineage) if lineage is not None else self.world_size > 1
        self.ray_tile_bytes = int(ray_tile_bytes if ray_tile_bytes is not None
                                  else os.environ.get("LPCL_RAY_TILE_BYTES", 8 << 30))
        # LPCL_PIN_HOST=1: page-locked host staging (the result / geometry pool, and source arrays met twice).  Off by
        # default: page-locking costs 1-2 ms per MB ONCE (measured: 45 ms to register 36 MB of rays, 35 ms to allocate
        # an 18 MB result buffer), which only a loop of many traces over the same sources earns back (4 ms per trace)
        self._pin_host = os.environ.get("LPCL_PIN_HOST", "0") == "1"
        self._host_pool = _HostPool(pinned=self._pin_host)
        self.shard_input = shard_input
        # LPCL_DEVICE_GATHER=0: gather the per-triangle vertex buffers on the host (flatten_meshes) instead of on the GPU
        self.device_gather = os.environ.get("LPCL_DEVICE_GATHER", "1") != "0"
        # LPCL_STREAM_UPLOAD=0: upload all source rays in prepare() instead of slice by slice under generation 0's search
        self.stream_upload = os.environ.get("LPCL_STREAM_UPLOAD", "1") != "0"
        self._copy_stream = None
        self._candidate_density = None  # filter candidates per (ray, triangle) pair of the last generation traced
        self._gen0_density, self._gen0_density_tris = None, -1
        self._tc_warps_set = 0
        self._geometry = None
        self.profile = profile
        self.kernel_times = []
        self.kernel_launches = 0
        self._job = None
        self._scene = None
        self._scene_tensors = None
        self._workspace = None
        self._scratch_per_ray = None
        self.timings = {}
        self.stats = {}

    # ---------------------------------------------------------------------------------- scene
    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _release_scene(self):
        if self._scene is not None:
            self._lib.lpcl_scene_destroy(self._scene)
            self._scene = None
            self._scene_tensors = None

    def __del__(self):
        try:
            self._release_scene()
        except Exception:
            pass

    @property
    def geometry(self):
        """``(m_v0, m_v1, m_v2)`` float32 host arrays like the reference's attribute (it:157).  With the device-side
        gather they are not needed for tracing and are built on first access."""
        if self._geometry is None and getattr(self, "meshes", None) is not None:
            flat = flatten_meshes(self.meshes, alloc=self._host_array)
            self._geometry = (flat["v0"], flat["v1"], flat["v2"])
        return self._geometry

    @geometry.setter
    def geometry(self, value):
        self._geometry = value

    def _upload_scene(self, flat):
        torch = self._torch
        self._release_scene()
        # (the flattened vertex arrays come from the page-locked pool and stay alive in self.geometry)
        if "parts" in flat:  # indexed scene: the three vertex buffers are gathered on the GPU (set-up path, not the trace step)
            total = len(flat["tri_mesh"])
            dev = {"v%d" % k: torch.empty((total, 4), dtype=torch.float32, device=self.device) for k in range(3)}
            parts = flat["parts"]
            if len(parts) > 16:  # many small meshes: one upload of the concatenated scene instead of two per mesh
                offs = np.cumsum([0] + [len(v) for v, _ in parts[:-1]])
                parts = [(np.concatenate([v for v, _ in parts]),
                          np.concatenate([t.astype(np.int64) + o for (_, t), o in zip(parts, offs)]))]
            off = 0
            for v, t in parts:
                vd, td = torch.from_numpy(v).to(self.device), torch.from_numpy(t).to(self.device)
                for k in range(3):
                    torch.index_select(vd, 0, td[:, k], out=dev["v%d" % k][off: off + len(t)])
                off += len(t)
        else:
            dev = {k: (lambda h: h.to(self.device, non_blocking=h.is_pinned()))(torch.from_numpy(flat[k])) for k in ("v0", "v1", "v2")}
        # the handle's device memory comes from torch's caching allocator (no cudaMalloc/cudaFree per trace)
        n_tri, n_mesh = len(flat["tri_mesh"]), len(flat["mat_type"])
        need = int(self._lib.lpcl_scene_bytes(n_tri, n_mesh))
        dev["scene_mem"] = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = (dev["scene_mem"].data_ptr() + 255) // 256 * 256
        handle = C.c_void_p()
        rc = self._lib.lpcl_scene_create_in(
            C.byref(handle), dev["v0"].data_ptr(), dev["v1"].data_ptr(), dev["v2"].data_ptr(),
            flat["tri_mesh"].ctypes.data, n_tri, flat["mat_type"].ctypes.data, flat["ior"].ctypes.data,
            flat["refl"].ctypes.data, flat["diss"].ctypes.data, n_mesh, base, need, self._stream())
        _cabi.check(rc, "lpcl_scene_create_in")
        self._scene = handle
        self._tc_warps_set = 0  # a fresh handle uses the library default until _pick_search_variant says otherwise
        self._scratch_per_ray = None
        self._scene_tensors = dev  # the handle borrows the vertex buffers and its own memory: keep them alive

    def _get_workspace(self, n):
        torch = self._torch
        need = int(self._lib.lpcl_step_workspace_bytes(self._scene, n))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=self.device)
        return self._workspace

    # ---------------------------------------------------------------------------------- the loop
    def iterative_tracer(self, light_source, meshes, trace_iterations=100, trace_until_dissipated=0.99,
                         max_ray_len=np.float32(1e3), ior_env=np.float32(1.0)):
        """Trace ``light_source`` (list of sources) through ``meshes`` (list of GeoObject-like).

        Same contract as the reference (it:77-393): repeat intersect -> resolve media -> Fresnel
        split -> prune, until ``trace_iterations`` generations were processed, the surviving
        power falls below ``(1 - trace_until_dissipated) * input_power`` or no ray is left.
        Returns and stores ``self.results``; also sets ``self.tri_count``, ``self.meshes`` and
        ``self.geometry`` like the reference (it:152-158).

        Equivalent to ``prepare(light_source, meshes)`` (host flattening, sharding, upload, scene
        build) followed by ``run(...)`` (the device-resident loop).
        """
        t_start = time()
        self.prepare(light_source, meshes)
        out = self.run(trace_iterations, trace_until_dissipated, max_ray_len, ior_env)
        self.timings["iterative_tracer_s"] = time() - t_start
        return out

    def prepare(self, light_source, meshes):
        """Host side of it:99-169: flatten sources and meshes to float32 arrays, pick this rank's
        ray shard, copy rays and geometry to HBM and build the scene handle."""
        torch = self._torch
        t0 = time()
        flat = index_meshes(meshes) if self.device_gather else None
        if flat is None:
            flat = flatten_meshes(meshes, alloc=self._host_array)
            self._geometry = (flat["v0"], flat["v1"], flat["v2"])
        else:
            self._geometry = None  # materialised on first access (the `geometry` property)
        self.tri_count = np.int32(len(flat["tri_mesh"]))
        self.mesh_count = np.int32(len(flat["mat_type"]))
        self.meshes = meshes
        if self.verbose:
            print("Triangle count:              ", self.tri_count)
            print("Mesh count:                  ", self.mesh_count)

        # shard the rays over the ranks (contiguous blocks keep the reference order inside a rank): a rank
        # flattens only its own window of the concatenated sources
        n_all = int(sum(int(light.ray_count) for light in light_source))
        lo, hi, lineage_base = 0, n_all, 0
        if self.world_size > 1 and self.shard_input:
            lo, hi = _dist.shard_bounds(n_all, self.world_size, self.rank)
            lineage_base = lo
        origin, direc, power = flatten_light_sources(light_source, lo, hi)
        # it:115 input power = sum(np.sort(power)), a sequential float32 sum over ALL rays: only the
        # termination test needs it, so it is evaluated in run() while generation 0 is on the GPU
        self._power_all = source_powers(light_source)
        # ... on a helper thread (NumPy releases the GIL in the sort / cumulative sum), started here so that it also
        # overlaps the upload: on a rank of an 8-GPU run generation 0 is shorter than this sum over ALL rays
        self._power_thread_result = []
        self._power_thread = threading.Thread(
            target=lambda out=self._power_thread_result, p=self._power_all: out.append(float(sorted_sequential_sum_f32(p))),
            daemon=True)
        self._power_thread.start()
        self._power_needs_allreduce = False
        self.input_power = None
        pow_shape = None if power.ndim == 1 else power.shape
        if self.world_size > 1 and not self.shard_input:  # every rank brought its own rays
            counts = _dist.all_gather_int(len(origin), self._group, self.device)
            lineage_base = int(sum(counts[: self.rank]))
            self._power_needs_allreduce = True
        self.timings["host_flatten_s"] = time() - t0

        t0 = time()
        with torch.cuda.device(self.device):
            self._upload_scene(flat)
            n = len(origin)
            job = dict(n=n, lineage_base=lineage_base, pow_shape=pow_shape)
            if n > 0:
                # page-locked source arrays go by direct DMA (arrays seen before are page-locked in place, see
                # _page_lock_if_seen_again); the host arrays stay referenced by the job until run() has synchronised
                arrays = [np.ascontiguousarray(a) for a in (origin, direc, power.reshape(-1))]
                for a in arrays:
                    self._page_lock_if_seen_again(a)
                host = [torch.from_numpy(a) for a in arrays]
                job["host"] = host
                # Large ray sets from pageable memory are not uploaded here: run() uploads generation 0 in slices on a copy
                # stream while the previous slice is being searched (self._stream_slices), which hides ~3 of the 4 ms
                job["resident"] = not (self.stream_upload and n >= self.STREAM_MIN_RAYS and self.world_size == 1
                                       and not any(h.is_pinned() for h in host))
                if job["resident"]:
                    job["origin"], job["dir"], job["power"] = (h.to(self.device, non_blocking=h.is_pinned()) for h in host)
                else:
                    job["origin"], job["dir"], job["power"] = (torch.empty(h.shape, dtype=h.dtype, device=self.device)
                                                               for h in host)
            self._job = job
        geo_bytes = (sum(v.nbytes + t.nbytes for v, t in flat["parts"]) if "parts" in flat
                     else sum(flat[k].nbytes for k in ("v0", "v1", "v2")))
        self.h2d_bytes = int(geo_bytes + n * (16 + 16 + 4))
        self.timings["upload_s"] = time() - t0

    # page-locking of caller-owned source arrays ------------------------------------------------------------
    PAGE_LOCK_BUDGET = 1 << 30  # bytes of caller memory this tracer keeps registered at most

    def _page_lock_if_seen_again(self, arr):
        """Opt-in (LPCL_PIN_HOST=1).  Uploading rays from pageable memory runs at ~1/4 of the DMA rate (4.1 vs 0.9 ms per
        1M rays).  The light sources stay plain NumPy; instead, when prepare() meets the SAME source array a
        second time (a sweep over geometries, a loop of traces) the tracer page-locks that memory in place
        (cudaHostRegister, on its own device) so that later uploads are direct DMA.  The registration is undone when the array's owner is
        garbage collected; at most PAGE_LOCK_BUDGET bytes are registered; a refusal by the driver just leaves the
        array pageable."""
        if not self._pin_host or arr.nbytes < (1 << 20):
            return
        state = self.__dict__.setdefault("_page_locks", dict(seen={}, locked={}, bytes=0))
        key = (arr.ctypes.data, arr.nbytes)
        if key in state["locked"]:
            return
        state["seen"][key] = state["seen"].get(key, 0) + 1
        if state["seen"][key] < 2 or state["bytes"] + arr.nbytes > self.PAGE_LOCK_BUDGET:
            if len(state["seen"]) > 64:
                state["seen"].clear()
            return
        cudart = self._torch.cuda.cudart()
        rc = cudart.cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
        if int(rc) != 0:  # e.g. the range overlaps memory registered elsewhere: stay pageable
            state["seen"][key] = -(1 << 30)
            return
        owner = arr
        while isinstance(getattr(owner, "base", None), np.ndarray):
            owner = owner.base
        state["locked"][key] = True
        state["bytes"] += arr.nbytes

        def _unlock(ptr=arr.ctypes.data, nbytes=arr.nbytes, state=state, key=key, cudart=cudart):
            cudart.cudaHostUnregister(ptr)
            state["locked"].pop(key, None)
            state["bytes"] -= nbytes

        weakref.finalize(owner, _unlock)

    def _ensure_input_power(self):
        """Total source power as the reference sums it (it:115); with rays brought per rank, summed
        over the ranks (a collective: every rank calls this at the same point of run())."""
        if self.input_power is None:
            thread = getattr(self, "_power_thread", None)
            if thread is not None:
                thread.join()
                self._power_thread = None
            p = self._power_thread_result[0] if getattr(self, "_power_thread_result", None) else \
                float(sorted_sequential_sum_f32(self._power_all))
            if self._power_needs_allreduce:
                p = _dist.all_reduce_sum(p, self._group, self.device)
            self.input_power = p
        return self.input_power

    def run(self, trace_iterations=100, trace_until_dissipated=0.99, max_ray_len=np.float32(1e3),
            ior_env=np.float32(1.0)):
        """The device-resident loop (replaces it:241-391).  Can be called repeatedly after one
        ``prepare``; generation 0 is re-packed from the uploaded source rays every time."""
        torch = self._torch
        lib = self._lib
        t_start = time()
        max_ray_len = np.float32(max_ray_len)
        ior_env = np.float32(ior_env)
        job = self._job
        pow_shape = job["pow_shape"]
        prof = self.profile
        events = []

        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            n = job["n"]
            results = TraceResults()
            self.results = results
            self.hit_records = []
            self._measured = None
            self._measured_lineage = None
            self._measured_count = 0
            self._measured_per_iter = []
            self.kernel_launches = 0
            self.stats = dict(rays=[], live=[], live_local=[], power_left=[], candidates=[], measured=[], terminated=[],
                              tiles=[])
            counters_dev = torch.zeros(8, dtype=torch.int64, device=self.device)
            if getattr(self, "_counters_host", None) is None:  # pinned once per tracer (cudaHostAlloc is slow)
                self._counters_host = torch.zeros(8, dtype=torch.int64).pin_memory()
            counters_host = self._counters_host

            gen = None
            t_first, finished = 0, False
            small = self._small_plan(n, int(trace_iterations))
            if n > 0:
                cap0 = small["arena_rays"] if small else n
                org_pow = torch.empty((cap0, 4), dtype=torch.float32, device=self.device)
                dir_prev = torch.empty((cap0, 4), dtype=torch.float32, device=self.device)
                lineage = torch.empty(cap0, dtype=torch.int64, device=self.device) if self._lineage else None
                streamed = not job.get("resident", True) and not small
                if streamed:
                    # generation 0 arrives slice by slice: slice 0 now, slice k + 1 while slice k is searched (tile loop below)
                    slices = self._stream_slices(n)
                    self._upload_slice(job, slices[0], org_pow, dir_prev, lineage, stream)
                else:
                    if not job.get("resident", True):  # (a small scene never streams)
                        for key, h in zip(("origin", "dir", "power"), job["host"]):
                            job[key].copy_(h)
                        job["resident"] = True
                    rc = lib.lpcl_pack_rays(job["origin"].data_ptr(), job["dir"].data_ptr(), job["power"].data_ptr(), None, n,
                                            job["lineage_base"], org_pow.data_ptr(), dir_prev.data_ptr(),
                                            lineage.data_ptr() if lineage is not None else None, self._stream())
                    _cabi.check(rc, "lpcl_pack_rays")
                    self.kernel_launches += 1
                if small:
                    # small scene: every generation in ONE launch (csrc/lpcl_small.cuh); the general loop below only
                    # runs if the kernel hands a generation back (the rays multiplied beyond what one SM should trace)
                    t_first, gen, finished = self._run_small(small, org_pow, dir_prev, lineage, n, int(trace_iterations),
                                                             float(trace_until_dissipated), max_ray_len, ior_env, pow_shape,
                                                             results, stream)
                else:
                    gen = _Generation(org_pow, dir_prev, lineage, n, n, n)

            epoch_start = 0  # iteration at which the current lineage numbering started (dist.renumber_lineage)
            for t_iter in range(t_first, 0 if finished else int(trace_iterations)):
                n = gen.n if gen is not None else 0
                if self.verbose:
                    print("ITERATION", t_iter + 1, "with ", n, "rays.")
                if self._lineage and t_iter - epoch_start >= _dist.PATH_BITS:
                    self._renumber_lineage(gen)  # a collective with world_size > 1: every rank gets here
                    epoch_start = t_iter
                c = dict(n_reflect=0, n_refract=0, n_measured=0, n_terminated=0, live_power=0.0, measured_power=0.0,
                         candidates=0)
                children = None
                if n > 0:
                    cap = 2 * n
                    ch_org = torch.empty((cap, 4), dtype=torch.float32, device=self.device)
                    ch_dir = torch.empty((cap, 4), dtype=torch.float32, device=self.device)
                    ch_lin = torch.empty(cap, dtype=torch.int64, device=self.device) if self._lineage else None
                    children = _Generation(ch_org, ch_dir, ch_lin, 0, 0, n)
                    tile = self._tile_rays(n)
                    tiles = [(lo, min(lo + tile, n)) for lo in range(0, n, tile)]
                    upload_next = None
                    if t_iter == 0 and n > 0 and not job.get("resident", True):
                        tiles = slices  # one ray tile per uploaded slice
                        tile = max(hi_ - lo_ for lo_, hi_ in tiles)
                        upload_next = {lo_: sl for (lo_, _), sl in zip(slices, slices[1:])}
                    keep = self.keep_results
                    dest_meas = torch.empty((n if keep else min(n, tile), 4), dtype=torch.float32, device=self.device)
                    hit = torch.empty((n, 4), dtype=torch.int32, device=self.device) if self.record_hits else None
                    self._reserve_measured(self._measured_count + n)
                    ws = self._get_workspace(min(n, tile))
                    self._pick_search_variant(t_iter)
                    ev = []
                    n_tiles_run = len(tiles)
                    for lo, hi in tiles:
                        pg = gen.c_struct(lo, hi)
                        cg = children.child_struct(n, c["n_reflect"], c["n_refract"])
                        dm_ptr = dest_meas.data_ptr() + (16 * lo if keep else 0)
                        if prof:
                            ev.append([torch.cuda.Event(enable_timing=True) for _ in range(3)])
                            ev[-1][0].record(stream)
                        rc = lib.lpcl_trace_search(self._scene, C.byref(pg), ws.data_ptr(), C.c_float(max_ray_len),
                                                   counters_dev.data_ptr(), self._stream())
                        _cabi.check(rc, "lpcl_trace_search")
                        if prof:
                            ev[-1][1].record(stream)
                        rc = lib.lpcl_trace_resolve(
                            self._scene, C.byref(pg), C.byref(cg), dm_ptr,
                            hit.data_ptr() + 16 * lo if hit is not None else None, self._measured.data_ptr(),
                            self._measured_lineage.data_ptr() if self._measured_lineage is not None else None,
                            self._measured_count + c["n_measured"], self._measured.shape[0], counters_dev.data_ptr(),
                            ws.data_ptr(), C.c_float(max_ray_len), C.c_float(ior_env), t_iter - epoch_start,
                            self._stream())
                        _cabi.check(rc, "lpcl_trace_resolve")
                        if prof:
                            ev[-1][2].record(stream)
                        # scratch init, search (+ its second pass), fused resolve/Fresnel/compaction
                        self.kernel_launches += int(lib.lpcl_search_launch_count(self._scene)) + 1
                        counters_host.copy_(counters_dev, non_blocking=True)
                        if upload_next is not None:
                            if lo in upload_next:  # the next slice travels while this one is searched
                                self._upload_slice(job, upload_next[lo], gen.org_pow, gen.dir_prev, gen.lineage, stream)
                            else:
                                job["resident"] = True  # last slice launched: later run() calls find the rays in HBM
                        if lo == 0:
                            self._ensure_input_power()  # host work hidden behind the kernels just launched
                        stream.synchronize()
                        sc = _cabi.StepCounters.from_buffer_copy(counters_host.numpy().tobytes())
                        for k, _ in _cabi.StepCounters._fields_:
                            if k != "reserved":
                                c[k] += getattr(sc, k)
                    if prof:
                        events.append((n, ev))
                    children.n = c["n_reflect"] + c["n_refract"]
                    children.split = c["n_reflect"]
                    if self.keep_results:
                        results._append(gen, dest_meas, pow_shape if t_iter == 0 else None)
                    else:
                        results.ray_counts.append(n)
                    if self.record_hits:
                        self.hit_records.append(hit)
                    self._measured_count += c["n_measured"]
                    self._measured_per_iter.append(c["n_measured"])
                    self._candidate_density = c["candidates"] / (float(n) * float(self.tri_count))
                    if t_iter == 0:
                        self._gen0_density, self._gen0_density_tris = self._candidate_density, int(self.tri_count)
                else:
                    if self.keep_results:
                        empty = torch.zeros((0, 4), dtype=torch.float32, device=self.device)
                        results._append(_Generation(empty, empty, None, 0, 0, 0), empty,
                                        pow_shape if t_iter == 0 else None)
                    else:
                        results.ray_counts.append(0)
                    self._measured_per_iter.append(0)

                # it:372-391, evaluated over ALL ranks
                input_power = self._ensure_input_power()
                live_local = c["n_reflect"] + c["n_refract"]
                live, power_left = live_local, c["live_power"]
                if self.world_size > 1:
                    live, power_left = _dist.all_reduce_pair(live, power_left, self._group, self.device)
                for k, v in (("rays", n), ("live", live), ("live_local", live_local), ("power_left", power_left),
                             ("candidates", c["candidates"]), ("measured", c["n_measured"]),
                             ("terminated", c["n_terminated"]), ("tiles", n_tiles_run if n else 0)):
                    self.stats[k].append(v)
                gen = children if (children is not None and children.n > 0) else None
                if self.verbose:
                    print("Power left in scene:           ", power_left / input_power * 100.0 if input_power else 0.0)
                if power_left < (1.0 - float(trace_until_dissipated)) * input_power:
                    break
                if live == 0:
                    break
            if prof:
                stream.synchronize()
                general = [dict(rays=n_, search_ms=sum(e[0].elapsed_time(e[1]) for e in ev_),
                                resolve_ms=sum(e[1].elapsed_time(e[2]) for e in ev_), tiles=len(ev_)) for n_, ev_ in events]
                self.kernel_times = (self.kernel_times if self.stats.get("one_launch_generations") else []) + general
        self.timings["run_s"] = time() - t_start
        return self.results

    def _reserve_measured(self, need):
        torch = self._torch
        cap = 0 if self._measured is None else self._measured.shape[0]
        if cap >= need:
            return
        new_cap = max(need, int(cap * 1.5), 1024)
        buf = torch.empty((new_cap, 4), dtype=torch.float32, device=self.device)
        lin = torch.empty(new_cap, dtype=torch.int64, device=self.device) if self._lineage else None
        if self._measured is not None and self._measured_count:
            buf[: self._measured_count] = self._measured[: self._measured_count]
            if lin is not None:
                lin[: self._measured_count] = self._measured_lineage[: self._measured_count]
        self._measured, self._measured_lineage = buf, lin

    # ---------------------------------------------------------------------------------- small scenes
    SMALL_PAIR_LIMIT = 1 << 16  # ray x triangle pairs per generation one CTA traces in ~10 us

    def _small_plan(self, n, trace_iterations):
        """Arena sizes for the one-launch trace of a small scene (csrc/lpcl_small.cuh), or None when the general
        loop should run: several ranks (the termination tests are global), more than 8 meshes, a first generation
        beyond SMALL_PAIR_LIMIT ray x triangle pairs, or LPCL_SMALL_PATH=0."""
        tri = int(self.tri_count)
        if (n <= 0 or trace_iterations < 1 or self.world_size > 1 or int(self.mesh_count) > 8
                or n * tri > self.SMALL_PAIR_LIMIT or os.environ.get("LPCL_SMALL_PATH", "1") == "0"):
            return None
        per_gen = self.SMALL_PAIR_LIMIT // tri + 1                 # the largest generation the kernel accepts
        iters = min(trace_iterations, 31)
        return dict(scratch_rays=per_gen, iters=iters, arena_rays=min(n + 2 * iters * per_gen, 1 << 18))

    def _run_small(self, plan, org_pow, dir_prev, lineage, n0, trace_iterations, trace_until_dissipated, max_ray_len, ior_env,
                   pow_shape, results, stream):
        """Launch the one-CTA trace, read its table back and turn it into the same state the general loop builds
        (results, stats, measured buffer).  Returns (next iteration, generation to continue with, trace finished)."""
        torch, lib = self._torch, self._lib
        arena, iters = plan["arena_rays"], plan["iters"]
        dest_meas = torch.empty((arena, 4), dtype=torch.float32, device=self.device)
        hit = torch.empty((arena, 4), dtype=torch.int32, device=self.device) if self.record_hits else None
        scratch = torch.empty(int(lib.lpcl_small_scratch_bytes(self._scene, plan["scratch_rays"])) + 256, dtype=torch.uint8,
                              device=self.device)
        table = torch.empty((iters + 1, 8), dtype=torch.int64, device=self.device)  # row 0: header, rows 1..: generations
        self._reserve_measured(arena)
        if getattr(self, "_small_host", None) is None or self._small_host.shape[0] < iters + 1:
            self._small_host = torch.zeros((32, 8), dtype=torch.int64).pin_memory()
        host = self._small_host[: iters + 1]
        input_power = self._ensure_input_power()
        threshold = (1.0 - trace_until_dissipated) * input_power
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if self.profile else None
        if ev:
            ev[0].record(stream)
        rc = lib.lpcl_trace_small(
            self._scene, org_pow.data_ptr(), dir_prev.data_ptr(), lineage.data_ptr() if lineage is not None else None,
            dest_meas.data_ptr(), hit.data_ptr() if hit is not None else None, arena, n0, scratch.data_ptr(),
            plan["scratch_rays"], self._measured.data_ptr(),
            self._measured_lineage.data_ptr() if self._measured_lineage is not None else None, self._measured.shape[0],
            C.c_float(max_ray_len), C.c_float(ior_env), trace_iterations, C.c_double(threshold), self.SMALL_PAIR_LIMIT,
            table[1:].data_ptr(), table.data_ptr(), self._stream())
        _cabi.check(rc, "lpcl_trace_small")
        if ev:
            ev[1].record(stream)
        self.kernel_launches += 1
        host.copy_(table, non_blocking=True)
        stream.synchronize()
        raw = host.numpy()
        hdr = _cabi.SmallHeader.from_buffer_copy(raw[0].tobytes())
        done = int(hdr.iterations)
        split, base, cap = n0, n0, n0
        tri = int(self.tri_count)
        for k in range(done):
            rec = _cabi.SmallIter.from_buffer_copy(raw[1 + k].tobytes())
            n, off = int(rec.n_rays), int(rec.offset)
            if self.keep_results:
                gen = _Generation(org_pow[off: off + cap], dir_prev[off: off + cap],
                                  lineage[off: off + cap] if lineage is not None else None, n, split, base)
                results._append(gen, dest_meas[off: off + n], pow_shape if k == 0 else None)
            else:
                results.ray_counts.append(n)
            if self.record_hits:
                self.hit_records.append(hit[off: off + n])
            live = int(rec.n_reflect + rec.n_refract)
            for key, v in (("rays", n), ("live", live), ("live_local", live), ("power_left", float(rec.live_power)),
                           ("candidates", n * tri), ("measured", int(rec.n_measured)), ("terminated", int(rec.n_terminated)),
                           ("tiles", 1)):
                self.stats[key].append(v)
            self._measured_per_iter.append(int(rec.n_measured))
            self._measured_count += int(rec.n_measured)
            split, base, cap = int(rec.n_reflect), n, 2 * n
            if self.verbose:
                print("ITERATION", k + 1, "with ", n, "rays.")
                print("Power left in scene:           ", rec.live_power / input_power * 100.0 if input_power else 0.0)
        if ev:
            self.kernel_times = [dict(rays=sum(self.stats["rays"]), search_ms=ev[0].elapsed_time(ev[1]), resolve_ms=0.0,
                                      tiles=1, one_launch_trace=True)]
        self.stats["one_launch_generations"] = done
        if int(hdr.stop_reason) == 0:
            return done, None, True
        off, n = int(hdr.next_offset), int(hdr.next_n)
        gen = _Generation(org_pow[off:], dir_prev[off:], lineage[off:] if lineage is not None else None, n,
                          int(hdr.next_split), int(hdr.next_base))
        return done, (gen if n > 0 else None), False

    # Epilogue variant of the tensor-core search (lpcl_scene_set_tc_warps): 0 = the library default, the deferred-flag
    # sweep + second pass, which is the fastest one at every candidate density measured (C2 - C5); 8 / 16 = the
    # variants that vote after every chunk test (fat / thin epilogue warps), kept for cross-checks.
    SEARCH_VARIANT = 0

    def _pick_search_variant(self, t_iter):
        """Tell the scene handle which epilogue variant the next generation's search uses (all give bit-identical
        results).  Until round 2 the choice followed the candidate density of the previous generation (thin warps above
        8 % flagged chunk tests); the deferred-flag sweep made it moot."""
        want = int(self.SEARCH_VARIANT)
        if want != self._tc_warps_set:
            _cabi.check(self._lib.lpcl_scene_set_tc_warps(self._scene, want), "lpcl_scene_set_tc_warps")
            self._tc_warps_set = want
        self.stats.setdefault("tc_warps", []).append(want)

    STREAM_MIN_RAYS = 1 << 18   # below this a generation 0 is uploaded in one piece
    STREAM_FIRST = 0.25         # share of the rays in the first slice (its upload is the part nothing hides)

    def _stream_slices(self, n):
        """[lo, hi) of the slices generation 0 is uploaded and traced in: a first slice of a quarter of the rays (a
        multiple of 4096), then the rest - its upload (3 ms per 1M rays) hides behind the first slice's search (6 ms) -
        cut further only where the scratch budget demands it.  Every extra slice costs ~0.5 ms of launch tail and
        synchronisation, so two beat four (measured: device loop +1.7 ms with four equal slices)."""
        first = min(max(4096, int(n * self.STREAM_FIRST) // 4096 * 4096), self._tile_rays(n))
        cuts, lo, cap = [(0, first)], first, self._tile_rays(n)
        while lo < n:
            hi = min(lo + cap, n)
            cuts.append((lo, hi))
            lo = hi
        return cuts

    def _upload_slice(self, job, sl, org_pow, dir_prev, lineage, stream):
        """Host -> device copy of source rays [lo, hi) on the copy stream, packed into the generation-0 buffers there;
        the compute stream waits for it before the slice's search.  A copy from pageable memory keeps the HOST busy
        until it is staged, but the GPU goes on with the kernels already queued on the compute stream."""
        torch = self._torch
        lo, hi = sl
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        if lo == 0:
            cs.wait_stream(stream)  # the buffers were allocated on the compute stream (idle at this point); the later
            # slices must NOT wait for the compute stream - the point is to travel while it is busy
        with torch.cuda.stream(cs):
            for key, h in zip(("origin", "dir", "power"), job["host"]):
                job[key][lo:hi].copy_(h[lo:hi], non_blocking=True)
            rc = self._lib.lpcl_pack_rays(job["origin"].data_ptr() + 16 * lo, job["dir"].data_ptr() + 16 * lo,
                                          job["power"].data_ptr() + 4 * lo, None, hi - lo, job["lineage_base"] + lo,
                                          org_pow.data_ptr() + 16 * lo, dir_prev.data_ptr() + 16 * lo,
                                          lineage.data_ptr() + 8 * lo if lineage is not None else None,
                                          C.c_void_p(cs.cuda_stream))
            _cabi.check(rc, "lpcl_pack_rays")
        self.kernel_launches += 1
        stream.wait_stream(cs)

    def _tile_rays(self, n):
        """Rays per tile of a generation of n rays: as many as the scratch budget allows (per ray: 12 B per mesh of
        search results + the record budget of the deferred-flag search, as the library sizes them, + the per-tile end
        points), in multiples of 4096, at least 65,536."""
        if self._scratch_per_ray is None:  # marginal workspace bytes per ray, from the library's own sizing
            m = 1 << 20
            grow = int(self._lib.lpcl_step_workspace_bytes(self._scene, 2 * m)) - int(self._lib.lpcl_step_workspace_bytes(self._scene, m))
            self._scratch_per_ray = max(12 * int(self.mesh_count), -(-grow // m))
        per_ray = self._scratch_per_ray + 16 + 1
        tile = max(65536, (self.ray_tile_bytes // per_ray) // 4096 * 4096)
        return n if n <= tile else tile

    def _renumber_lineage(self, gen):
        """Start a new lineage epoch (the 31 branch bits of the 64-bit id are used up): every live
        ray's id becomes its rank in the generation's global reference order."""
        torch = self._torch
        keys = gen.logical(gen.lineage) if gen is not None else torch.zeros(0, dtype=torch.int64, device=self.device)
        new = _dist.renumber_lineage(keys, self._group, self.device)
        if gen is not None:
            gen.lineage[: gen.split] = new[: gen.split]
            gen.lineage[gen.base: gen.base + (gen.n - gen.split)] = new[gen.split:]

    # ---------------------------------------------------------------------------------- results
    def get_measured_rays(self, gather=True):
        """it:395-411: ``(pos (K,4) f32, pwr (K,) f32)`` of every ray that ended on a measure
        surface, iteration-major, ray order inside an iteration.  Read from the per-GPU
        measured-ray buffer the step kernel appends to (no pass over ``self.results``).

        With world_size > 1 and ``gather=True`` every rank returns the merged global arrays in
        the reference order (records are merged by (iteration, lineage))."""
        pos_pwr = self.measured_device()
        per_iter = list(self._measured_per_iter)
        if self.world_size > 1 and gather:
            lin = (self._measured_lineage[: self._measured_count] if self._measured_lineage is not None
                   else self._torch.zeros(0, dtype=self._torch.int64, device=self.device))  # a rank that never had rays
            pos_pwr = _dist.gather_measured(pos_pwr, lin, per_iter, self._group, self.device)
        if pos_pwr.is_cuda:  # split (x, y, z, power) on the GPU: two device->host copies, no host passes
            pwr_dev = pos_pwr[:, 3].contiguous()
            pos_dev = pos_pwr.clone()
            pos_dev[:, 3] = 0.0
            pos, pwr = self._host_result_arrays(int(pos_pwr.shape[0]))
            self._torch.from_numpy(pos).copy_(pos_dev)
            self._torch.from_numpy(pwr).copy_(pwr_dev)
        else:
            pos = np.ascontiguousarray(pos_pwr.numpy().copy())
            pwr = np.ascontiguousarray(pos[:, 3].copy())
            pos[:, 3] = 0.0
        if len(self._measured_per_iter) == 1 and self.results is not None and len(self.results._dev) == 1 \
                and self.results._dev[0][2] is not None:
            pwr = pwr.reshape((-1,) + tuple(self.results._dev[0][2][1:]))  # it:406: no flatten for k == 0
        return (pos, pwr)

    def _host_array(self, name, rows, cols, dtype):
        """A ``(rows, cols)`` (or ``(rows,)`` when cols == 0) host array from the recycling pool
        (``_HostPool``): its memory is reused only after the caller dropped every array derived from it."""
        return self._host_pool.array((name, cols, np.dtype(dtype).str), (rows, cols) if cols else (rows,), dtype)

    def _host_result_arrays(self, k):
        """``(pos (k,4), pwr (k,))`` float32 arrays for get_measured_rays, from the recycling pool."""
        return self._host_array("pos", k, 4, np.float32), self._host_array("pwr", k, 0, np.float32)

    def measured_device(self):
        """The measured-ray records of THIS GPU as a (K,4) float32 CUDA tensor (x, y, z, power)."""
        torch = self._torch
        if self._measured is None:
            return torch.zeros((0, 4), dtype=torch.float32, device=self.device)
        return self._measured[: self._measured_count]

    def _binned(self, kind, limits, points, rot=None):
        """Projection + weighted 2-D histogram on the GPU (csrc/lpcl_project.cuh): the measured-ray
        buffer stays in HBM, only the (points x points) float64 histogram is downloaded.  With
        world_size > 1 every rank bins its own records and the histograms are all-reduced."""
        from . import projection

        H, xe, ye = projection.project_bin_device(self._lib, self.measured_device(), kind, limits, int(points), rot)
        if self.world_size > 1:
            H = _dist.all_reduce_array(H, self._group, self.device)
        self.hist_data = (H, xe, ye)
        return self.hist_data

    def get_binned_data(self, limits=((-10, 10), (-10, 10)), points=500):
        """it:413-418: 2-D histogram of measured end points (x, y) weighted by power."""
        return self._binned("xy", limits, points)

    def get_binned_data_stereographic(self, limits=((-1, 1), (-1, 1)), points=500):
        """it:503-531: stereographic projection (cl:488-506) + binning, weights / (dx dy)."""
        return self._binned("stereographic", limits, points)

    def get_binned_data_angular(self, limits=((-1, 1), (-1, 1)), points=500):
        """it:534-562: azimuth/elevation map (cl:509-538) + binning, weights / (dx dy)."""
        return self._binned("angular", limits, points)

    def get_binned_data_replicated(self, limits=((-10, 10), (-10, 10)), points=500, axis="z", sources=36):
        """Data part of replicate_lightsources_and_plot (it:564-628): emulate ``sources`` light
        sources by rotating the measured rays about the symmetry ``axis`` in steps of 2 pi / sources
        (angular projection per step) and summing the histograms."""
        from . import light_source as _ls

        R = {"x": _ls._Rx, "X": _ls._Rx, "y": _ls._Ry, "Y": _ls._Ry, "z": _ls._Rz, "Z": _ls._Rz}.get(axis, _ls._Rx)
        total = None
        for k in range(int(sources)):
            H, xe, ye = self._binned("angular", limits, points, rot=R(k * 2.0 * np.pi / sources).astype(np.float32))
            total = H if total is None else total + H
        self.hist_data = (total, xe, ye)
        return self.hist_data

    # ---------------------------------------------------------------------------------- beam metrics
    def _elevations(self, pole):
        (pos, pwr) = self.get_measured_rays()
        p = np.asarray(pos, dtype=np.float32)
        pos0 = p / np.linalg.norm(p, axis=1)[:, None]  # it:424,447,472
        return np.arccos(np.dot(pos0, np.asarray(pole))), np.float64(np.asarray(pwr).reshape(-1))

    def get_elevation_histogram(self, points=500, pole=[0, 0, 1, 0]):
        """Data part of plot_elevation_histogram (it:420-441): power per elevation bin divided by
        sin(elevation) * d_elevation.  Returns (elevation bin centres [rad], intensity)."""
        elevation, pwr = self._elevations(pole)
        (H, x) = np.histogram(elevation, bins=points, weights=pwr)
        x = (x[0:-1] + x[1:]) / 2.0
        dx = x[1] - x[0]
        return x, H / (np.sin(x) * dx)

    def get_beam_width_half_power(self, points=500, pole=[0, 0, 1, 0]):
        """it:443-466: elevation inside which half of the measured power falls.  Prints the
        reference's three lines and returns (total power, half-power angle [rad], throughput)."""
        elevation0, pwr0 = self._elevations(pole)
        order = np.argsort(elevation0, kind="stable")
        elevation, pwr = elevation0[order], pwr0[order]
        pwr_cumsum = np.cumsum(pwr)
        pwr_hmax = pwr_cumsum[-1] / 2.0
        dev = np.absolute(pwr_cumsum - pwr_hmax)
        idx = np.where(dev == dev.min())
        elevation_hmax, pwr_sum_hmax = elevation[idx], pwr_cumsum[idx]
        print("")
        print("Throughput results:")
        print("===================")
        print("Total measured power: ", pwr_cumsum[-1])
        print("Beam halfpower angle: ", elevation_hmax / np.pi * 180.0, "Deg")
        print("Halfwidth throughput: ", pwr_sum_hmax / pwr_cumsum[-1] * 100.0, "%")
        return pwr_cumsum[-1], elevation_hmax, pwr_sum_hmax / pwr_cumsum[-1]

    def get_beam_HWHM(self, points=500, pole=[0, 0, 1, 0]):
        """it:468-501: half width at half maximum of the elevation intensity histogram.  Prints the
        reference's three lines and returns (total power, HWHM angle [rad], throughput)."""
        elevation0, pwr0 = self._elevations(pole)
        order = np.argsort(elevation0, kind="stable")
        (H, x) = np.histogram(a=elevation0[order], bins=points, weights=pwr0[order])
        x = (x[0:-1] + x[1:]) / 2.0
        H = (H.T / np.sin(x)).flatten()
        pwr_hmax = max(H) / 2.0
        dev = np.absolute(H - pwr_hmax)
        idx = np.where(dev == dev.min())[0]
        elevation_hmax = x[idx]
        total = pwr0.sum()
        pwr_sum_hmax = np.sum(pwr0[np.where(elevation0 < elevation_hmax[0])]) if len(elevation_hmax) else 0.0
        print("")
        print("Throughput results HWHM:")
        print("========================")
        print("Total measured power: ", total)
        print("Beam HWHM angle:      ", elevation_hmax / np.pi * 180.0, "Deg")
        print("HWHM throughput:      ", pwr_sum_hmax / total * 100.0, "%")
        return total, elevation_hmax, pwr_sum_hmax / total

    # ---------------------------------------------------------------------------------- persistence
    def pickle_results(self):
        """it:711-733: dump (results, meshes) to ./<timestamp>-tracer_results.txt."""
        try:
            payload = pickle.dumps((list(self.results), self.meshes), 1)
            fname = "./{0}-tracer_results.txt".format(strftime("%Y.%m.%d.%H.%M.%S"))
            with open(fname, "wb") as f:
                f.write(payload)
        except Exception:
            print("Pickling results failed.")
            return None
        print("Pickled results and meshes to ", fname)
        return fname

    def load_pickle_results(self, path):
        """it:735-751."""
        try:
            with open(path, "rb") as f:
                (self.results, self.meshes) = pickle.loads(f.read())
        except Exception:
            print("UnPickling results failed.")
        else:
            print("UnPickled results and meshes from ", path)

    def save_traced_scene(self, dxf_file):
        """it:754-779: rays and geometry as DXF 3DFACE entities (dependency-free writer)."""
        from ._dxf import write_faces

        o = np.concatenate([r[0][:, :3] for r in self.results]) if len(self.results) else np.zeros((0, 3))
        d = np.concatenate([r[1][:, :3] for r in self.results]) if len(self.results) else np.zeros((0, 3))
        (v0, v1, v2) = self.geometry
        write_faces(dxf_file, [("Rays", 3, o, d, d), ("Geometry", 5, v0[:, :3], v1[:, :3], v2[:, :3])])
