"""Parity at BASELINE size (VERDICT r1 "Next" item 1), two independent ways.

(a) ``test_filtered_searches_equal_exact_search``: the two filtered nearest-hit searches (tensor-core
    filter, packed-FP32 filter) against ``LPCL_SEARCH_IMPL=exact`` - a kernel with NO filter that runs
    the reference-ordered evaluation (cl:50-101, 277-278) on EVERY (ray, triangle) pair - on the
    full-size scenes and ray sets: per (ray, mesh) nearest t, crossing count and nearest triangle
    must be bit-identical.  A conservative filter that ever dropped a pair the reference accepts
    would show up here as a difference; the count is asserted to be zero and printed.

(b) ``test_sampled_rays_of_full_size_trace_equal_reference_oracle``: rays are independent
    (cl:247-249), so the records of any subset of the source rays in a full-size trace must equal a
    trace of just that subset.  The full-size trace carries a 64-bit lineage per ray; a random
    sample of source rays is traced through the CPU oracle (the reference kernels themselves where
    oracle/_ref is present) and every generation's origin / dest / power / measured flag / hit mesh /
    hit triangle / n1 / n2 are compared bit for bit, ray by ray, in the reference's order.
"""

import ctypes as C
import os

import numpy as np
import pytest

import util  # noqa: F401  (sys.path side effect of the shared helpers)
from oracle import oracle

pytestmark = pytest.mark.gpu


def _tracer(**kw):
    from lightpycl_b200 import iterative_tracer as it

    return it.CL_Tracer(platform_name="NVIDIA", device_name="B200", **kw)


def _search(tr, origin, direc, max_ray_len, impl):
    """One call of the reference-seam entry point lpcl_intersect (it:288) under a given search
    implementation -> (near_t bits, hits, near_tri) as CUDA tensors of shape (n, mesh_count)."""
    import torch

    n, m = int(origin.shape[0]), int(tr.mesh_count)
    near_t = torch.empty((n, m), dtype=torch.float32, device=tr.device)
    hits = torch.empty((n, m), dtype=torch.int32, device=tr.device)
    near_tri = torch.empty((n, m), dtype=torch.int32, device=tr.device)
    scratch = torch.empty(int(tr._lib.lpcl_scratch_bytes(tr._scene, n)), dtype=torch.uint8, device=tr.device)
    old = os.environ.get("LPCL_SEARCH_IMPL")
    os.environ["LPCL_SEARCH_IMPL"] = "tc" if impl.startswith("tc") else impl
    if impl in ("tc8", "tc16"):  # the tensor-core search variants that vote per chunk test (fat / thin warps)
        os.environ["LPCL_TC_WARPS"] = impl[2:]
    try:
        rc = tr._lib.lpcl_intersect(tr._scene, origin.data_ptr(), direc.data_ptr(), n, C.c_float(max_ray_len),
                                    near_t.data_ptr(), hits.data_ptr(), near_tri.data_ptr(), scratch.data_ptr(),
                                    tr._stream())
    finally:
        os.environ.pop("LPCL_TC_WARPS", None)
        if old is None:
            os.environ.pop("LPCL_SEARCH_IMPL", None)
        else:
            os.environ["LPCL_SEARCH_IMPL"] = old
    assert rc == 0, tr._lib.lpcl_last_error()
    torch.cuda.synchronize(tr.device)
    return near_t.view(torch.int32), hits, near_tri


# (config, builder kwargs, generations compared): C2/C3 = BASELINE's 1M rays; C4 = 4 x 1M rays, first
# three generations (4M, 8M, 8M rays); C5 = the full 1,008,778-triangle lenslet array at 1M rays
FULL_SIZE = [
    ("c2_directivity_lens", dict(ray_count=1_000_000), 4),
    ("c3_parabolic_mirror", dict(ray_count=1_000_000), 2),
    ("c4_human_eye", dict(ray_count=1_000_000, sources=4), 3),
    ("c5_lenslet_array", dict(ray_count=262_144, sources=4), 2),
]


@pytest.mark.parametrize("name,kw,gens", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_filtered_searches_equal_exact_search(name, kw, gens):
    import torch

    from lightpycl_b200 import scenes

    ls, meshes, params = scenes.CONFIGS[name](**kw)
    tr = _tracer(keep_results=True)
    tr.iterative_tracer(ls, meshes, **dict(params, trace_iterations=gens, trace_until_dissipated=1.0))
    if name == "c5_lenslet_array":
        assert int(tr.tri_count) == 1_008_778
    mrl = float(params["max_ray_len"])
    total_pairs = 0
    for k, (gen, _dm, _ps) in enumerate(tr.results._dev):
        if gen.n == 0:
            continue
        origin = gen.logical(gen.org_pow).contiguous()
        direc = gen.logical(gen.dir_prev).contiguous()
        want = _search(tr, origin, direc, mrl, "exact")
        for impl in ("tc", "tc8", "tc16", "simt"):
            got = _search(tr, origin, direc, mrl, impl)
            bad = [int((g != w).sum().item()) for g, w in zip(got, want)]
            print("%s generation %d: %d rays x %d triangles, %s vs exact: near_t %d, hits %d, near_tri %d differences"
                  % (name, k, gen.n, int(tr.tri_count), impl, *bad))
            assert bad == [0, 0, 0], (name, k, impl, bad)
            del got
        total_pairs += gen.n * int(tr.tri_count)
        del want, origin, direc
        torch.cuda.empty_cache()
    assert total_pairs > 0


# (config, builder kwargs, extra tracer params, sampled source rays)
SAMPLED = [
    ("c2_directivity_lens", dict(ray_count=1_000_000), {}, 20_000),
    ("c3_parabolic_mirror", dict(ray_count=1_000_000), {}, 20_000),
    ("c4_human_eye", dict(ray_count=1_000_000, sources=4), dict(trace_iterations=6), 4_000),
    ("c5_lenslet_array", dict(ray_count=1_000_000, sources=16), {}, 4_000),
]


@pytest.mark.parametrize("name,kw,extra,n_sample", SAMPLED, ids=[c[0] for c in SAMPLED])
def test_sampled_rays_of_full_size_trace_equal_reference_oracle(name, kw, extra, n_sample):
    import torch

    from lightpycl_b200 import iterative_tracer as it, scenes

    ls, meshes, params = scenes.CONFIGS[name](**kw)
    params = dict(params, **extra)
    tr = _tracer(keep_results=True, record_hits=True, lineage=True)
    tr.iterative_tracer(ls, meshes, **params)
    n_iter = len(tr.results)
    origin, direc, power = it.flatten_light_sources(ls)
    n_src = len(origin)

    rng = np.random.RandomState(20261018)
    sample = np.sort(rng.choice(n_src, size=n_sample, replace=False)).astype(np.int64)
    sub = util.ArraySource(origin[sample], direc[sample], power.reshape(n_src, -1)[sample].reshape(
        (n_sample,) + power.shape[1:]))
    backend = "reference" if oracle.available("reference") else "port"
    oracle.set_threads(backend)
    # the same number of generations as the full run; the power test cannot fire (threshold 0)
    ref = oracle.trace([sub], meshes, backend=backend, power_sum="f64", lineage0=sample,
                       **dict(params, trace_iterations=n_iter, trace_until_dissipated=1.0))

    in_sample = torch.zeros(n_src, dtype=torch.bool, device=tr.device)
    in_sample[torch.from_numpy(sample).to(tr.device)] = True
    compared, mismatches = 0, {}
    for k in range(n_iter):
        gen, dest_meas, _ps = tr.results._dev[k]
        if gen.n == 0:
            assert k >= len(ref["results"]) or len(ref["results"][k][0]) == 0
            continue
        lin = gen.logical(gen.lineage)
        idx = torch.nonzero(in_sample[lin & 0xFFFFFFFF]).reshape(-1)
        got_lin = lin[idx].cpu().numpy()
        if k >= len(ref["results"]):
            assert len(got_lin) == 0, (name, k, "the oracle's sample died out, the GPU's did not")
            continue
        want_lin = ref["lineage"][k]
        # same rays, same ORDER (the reference's, it:366-373)
        assert np.array_equal(got_lin, want_lin), (name, k, len(got_lin), len(want_lin))
        op = gen.logical(gen.org_pow)[idx].cpu().numpy()
        dm = dest_meas[idx].cpu().numpy()
        hit = tr.hit_records[k][idx].cpu().numpy()
        o_ref, d_ref, p_ref, m_ref = ref["results"][k]
        dbg = ref["debug"][k]
        pairs = {
            "origin": (op[:, :3], o_ref[:, :3]),
            "dest": (dm[:, :3], d_ref[:, :3]),
            "power": (op[:, 3], np.asarray(p_ref).reshape(-1)),
            "measured": (np.ascontiguousarray(dm[:, 3]).view(np.int32), m_ref),
            "hit_mesh": (hit[:, 0], dbg["hit_mesh"]),
            "hit_tri": (hit[:, 1], dbg["hit_tri"]),
            "n1": (hit[:, 2], dbg["n1"]),
            "n2": (hit[:, 3], dbg["n2"]),
        }
        for what, (a, b) in pairs.items():
            bad = int(np.count_nonzero(util.bits(np.ascontiguousarray(a)) != util.bits(np.ascontiguousarray(b))))
            mismatches[what] = mismatches.get(what, 0) + bad
        compared += len(got_lin)
    print("%s: %d source rays sampled out of %d, %d generations, %d ray records compared with the %s oracle; "
          "mismatches %s" % (name, n_sample, n_src, n_iter, compared, backend, mismatches))
    assert compared >= n_sample
    assert all(v == 0 for v in mismatches.values()), mismatches
    # measured power of the sample: the north_star's third criterion on the sampled rays (<= 1e-6 relative)
    rpos, rpwr = oracle.measured_rays(ref["results"])
    if rpwr is not None and np.asarray(rpwr).size:
        mine = 0.0
        for k in range(min(n_iter, len(ref["results"]))):
            gen, dest_meas, _ps = tr.results._dev[k]
            if gen.n == 0:
                continue
            lin = gen.logical(gen.lineage)
            sel = in_sample[lin & 0xFFFFFFFF] & (dest_meas[:, 3].view(torch.int32) == 1)
            mine += float(gen.logical(gen.org_pow)[sel][:, 3].double().sum().item())
        want = float(np.sum(np.asarray(rpwr, np.float64)))
        assert abs(mine - want) <= 1e-6 * abs(want), (mine, want)
