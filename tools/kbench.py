#!/usr/bin/env python
"""Kernel A/B harness: times the device-resident trace of one or more workloads under several builds of
liblightpycl_b200.so in ONE process (one torch import, one scene build per workload), and checks that every
build measures exactly the same rays.

    python tools/kbench.py --libs a.so b.so ... [--workloads c3_parabolic_mirror c2_directivity_lens] [--reps 3]

Output: one line per (workload, lib): tests/s of the whole trace, search-kernel ms per generation.
Development tool (not part of the product path, not a bench line)."""
import argparse
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", nargs="+", required=True)
    ap.add_argument("--workloads", nargs="+", default=["c3_parabolic_mirror"])
    ap.add_argument("--rays", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--iters", type=int, default=0, help="cap trace_iterations (0 = the workload's)")
    ap.add_argument("--trace", action="store_true", help="print the clock64 trace of a -DLPCL_TC_TRACE build")
    args = ap.parse_args()

    import numpy as np
    import torch

    from lightpycl_b200 import _cabi, iterative_tracer as it, scenes

    base_env = {k: v for k, v in os.environ.items() if k.startswith("LPCL_")}
    for wl in args.workloads:
        kw = dict(ray_count=args.rays)
        if wl == "c5_lenslet_array":
            kw = dict(ray_count=args.rays // 4, sources=4)
        ls, meshes, params = scenes.CONFIGS[wl](**kw)
        if args.iters:
            params = dict(params, trace_iterations=args.iters)
        ref_sig = None
        for spec in args.libs:
            # "path.so@NAME=VALUE,NAME2=VALUE2": run-time switches of the library for this run only
            lib, _, envs = spec.partition("@")
            for old in [k for k in os.environ if k.startswith("LPCL_") and k not in base_env]:
                del os.environ[old]
            os.environ.update(base_env)
            for kv in filter(None, envs.split(",")):
                os.environ[kv.split("=")[0]] = kv.split("=")[1]
            _cabi._lib, _cabi.LIB_PATH = None, os.path.abspath(lib)
            tr = it.CL_Tracer(keep_results=False, profile=True)
            tr.prepare(ls, meshes)
            tr.run(**params)
            best, search = None, None
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                tr.run(**params)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1)
                if best is None or ms < best:
                    best, search = ms, [round(k["search_ms"], 3) for k in tr.kernel_times]
            tests = sum(tr.stats["rays"]) * int(tr.tri_count)
            pos, pwr = tr.get_measured_rays()
            sig = hashlib.sha1(np.ascontiguousarray(pos).tobytes() + np.ascontiguousarray(pwr).tobytes()).hexdigest()[:12]
            ref_sig = ref_sig or sig
            cand = [round(c / max(r, 1), 2) for c, r in zip(tr.stats["candidates"], tr.stats["rays"])]
            print("%-22s %-28s %.3e tests/s  %8.2f ms  search ms/gen %s  cand/ray %s  measured %s %s" % (
                wl, os.path.basename(spec), tests / (best * 1e-3), best, search[:6], cand[:4], sig,
                "OK" if sig == ref_sig else "MISMATCH vs first lib"), flush=True)
            if args.trace and hasattr(tr._lib, "lpcl_debug_read_trace"):
                import ctypes as C

                buf = np.zeros((3, 48, 8), dtype=np.int64)
                tr._lib.lpcl_debug_read_trace(C.c_void_p(buf.ctypes.data), C.c_int64(buf.nbytes))
                names = ("epilogue warp 0 (half 0)", "epilogue warp 4 (half 1)")
                for w in range(2):
                    t = buf[w]
                    seg = np.diff(t, axis=1)  # slots: 0 top, 1 after wait::ld, 2 after arrive/queue, 3 after test A, 4 after tfull, 5 after ld A', 6 after test B, 7 after ld B'
                    period = np.diff(t[:, 0])
                    print("  trace %s: step period median %d cycles; segments (median cycles): wait::ld %d, arrive+queue %d, test A %d, "
                          "tfull wait %d, fence+ld A' %d, test B %d, ld B' %d, loop back %d" % (
                              names[w], np.median(period), *[np.median(seg[:, k]) for k in range(7)],
                              np.median(t[1:, 0] - t[:-1, 7])))
                m = buf[2]
                print("  trace MMA warp: step period median %d; tempty wait %d, issue %d, loop %d" % (
                    np.median(np.diff(m[:, 0])), np.median(m[:, 1] - m[:, 0]), np.median(m[:, 2] - m[:, 1]),
                    np.median(m[1:, 0] - m[:-1, 2])))
                print("  offset warp4 - warp0 at step top (median): %d" % np.median(buf[1][:, 0] - buf[0][:, 0]))
            del tr
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
