#!/usr/bin/env python
"""Per-step timeline and SM clock of the deferred-flag search (development tool; needs a -DLPCL_TC_TRACE build):

    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DLPCL_TC_TRACE \
         -o /tmp/trace.so lightpycl_b200/csrc/lpcl_cabi.cu
    python tools/tc_trace.py /tmp/trace.so [workload]

* clock64 stamps of CTA (0,0)'s epilogue warps 0 and 4 (one SM sub-partition) and of its MMA warp over 48 steps:
  per step and per position of the step in its triangle tile (j = 0 .. 3), the median cycles of every segment;
* clock64 against globaltimer over the whole sweep of one CTA per 64 ray blocks: the SM clock the kernel really runs at
  (measured: 1.84 GHz of 1.965 once settled, 1.64 - 1.73 GHz during the first ~1.5 ms of a launch).
The stamps serialise the step (a trace build is ~15 % slower): read proportions, not absolutes."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightpycl_b200 import _cabi, iterative_tracer as it, scenes  # noqa: E402


def main():
    _cabi._lib, _cabi.LIB_PATH = None, os.path.abspath(sys.argv[1])
    wl = sys.argv[2] if len(sys.argv) > 2 else "c3_parabolic_mirror"
    ls, meshes, params = scenes.CONFIGS[wl](ray_count=1_000_000)
    params = dict(params, trace_iterations=1)
    tr = it.CL_Tracer(keep_results=False, profile=True)
    tr.prepare(ls, meshes)
    tr.run(**params)
    tr.run(**params)
    print("search ms", [round(k["search_ms"], 3) for k in tr.kernel_times])
    buf = np.zeros((3, 48, 8), dtype=np.int64)
    tr._lib.lpcl_debug_read_trace(C.c_void_p(buf.ctypes.data), C.c_int64(buf.nbytes))
    names = ["wait::ld", "release", "acc wait", "fold A", "ld A'", "fold B", "ld B'"]
    for w in range(2):
        t = buf[w]
        seg = np.diff(t, axis=1)
        print("epilogue warp %d: step period median %d cycles" % (4 * w, np.median(np.diff(t[:, 0]))))
        for j in range(4):
            rows, loop = seg[j::4], (t[1:, 0] - t[:-1, 7])[j::4]
            print("   j=%d " % j, " ".join("%s %4d" % (n, np.median(rows[:, k])) for k, n in enumerate(names)),
                  " to next step %4d (p90 %4d)  period %d" % (np.median(loop), np.percentile(loop, 90),
                                                            np.median(np.diff(t[:, 0])[j::4])))
    m = buf[2]
    for j in range(4):
        print("MMA warp j=%d: accumulator wait %4d (p90 %4d)  issue %3d  to next step %4d" % (
            j, np.median((m[:, 1] - m[:, 0])[j::4]), np.percentile((m[:, 1] - m[:, 0])[j::4], 90),
            np.median((m[:, 2] - m[:, 1])[j::4]), np.median((m[1:, 0] - m[:-1, 2])[j::4])))
    clk = (C.c_longlong * 128)()
    tr._lib.lpcl_debug_read_clk(clk)
    c = np.array(list(clk)).reshape(32, 4)
    live = c[:, 3] > c[:, 2]
    t0 = c[live, 2].min() if live.any() else 0
    for k in range(32):
        if live[k]:
            print("ray block %4d: starts at %8.1f us, sweep of %7d cycles in %7d ns -> %.0f MHz" % (
                64 * k, (c[k, 2] - t0) / 1e3, c[k, 1] - c[k, 0], c[k, 3] - c[k, 2], (c[k, 1] - c[k, 0]) * 1e3 / (c[k, 3] - c[k, 2])))


if __name__ == "__main__":
    main()
