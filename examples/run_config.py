#!/usr/bin/env python
"""Run one of the BASELINE configurations (C1-C5, see lightpycl_b200/scenes.py) the way the
reference's example scripts do (example_*.py: build sources and meshes, CL_Tracer(...),
iterative_tracer(...), get_measured_rays(), print the RI/s summary) - on a B200.

    python examples/run_config.py --config c2_directivity_lens --rays 1000000
"""

import argparse
import os
import sys
from time import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lightpycl_b200 import iterative_tracer as it  # noqa: E402
from lightpycl_b200 import scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2_directivity_lens", choices=sorted(scenes.CONFIGS))
    ap.add_argument("--rays", type=int, default=100_000, help="rays per source")
    ap.add_argument("--bins", type=int, default=30)
    a = ap.parse_args()

    time1 = time()
    ls, meshes, params = scenes.CONFIGS[a.config](ray_count=a.rays)
    prep_time = time() - time1

    tr = it.CL_Tracer(platform_name="NVIDIA", device_name="B200")
    time1 = time()
    tr.iterative_tracer(light_source=ls, meshes=meshes, **params)
    measured_rays = tr.get_measured_rays()
    sim_time = time() - time1

    proc_ray_count = sum(tr.results.ray_counts)  # == sum(len(res[3]) for res in tr.results), without a download
    print("Scene setup              ", prep_time, "s;", int(tr.tri_count), "triangles in", int(tr.mesh_count), "meshes")
    print("Rays per iteration       ", tr.results.ray_counts)
    print("Processed                ", proc_ray_count, "rays in ", sim_time, "s")
    print("Measured                 ", len(np.asarray(measured_rays[1]).reshape(-1)), "rays, power",
          float(np.sum(np.asarray(measured_rays[1], dtype=np.float64))), "of", tr.input_power)
    print("On average performed     ", proc_ray_count * np.int64(tr.tri_count) / np.float64(sim_time), "RI/s")
    if len(np.asarray(measured_rays[1]).reshape(-1)):
        lim = ((-np.pi / 2.0, np.pi / 2.0), (-np.pi / 2.0, np.pi / 2.0))
        H, xe, ye = tr.get_binned_data_angular(limits=lim, points=a.bins)
        print("Angular map %dx%d: total binned intensity %.6g, peak bin %.6g" % (a.bins, a.bins, H.sum(), H.max()))


if __name__ == "__main__":
    main()
