"""Projection of measured end points onto the plane + weighted 2-D binning
(post-processing).

``project_bin_device`` is the product path: the measured-ray buffer never leaves the GPU; the
``lpcl_project`` and ``lpcl_bin2d`` kernels (csrc/lpcl_project.cuh) replace the reference's
``stereograph_project`` / ``angular_project`` launches (it:515,546,602) and ``np.histogram2d``
(it:416,529,560); only the (points x points) float64 histogram is downloaded.

The checker for these kernels lives outside the package, in oracle/oracle.py (``project``,
``bin2d``: the reference's own projection kernels compiled from the reference source).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

KINDS = {"xy": 0, "stereographic": 1, "angular": 2}


# ------------------------------------------------------------------------------------------ device path
def project_bin_device(lib, records, kind, limits, points, rot=None, stream=None, weight_denominator=None):
    """records: (K,4) float32 CUDA tensor (x, y, z, power).  Returns (H (points,points) float64
    ndarray, x_edges, y_edges) exactly shaped like np.histogram2d's result."""
    import torch

    from . import _cabi

    dev = records.device
    k = int(records.shape[0])
    xe = np.linspace(limits[0][0], limits[0][1], points + 1)
    ye = np.linspace(limits[1][0], limits[1][1], points + 1)
    if weight_denominator is None:
        dx = np.float64(limits[0][1] - limits[0][0]) / np.float64(points)
        dy = np.float64(limits[1][1] - limits[1][0]) / np.float64(points)
        weight_denominator = 1.0 if kind == "xy" else float(dx * dy)  # it:525-527, 556-558
    hist = torch.empty((points, points), dtype=torch.float64, device=dev)
    x = torch.empty(max(k, 1), dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    w = torch.empty_like(x)
    s = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream)
    if k:
        rot9 = None
        if rot is not None:
            rot9 = np.ascontiguousarray(np.asarray(rot, dtype=np.float32)[:3, :3])
        rc = lib.lpcl_project(records.data_ptr(), k, rot9.ctypes.data if rot9 is not None else None, KINDS[kind],
                              x.data_ptr(), y.data_ptr(), w.data_ptr(), s)
        _cabi.check(rc, "lpcl_project")
    xe_d, ye_d = torch.from_numpy(xe).to(dev), torch.from_numpy(ye).to(dev)
    rc = lib.lpcl_bin2d(x.data_ptr(), y.data_ptr(), w.data_ptr(), k, xe_d.data_ptr(), ye_d.data_ptr(), int(points),
                        C.c_double(weight_denominator), hist.data_ptr(), s)
    _cabi.check(rc, "lpcl_bin2d")
    return hist.cpu().numpy(), xe, ye
