"""Projection of measured end points onto the plane + weighted 2-D binning
(post-processing).

``project_bin_device`` is the product path: the measured-ray buffer never leaves the GPU; the
``lpcl_project`` and ``lpcl_bin2d`` kernels (csrc/lpcl_project.cuh) replace the reference's
``stereograph_project`` / ``angular_project`` launches (it:515,546,602) and ``np.histogram2d``
(it:416,529,560); only the (points x points) float64 histogram is downloaded.

``project_numpy`` / ``bin_numpy`` restate the same two steps in NumPy (float32 in the kernels'
operation order, double where OpenCL C promotes).  They are the checker for the GPU kernels
(tests/test_gpu_projection.py) and are not called by the tracer.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

_F = np.float32
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


# ------------------------------------------------------------------------------------------ NumPy checker
def _rotate(pos, rot):
    p = np.asarray(pos, dtype=_F)[:, :3]
    r = np.eye(3, dtype=_F) if rot is None else np.asarray(rot, dtype=_F)[:3, :3]
    # dot(vec, rot_mtx[k]) = (x*r0 + y*r1) + z*r2, each product and sum rounded to float32
    u, v, w = [(p[:, 0] * r[k, 0] + p[:, 1] * r[k, 1]) + p[:, 2] * r[k, 2] for k in range(3)]
    l = np.sqrt((u * u + v * v) + w * w, dtype=_F)
    return u, v, w, l


def project_numpy(kind, pos, pwr, rot=None):
    """-> (x, y, corrected power), float32, like the kernels' x / y / pwrs_cor outputs."""
    pwr = np.asarray(pwr, dtype=_F).reshape(-1)
    if kind == "xy":
        p = np.asarray(pos, dtype=_F)
        return p[:, 0].copy(), p[:, 1].copy(), pwr
    u, v, w, l = _rotate(pos, rot)
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == "stereographic":  # cl:499-505
            xt = u / (l + w)
            yt = v / (l + w)
            # A = 4.0 / pown(1.0 + pown(xt,2) + pown(yt,2), 2): pown(float,2) is float; the sum with the
            # literal 1.0 is double, so pown(double,2) and the division are double; A is stored as float
            s = (1.0 + (xt * xt).astype(np.float64)) + (yt * yt).astype(np.float64)
            a = (4.0 / (s * s)).astype(_F)
            return xt.astype(_F), yt.astype(_F), (pwr / a).astype(_F)
        if kind == "angular":  # cl:519-537
            cos_t = w / l
            phi = np.arctan2(v, u, dtype=_F)
            r = np.arccos(cos_t, dtype=_F)
            xt = r * np.cos(phi, dtype=_F)
            yt = r * np.sin(phi, dtype=_F)
            a = np.ones_like(r)
            big = r > _F(0.000001)
            a[big] = np.sin(r[big], dtype=_F) / r[big]
            return xt.astype(_F), yt.astype(_F), (pwr / a).astype(_F)
    raise ValueError("unknown projection %r" % (kind,))


def bin_numpy(x, y, w, limits, points, kind):
    """np.histogram2d exactly as the reference calls it (it:416,529,560)."""
    if kind == "xy":
        weights = np.asarray(w)
    else:
        dx = np.float64(limits[0][1] - limits[0][0]) / np.float64(points)
        dy = np.float64(limits[1][1] - limits[1][0]) / np.float64(points)
        weights = np.float64(w) / (dx * dy)
    return np.histogram2d(x=np.asarray(x).flatten(), y=np.asarray(y).flatten(), bins=points, range=limits,
                          weights=np.asarray(weights).flatten())
