"""Projection of measured end points onto the plane (post-processing).

Restates the reference's ``stereograph_project`` (cl:488-506) and ``angular_project``
(cl:509-538) kernels in float32 NumPy: O(K) work once per run on the already-measured rays,
outside the trace step.  ``rot`` is the (3..4, 3..4) rotation the reference uploads as R_dev
(identity at it:514,545; Rz(k*2pi/sources) in replicate_lightsources_and_plot it:598-602);
the pivot is the origin as at it:511,542.
"""

from __future__ import annotations

import numpy as np

_F = np.float32


def _rotate(pos, rot):
    p = np.asarray(pos, dtype=_F)[:, :3]
    r = np.eye(3, dtype=_F) if rot is None else np.asarray(rot, dtype=_F)[:3, :3]
    # dot(vec, rot_mtx[k]) = (x*r0 + y*r1) + z*r2, each product and sum rounded to float32
    cols = []
    for k in range(3):
        cols.append((p[:, 0] * r[k, 0] + p[:, 1] * r[k, 1]) + p[:, 2] * r[k, 2])
    u, v, w = cols
    l = np.sqrt((u * u + v * v) + w * w, dtype=_F)
    return u, v, w, l


def project(kind, pos, pwr, rot=None):
    """-> (x, y, corrected power), all float32, like the kernels' x / y / pwrs_cor outputs."""
    pwr = np.asarray(pwr, dtype=_F).reshape(-1)
    u, v, w, l = _rotate(pos, rot)
    with np.errstate(divide="ignore", invalid="ignore"):
        if kind == "stereographic":  # cl:499-505
            xt = u / (l + w)
            yt = v / (l + w)
            # A = 4.0 / pown(1.0 + xt^2 + yt^2, 2): the sum is formed in double (literal 1.0) and
            # passed to pown(float, int), i.e. rounded to float before squaring
            s = (1.0 + (xt * xt).astype(np.float64) + (yt * yt).astype(np.float64)).astype(_F)
            a = (4.0 / (s * s).astype(np.float64)).astype(_F)
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
