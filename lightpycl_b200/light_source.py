"""Ray emitters.

Python-3 / NumPy-2 host-side mirror of the reference's ``light_source.py``
(reference ls:28-204): same class name, constructor keywords, attributes
(``ray_count, rays_origin, rays_dir, rays_power``) and - so that
``np.random.seed(s)`` reproduces the same rays - the same draws from NumPy's
legacy global RNG in the same order (reference ls:119-120, 153-154).

The tracer reads only the four attributes above (reference it:102-105).
"""

from __future__ import annotations

import numpy as np
import numpy.linalg as la

__all__ = ["light_source"]


def _Rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 0]], dtype=np.float64)


def _Ry(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 0]], dtype=np.float64)


def _Rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 0]], dtype=np.float64)


def _pinned(a):
    """Opt-in (``LPCL_PIN_SOURCES=1``): the same float32 values in page-locked host memory, so that
    the tracer's host->device copy of the rays is a direct DMA (pageable memory is staged at ~1/4 of
    the rate; 4.1 -> 0.9 ms for 1M rays).  By default this class is plain NumPy like the reference's
    and touches neither torch nor CUDA: no context is created before the tracer has picked its
    device, fork-based multiprocessing keeps working, nothing is page-locked behind the caller's back."""
    import os

    if a is None or os.environ.get("LPCL_PIN_SOURCES", "0") != "1":
        return a
    import torch

    if not torch.cuda.is_available():
        return a
    a = np.ascontiguousarray(a)
    if a.dtype != np.float32 or a.size == 0:
        return a
    return torch.from_numpy(a).pin_memory().numpy()


class light_source:
    """A point-like emitter with a directivity pattern (reference ls:28-43).

    ``center``      (4,) float32 position, w = 0
    ``direction``   main pointing direction (need not be unit length)
    ``directivity`` callable (azimuth, elevation) -> relative intensity
    ``power``       total power shared by all rays
    ``ray_count``   number of rays

    The constructor always calls ``random_rays()`` (reference ls:43), which consumes two
    ``rand(N, 1)`` draws even when the caller later replaces the rays with
    ``random_collimated_rays``.
    """

    center = None
    direction = None
    directivity = None
    power = 1
    ray_count = None
    rays_origin = None
    rays_dir = None

    def __init__(self, center=np.array([0, 0, 0, 0], dtype=np.float32), direction=(0, 0, 1),
                 directivity=lambda x, y: np.cos(y), power=1, ray_count=500):
        self.center = center
        self.direction = direction
        self.directivity = directivity
        self.power = power
        self.ray_count = ray_count
        self.random_rays()

    # -------------------------------------------------------------- pointing helpers
    def _elevation0(self):
        return np.arccos(self.direction[2] / la.norm(self.direction))

    def _azimuth0_guarded(self):
        # reference ls:107-113 (and ls:56-62): +-pi/2 when direction[0] == 0
        if self.direction[0] == 0:
            return np.pi / 2.0 if self.direction[1] >= 0 else -np.pi / 2.0
        return np.arctan2(self.direction[1], self.direction[0])

    # -------------------------------------------------------------- emitters
    def random_rays(self):
        """Cosine-weighted-in-elevation random rays over the +z hemisphere, rotated to
        ``direction`` (reference ls:98-137).

        Draw order: ``u = rand(N,1)`` then ``v = rand(N,1)``; elevation = arccos(u),
        azimuth = 2 pi v; directions are built in float64, multiplied as ROW vectors by
        Rx(elevation0) then Rz(azimuth0), and cast to float32 last.  ``rays_power`` has shape
        (N, 1): float32(directivity) * power / sum.
        """
        n = self.ray_count
        self.rays_origin = np.zeros((n, 4)).astype(np.float32) + self.center
        self.rays_dest = np.zeros((n, 4)).astype(np.float32)
        elevation0 = self._elevation0()
        azimuth0 = self._azimuth0_guarded()

        u = np.random.rand(n, 1)
        v = np.random.rand(n, 1)
        elevation = np.arccos(u)  # +z half space only
        azimuth = 2.0 * np.pi * v
        rd = np.concatenate((np.sin(elevation) * np.cos(azimuth), np.sin(elevation) * np.sin(azimuth),
                             np.cos(elevation), np.zeros((n, 1))), axis=1)

        self.rays_power = np.float32(self.directivity(azimuth, elevation))
        self.rays_power = self.rays_power * self.power / np.sum(self.rays_power)

        rd = np.dot(rd, _Rx(elevation0))
        self.rays_dir = np.dot(rd, _Rz(azimuth0)).astype(np.float32)
        self._pin_rays()

    def _pin_rays(self):
        self.rays_origin, self.rays_dir, self.rays_power = _pinned(self.rays_origin), _pinned(self.rays_dir), _pinned(self.rays_power)

    def random_collimated_rays(self, diameter=1.0):
        """Parallel rays from a random SQUARE footprint of side ``diameter`` (reference
        ls:139-170).  Draw order: x then y, each ``rand(N,1)``.  ``rays_power`` has shape (N,)."""
        n = self.ray_count
        self.rays_dest = np.zeros((n, 4)).astype(np.float32)
        elevation0 = self._elevation0()
        azimuth0 = np.arctan2(self.direction[1], self.direction[0])  # no direction[0]==0 guard here (ls:147)

        x = (np.random.rand(n, 1) - 0.5) * diameter
        y = (np.random.rand(n, 1) - 0.5) * diameter
        r_orig = np.concatenate((x, y, np.zeros((n, 1)), np.zeros((n, 1))), axis=1)

        rays_dir = np.zeros((n, 4)).astype(np.float32)
        rays_dir[:, 2] = 1.0

        self.rays_power = np.ones(n).astype(np.float32)
        self.rays_power = self.rays_power * self.power / np.sum(self.rays_power)

        ro = np.dot(r_orig, _Rx(elevation0))
        self.rays_origin = np.array(np.dot(ro, _Rz(azimuth0)).astype(np.float32) + self.center).astype(np.float32)

        rays_dir = np.dot(rays_dir, _Rx(elevation0))
        self.rays_dir = np.array(np.dot(rays_dir, _Rz(azimuth0))).astype(np.float32)
        self._pin_rays()

    def grid_rays(self):
        """Regular elevation x azimuth grid (reference ls:45-95).  The reference version no
        longer runs on NumPy >= 1.24 (``np.float``, float ``linspace`` counts and subscripts); this is the
        same construction with integer counts, bit-identical to it (tests/test_refhost_golden.py).
        ``ray_count`` is rounded down to a square AFTER ``rays_origin`` / ``rays_dest`` were sized (ls:46-49):
        like the reference's, they keep the original row count - the tracer reads the first ``ray_count``."""
        m = int(np.floor(np.sqrt(self.ray_count)))
        self.rays_origin = np.zeros((self.ray_count, 4)).astype(np.float32) + self.center
        self.rays_dest = np.zeros((self.ray_count, 4)).astype(np.float32)
        self.ray_count = m * m
        elevation0 = self._elevation0()
        if self.direction[0] == 0:
            azimuth0 = np.pi / 2.0 if self.direction[1] >= 0 else -np.pi / 2.0
        else:
            azimuth0 = np.arctan(self.direction[1] / self.direction[0])  # arctan, not arctan2 (ls:62)
        el = np.linspace(0.0, np.pi / 2.0, m)
        az = np.linspace(0.0, 2.0 * np.pi, m)
        elevation = np.tile(el, m)[:, None]  # elevation cycles fastest
        azimuth = np.repeat(az, m)
        azimuth[:m] = az  # ls:72-77: the first block takes the azimuth RAMP, blocks k >= 1 the constant az[k]
        azimuth = azimuth[:, None]
        self.rays_power = np.array(self.directivity(azimuth, elevation), dtype=np.float32) * self.power
        rd = np.concatenate((np.sin(elevation) * np.cos(azimuth), np.sin(elevation) * np.sin(azimuth),
                             np.cos(elevation), np.zeros((self.ray_count, 1))), axis=1).astype(np.float32)
        rd = np.dot(rd, _Rx(elevation0))
        self.rays_dir = np.dot(rd, _Rz(azimuth0))

    def rotate_rays(self, axis="z", pivot=[0, 0, 0, 0], ang=np.pi / 2.0):
        """Rotate origins about ``pivot`` and directions about 0, row-vector convention
        (reference ls:173-188)."""
        R = {"x": _Rx, "X": _Rx, "y": _Ry, "Y": _Ry, "z": _Rz, "Z": _Rz}.get(axis, _Rx)(ang)
        piv = np.array(pivot)
        self.rays_origin = np.array(np.add(np.dot(np.subtract(self.rays_origin, piv), R), piv)).astype(np.float32)
        self.rays_dir = np.array(np.dot(self.rays_dir, R)).astype(np.float32)
        self._pin_rays()

    def save_dxf(self, dxf_file):
        """reference ls:191-204 (needs ``dxfwrite``); here a dependency-free ASCII DXF with one
        degenerate 3DFACE per ray on layer "Rays"."""
        from ._dxf import write_faces

        o = np.asarray(self.rays_origin)[:, :3]
        e = (np.asarray(self.rays_origin) + np.asarray(self.rays_dir))[:, :3]
        write_faces(dxf_file, [("Rays", 3, o, e, e)])
