"""Scene description: triangle-mesh container and optical-element generators.

Python-3 / NumPy-2 host-side mirror of the reference's ``geo_optical_elements.py``
(reference geo:25-504).  Same class names, method names, keyword arguments and -
because parity is checked "on the same inputs" - the same float values: every
generator evaluates the same float64 expressions as the reference and rounds to
float32 at the same point.  The code is written fresh and vectorised (the
reference builds meshes with per-vertex ``np.matrix`` products in Python loops,
which takes minutes for a 1M-triangle scene); it feeds the tracer through the
same two calls the reference's host loop makes: ``tribuf()`` and
``getMaterialBuf()`` (reference it:130,137).

Nothing here runs on the GPU; the mesh is uploaded once per trace.
"""

from __future__ import annotations

import numpy as np

__all__ = ["GeoObject", "optical_elements"]


def _rot4(axis, ang):
    """4x4 rotation used by ``GeoObject.rotate`` (reference geo:73-75).

    The last row/column are zero, as in the reference, so the w lane of every
    rotated vertex is 0.
    """
    c, s = np.cos(ang), np.sin(ang)
    if axis in ("y", "Y"):
        return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 0]], dtype=np.float64)
    if axis in ("z", "Z"):
        return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 0]], dtype=np.float64)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 0]], dtype=np.float64)


class GeoObject:
    """A triangle mesh with one material (reference geo:25-148).

    ``vertices``  - sequence/array of 4-vectors (x, y, z, 0)
    ``triangles`` - sequence/array of 3 vertex indices
    Material codes (``matTypes``) are the ones the kernels switch on
    (reference cl:364): refractive 0, mirror 1, terminator 2, measure 3,
    refractive_anisotropic 4.
    """

    vertices = None
    triangles = None
    IOR = 1.0
    reflectivity = 1.0
    dissipation = 0.0  # 1/length-unit
    AR_IOR = 1.0  # carried, unused by the tracer (as in the reference)
    AR_thickness = 0.0  # carried, unused
    anisotropy = None  # carried, unused

    matTypes = {"refractive": 0, "mirror": 1, "terminator": 2, "measure": 3, "refractive_anisotropic": 4}
    matType = "refractive"

    def __init__(self, verts, tris, mat_type="refractive", IOR=1.0, reflectivity=1.0, dissipation=0.0,
                 AR_IOR=1.0, AR_thickness=0.0, anisotropy=None):
        self.vertices = verts
        self.triangles = tris
        self.setMaterial(mat_type, IOR, reflectivity, dissipation, AR_IOR, AR_thickness, anisotropy)

    def setMaterial(self, mat_type="refractive", IOR=1.0, reflectivity=1.0, dissipation=0.0, AR_IOR=1.0,
                    AR_thickness=0.0, anisotropy=None):
        """Set the material (reference geo:46-62).

        Reference behaviour that is part of the spec: only the
        ``"refractive"`` branch stores parameters.  For ``"mirror"`` neither
        ``reflectivity`` nor ``IOR`` is stored - the reference tests
        ``"refractive"`` twice (geo:54,59) - so a mirror traces with whatever
        ``self.reflectivity`` already holds (class default 1.0).  Set
        ``mesh.reflectivity = 0.98`` by attribute to get a lossy mirror.
        An unknown material name raises ``NameError`` like the reference does
        (geo:52 formats an undefined name).
        """
        if mat_type in self.matTypes:
            self.matType = mat_type
        else:
            raise NameError("material %r unknown (reference geo:52 fails with NameError here)" % (mat_type,))

        if self.matType == "refractive":
            self.IOR = IOR
            self.dissipation = dissipation
            self.AR_IOR = AR_IOR
            self.AR_thickness = AR_thickness
        elif self.matType == "refractive_anisotropic":
            print("Warning: anisotropic materials are not yet supported.")

    def getMaterialBuf(self):
        """Material parameters in the form the tracer consumes (reference geo:64-66)."""
        return {"type": self.matTypes.get(self.matType), "IOR": self.IOR, "R": self.reflectivity,
                "dissipation": self.dissipation}

    # ------------------------------------------------------------------ geometry ops
    def translate(self, vec):
        """reference geo:68-69: array addition; the dtype promotes with ``vec`` (float64 for
        Python numbers), the float32 cast happens when the tracer flattens the scene."""
        self.vertices = np.asarray(self.vertices) + np.array(vec)

    def rotate(self, axis="x", angle=np.pi / 2, pivot=(0, 0, 0, 0)):
        """reference geo:72-93: R(angle) @ (V - pivot)^T in float64, each vertex rewritten as
        float32 (x, y, z, 0)."""
        piv = np.array(pivot)
        v = np.asarray(self.vertices)
        vt = np.dot(_rot4(axis, angle), np.transpose(v - piv))
        out = (np.transpose(vt) + piv).astype(np.float32)
        out[:, 3] = 0
        if isinstance(self.vertices, np.ndarray):
            self.vertices[...] = out  # keeps the container dtype, values are float32-exact
        else:
            self.vertices = out

    def trimesh(self):
        """List of triangles, each a list of its three vertices (reference geo:95-103)."""
        v = np.asarray(self.vertices)
        return [[v[i] for i in tri] for tri in np.asarray(self.triangles, dtype=np.int64)]

    def trimesh_to_geoObject(self, trimesh):
        """Inverse of ``trimesh`` with duplicated vertices (reference geo:105-119)."""
        verts, tris, k = [], [], 0
        for tri in trimesh:
            verts.extend(tri)
            tris.append([k, k + 1, k + 2])
            k += 3
        self.vertices = verts
        self.triangles = tris

    def tribuf(self):
        """Three per-triangle vertex buffers ``(m_v0, m_v1, m_v2)``, one row per triangle
        (reference geo:121-132).  Returned as arrays (fancy-indexed) instead of Python lists;
        the tracer casts them to float32 exactly where the reference does (it:141-149)."""
        v = np.asarray(self.vertices)
        t = np.asarray(self.triangles).astype(np.int64)
        if t.size == 0:
            e = np.zeros((0, 4), dtype=v.dtype if v.size else np.float32)
            return (e, e.copy(), e.copy())
        return (v[t[:, 0]], v[t[:, 1]], v[t[:, 2]])

    def append(self, verts, tris):
        """Merge another mesh into this one, re-basing its indices (reference geo:134-137)."""
        nv = len(self.vertices)
        self.triangles = np.append(np.asarray(self.triangles), np.array(tris).astype(np.int32) + nv, axis=0)
        self.vertices = np.append(np.asarray(self.vertices), verts, axis=0)

    def write_dxf(self, dxf_file):
        """reference geo:139-148 (needs the third-party ``dxfwrite``).  Here: a dependency-free
        ASCII DXF with one 3DFACE per triangle on layer "0"."""
        from ._dxf import write_faces

        v0, v1, v2 = self.tribuf()
        write_faces(dxf_file, [("0", 2, np.asarray(v0)[:, :3], np.asarray(v1)[:, :3], np.asarray(v2)[:, :3])])


class optical_elements:
    """Mesh generators (reference geo:150-504).  Each returns a ``GeoObject``."""

    # ------------------------------------------------------------------ primitives
    def cube(self, center, size):
        """Axis-aligned box, 12 triangles (reference geo:153-170).  float32 unit corners /2.0
        * size + center: with Python sequences the result is float64, as in the reference."""
        corners = np.array([[-1, -1, -1, 0], [1, -1, -1, 0], [-1, 1, -1, 0], [1, 1, -1, 0],
                            [-1, -1, 1, 0], [1, -1, 1, 0], [-1, 1, 1, 0], [1, 1, 1, 0]], dtype=np.float32)
        verts = corners / 2.0 * size + center
        tris = [[0, 1, 2], [2, 3, 1],  # bottom
                [4, 5, 6], [6, 7, 5],  # top
                [0, 1, 4], [4, 5, 1],  # rear
                [2, 3, 6], [6, 7, 3],  # front
                [0, 2, 4], [4, 6, 2],  # left
                [1, 3, 5], [5, 7, 3]]  # right
        return GeoObject(verts, tris)

    def revolve_curve(self, curve, axis="x", ang=2 * np.pi, ang_pts=36):
        """Surface of revolution of a 2-D polyline (reference geo:259-302).

        ``ang_pts`` wedges start at ``linspace(0, ang - d_ang, ang_pts)`` and span
        ``d_ang = ang / (ang_pts - 1)`` - i.e. neighbouring wedges overlap slightly and the
        sweep ends at ``ang`` (same generator => same behaviour).  Per wedge k and
        segment i: four float32 vertices R(phi1)c[i], R(phi1)c[i+1], R(phi2)c[i], R(phi2)c[i+1]
        and triangles [0,1,2], [2,3,1] offset by 4*(i + k*nseg).  Output: 2*nseg*ang_pts
        triangles, vertex/triangle order identical to the reference loops (k outer, i inner).
        """
        c = np.asarray(curve, dtype=np.float64)
        nseg = len(c) - 1
        d_ang = ang / (ang_pts - 1.0)
        phi1 = np.linspace(0.0, ang - d_ang, ang_pts)
        phi2 = phi1 + d_ang

        def rot(phi, pts):
            # R(phi) @ [x, y, 0]^T for every (phi, point) pair -> (len(phi), len(pts), 3).  The reference forms the
            # full 3 x 3 product (geo:289-292): a sum that starts from +0.0, so a component that comes out as -0.0
            # here (cos(phi) * 0 with cos < 0) is +0.0 there - the final `+ 0.0` maps -0.0 to +0.0, nothing else
            cs, sn = np.cos(phi)[:, None], np.sin(phi)[:, None]
            x, y = pts[None, :, 0], pts[None, :, 1]
            if axis in ("y", "Y"):
                return np.stack([cs * x, y + 0 * cs, -sn * x], axis=-1) + 0.0
            if axis in ("z", "Z"):
                return np.stack([cs * x - sn * y, sn * x + cs * y, 0 * (cs * x)], axis=-1) + 0.0
            return np.stack([x + 0 * cs, cs * y, sn * y], axis=-1) + 0.0

        a1, a2 = rot(phi1, c), rot(phi2, c)  # (K, P, 3)
        quad = np.stack([a1[:, :-1], a1[:, 1:], a2[:, :-1], a2[:, 1:]], axis=2)  # (K, nseg, 4, 3)
        verts = np.zeros((ang_pts * nseg * 4, 4), dtype=np.float32)
        verts[:, :3] = quad.reshape(-1, 3)  # float64 -> float32 here, like np.array(..., float32)
        base = 4 * np.arange(ang_pts * nseg, dtype=np.int64)[:, None]
        tris = np.empty((ang_pts * nseg, 2, 3), dtype=np.int64)
        tris[:, 0, :] = base + np.array([0, 1, 2])
        tris[:, 1, :] = base + np.array([2, 3, 1])
        return GeoObject(verts, tris.reshape(-1, 3))

    def extrude_by_vector(self, curve, vector, capped=True):
        """Extrude a closed 2-D polyline along ``vector`` (reference geo:304-338).  Capping
        needs ``curve_to_mesh`` and therefore the third-party ``triangle`` package."""
        c = np.asarray(curve, dtype=np.float64)
        n = len(c)
        verts = np.zeros((2 * n, 4), dtype=np.float32)
        verts[0::2, 0], verts[0::2, 1] = c[:, 0], c[:, 1]
        verts[1::2, 0] = c[:, 0] + vector[0]
        verts[1::2, 1] = c[:, 1] + vector[1]
        verts[1::2, 2] = 0.0 + vector[2]
        k = np.arange(n - 1, dtype=np.int64)[:, None] * 2
        tris = np.empty((n - 1, 2, 3), dtype=np.int64)
        tris[:, 0, :] = k + np.array([0, 1, 2])
        tris[:, 1, :] = k + np.array([2, 3, 1])
        gobj = GeoObject(verts, tris.reshape(-1, 3))
        if capped:
            cap = self.curve_to_mesh(curve)
            gobj.append(cap.vertices, cap.triangles)
            gobj.append(cap.vertices + np.array(vector), cap.triangles)
        return gobj

    def curve_to_mesh(self, curve):
        """Constrained Delaunay fill of a closed polyline (reference geo:340-364); needs the
        third-party ``triangle`` package exactly like the reference."""
        try:
            import triangle
        except ImportError as e:  # pragma: no cover - optional dependency absent in this image
            raise ImportError("curve_to_mesh needs the 'triangle' package (as the reference does)") from e
        m = len(curve)
        segs = np.zeros((m, 2), dtype=np.int32)
        segs[:, 0] = np.arange(m)
        segs[:, 1] = segs[:, 0] + 1
        segs[-1, -1] = 0
        tri = triangle.triangulate({"vertices": np.array(curve).astype(np.float32), "segments": segs}, "pq10")
        v2, t2 = tri["vertices"], tri["triangles"]
        verts = np.zeros((len(v2), 4), dtype=np.float32)
        verts[:, :2] = v2[:, :2]
        tris = np.zeros((len(t2), 3), dtype=np.float32)
        tris[:, :] = t2[:, :3]
        return GeoObject(verts=verts, tris=tris)

    # ------------------------------------------------------------------ revolved elements
    @staticmethod
    def _xy(x, y):
        return [[a, b] for (a, b) in zip(x, y)]

    def sphere(self, center, radius):
        """reference geo:202-214 (full circle profile revolved by pi)."""
        n = 72
        phi = np.linspace(0.0, 2.0 * np.pi, n)
        mesh = self.revolve_curve(self._xy(np.cos(phi) * radius, np.sin(phi) * radius), axis="x", ang=np.pi,
                                  ang_pts=n + 1)
        mesh.translate(center)
        return mesh

    def hemisphere(self, center, radius):
        """reference geo:216-228 (half circle revolved by pi => z >= 0 dome, 10,366 triangles)."""
        n = 72
        phi = np.linspace(0.0, np.pi, n)
        mesh = self.revolve_curve(self._xy(np.cos(phi) * radius, np.sin(phi) * radius), axis="x", ang=np.pi,
                                  ang_pts=n + 1)
        mesh.translate(center)
        return mesh

    def parabolic_mirror(self, focus=(0, 0, 0), focal_length=5.0, diameter=20.0, reflectivity=0.98):
        """reference geo:230-245; ``reflectivity`` is accepted and - like the reference - NOT
        stored."""
        n, m = 72, 200
        yn = np.linspace(0.0, diameter / 2.0, m)
        xn = yn ** 2 / (4.0 * focal_length) - focal_length
        mesh = self.revolve_curve(self._xy(focus[0] + xn, focus[1] + yn), axis="x", ang=2.0 * np.pi,
                                  ang_pts=n + 1)
        mesh.setMaterial(mat_type="mirror", reflectivity=reflectivity)
        return mesh

    def topless_cylinder(self, center=(0, 0, 0), diameter=20.0, height=10.0):
        """reference geo:247-257."""
        n = 72
        x = np.array([0.0, 0.0, 1.0]) * height + center[1]
        y = np.array([0.0, 0.5, 0.5]) * diameter + center[0]
        return self.revolve_curve(self._xy(x, y), axis="x", ang=2.0 * np.pi, ang_pts=n + 1)

    def spherical_lens_nofoc(self, r1, r2, x1, x2, d, d2=None, sign1_arcsin=1.0, sign2_arcsin=1.0):
        """Two spherical caps joined into a closed lens body (reference geo:172-200); used by
        the eye model."""
        n = 50
        if d2 is None:
            d2 = d
        z1, z2 = r1 + x1, r2 + x2
        dphi1 = np.pi / 2.0 if sign1_arcsin < 0 else 0.0
        dphi2 = np.pi / 2.0 if sign2_arcsin < 0 else 0.0
        phi1 = np.linspace(0.0, np.absolute(np.arcsin(d / r1)) + dphi1, n)
        phi2 = np.linspace(np.absolute(np.arcsin(d2 / r2)) + dphi2, 0.0, n)
        xc = np.append(z1 - r1 * np.cos(phi1), z2 - r2 * np.cos(phi2))
        yc = np.append(r1 * np.sin(phi1), r2 * np.sin(phi2))
        mesh = self.revolve_curve(self._xy(xc, yc), axis="x", ang=2.0 * np.pi, ang_pts=72)
        mesh.rotate(axis="y", angle=-np.pi / 2.0, pivot=(0, 0, 0, 0))
        return mesh

    # ------------------------------------------------------------------ lens profiles
    def lens_spherical_2r(self, focus, r1, r2, diameter, lens_sign, n):
        """2-D profile of a two-radius spherical lens (reference geo:408-443); 2*60+1 points,
        closed (last == first)."""
        N = 60
        fx0, fy0 = focus[0], focus[1]
        ab = np.absolute
        phi_r1 = np.arcsin(diameter / r1)
        phi_r2 = np.arcsin(diameter / r2)
        d = ab(r1 - r1 * np.cos(phi_r1)) + ab(r2 - r2 * np.cos(phi_r2))
        f = ab(1 / ((n - 1) * (1 / r1 - 1 / r2 + (n - 1) * d / (n * r1 * r2))))
        q = f - r1
        r1x0, r1y0 = fx0 + q, fy0
        r2x0, r2y0 = fx0 + q + r1 - lens_sign * d + r2, fy0
        x = np.zeros(2 * N + 1)
        y = np.zeros(2 * N + 1)
        p2 = np.linspace(np.pi - ab(phi_r2), np.pi + ab(phi_r2), N)
        p1 = np.linspace(-ab(phi_r1), ab(phi_r1), N)
        for i in range(N):  # scalar cos/sin per point, as the reference evaluates them
            x[i] = r2x0 + r2 * np.cos(p2[i])
            y[i] = r2y0 + r2 * np.sin(p2[i])
            x[N + i] = r1x0 + r1 * np.cos(p1[i])
            y[N + i] = r1y0 + r1 * np.sin(p1[i])
        x[2 * N], y[2 * N] = x[0], y[0]
        return self._xy(x, y)

    def lens_spherical_biconcave(self, focus, r1, r2, diameter, IOR):
        """reference geo:366-370: the 2r profile revolved by pi with 36 wedges => 8,640 triangles."""
        mesh = self.revolve_curve(self.lens_spherical_2r(focus, r1, r2, diameter, 1, IOR), axis="x", ang=np.pi,
                                  ang_pts=36)
        mesh.setMaterial(mat_type="refractive", IOR=IOR)
        return mesh

    def curve_lens_spherical(self, focus, r1, diameter, lens_direction, lens_sign, IOR):
        """Single-surface lens profile (reference geo:445-504) -> (curve, f, d)."""
        N = 20
        n = IOR
        pi = np.pi
        ab = np.absolute
        x = np.zeros(N + 3)
        y = np.zeros(N + 3)
        phi_r1 = np.arcsin(diameter / (2 * r1))
        d = ab(r1 - r1 * np.cos(phi_r1))
        f = ab(1 / ((n - 1) * (1 / r1)))
        q = f + r1
        fx0, fy0 = focus[0], focus[1]
        ang_offset = 0 if lens_sign > 0 else -pi
        if lens_direction >= 0:
            r1x0, r1y0 = fx0 - lens_sign * q, fy0
            angles = np.linspace(ang_offset - lens_sign * ab(phi_r1), ang_offset + lens_sign * ab(phi_r1), N)
        else:
            r1x0, r1y0 = fx0 + lens_sign * q, fy0
            angles = np.linspace(ang_offset + pi - lens_sign * ab(phi_r1), ang_offset + pi + lens_sign * ab(phi_r1), N)
        i = 0
        for phi in angles:
            x[i] = r1x0 + r1 * np.cos(phi)
            y[i] = r1y0 + r1 * np.sin(phi)
            i += 1
        if lens_sign < 0:
            x[i], y[i] = x[i - 1] + lens_direction * d, y[i - 1]
            i += 1
            x[i], y[i] = x[i - 2] + lens_direction * d, y[0]
            i += 1
        x[i], y[i] = x[0], y[0]
        return (self._xy(x, y), f, d)

    def curve_lens_spherical_biconcave(self, focus, r1, r2, d, diameter, axis, IOR):
        """Two single-surface profiles joined (reference geo:373-406) -> (curve, f0, d0)."""
        ab = np.absolute
        n = IOR
        f1 = ab(1 / ((n - 1) * (1 / r1)))
        f2 = ab(1 / ((n - 1) * (1 / r2)))
        f0 = 1 / (1 / f1 + 1 / f2)
        fx0, fy0 = focus[0], focus[1]
        (poly1, f1, _d1) = self.curve_lens_spherical((fx0 + f1 + f0, fy0), r1, diameter, -1, -1, IOR)
        (poly2, f2, _d2) = self.curve_lens_spherical((fx0 + d - f1 + f0, fy0), r2, diameter, 1, -1, IOR)
        curve = poly1[0:-3]
        curve.extend(poly2[0:-3])
        curve.extend([poly1[1]])
        return (curve, f0, d)
