"""Dependency-free ASCII DXF writer (3DFACE entities only).

Stands in for the third-party ``dxfwrite`` package the reference uses for its
visual-debugging exports (reference geo:139-148, ls:191-204, it:754-779).  Host
side, not on the hot path.
"""

from __future__ import annotations

import numpy as np


def write_faces(path, layers):
    """``layers``: iterable of (layer_name, color, p0 (n,3), p1 (n,3), p2 (n,3)).

    Each triangle becomes a 3DFACE whose 4th corner repeats the 3rd (DXF convention for
    triangles; the reference passes three points to ``dxf.face3d``).
    """
    out = ["0", "SECTION", "2", "TABLES", "0", "TABLE", "2", "LAYER"]
    layers = list(layers)
    for name, color, *_ in layers:
        out += ["0", "LAYER", "2", str(name), "70", "0", "62", str(int(color)), "6", "CONTINUOUS"]
    out += ["0", "ENDTAB", "0", "ENDSEC", "0", "SECTION", "2", "ENTITIES"]
    for name, _color, p0, p1, p2 in layers:
        p0, p1, p2 = (np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in (p0, p1, p2))
        for a, b, c in zip(p0, p1, p2):
            out += ["0", "3DFACE", "8", str(name)]
            for k, p in enumerate((a, b, c, c)):
                out += [str(10 + k), repr(float(p[0])), str(20 + k), repr(float(p[1])), str(30 + k), repr(float(p[2]))]
    out += ["0", "ENDSEC", "0", "EOF"]
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
