"""Drop-in shim: put this directory on PYTHONPATH and a LightPyCL script's `import geo_optical_elements` resolves to the
B200 implementation (lightpycl_b200.geo_optical_elements)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from lightpycl_b200.geo_optical_elements import *  # noqa: F401,F403,E402
from lightpycl_b200 import geo_optical_elements as _impl  # noqa: E402

for _k in dir(_impl):  # also re-export names without __all__ filtering (e.g. CL_Tracer, GeoObject)
    if not _k.startswith("__"):
        globals().setdefault(_k, getattr(_impl, _k))
