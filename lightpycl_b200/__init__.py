"""lightpycl_b200 - B200-native trace step behind LightPyCL's Python API.

Modules mirror the reference's three flat modules (same names, classes, kwargs):

    from lightpycl_b200 import iterative_tracer as it, light_source, geo_optical_elements as goe

``iterative_tracer.CL_Tracer`` needs the CUDA extension (lightpycl_b200/_lib/liblightpycl_b200.so)
and a CUDA device; there is no CPU fallback.  ``light_source`` and ``geo_optical_elements`` are
host-side NumPy and import anywhere.
"""

__version__ = "0.1.0"
