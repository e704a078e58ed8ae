"""ctypes binding of liblightpycl_b200.so (the C ABI declared in include/lightpycl_b200.h).

This is the only place the package touches native code.  There is NO fallback: if the
library is missing or fails to load, importing the tracer raises.  Device memory is owned by
torch tensors on the Python side; only raw ``data_ptr()`` values, sizes and scalars cross the
boundary (no torch types in any signature).
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LPCL_LIB points at an alternative build of the same ABI (kernel experiments); default: in-tree
LIB_PATH = os.environ.get("LPCL_LIB") or os.path.join(_HERE, "_lib", "liblightpycl_b200.so")

ABI_VERSION = 1


class StepCounters(C.Structure):
    """Mirror of ``lpcl_step_counters``."""

    _fields_ = [
        ("n_reflect", C.c_int64),
        ("n_refract", C.c_int64),
        ("n_measured", C.c_int64),
        ("n_terminated", C.c_int64),
        ("live_power", C.c_double),
        ("measured_power", C.c_double),
        ("candidates", C.c_int64),
        ("reserved", C.c_int64),
    ]


class Generation(C.Structure):
    """Mirror of ``lpcl_generation``."""

    _fields_ = [
        ("d_org_pow", C.c_void_p),
        ("d_dir_prev", C.c_void_p),
        ("d_lineage", C.c_void_p),
        ("n", C.c_int64),
        ("seg_split", C.c_int64),
        ("seg_base", C.c_int64),
        ("capacity", C.c_int64),
    ]


class SmallIter(C.Structure):
    """Mirror of ``lpcl_small_iter``."""

    _fields_ = [("n_rays", C.c_int64), ("n_reflect", C.c_int64), ("n_refract", C.c_int64), ("n_measured", C.c_int64),
                ("n_terminated", C.c_int64), ("offset", C.c_int64), ("live_power", C.c_double),
                ("measured_power", C.c_double)]


class SmallHeader(C.Structure):
    """Mirror of ``lpcl_small_header``."""

    _fields_ = [("iterations", C.c_int64), ("stop_reason", C.c_int64), ("next_offset", C.c_int64), ("next_n", C.c_int64),
                ("next_split", C.c_int64), ("next_base", C.c_int64), ("measured_total", C.c_int64), ("reserved", C.c_int64)]


# every symbol include/lightpycl_b200.h declares: (restype, argtypes)
_p, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "lpcl_abi_version": (C.c_int, []),
    "lpcl_last_error": (C.c_char_p, []),
    "lpcl_scene_create": (C.c_int, [C.POINTER(_p), _p, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _p]),
    "lpcl_scene_bytes": (_i64, [_i32, _i32]),
    "lpcl_scene_create_in": (C.c_int, [C.POINTER(_p), _p, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _p, _i64, _p]),
    "lpcl_scene_set_tc_warps": (C.c_int, [_p, _i32]),
    "lpcl_search_launch_count": (_i32, [_p]),
    "lpcl_scene_destroy": (None, [_p]),
    "lpcl_scene_tri_count": (_i32, [_p]),
    "lpcl_scene_mesh_count": (_i32, [_p]),
    "lpcl_scene_device_bytes": (_i64, [_p]),
    "lpcl_scratch_bytes": (_i64, [_p, _i64]),
    "lpcl_intersect": (C.c_int, [_p, _p, _p, _i64, _f, _p, _p, _p, _p, _p]),
    "lpcl_intersect_postproc": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _f, _p]),
    "lpcl_reflect_refract_rays": (C.c_int, [_p] * 18 + [_i64, _f, _p]),
    "lpcl_step_workspace_bytes": (_i64, [_p, _i64]),
    "lpcl_trace_step": (C.c_int, [_p, C.POINTER(Generation), C.POINTER(Generation), _p, _p, _p, _p, _i64, _i64, _p,
                                  _p, _f, _f, _i32, _p]),
    "lpcl_trace_resolve": (C.c_int, [_p, C.POINTER(Generation), C.POINTER(Generation), _p, _p, _p, _p, _i64, _i64, _p,
                                     _p, _f, _f, _i32, _p]),
    "lpcl_trace_search": (C.c_int, [_p, C.POINTER(Generation), _p, _f, _p, _p]),
    "lpcl_pack_rays": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "lpcl_small_scratch_bytes": (_i64, [_p, _i64]),
    "lpcl_trace_small": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _p, _i64, _p, _p, _i64, _f, _f, _i32, C.c_double, _i64,
                                   _p, _p, _p]),
    "lpcl_project": (C.c_int, [_p, _i64, _p, _i32, _p, _p, _p, _p]),
    "lpcl_bin2d": (C.c_int, [_p, _p, _p, _i64, _p, _p, _i32, C.c_double, _p, _p]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "lightpycl_b200: CUDA extension not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lightpycl_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    got = lib.lpcl_abi_version()
    if got != ABI_VERSION:
        raise NativeError("lightpycl_b200: ABI version mismatch (library %d, binding %d)" % (got, ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().lpcl_last_error()
        raise NativeError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))
