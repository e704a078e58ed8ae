"""TEST INFRASTRUCTURE - runs the reference's own host-side Python modules (light_source.py,
geo_optical_elements.py, iterative_tracer.py) in THIS container to generate golden vectors.

The reference is Python-2 source: it cannot be imported by the Python 3 here, and `lib2to3` is not in this image.  This
loader reads a module from /root/reference BY PATH, rewrites - in memory, nothing is stored - only what the two
languages parse differently, and executes it with Python-2 semantics where they differ:

  * `print a, b` statements -> `print(a, b)` (a line-level rewrite; the reference has no `print >>f` and no
    trailing-comma prints on the lines used here);
  * every `/` becomes a call of `_py2_div`: floor division when both operands are integers (Python ints or NumPy
    integer scalars/arrays), true division otherwise - what Python 2 did without `from __future__ import division`;
  * `pyopencl`, `matplotlib`, `pylab`, `dxfwrite`, `triangle`, `mpl_toolkits` are stubbed (none is in the image; none
    is touched by the functions the fixtures call: ray generators, mesh generators, beam metrics);
  * NumPy aliases removed since the reference was written (`np.float`, `np.int`) are restored inside the module's
    view of NumPy only, and `np.linspace` / `np.zeros` / `np.arange` accept the float counts and float subscripts the
    NumPy of the time truncated (light_source.grid_rays passes np.sqrt(ray_count) to all three).

Only tests/golden/make_golden_host.py uses it (the GPU box has no /root/reference: tests read the committed
fixtures).  Nothing in the product package imports this file."""

from __future__ import annotations

import ast
import contextlib
import io
import os
import re
import sys
import types

import numpy as np

REFERENCE_DIR = "/root/reference"
_PRINT = re.compile(r"^(\s*)print\s+(.*?)\s*$")
_PRINT_BARE = re.compile(r"^(\s*)print\s*$")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "iterative_tracer.py"))


def _py2_div(a, b):
    def is_int(x):
        if isinstance(x, (bool, np.bool_)):
            return True
        if isinstance(x, (int, np.integer)):
            return True
        return isinstance(x, np.ndarray) and np.issubdtype(x.dtype, np.integer)

    if is_int(a) and is_int(b):
        return a // b
    return a / b


class _Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id="_py2_div", ctx=ast.Load()), args=[node.left, node.right],
                                              keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            load = ast.parse(ast.unparse(node.target), mode="eval").body
            return ast.copy_location(ast.Assign(targets=[node.target], value=ast.Call(
                func=ast.Name(id="_py2_div", ctx=ast.Load()), args=[load, node.value], keywords=[])), node)
        return node


def _split_comment(code):
    """(code, comment) of one physical line: the first `#` outside a string literal starts the comment."""
    quote = None
    i = 0
    while i < len(code):
        c = code[i]
        if quote:
            if c == "\\":
                i += 1
            elif c == quote:
                quote = None
        elif c in "\"'":
            quote = c
        elif c == "#":
            return code[:i].rstrip(), "  " + code[i:]
        i += 1
    return code, ""


def _rewrite_prints(src):
    out = []
    for line in src.expandtabs(8).split("\n"):  # Python 2 counts a tab as the next multiple of eight columns
        code, comment = _split_comment(line)
        m = _PRINT_BARE.match(code)
        if m:
            out.append(m.group(1) + "print()" + comment)
            continue
        m = _PRINT.match(code)
        if m and not m.group(2).startswith(("(", "=")):
            out.append("%sprint(%s)%s" % (m.group(1), m.group(2), comment))
        else:
            out.append(line)
    return "\n".join(out)


class _Anything(types.ModuleType):
    """Stub module: any attribute is another stub, calling it raises (the fixtures never get that far)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        raise RuntimeError("reference called the stubbed module %s" % self.__name__)


_STUBS = ("pyopencl", "pyopencl.array", "pyopencl.tools", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
          "matplotlib.colors", "matplotlib.ticker", "pylab", "dxfwrite", "triangle", "mpl_toolkits",
          "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1")


class _NumpyView(types.ModuleType):
    """The module's view of numpy: today's numpy plus the aliases the reference still uses."""

    def __init__(self, real):
        super().__init__("numpy")
        self.__dict__["_real"] = real
        self.__dict__["float"] = float
        self.__dict__["int"] = int
        self.__dict__["bool"] = bool

        def linspace(start, stop, num=50, *a, **k):  # NumPy of the reference's time truncated a float `num`
            return real.linspace(start, stop, int(num), *a, **k)

        def _count(x):
            return int(x) if isinstance(x, (float, real.floating)) else x

        def zeros(shape, *a, **k):  # ... and float sizes
            return real.zeros(tuple(_count(x) for x in shape) if isinstance(shape, (tuple, list)) else _count(shape), *a, **k)

        def arange(*a, **k):
            # `for k in np.arange(np.sqrt(n)): ... azimuth_[k]` (ls:71-77): float loop indices were truncated when used
            # as subscripts; an integral float range is handed out as the integer range (k == 0 and the values agree)
            if a and all(isinstance(x, (float, real.floating)) and float(x).is_integer() for x in a):
                a = tuple(int(x) for x in a)
            return real.arange(*a, **k)

        self.__dict__["linspace"] = linspace
        self.__dict__["zeros"] = zeros
        self.__dict__["arange"] = arange

    def __getattr__(self, name):
        return getattr(self.__dict__["_real"], name)


_loaded = {}


def load(name):
    """Execute /root/reference/<name>.py under the shim and return it as a module object."""
    if name in _loaded:
        return _loaded[name]
    path = os.path.join(REFERENCE_DIR, name + ".py")
    src = _rewrite_prints(open(path).read())
    tree = _Div().visit(ast.parse(src, filename=path))
    ast.fix_missing_locations(tree)
    mod = types.ModuleType("reference_" + name)
    mod.__file__ = path
    mod.__dict__["_py2_div"] = _py2_div
    saved = {}
    for s in _STUBS:
        saved[s] = sys.modules.get(s)
        sys.modules[s] = _Anything(s)
    import pickle

    saved["cPickle"] = sys.modules.get("cPickle")
    sys.modules["cPickle"] = pickle  # Python 2's C pickle
    # the reference's modules import each other by their bare names
    for dep in ("light_source", "geo_optical_elements"):
        if dep != name and name == "iterative_tracer":
            saved[dep] = sys.modules.get(dep)
            sys.modules[dep] = load(dep)
    try:
        exec(compile(tree, path, "exec"), mod.__dict__)
    finally:
        for s, old in saved.items():
            if old is None:
                sys.modules.pop(s, None)
            else:
                sys.modules[s] = old
    if "np" in mod.__dict__:
        mod.__dict__["np"] = _NumpyView(np)
    _loaded[name] = mod
    return mod


@contextlib.contextmanager
def captured_stdout():
    buf, old = io.StringIO(), sys.stdout
    sys.stdout = buf
    try:
        yield buf
    finally:
        sys.stdout = old


# --------------------------------------------------------------------------------------------------------------------
# A stand-in for the few pyopencl entry points the reference's CL_Tracer touches, so that the reference's OWN host code
# (iterative_tracer.py: the trace loop it:77-393, get_measured_rays, the projection / replication functions) runs here
# on top of the reference's OWN kernels: oracle/_ref/libref_strict.so and libref_proj.so are the reference .cl compiled
# by path (oracle/Makefile).  "Device" arrays are host NumPy arrays; a kernel launch of global size (n,) is the
# launcher's loop over n work-items.  This is the closest thing to running the reference this image allows.
import ctypes as _C

_HERE = os.path.dirname(os.path.abspath(__file__))


class _DevArray:
    """pyopencl.array.Array stand-in (host memory)."""

    def __init__(self, a):
        self.a = np.array(np.asarray(a), order="C", copy=True)

    @property
    def data(self):
        return self

    @property
    def shape(self):
        return self.a.shape

    @property
    def dtype(self):
        return self.a.dtype

    def get(self, queue=None):
        return self.a.copy()

    def __add__(self, other):
        return _DevArray((self.a + other).astype(self.a.dtype))


class _ClArrayModule:
    @staticmethod
    def to_device(queue, ary):
        return _DevArray(ary)

    @staticmethod
    def zeros(queue, shape, dtype=np.float32):
        return _DevArray(np.zeros(shape, dtype=dtype))

    @staticmethod
    def empty(queue, shape, dtype=np.float32):
        return _DevArray(np.zeros(shape, dtype=dtype))


class _Event:
    def wait(self):
        return None


class _Program:
    """cl.Program(context, source).build(): the kernels of the reference .cl, precompiled by oracle/Makefile."""

    KERNELS = {"intersect": "strict", "intersect_postproc": "strict", "reflect_refract_rays": "strict",
               "stereograph_project": "proj", "angular_project": "proj", "rotate_vec": "proj"}

    def __init__(self, context=None, source=None):
        self.source = source
        self._libs = {"strict": _C.CDLL(os.path.join(_HERE, "_ref", "libref_strict.so")),
                      "proj": _C.CDLL(os.path.join(_HERE, "_ref", "libref_proj.so"))}

    def build(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name not in self.KERNELS:
            raise AttributeError(name)
        fn = getattr(self._libs[self.KERNELS[name]], "ref_launch_" + name)
        fn.restype = None

        def launch(queue, global_size, local_size, *args):
            cargs = [_C.c_int(int(global_size[0]))]
            for a in args:
                if isinstance(a, _DevArray):
                    cargs.append(_C.c_void_p(a.a.ctypes.data))
                elif isinstance(a, (np.integer, int)):
                    cargs.append(_C.c_int(int(a)))
                elif isinstance(a, (np.floating, float)):
                    cargs.append(_C.c_float(float(a)))
                else:
                    raise TypeError("kernel argument %r" % (a,))
            fn(*cargs)
            return _Event()

        return launch


class _Device:
    name = "reference kernels compiled by g++ (oracle/_ref)"
    global_mem_size = 8 * 1024 ** 3


class _Platform:
    name = "host"

    def get_devices(self):
        return [_Device()]


class _ClModule:
    class mem_flags:
        READ_ONLY, WRITE_ONLY, READ_WRITE, COPY_HOST_PTR = 1, 2, 4, 8

    Program = _Program

    @staticmethod
    def get_platforms():
        return [_Platform()]

    @staticmethod
    def Context(devices=None):
        return object()

    @staticmethod
    def CommandQueue(context=None):
        return object()


def kernels_available():
    return all(os.path.isfile(os.path.join(_HERE, "_ref", f)) for f in ("libref_strict.so", "libref_proj.so"))


def reference_tracer():
    """An instance of the reference's CL_Tracer (its own __init__ ran: it reads the kernel source from the current
    directory, so the call happens inside /root/reference) whose `cl` / `cl_array` are the stand-ins above."""
    mod = load("iterative_tracer")
    mod.__dict__["cl"] = _ClModule
    mod.__dict__["cl_array"] = _ClArrayModule
    cwd = os.getcwd()
    os.chdir(REFERENCE_DIR)
    try:
        with captured_stdout():
            tr = mod.CL_Tracer()
    finally:
        os.chdir(cwd)
    return tr


@contextlib.contextmanager
def reference_scene_modules():
    """Inside the block lightpycl_b200.scenes builds its scenes with the REFERENCE's light_source and
    geo_optical_elements classes (scenes.py names them through two module globals)."""
    from lightpycl_b200 import scenes

    old = scenes.goe, scenes.ls_mod
    scenes.goe, scenes.ls_mod = load("geo_optical_elements"), load("light_source")
    try:
        with captured_stdout():
            yield scenes
    finally:
        scenes.goe, scenes.ls_mod = old
