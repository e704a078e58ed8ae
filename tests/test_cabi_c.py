"""The boundary is a C ABI: include/lightpycl_b200.h must compile as plain C (no C++, no CUDA, no
torch headers) and every function it declares must resolve against the shared library at link
time.  No GPU needed: the program is linked, not run."""

import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lightpycl_b200.h")
LIB_DIR = os.path.join(ROOT, "lightpycl_b200", "_lib")


def test_header_is_plain_c_and_links(tmp_path):
    from lightpycl_b200 import _cabi

    _cabi.load()  # make sure the library exists
    names = sorted(set(re.findall(r"\b(lpcl_[a-z_0-9]+)\s*\(", open(HEADER).read())) - {"lpcl_step_counters", "lpcl_generation"})
    assert len(names) >= 15
    src = tmp_path / "link_all.c"
    body = "\n".join("    p[%d] = (fn_t)%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include "lightpycl_b200.h"\n#include <stdio.h>\n'
                   "typedef void (*fn_t)(void);\nint main(void) {\n    fn_t p[%d];\n%s\n"
                   "    lpcl_step_counters c; lpcl_generation g; (void)c; (void)g;\n"
                   '    printf("%%d %%d\\n", LPCL_ABI_VERSION, p[0] != 0);\n    return 0;\n}\n' % (len(names), body))
    exe = tmp_path / "link_all"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
           "-L", LIB_DIR, "-llightpycl_b200", "-Wl,-rpath," + LIB_DIR, "-o", str(exe)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    # struct layouts the Python binding mirrors
    chk = tmp_path / "sizes.c"
    chk.write_text('#include "lightpycl_b200.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu\\n", '
                   "sizeof(lpcl_step_counters), sizeof(lpcl_generation)); return 0;}\n")
    exe2 = tmp_path / "sizes"
    assert subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(chk), "-o", str(exe2)]).returncode == 0
    sizes = subprocess.run([str(exe2)], capture_output=True, text=True).stdout.split()
    import ctypes

    assert [int(x) for x in sizes] == [ctypes.sizeof(_cabi.StepCounters), ctypes.sizeof(_cabi.Generation)]
