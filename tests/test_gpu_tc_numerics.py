"""The tensor-core search filter (lightpycl_b200/csrc/lpcl_search_tc.cuh) is conservative only if the
kind::tf32 tcgen05.mma keeps the error bound its margin budgets: <= 16 u * sum|terms| per output,
u = 2^-24 (derivation in lpcl_math.cuh).  The bound is not documented by the vendor, so it is
RE-MEASURED here on the GPU the tests run on with tools/ubench_tmem.cu (built by
__graft_entry__.build()): operand layout against a CPU GEMM, truncation probes, and random TF32
operands over 24 binades."""

import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "ubench_tmem")


@pytest.fixture(scope="module")
def mma_report():
    if not os.path.exists(BIN):
        subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", BIN,
                               os.path.join(ROOT, "tools", "ubench_tmem.cu")])
    out = subprocess.run([BIN, "mma"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def _line(report, prefix, ksteps):
    for ln in report.splitlines():
        if ln.startswith(prefix) and ("ksteps %d:" % ksteps) in ln:
            return ln
    raise AssertionError("no %r line (ksteps %d) in:\n%s" % (prefix, ksteps, report))


def test_operand_layout_matches_a_cpu_gemm(mma_report):
    """no-swizzle K-major descriptors + lane/column mapping of the accumulator: exact on small integers"""
    ln = _line(mma_report, "layout (small ints)", 3)
    assert "inexact 0 /" in ln, ln


def test_tf32_mma_error_stays_inside_the_margin_budget(mma_report):
    for ksteps, limit in ((1, 4.0), (3, 8.0)):  # budget 16 u over the three chained MMAs of the filter
        ln = _line(mma_report, "random tf32 operands", ksteps)
        units = float(re.search(r"= ([0-9.]+) u", ln).group(1))
        assert units <= limit, "tf32 MMA error %.2f u exceeds %.1f u (%s)" % (units, limit, ln)


def test_accumulation_is_wide_then_truncated(mma_report):
    """2^24 + 7 ones: each addend alone is below half an ulp; a unit that truncated addends one by one
    would return 16777216.  Observed: the exact sum, truncated once (16777222)."""
    rows = re.findall(r"row 0: d\[0\]=([0-9.e+]+)", mma_report)
    assert rows and float(rows[0]) in (16777222.0, 16777224.0), rows
    ln = _line(mma_report, "non-tf32 operands (trunc?)", 1)
    assert "inexact 0 /" in ln, "fp32 operands are expected to be truncated to tf32: " + ln
