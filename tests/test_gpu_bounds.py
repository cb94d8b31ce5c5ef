"""Memory-safety stand-in for compute-sanitizer (closed on this pool): lib/variants/libdpomp_bounds.so is the
library compiled with -DDPOMP_BOUNDS_CHECK -- every shared / global index of the resample phase (windows, heavy-tile path,
offspring rows, ancestor staging), of the warp work queue of the simulate kernel and of the MBP trajectory windows is
checked against its buffer and traps.  The parity tests of the ragged, weight-collapse, multi-chunk (1029 tiles), event-cap
and trajectory-overflow cases are re-run against that build in a subprocess: they must still pass bit for bit, i.e. no
index ever left its buffer.  A second subprocess proves the checker is live."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BOUNDS_LIB = os.path.join(ROOT, "discretepomp.jl_b200", "lib", "variants", "libdpomp_bounds.so")
CASES = ("bit_exact or collapse or scan_chunk or fused_step or overflow or interleaved or multinomial_seam or zero_rate "
         "or weight_pass or batched_filters or warp_and_thread or iterate_and_propose or export_import or obs_id_zero")


def _env():
    if not os.path.exists(BOUNDS_LIB):
        import __graft_entry__ as ge

        ge.build()
    env = dict(os.environ)
    env["DPOMP_LIB_PATH"] = BOUNDS_LIB
    return env


def test_checker_is_live():
    code = ("import ctypes, sys; lib = ctypes.CDLL(sys.argv[1]); "
            "assert lib.dpomp_debug_bounds_selftest(3, 4) == 0; "   # in range: nothing happens
            "sys.exit(0 if lib.dpomp_debug_bounds_selftest(4, 4) != 0 else 1)")  # out of range: the kernel traps
    res = subprocess.run([sys.executable, "-c", code, BOUNDS_LIB], env=_env(), capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "dpomp bounds:" in res.stdout + res.stderr


def test_edge_case_parity_suite_passes_on_the_bounds_checked_build():
    res = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", CASES, "-p", "no:cacheprovider",
                          os.path.join(ROOT, "tests", "test_gpu_pf.py"), os.path.join(ROOT, "tests", "test_gpu_mbp.py"),
                          os.path.join(ROOT, "tests", "test_gpu_resample.py")],
                         env=_env(), capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    assert "dpomp bounds:" not in res.stdout + res.stderr, tail
    assert " passed" in res.stdout, tail
