"""tests/c_driver.c replays config C1 through include/dpomp.h with no Python in the process: compiled with gcc against the
in-tree libdpomp.so and run as a subprocess (the boundary must be usable by a host that is not Python)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_c_driver_replays_config_c1_without_python(tmp_path, built_lib):
    gcc = shutil.which("gcc")
    assert gcc, "gcc not found"
    libdir = os.path.join(ROOT, "discretepomp.jl_b200", "lib")
    exe = str(tmp_path / "c_driver")
    subprocess.run([gcc, "-O1", "-std=c11", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_driver.c"), "-o", exe,
                    "-L", libdir, "-ldpomp", "-lm", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "C DRIVER OK" in res.stdout, res.stdout
