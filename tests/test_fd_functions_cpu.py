"""CPU suite: the reference's own finite-difference checks, applied to OUR hand-derived adjoints.

north_star: "gradients pass the repo's own finite-difference checks".  Those checks are the reference's C++ unit tests
(test_d_intersect, test_d_sample_shape, test_d_bsdf, test_d_sample_primary_rays, test_d_camera_to_screen; run on the reference itself
by tests/test_oracle_cpu.py::test_reference_unit_tests).  tests/fd_functions.cpp restates them for the device functions of
redner_b200/csrc/*.cuh: the headers are compiled for the host with Real = double and every adjoint is compared with central
differences of its primal at the reference's inputs and tolerance (1e-3, src/test_utils.h:15-23).  75 scalar checks."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adjoints_against_finite_differences(tmp_path):
    if shutil.which("g++") is None or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path / "fd_functions")
    data = os.path.join(ROOT, "redner_b200", "data")
    cmd = ["g++", "-O1", "-std=c++17", "-w", "-DRB_REAL_DOUBLE", "-include", os.path.join(ROOT, "tools", "cpu_emu", "emu_shim.h"), "-I/usr/local/cuda/include",
           "-I" + os.path.join(ROOT, "include"), '-DRB_DATA_DIR="%s"' % data, os.path.join(ROOT, "tests", "fd_functions.cpp"), "-o", exe, "-lpthread"]
    subprocess.run(cmd, check=True, timeout=900)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert len([l for l in lines if l.startswith("ok ")]) == 5 and lines[-1].startswith("checks "), r.stdout
    assert int(lines[-1].split()[1]) >= 75
