"""Unsteady driver of the host class surface (TVDRKSolver in fvens_b200/host/fvens_b200.hpp; reference
ode/aodesolver.hpp:195-258, aodesolver.cpp:45-67, 647-785) on a model problem through its generic (host Vec) loop:
no GPU needed. The cases are in tests/cpp/test_ode_host.cpp."""
import os
import subprocess

import numpy as np

from common import ROOT, mesh_path
from fvens_b200 import lib

BIN = os.path.join(ROOT, "tests", "cpp", "test_ode_host")


def test_tvdrk_host_driver(tmp_path):
    r = subprocess.run([BIN, mesh_path("testhybrid.msh"), str(tmp_path / "log")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout


def test_coefficient_table_is_the_reference_s():
    # initialize_TVDRK_Coeffs, ode/aodesolver.cpp:45-67; consistency: a + b = 1 in every stage (convex combination)
    assert np.array_equal(lib.tvdrk_coefficients(1), [[1.0, 0.0, 1.0]])
    assert np.array_equal(lib.tvdrk_coefficients(2), [[1.0, 0.0, 1.0], [0.5, 0.5, 0.5]])
    c3 = lib.tvdrk_coefficients(3)
    assert np.array_equal(c3, [[1.0, 0.0, 1.0], [0.75, 0.25, 0.25], [0.3333333333333333, 0.6666666666666667, 0.6666666666666667]])
    assert np.abs(c3[:, 0] + c3[:, 1] - 1).max() < 1e-15


def test_argument_validation_needs_no_gpu():
    import ctypes as C
    L = lib.load()
    steps, time = C.c_int(7), C.c_double(7.0)
    assert L.fvg_tvdrk_solve(None, None, 2, C.c_double(0.5), C.c_double(1.0), 0, C.byref(steps), C.byref(time)) != 0
    assert "fvg_tvdrk_solve: bad argument" in L.fvg_last_error().decode()
    c = (C.c_double*9)()
    assert L.fvg_tvdrk_coefficients(0, c) != 0 and "temporal order 0 not available" in L.fvg_last_error().decode()
    assert L.fvg_tvdrk_coefficients(3, None) != 0
