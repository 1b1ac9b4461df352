"""The C++ class surface (fvens_b200/host/fvens_b200.hpp) exercised by a C++ program that reads like the
reference's own unit tests (tests/cpp/test_host_surface.cpp): flux known-answer, wall BCs, 1-exact gradients,
the FlowFV::compute_residual contract and the forward-Euler driver's error behaviour."""
import os
import subprocess
import pytest
from common import ROOT, MESHDIR

BIN = os.path.join(ROOT, "tests", "cpp", "test_host_surface")


def test_cpp_surface_builds_and_links():
    """CPU side: the header-only surface compiles as C++14 against the ABI and the binary resolves the library."""
    assert os.path.exists(BIN)
    out = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libfvens_b200.so" in out and "not found" not in out.split("libfvens_b200.so")[1].splitlines()[0]


@pytest.mark.gpu
def test_cpp_surface_on_gpu():
    r = subprocess.run([BIN, MESHDIR], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "HOST_SURFACE OK" in r.stdout
