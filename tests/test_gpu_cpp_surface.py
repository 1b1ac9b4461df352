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


DIST_BIN = os.path.join(ROOT, "tests", "cpp", "test_dist_surface")


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 3])
def test_cpp_distributed_surface_on_one_gpu(nranks, tmp_path):
    """DistributedFlowFV + SteadyForwardEulerSolver (the multi-GPU half of the class surface): nranks processes share
    cuda:0, the set-up all-gather goes through files; the merged state after 60 steps and its residual equal the
    single-GPU run bit for bit, the history (kept in single precision, as the reference does) to 1e-6."""
    import numpy as np
    import torch
    from fvens_b200 import lib
    mesh = os.path.join(MESHDIR, "naca0012luo.msh")
    procs = [subprocess.Popen([DIST_BIN, mesh, str(r), str(nranks), str(tmp_path), "0"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(nranks)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        print(o[-1500:])
        assert p.returncode == 0 and f"DIST_SURFACE OK rank {r} steps 60" in o
    um = lib.UMesh.read(mesh)
    state = np.zeros((um.nelem, 4)); resid = np.zeros((um.nelem, 4)); seen = np.zeros(um.nelem, dtype=int)
    hists = []
    for r in range(nranks):
        raw = open(os.path.join(tmp_path, f"state_{r}.bin"), "rb").read()
        nown, nst, flag = np.frombuffer(raw, dtype=np.int32, count=3)
        assert flag == 1 and nst == 60          # Tolerance_error at maxiter, as the reference's driver reports it
        ids = np.frombuffer(raw, dtype=np.int32, count=nown, offset=12)
        off = 12 + 4*nown
        state[ids] = np.frombuffer(raw, dtype=np.float64, count=4*nown, offset=off).reshape(nown, 4)
        resid[ids] = np.frombuffer(raw, dtype=np.float64, count=4*nown, offset=off + 32*nown).reshape(nown, 4)
        hists.append(np.frombuffer(raw, dtype=np.float64, count=nst, offset=off + 64*nown))
        seen[ids] += 1
    assert (seen == 1).all()
    phys = lib.make_physics(1.4, 0.8, 288.15, 5000.0, 0.72, 1.25*np.pi/180.0)
    dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=256)
    fl = lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "VENKATAKRISHNAN", 2.0, True, 0, [(2, "slipwall", ()), (4, "farfield", ())])
    du = torch.from_numpy(np.tile(lib.freestream(phys), (um.nelem, 1))).cuda()
    code, steps, h1 = fl.solve_forward_euler(du, 0.4, 1e-30, 60)
    r1 = torch.zeros_like(du); d1 = torch.zeros(um.nelem, dtype=torch.float64, device="cuda")
    fl.compute_residual(du, r1, True, d1, accumulate=False)
    assert np.array_equal(state, du.cpu().numpy())
    assert np.array_equal(resid, r1.cpu().numpy())
    for h in hists:
        assert np.abs(h/h1 - 1).max() < 1e-6 and np.array_equal(h, hists[0])
