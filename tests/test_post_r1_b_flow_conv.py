"""The reference's grid-convergence acceptance test for the explicit solver (tests/flow_conv.cpp, registered as
Flow_Explicit_Euler_Cylinder_GreenGauss_Roe_Tri_EntropyConvergence in tests/inv-2dcyl/CMakeLists.txt): inviscid flow
past a cylinder at Mach 0.38 on the reference's meshes 2dcylinder{0,1,2}.msh with its control file
expl-inv-cyl-gg-roe_tri.ctrl (first-order starter, then Green-Gauss + Roe to a relative residual of 1e-4); the
entropy error must fall with an observed order in [1.65, 2.1] between the two finest meshes.

The same procedure is run on the oracle (CPU; the two coarser meshes in the CPU suite, all three once by hand:
entropy errors 0.0655314050, 0.0195118119, 0.00496115021, orders 1.748, 1.976) and by the flow_conv program on the
GPU (tests/cpp/flow_conv.cpp). The GPU case was written after the round's GPU minutes were spent (the test_post_r1_* files sort after
the verified GPU tests on purpose)."""
import os
import re
import subprocess

import numpy as np
import pytest

import orc
from common import ROOT, MESHDIR, mesh_path
from fvens_b200 import lib

CTRL = os.path.join(ROOT, "tests", "golden", "ctrl")
FLOW_CONV = os.path.join(ROOT, "tests", "cpp", "flow_conv")
ENTROPY = [0.06553140498914722, 0.019511811850745065, 0.004961150206335574]      # oracle, meshes 0, 1, 2


def oracle_case(imesh):
    """SteadyFlowCase::run_output of the control file on the oracle: starter (first order, cfl 0.5, tol 1e-1, <= 15000
    steps), main solve (GG + Roe, cfl 0.25, tol 1e-4, <= 60000 steps), entropy error."""
    phys = lib.make_physics(1.4, 0.38, 288.15, 5000.0, 0.72, 0.0, False, False)
    bcs = [(2, lib.BC["slipwall"], ()), (4, lib.BC["farfield"], ())]
    om = orc.Mesh.read(mesh_path(f"2dcylinder{imesh}.msh"))
    n = len(om.arrays()["area"])
    u = np.tile(lib.freestream(phys), (n, 1))
    f1 = orc.Flow(om, phys, lib.FLUX["ROE"], lib.GRAD["NONE"], lib.RECON["NONE"], 1.0, False, 0, bcs)
    c1, _, _, u = f1.forward_euler(u, 0.5, 1e-1, 15000)
    f2 = orc.Flow(om, phys, lib.FLUX["ROE"], lib.GRAD["GREENGAUSS"], lib.RECON["NONE"], 1.0, True, 0, bcs)
    c2, _, _, u = f2.forward_euler(u, 0.25, 1e-4, 60000)
    assert c1 == 0 and c2 == 0
    return 1.0/np.sqrt(n), f2.entropy_error(u)


def test_oracle_entropy_convergence_on_the_two_coarser_meshes():
    orc.set_threads(os.cpu_count() or 1)
    (h0, e0), (h1, e1) = oracle_case(0), oracle_case(1)
    assert abs(e0/ENTROPY[0] - 1) < 1e-6 and abs(e1/ENTROPY[1] - 1) < 1e-6
    assert abs(np.log10(e1/e0)/np.log10(h1/h0) - 1.7478) < 1e-3
    # with the recorded value of the finest mesh: the order the reference's test accepts
    assert 1.65 <= np.log10(ENTROPY[2]/ENTROPY[1])/np.log10(0.5) <= 2.1


@pytest.mark.gpu
def test_flow_conv_program_passes_the_reference_s_acceptance_window(tmp_path):
    r = subprocess.run([FLOW_CONV, os.path.join(CTRL, "expl-inv-cyl-gg-roe_tri.ctrl"), "--source_dir", CTRL,
                        "--number_of_meshes", "3", "--mesh_file", os.path.join(MESHDIR, "2dcylinder"),
                        "--log_file_prefix", str(tmp_path / "2dcyl")], capture_output=True, text=True, timeout=1200,
                       cwd=str(tmp_path))
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "--------------- End" in r.stdout
    pairs = re.findall(r"Log of Mesh size and error are (\S+)\s+(\S+)", r.stdout)
    assert len(pairs) == 3
    for (lh, le), n, e in zip(pairs, (128, 512, 2048), ENTROPY):
        assert abs(float(lh) - np.log10(1/np.sqrt(n))) < 1e-10
        assert abs(10**float(le)/e - 1) < 1e-3            # same converged state as the oracle's run
    orders = [float(x) for x in re.findall(r"^\s+(\S+)\s*$", r.stdout.split(">> Spatial orders =")[1].split("---")[0], re.M)]
    assert len(orders) == 2 and abs(orders[0] - 1.748) < 5e-3 and 1.65 <= orders[1] <= 2.1


def test_flow_conv_rejects_what_it_cannot_run():
    """No GPU needed: usage errors and an implicit control file are refused before any device work."""
    assert subprocess.run([FLOW_CONV], capture_output=True).returncode == 2
    r = subprocess.run([FLOW_CONV, os.path.join(CTRL, "expl-inv-cyl-gg-roe_tri.ctrl"), "--source_dir", CTRL, "--number_of_meshes", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "at least 2" in r.stderr
    r = subprocess.run([FLOW_CONV, os.path.join(CTRL, "naca0012-transonic-implicit.ctrl"), "--number_of_meshes", "3"],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "implicit" in r.stderr
