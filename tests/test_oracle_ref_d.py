"""Tier D: the explicit pseudo-time driver. The oracle's restatement of SteadyForwardEulerSolver::solve
(oracle/orc_spatial.hpp, the checker of the GPU solver tests) against the REFERENCE'S OWN solver object code
(ode/aodesolver.cpp compiled unmodified on top of its own FlowFV, oracle/ref_tier_d.cpp), and the residual-history
writer of fvens_b200/host/casesolvers.hpp against the reference's writer (spatial/aoutput.cpp:617-636) character by
character."""
import os
import subprocess
import numpy as np
import pytest

import orc
from common import ROOT, mesh_path, rel_err_by_component, INVISCID_BCS, VISCOUS_BCS
from fvens_b200 import lib, synth

pytestmark = pytest.mark.skipif(not orc.have_ref_d(), reason="oracle/_ref/libfvens_ref_d.so not built (needs /root/reference)")


def setup(mesh, flux, gradient, recon, order2=True, viscous=False, lp=1.0):
    om = orc.Mesh.read(mesh_path(mesh))
    a = om.arrays()
    phys = lib.make_physics(1.4, 0.5, 288.15, 100.0, 0.72, 0.02, viscous, False)
    tags = set(a["btags"].tolist())
    bcs = [(t, lib.BC[ty], v) for (t, ty, v) in (VISCOUS_BCS if viscous else INVISCID_BCS) if t in tags]
    of = orc.Flow(om, phys, lib.FLUX[flux], lib.GRAD[gradient], lib.RECON[recon], lp, order2, 0, bcs)
    rs = orc.RefSolver(a, phys, flux, gradient if order2 else "NONE", recon if order2 else "NONE", lp, order2, bcs)
    rc, _, _ = of.geometry()
    u = synth.perturbed_state(rc, 1.4, 0.5, 0.02, amp=0.05)
    return of, rs, u


@pytest.mark.parametrize("cfg", [("2dcylinderhybrid.msh", "ROE", "LEASTSQUARES", "VANALBADA", True, False),
                                 ("naca0012luo.msh", "HLLC", "GREENGAUSS", "NONE", True, False),
                                 ("2dcylinderhybrid.msh", "HLL", "NONE", "NONE", False, False),
                                 ("2dcylinderhybrid.msh", "ROE", "LEASTSQUARES", "NONE", True, True)])
def test_forward_euler_loop_against_reference_solver(cfg):
    mesh, flux, gradient, recon, order2, viscous = cfg
    of, rs, u = setup(mesh, flux, gradient, recon, order2, viscous)
    nsteps = 60
    code0, steps0, hist0, u0 = of.forward_euler(u, 0.4, 1e-30, nsteps)
    code1, steps1, rel1, abs1, u1 = rs.forward_euler(u, 0.4, 1e-30, nsteps)
    assert (code0, steps0) == (1, nsteps) and (code1, steps1) == (1, nsteps)         # both: Tolerance_error at max iterations
    # the state in double precision; round-off grows through 60 non-linear steps
    assert rel_err_by_component(u0, u1) < 1e-10
    # the reference keeps its history in single precision
    assert np.abs(hist0/abs1 - 1).max() < 3e-7 and np.abs(hist0/hist0[0]/rel1 - 1).max() < 3e-7


def test_converged_exit_and_tolerance():
    of, rs, u = setup("2dcylinderhybrid.msh", "ROE", "NONE", "NONE", False)
    code0, steps0, hist0, u0 = of.forward_euler(u, 0.5, 0.5, 2000)        # stop once the residual has halved
    code1, steps1, rel1, abs1, u1 = rs.forward_euler(u, 0.5, 0.5, 2000)
    assert code0 == 0 and code1 == 0 and steps0 == steps1 and 1 < steps0 < 2000
    assert rel_err_by_component(u0, u1) < 1e-10 and rel1[-1] <= 0.5 < rel1[-2]


def test_residual_history_writer_matches_the_reference_writer():
    of, rs, u = setup("2dcylinderhybrid.msh", "ROE", "NONE", "NONE", False)
    steps = [1, 2, 50, 51, 1234, 200000]
    rel = [1.0, 0.731234, 3.4e-3, 1.0e-5, 9.87654321e-9, 1.2e-12]
    abs_ = [0.0123, 0.0091, 4.2e-5, 1.2e-7, 1.1e-10, 1.5e-14]
    wt = [0.001, 0.0025, 1.25, 1.3, 321.5, 98765.4]
    cfl = [0.25, 0.25, 0.5, 0.5, 1.0, 1000.0]
    want = rs.history_text(steps, rel, abs_, wt, cfl)
    args = [str(x) for row in zip(steps, rel, abs_, wt, cfl) for x in row]
    got = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_controlparser"), "--history", *args], capture_output=True, text=True).stdout
    assert got == want and want.count("\n") == 2 + len(steps)


def _stage_table(order):
    return lib.tvdrk_coefficients(order)


def test_reference_tvdrk_loop_as_it_is_and_what_the_product_does_instead():
    """TVDRKSolver::solve of the reference (ode/aodesolver.cpp:672-785), run from its own object code, equals its loop
    restated literally: every stage takes compute_residual at the step's INITIAL state (`uvec`, :719) and the update is
    SUBTRACTED (:740) although compute_residual leaves -r(u) (the forward-Euler loop adds it, :208). So one order-1 step
    of the reference is exactly the mirror image of a forward-Euler step with the global time step - it integrates
    backwards in time. The product's fvg_tvdrk_solve / TVDRKSolver keep the coefficient table, the time step
    (cfl * min dtm of the first stage) and the loop condition, evaluate at the stage state and add the update; that
    scheme's order of accuracy is shown in tests/cpp/test_ode_host.cpp. Here: the shared pieces (table, dt) are pinned
    by the reference's object code."""
    of, rs, u = setup("2dcylinderhybrid.msh", "ROE", "LEASTSQUARES", "VANALBADA")     # (no H1 hazard with MUSCL)
    area = orc.Mesh.read(mesh_path("2dcylinderhybrid.msh")).arrays()["area"]
    cfl = 0.4
    r, dtm, _, _ = of.residual(u)                       # what compute_residual leaves: -r(u), local time steps
    dt = cfl*dtm.min()
    # order 1, one step (finaltime tiny: the loop runs once, the step is not clipped)
    code, u1 = rs.tvdrk(u, 1, cfl, 1e-9)
    assert code == 0
    assert rel_err_by_component(u1, u - (dt/area)[:, None]*r) < 1e-12          # the reference: minus
    forward = u + (dt/area)[:, None]*r                                         # the product's order-1 step: plus
    assert rel_err_by_component(2*u - u1, forward) < 1e-12
    # order 3, two steps: the literal loop in numpy on the oracle's residual
    c = _stage_table(3)
    lit = u.copy()
    time = 0.0
    for _ in range(2):
        rr, dd, _, _ = of.residual(lit)
        dts = cfl*dd.min()
        us = lit.copy()
        for i in range(3):
            us = c[i, 0]*lit + c[i, 1]*us - (c[i, 2]*dts/area)[:, None]*rr     # rr: always at the step's initial state
        lit = us
        time += dts
    code, u3 = rs.tvdrk(u, 3, cfl, time - 0.25*dts)
    assert code == 0 and rel_err_by_component(u3, lit) < 1e-11
    # ... which is not a third-order scheme: with the residual frozen, the three stages collapse to ONE Euler-like step
    # of weight sum_i (prod of later b's) c_i = 1 in the wrong direction
    rr, dd, _, _ = of.residual(u)
    one = u - (cfl*dd.min()/area)[:, None]*rr
    code, u3_1 = rs.tvdrk(u, 3, cfl, 1e-9)
    assert rel_err_by_component(u3_1, one) < 1e-12
