"""The explicit time-accurate driver fvg_tvdrk_solve (TVDRKSolver of the reference, ode/aodesolver.cpp:672-785, as a
strong-stability-preserving Runge-Kutta scheme) against the same scheme stepped in numpy on the
oracle's residual: state after a few steps, step count, physical time, the loop condition, the divergence exit.
The scheme itself (coefficients, observed order 1/2/3, dt = cfl*min(dtm) from the first stage) is tested without a
GPU through the host class's generic loop (tests/test_ode_host.py).

This file was written after the round's GPU minutes were spent; the test_post_r1_* files sort after the verified GPU tests on purpose."""
import numpy as np
import pytest
import torch
from common import rel_err_by_component
from gpu_common import make_case
from fvens_b200 import lib

pytestmark = [pytest.mark.gpu]


def oracle_tvdrk(of, u, area, order, cfl, finaltime, maxsteps=0):
    c = lib.tvdrk_coefficients(order)
    u = u.copy()
    time, step = 0.0, 0
    while time <= finaltime - 1e-12 and (maxsteps <= 0 or step < maxsteps):
        us = u.copy()
        for i in range(order):
            r, dtm, _, _ = of.residual(us)
            if i == 0:
                dtmin = dtm.min()
            us = c[i, 0]*u + c[i, 1]*us + (c[i, 2]*cfl*dtmin/area)[:, None]*r
        u = us
        step += 1
        time += dtmin*cfl
    return u, step, time


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("cfg", [
    dict(mesh="2dcylinderhybrid.msh", flux="ROE", recon="VENKATAKRISHNAN", reorder="hilbert", tile=64),
    dict(mesh="bump:40:15", flux="HLLC", gradient="GREENGAUSS", recon="BARTHJESPERSEN", reorder="none", tile=128),
])
def test_state_steps_and_time_follow_the_oracle(cfg, order):
    fl, of, u, um = make_case(limiter_param=3.0, Minf=0.5, **cfg)
    area = um.arrays()["area"]
    u0, steps0, time0 = oracle_tvdrk(of, u, area, order, 0.4, 1e30, maxsteps=6)
    du = torch.from_numpy(u).cuda()
    code, steps, time = fl.solve_tvdrk(du, order, 0.4, 1e30, maxsteps=6)
    assert code == 0 and steps == steps0 == 6
    assert abs(time/time0 - 1) < 1e-12
    assert rel_err_by_component(du.cpu().numpy(), u0) < 1e-11
    assert np.abs(u0 - u).max() > 1e-6            # the state did move


def test_stops_on_the_final_time_without_clipping_the_last_step():
    fl, of, u, um = make_case("2dcylinderhybrid.msh", flux="ROE", recon="NONE", tile=64, Minf=0.5)
    area = um.arrays()["area"]
    _, dtm, _, _ = of.residual(u)
    tfinal = 2.5*0.5*dtm.min()
    u0, steps0, time0 = oracle_tvdrk(of, u, area, 2, 0.5, tfinal)
    du = torch.from_numpy(u).cuda()
    code, steps, time = fl.solve_tvdrk(du, 2, 0.5, tfinal)
    assert code == 0 and steps == steps0 == 3 and time > tfinal and abs(time/time0 - 1) < 1e-12
    assert rel_err_by_component(du.cpu().numpy(), u0) < 1e-12
    # nothing to do: state untouched
    du2 = torch.from_numpy(u).cuda()
    assert fl.solve_tvdrk(du2, 3, 0.5, 0.0) == (0, 0, 0.0) and np.array_equal(du2.cpu().numpy(), u)


def test_order_one_step_is_a_global_time_step_forward_euler_update():
    fl, of, u, um = make_case("bump:40:15", recon="VENKATAKRISHNAN", tile=64)
    du = torch.from_numpy(u).cuda()
    res = torch.zeros_like(du); dt = torch.zeros(len(u), dtype=torch.float64, device="cuda")
    fl.compute_residual(du, res, True, dt, accumulate=False)
    area = torch.from_numpy(um.arrays()["area"]).cuda()
    expect = du + (0.3*dt.min()/area)[:, None]*res
    code, steps, time = fl.solve_tvdrk(du, 1, 0.3, 1e30, maxsteps=1)
    assert (code, steps) == (0, 1) and abs(time/(0.3*dt.min().item()) - 1) < 1e-15
    assert rel_err_by_component(du.cpu().numpy(), expect.cpu().numpy()) < 1e-14


def test_divergence_and_bad_arguments():
    fl, of, u, um = make_case("2dcylinderhybrid.msh", flux="ROE", order2=True, recon="NONE", tile=64)
    du = torch.from_numpy(u).cuda()
    code, steps, time = fl.solve_tvdrk(du, 1, 1000.0, 1e30, maxsteps=2000)     # far beyond the stability limit
    assert code == 6 and steps < 2000 and "dtmin is Nan or inf" in lib.load().fvg_last_error().decode()
    with pytest.raises(lib.FvgError):
        fl.solve_tvdrk(du, 4, 0.5, 1.0)
    with pytest.raises(lib.FvgError):
        fl.solve_tvdrk(du, 2, -1.0, 1.0)
