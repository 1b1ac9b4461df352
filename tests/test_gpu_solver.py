"""The explicit pseudo-time driver (fvg_euler_step / fvg_forward_euler_solve) against the oracle's
restatement of SteadyForwardEulerSolver::solve (reference src/ode/aodesolver.cpp:136-282): step-by-step
residual history, final state, lift/drag, and the error taxonomy (Tolerance_error, Numerical_error)."""
import numpy as np
import pytest
import torch
import orc
from common import rel_err_by_component
from gpu_common import make_case
from fvens_b200 import lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [
    dict(mesh="2dcylinderhybrid.msh", flux="ROE", recon="VENKATAKRISHNAN", reorder="hilbert", tile=64),
    dict(mesh="bump:40:15", flux="HLLC", gradient="GREENGAUSS", recon="VANALBADA", reorder="none", tile=128),
    dict(mesh="2dcylinderhybrid.msh", flux="ROE", order2=False, reorder="rcm", tile=32),
    dict(mesh="2dcylinderhybrid.msh", flux="ROE", recon="NONE", viscous=True, Reinf=100.0, tile=64),
])
def test_history_and_state_follow_the_oracle(cfg):
    fl, of, u, um = make_case(limiter_param=3.0, Minf=0.5, **cfg)
    nsteps = 60
    code0, steps0, hist0, u0 = of.forward_euler(u, 0.4, 1e-30, nsteps)
    du = torch.from_numpy(u).cuda()
    code, steps, hist = fl.solve_forward_euler(du, 0.4, 1e-30, nsteps)
    assert (code, steps) == (5, nsteps) and (code0, steps0) == (1, nsteps)      # maxiter reached in both
    # round-off grows with the step count through the non-linear update; 60 steps stay well inside 1e-10
    assert np.abs(hist/hist0 - 1).max() < 1e-10
    assert rel_err_by_component(du.cpu().numpy(), u0) < 1e-10
    assert np.abs(hist[0]/hist0[0] - 1) < 1e-12


def test_single_step_entry_point_equals_solver_step():
    fl, of, u, um = make_case("bump:40:15", recon="VENKATAKRISHNAN", tile=64)
    du = torch.from_numpy(u).cuda()
    n2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    fl.euler_step(du, 0.5, n2)
    torch.cuda.synchronize()
    code0, steps0, hist0, u0 = of.forward_euler(u, 0.5, 1e-30, 1)
    assert rel_err_by_component(du.cpu().numpy(), u0) < 1e-12
    assert abs(np.sqrt(n2.item())/hist0[0] - 1) < 1e-12
    # and the fused step equals residual + explicit update done by hand
    du2 = torch.from_numpy(u).cuda()
    res = torch.zeros_like(du2); dt = torch.zeros(len(u), dtype=torch.float64, device="cuda")
    fl.compute_residual(du2, res, True, dt, accumulate=False)
    area = torch.from_numpy(um.arrays()["area"]).cuda()
    unew = du2 + (0.5*dt/area)[:, None]*res
    assert rel_err_by_component(du.cpu().numpy(), unew.cpu().numpy()) < 1e-14


def test_converges_and_reports_steps():
    fl, of, u, um = make_case("2dcylinderhybrid.msh", flux="HLLC", order2=False, tile=64, amp=0.01)
    du = torch.from_numpy(u).cuda()
    code, steps, hist = fl.solve_forward_euler(du, 0.5, 1e-2, 5000)
    code0, steps0, hist0, u0 = of.forward_euler(u, 0.5, 1e-2, 5000)
    assert code == 0 and code0 == 0
    assert steps == steps0
    assert hist[-1]/hist[0] <= 1e-2
    # batched norm read-back runs whole batches and stops after the batch that converged
    du = torch.from_numpy(u).cuda()
    code, steps_b, hist_b = fl.solve_forward_euler(du, 0.5, 1e-2, 5000, check_every=16)
    assert code == 0 and steps <= steps_b < steps + 16
    assert np.array_equal(hist_b[:steps], hist)


def test_numerical_error_on_blow_up():
    # tests/flowpseudotime.cpp: an unstable CFL must end in Numerical_error, not in a silent NaN state
    fl, of, u, um = make_case("2dcylinderhybrid.msh", flux="ROE", order2=True, recon="NONE", tile=64)
    du = torch.from_numpy(u).cuda()
    code, steps, hist = fl.solve_forward_euler(du, 500.0, 1e-12, 400)
    assert code == 6 and not np.isfinite(hist[-1])


def test_lift_and_drag_after_identical_steps():
    fl, of, u, um = make_case("naca0012luo.msh", flux="ROE", recon="VENKATAKRISHNAN", Minf=0.8, aoa=1.25*np.pi/180,
                              tile=128, amp=0.0)
    nsteps = 200
    _, _, hist0, u0 = of.forward_euler(u, 0.3, 1e-30, nsteps)
    du = torch.from_numpy(u).cuda()
    fl.solve_forward_euler(du, 0.3, 1e-30, nsteps)
    g = torch.zeros(um.nelem, 8, dtype=torch.float64, device="cuda")
    fl.getGradients(du, g)
    cl, cdp, cdf = fl.computeSurfaceData(du, g, 2)
    cl0, cdp0, cdf0 = of.surface_data(u0, of.get_gradients(u0), 2)
    assert abs(cl-cl0) < 1e-10 and abs(cdp-cdp0) < 1e-10 and abs(cdf-cdf0) < 1e-10


@pytest.mark.parametrize("cfg", [dict(recon="VENKATAKRISHNAN"), dict(order2=False, flux="HLLC"),
                                 dict(mesh="2dcylinderhybrid.msh", recon="NONE", viscous=True, Reinf=100.0)])
def test_matrix_free_jacobian_vector_product(cfg):
    """fvg_jacobian_vector_product against MatrixFreeSpatialJacobian::apply (reference src/linalg/alinalg.cpp:143-230)
    restated with the oracle's residual: y = mdt x + (R(u) - R(u + h x))/h, h = eps/|x|_2, R = what compute_residual
    leaves. The difference quotient divides round-off of size 1e-16 |R| by h ~ 1e-7/|x|, so two correct residual
    implementations agree on y only to about 1e-7 relative: the tolerance below is 5e-6 of max|y|."""
    mesh = cfg.pop("mesh", "bump:40:15")
    fl, of, u, um = make_case(mesh, tile=64, **cfg)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(u.shape)*np.abs(u).mean(axis=0)*0.1
    eps = 1e-7
    res0, dt0, _, _ = of.residual(u)
    mdt = um.arrays()["area"]/dt0
    h = eps/np.linalg.norm(x.ravel())
    yg0, _, _, _ = of.residual(u + h*x, False)
    y0 = mdt[:, None]*x + (res0 - yg0)/h
    du, dx = torch.from_numpy(u).cuda(), torch.from_numpy(x).cuda()
    dres = torch.zeros_like(du); ddt = torch.zeros(len(u), dtype=torch.float64, device="cuda")
    fl.compute_residual(du, dres, True, ddt, accumulate=False)
    dmdt = torch.from_numpy(um.arrays()["area"]).cuda()/ddt
    dy = torch.zeros_like(du)
    fl.jacobian_vector_product(du, dres, dmdt, dx, dy, eps)
    torch.cuda.synchronize()
    y = dy.cpu().numpy()
    assert np.abs(y - y0).max() < 5e-6*np.abs(y0).max()
    # it IS the directional derivative: halving x halves the Jacobian part (linearity up to the difference error)
    dy2 = torch.zeros_like(du)
    fl.jacobian_vector_product(du, dres, dmdt, 0.5*dx, dy2, eps)
    torch.cuda.synchronize()
    jx = y - mdt[:, None]*x
    jx2 = dy2.cpu().numpy() - mdt[:, None]*0.5*x
    assert np.abs(jx2 - 0.5*jx).max() < 1e-5*np.abs(jx).max()
