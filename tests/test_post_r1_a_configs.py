"""BASELINE configs[0] and configs[3] as tests.

configs[0]: testcases/naca0012 inviscid (NACA0012_inv.su2, M 0.8, alpha 1.25 deg), Roe + weighted least
squares + Venkatakrishnan, explicit forward-Euler pseudo-time from the free stream - residual history, state and
lift/drag after 1000 identical steps against the oracle. The oracle run with 1 and with 8 threads (different summation
orders) differs by 1.5e-14 in the history after 1000 steps: round-off does not grow on this case, so the bounds of the
60- and 200-step tests in test_gpu_solver.py hold here as well.
Added after the round's GPU minutes were spent (hence a file of its own; the test_post_r1_* files sort after the verified GPU tests); it
uses only entry points those tests already exercise."""
import numpy as np
import pytest
import torch

from common import rel_err_by_component
from gpu_common import make_case, gpu_residual

pytestmark = [pytest.mark.gpu]


def test_config0_naca0012_thousand_step_convergence_check():
    fl, of, u, um = make_case("NACA0012_inv.su2", flux="ROE", recon="VENKATAKRISHNAN", limiter_param=3.0, Minf=0.8,
                              aoa=1.25*np.pi/180, tile=256, amp=0.0)
    nsteps = 1000
    code0, steps0, hist0, u0 = of.forward_euler(u, 0.2, 1e-30, nsteps)
    du = torch.from_numpy(u).cuda()
    code, steps, hist = fl.solve_forward_euler(du, 0.2, 1e-30, nsteps)
    assert steps == steps0 == nsteps and code == 5 and code0 == 1
    assert np.abs(hist/hist0 - 1).max() < 1e-10
    assert rel_err_by_component(du.cpu().numpy(), u0) < 1e-10
    g = torch.zeros(um.nelem, 8, dtype=torch.float64, device="cuda")
    fl.getGradients(du, g)
    cl, cdp, cdf = fl.computeSurfaceData(du, g, 2)
    cl0, cdp0, cdf0 = of.surface_data(u0, of.get_gradients(u0), 2)
    assert abs(cl-cl0) < 1e-10 and abs(cdp-cdp0) < 1e-10 and abs(cdf-cdf0) < 1e-10
    assert abs(cl0 - 0.0437715) < 1e-6 and abs(cdp0 - 0.0473290) < 1e-6          # oracle values of this run, for the record


@pytest.mark.parametrize("lam", [1.0, 20.0])
@pytest.mark.parametrize("flux", ["LLF", "VANLEER", "AUSM", "HLL", "HLLC", "ROE"])
def test_config3_cylinder_ogrid_weno_sweep_of_six_fluxes(flux, lam):
    """BASELINE configs[3] at test size: inviscid cylinder O-grid (SURVEY 8d config 4 generator, geometric radial
    stretching), weighted least squares + WENO with an explicit central weight (1 and 20), the six fluxes of the sweep,
    M 0.38 as testcases/2dcylinder: residual and time steps against the oracle. (The 50 M-cell / 8-GPU run of the
    config is a bench mode for round 2; the WENO gradient exchange across ranks is covered by test_gpu_partition.py.)"""
    fl, of, u, um = make_case("ogrid:96:40", flux=flux, gradient="LEASTSQUARES", recon="WENO", limiter_param=lam, Minf=0.38,
                              aoa=0.0, tile=128)
    r, dt = gpu_residual(fl, u)
    r0, dt0, _, _ = of.residual(u)
    assert rel_err_by_component(r, r0) < 1e-12 and np.abs(dt/dt0 - 1).max() < 1e-12
