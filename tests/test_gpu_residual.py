"""Parity of the CUDA residual path (through the C ABI) with the CPU oracle on the same mesh and state.
Bar (BASELINE.json north_star): 1e-12 relative per component, the component's infinity norm being the
scale; local time steps 1e-12 relative."""
import numpy as np
import pytest
import torch
import orc
from common import rel_err_by_component, rel_err
from gpu_common import make_case, gpu_residual, load_mesh
from fvens_b200 import lib

pytestmark = pytest.mark.gpu
TOL = 1e-12

FLUXES = ["LLF", "VANLEER", "AUSM", "AUSMPLUS", "ROE", "HLL", "HLLC"]


def check(fl, of, u, gettimesteps=True):
    r, dt = gpu_residual(fl, u, gettimesteps)
    r0, dt0, _, _ = of.residual(u, gettimesteps)
    e = rel_err_by_component(r, r0)
    assert e < TOL, f"residual rel err {e:.3e}"
    if gettimesteps:
        ed = np.abs(dt/dt0 - 1.0).max()
        assert ed < TOL, f"dt rel err {ed:.3e}"


@pytest.mark.parametrize("flux", FLUXES)
@pytest.mark.parametrize("mesh,reorder,tile", [("2dcylinderhybrid.msh", "hilbert", 64), ("bump:40:15", "none", 128),
                                               ("naca0012luo.msh", "rcm", 256)])
def test_first_order_all_fluxes(flux, mesh, reorder, tile):
    fl, of, u, _ = make_case(mesh, flux=flux, order2=False, reorder=reorder, tile=tile, Minf=0.8)
    check(fl, of, u)


@pytest.mark.parametrize("gradient", ["LEASTSQUARES", "GREENGAUSS"])
@pytest.mark.parametrize("recon", ["NONE", "WENO", "VANALBADA", "BARTHJESPERSEN", "VENKATAKRISHNAN"])
@pytest.mark.parametrize("flux,mesh,reorder,tile", [("ROE", "2dcylinderhybrid.msh", "hilbert", 32),
                                                    ("HLLC", "bump:40:15", "hilbert", 128),
                                                    ("VANLEER", "naca0012luo.msh", "none", 512)])
def test_second_order(gradient, recon, flux, mesh, reorder, tile):
    fl, of, u, _ = make_case(mesh, flux=flux, gradient=gradient, recon=recon, reorder=reorder, tile=tile,
                             limiter_param=3.0, Minf=0.8)
    check(fl, of, u)


@pytest.mark.parametrize("flux", ["LLF", "AUSM", "AUSMPLUS", "HLL"])
def test_second_order_remaining_fluxes(flux):
    fl, of, u, _ = make_case("bump:40:15", flux=flux, recon="VENKATAKRISHNAN", tile=64, limiter_param=1.0)
    check(fl, of, u)


@pytest.mark.parametrize("recon", ["BARTHJESPERSEN", "VENKATAKRISHNAN", "WENO", "VANALBADA"])
def test_limiters_with_a_shock(recon):
    fl, of, u, _ = make_case("bump:40:15", flux="HLLC", gradient="GREENGAUSS", recon=recon, shock=True,
                             limiter_param=5.0, Minf=0.6, tile=96)
    check(fl, of, u)


@pytest.mark.parametrize("bnd_policy", [0, 1])
@pytest.mark.parametrize("recon", ["BARTHJESPERSEN", "VENKATAKRISHNAN"])
def test_boundary_neighbour_policy(bnd_policy, recon):
    fl, of, u, _ = make_case("2dcylinderhybrid.msh", recon=recon, bnd_policy=bnd_policy, tile=64)
    check(fl, of, u)


def test_zero_gradients_second_order():
    fl, of, u, _ = make_case("bump:40:15", gradient="NONE", recon="NONE")
    check(fl, of, u)


@pytest.mark.parametrize("const_visc", [True, False])
@pytest.mark.parametrize("order2,recon", [(False, "NONE"), (True, "NONE"), (True, "VENKATAKRISHNAN"), (True, "VANALBADA"),
                                          (True, "WENO")])
def test_viscous(const_visc, order2, recon):
    fl, of, u, _ = make_case("2dcylinderhybrid.msh", flux="ROE", recon=recon, order2=order2, viscous=True,
                             const_visc=const_visc, Reinf=200.0, tile=64)
    check(fl, of, u)


def test_config2_laminar_naca_hybrid():
    # BASELINE.json configs[1]: visc-naca0012, Roe + WLS + no limiter + modified-average-gradient viscous flux
    fl, of, u, um = make_case("NACA0012_lam_hybrid_1.msh", flux="ROE", gradient="LEASTSQUARES", recon="NONE",
                              viscous=True, const_visc=False, Reinf=5000.0, Minf=0.5, aoa=0.0, tile=512)
    assert um.nelem == 13156
    check(fl, of, u)


def test_config1_inviscid_naca_roe_wls_venkat():
    # BASELINE.json configs[0]: naca0012 Euler, Roe + WLS + Venkatakrishnan
    fl, of, u, um = make_case("NACA0012_inv.su2", flux="ROE", gradient="LEASTSQUARES", recon="VENKATAKRISHNAN",
                              Minf=0.8, aoa=1.25*np.pi/180, limiter_param=2.0, tile=512)
    assert um.nelem == 10216
    check(fl, of, u)


def test_accumulate_adds_into_residual_and_overwrite_does_not():
    fl, of, u, _ = make_case("2dcylinderhybrid.msh", recon="NONE", tile=64)
    du = torch.from_numpy(u).cuda()
    base = torch.full_like(du, 3.0)
    res = base.clone()
    fl.compute_residual(du, res, False, None, accumulate=True)
    res2 = base.clone()
    fl.compute_residual(du, res2, False, None, accumulate=False)
    torch.cuda.synchronize()
    r0, _, _, _ = of.residual(u, False)
    assert rel_err_by_component(res2.cpu().numpy(), r0) < TOL
    assert rel_err_by_component((res-base).cpu().numpy(), r0) < 1e-11     # one rounding of r+3 on top


def test_host_buffer_entry_point():
    fl, of, u, _ = make_case("bump:40:15", recon="VENKATAKRISHNAN")
    res = np.zeros_like(u); dt = np.zeros(len(u))
    fl.compute_residual_host(u, res, True, dt)
    res2 = np.ones_like(u); fl.compute_residual_host(u, res2, False, None, accumulate=False)
    assert np.array_equal(res2, res)
    r0, dt0, _, _ = of.residual(u)
    assert rel_err_by_component(res, r0) < TOL and np.abs(dt/dt0-1).max() < TOL


@pytest.mark.parametrize("kw", [dict(recon="VENKATAKRISHNAN"), dict(recon="VANALBADA", viscous=True),
                                dict(order2=False), dict(recon="BARTHJESPERSEN", gradient="GREENGAUSS", flux="HLLC")])
def test_host_buffer_pipeline_matches_the_device_resident_path(kw):
    # device-ordered mesh with many tiles: fvg_residual_host runs its chunked upload / compute / download pipeline
    # (tile ranges launched as soon as the chunks they depend on have arrived); same bits as one full launch
    fl, of, u, _ = make_case("bump:96:36", reorder="none", tile=32, **kw)
    assert fl.dmesh.info.ntile >= 64
    r1, d1 = gpu_residual(fl, u)
    res = np.full_like(u, 7.0); dt = np.zeros(len(u))
    fl.compute_residual_host(u, res, True, dt, accumulate=False)
    assert np.array_equal(res, r1) and np.array_equal(dt, d1)
    r0, dt0, _, _ = of.residual(u)
    assert rel_err_by_component(res, r0) < TOL and np.abs(dt/dt0-1).max() < TOL
    # the class surface always accumulates (the reference adds -r(u) into the caller's vector): the pipeline
    # uploads the residual rows with the state rows and accumulates on the device - same bits as the device-resident call
    import torch
    base = np.random.default_rng(3).standard_normal(u.shape)
    acc = base.copy(); dt2 = np.zeros(len(u))
    fl.compute_residual_host(u, acc, True, dt2, accumulate=True)
    dacc = torch.from_numpy(base).cuda(); ddt = torch.zeros(len(u), dtype=torch.float64, device="cuda")
    fl.compute_residual(torch.from_numpy(u).cuda(), dacc, True, ddt, accumulate=True)
    assert np.array_equal(acc, dacc.cpu().numpy()) and np.array_equal(dt2, d1)


@pytest.mark.parametrize("reorder", ["none", "hilbert", "rcm"])
def test_results_do_not_depend_on_renumbering(reorder):
    fl, of, u, _ = make_case("naca0012luo.msh", flux="HLLC", recon="BARTHJESPERSEN", gradient="GREENGAUSS",
                             reorder=reorder, tile=64, Minf=0.8)
    check(fl, of, u)


def test_bitwise_reproducible():
    fl, _, u, _ = make_case("naca0012luo.msh", recon="VENKATAKRISHNAN", tile=64)
    r1, d1 = gpu_residual(fl, u)
    r2, d2 = gpu_residual(fl, u)
    assert np.array_equal(r1, r2) and np.array_equal(d1, d2)


def test_conservation_and_freestream_preservation():
    # uniform free stream, far-field everywhere on a closed domain => zero residual to round-off
    um, om, rc = load_mesh("square:20")
    phys = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.3)
    bcs = [(t, "farfield", (0, 0)) for t in (1, 2, 3, 4)]
    dm = lib.DeviceMesh(um, "hilbert", 64)
    fl = lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "VENKATAKRISHNAN", 1.0, True, 0, bcs)
    u = np.tile(lib.freestream(phys), (um.nelem, 1))
    r, _ = gpu_residual(fl, u)
    flux_scale = np.abs(orc.flux("orc", 0, phys, u[:1], u[:1], np.array([[0.6, 0.8]]))).max()*um.arrays()["facemetric"][:, 2].max()
    assert np.abs(r).max() < 1e-13*flux_scale*4


# ---- the plug-in level entry points ---------------------------------------------------------------

@pytest.mark.parametrize("gradient", ["LEASTSQUARES", "GREENGAUSS", "NONE"])
@pytest.mark.parametrize("reorder", ["none", "hilbert"])
def test_gradient_scheme_entry_point(gradient, reorder):
    fl, of, u, um = make_case("2dcylinderhybrid.msh", gradient=gradient, reorder=reorder, tile=64)
    up = orc.cons2prim("orc", fl.phys, u)
    ug = orc.cons2prim("orc", fl.phys, of.boundary_states(u[um.arrays()["intfac"][:um.nbface, 0]]))
    g0 = of.gradients(up, ug)
    g = torch.zeros(um.nelem, 8, dtype=torch.float64, device="cuda")
    fl.compute_gradients(torch.from_numpy(up).cuda(), torch.from_numpy(ug).cuda(), g)
    torch.cuda.synchronize()
    assert rel_err(g.cpu().numpy(), g0) < TOL


@pytest.mark.parametrize("recon", ["NONE", "WENO", "VANALBADA", "BARTHJESPERSEN", "VENKATAKRISHNAN"])
@pytest.mark.parametrize("reorder", ["none", "hilbert"])
def test_reconstruction_entry_point(recon, reorder):
    fl, of, u, um = make_case("bump:40:15", recon=recon, reorder=reorder, tile=64, limiter_param=2.0)
    nb = um.nbface
    up = orc.cons2prim("orc", fl.phys, u)
    ug = orc.cons2prim("orc", fl.phys, of.boundary_states(u[um.arrays()["intfac"][:nb, 0]]))
    g0 = of.gradients(up, ug)
    ufl0, ufr0 = of.face_values(up, ug, g0)
    ufl = torch.zeros(um.naface, 4, dtype=torch.float64, device="cuda"); ufr = torch.zeros_like(ufl)
    fl.compute_face_values(torch.from_numpy(up).cuda(), torch.from_numpy(ug).cuda(), torch.from_numpy(g0).cuda(), ufl, ufr)
    torch.cuda.synchronize()
    assert rel_err_by_component(ufl.cpu().numpy(), ufl0) < TOL
    assert rel_err_by_component(ufr.cpu().numpy()[nb:], ufr0[nb:]) < TOL


def test_wls_one_exact_on_gpu():
    # the reference's known-answer test (tests/finite-volume/testgradientschemes.cpp) on the CUDA kernels
    from test_oracle_kat import linear_fields, EPS
    for mesh in ("testperiodic.msh", "2dcylinderhybrid.msh", "squareunsquad0.msh"):
        um, om, _ = load_mesh(mesh)
        tags = sorted(set(um.arrays()["btags"][:, 0].tolist()))
        phys = lib.make_physics()
        dm = lib.DeviceMesh(um, "hilbert", 64)
        fl = lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "NONE", 1.0, True, 0, [(t, "farfield", (0, 0)) for t in tags])
        of = orc.Flow(om, phys, 4, 2, 0, bcs=[(t, 1, (0, 0)) for t in tags])
        rc, gr, rcbp = of.geometry()
        u = torch.from_numpy(linear_fields(rc)).cuda(); ug = torch.from_numpy(linear_fields(rcbp)).cuda()
        g = torch.zeros(um.nelem, 8, dtype=torch.float64, device="cuda")
        fl.compute_gradients(u, ug, g)
        ufl = torch.zeros(um.naface, 4, dtype=torch.float64, device="cuda"); ufr = torch.zeros_like(ufl)
        fl.compute_face_values(u, ug, g, ufl, ufr)
        torch.cuda.synchronize()
        exact = linear_fields(gr)
        ufl = ufl.cpu().numpy(); ufr = ufr.cpu().numpy(); nb = um.nbface
        for k in range(4):
            scale = max(1.0, np.abs(exact[:, k]).max()/3.0)
            assert np.sqrt(((ufl[:, k]-exact[:, k])**2).sum()/um.naface) < 10*EPS*scale
            assert np.sqrt(((ufl[nb:, k]-ufr[nb:, k])**2).sum()/um.naface) < 10*EPS*scale


def test_wall_bcs_zero_flux_on_gpu():
    # tests/flow-general/testwallbcs.cpp through compute_boundary_states + the pointwise flux hook
    from test_oracle_kat import wall_test_state
    um, om, _ = load_mesh("testperiodic.msh")
    a = um.arrays()
    phys = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.0, True)
    for wall in ("adiabaticwall", "slipwall"):
        dm = lib.DeviceMesh(um, "hilbert", 64)
        fl = lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "NONE", 1.0, True, 0,
                        [(4, "farfield", (0, 0)), (2, wall, (0.0, 0)), (3, "isothermalwall", (0.0, 290.0))])
        ins = np.tile(wall_test_state(), (um.nbface, 1))
        gs = torch.zeros(um.nbface, 4, dtype=torch.float64, device="cuda")
        fl.compute_boundary_states(torch.from_numpy(ins).cuda(), gs)
        torch.cuda.synchronize()
        gs = gs.cpu().numpy()
        sel = a["btags"][:, 0] == 2
        for flux in ("HLLC", "ROE", "AUSM", "AUSMPLUS", "HLL", "LLF"):
            f = lib.flux_pointwise(flux, phys, ins, gs, a["facemetric"][:um.nbface, :2])
            assert np.abs(f[sel, 0]).max() < 10*2.2e-16*10      # FMA contraction: one extra decade of slack
            assert np.abs(f[sel, 3]).max() < 100*2.2e-16*10


def test_get_gradients_surface_data_entropy():
    fl, of, u, um = make_case("NACA0012_lam_hybrid_1.msh", viscous=True, Reinf=5000.0, tile=256)
    g0 = of.get_gradients(u)
    du = torch.from_numpy(u).cuda()
    g = torch.zeros(um.nelem, 8, dtype=torch.float64, device="cuda")
    fl.getGradients(du, g)
    torch.cuda.synchronize()
    assert rel_err(g.cpu().numpy(), g0) < TOL
    s0 = of.surface_data(u, g0, 2)
    s = np.array(fl.computeSurfaceData(du, g, 2))
    assert np.abs(s-s0).max() < 1e-12*np.abs(s0).max()
    e0 = of.entropy_error(u)
    assert abs(fl.entropy_error(du)-e0) < 1e-12*e0


def test_missing_bc_marker_is_an_error():
    um, _, _ = load_mesh("2dcylinderhybrid.msh")
    dm = lib.DeviceMesh(um, "none", 64)
    with pytest.raises(lib.FvgError) as e:
        lib.FlowFV(dm, lib.make_physics(), bcs=[(2, "slipwall", (0, 0))])     # marker 4 has no BC
    assert e.value.code == 1
