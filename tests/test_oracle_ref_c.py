"""Tier C: the oracle's FlowFV::compute_residual (oracle/orc_spatial.hpp) against the REFERENCE'S OWN
compute_residual - spatial/flow_spatial.cpp with aspatial.cpp, tracevector.cpp, petscutils.cpp and the tier-A/B
sources compiled unmodified, in place, against the stand-ins of oracle/ref_shim_b (oracle/ref_tier_c.cpp ->
oracle/_ref/libfvens_ref_c.so). Residual and local time steps, every flux, gradient scheme and reconstruction, first
and second order, inviscid and viscous (Sutherland and constant viscosity), on hybrid meshes.
With Barth-Jespersen / Venkatakrishnan the reference reads past the end of the cell states at boundary cells
(zero-filled slack in the harness): those cells, and the cells next to them, are excluded - everywhere
else the limiter sees only defined data."""
import numpy as np
import pytest

import orc
from common import mesh_path, INVISCID_BCS, VISCOUS_BCS
from fvens_b200 import lib, synth

pytestmark = pytest.mark.skipif(not orc.have_ref_c(), reason="oracle/_ref/libfvens_ref_c.so not built (needs /root/reference)")
TOL = 1e-12


def run(mesh, flux="ROE", gradient="LEASTSQUARES", recon="NONE", lp=1.0, order2=True, viscous=False, const_visc=False, shock=False):
    om = orc.Mesh.from_arrays(*synth.bump_channel(36, 14)) if mesh == "bump" else orc.Mesh.read(mesh_path(mesh))
    a = om.arrays()
    phys = lib.make_physics(1.4, 0.6, 288.15, 800.0, 0.72, 0.03, viscous, const_visc)
    tags = set(a["btags"].tolist())
    bcs = [(t, lib.BC[ty], v) for (t, ty, v) in (VISCOUS_BCS if viscous else INVISCID_BCS) if t in tags]
    of = orc.Flow(om, phys, lib.FLUX[flux], lib.GRAD[gradient], lib.RECON[recon], lp, order2, 0, bcs)
    rc, _, _ = of.geometry()
    u = synth.perturbed_state(rc, 1.4, 0.6, 0.03, amp=0.08, shock=shock)
    r0, dt0, _, _ = of.residual(u)
    r1, dt1 = orc.ref_residual(a, phys, flux, gradient if order2 else "NONE", recon if order2 else "NONE", lp, order2, bcs, u)
    return om, a, r0, dt0, r1, dt1


def compare(a, r0, dt0, r1, dt1, clean=None):
    scale = np.abs(r1).max(axis=0)
    err = np.abs(r0 - r1)/scale
    edt = np.abs(dt0/dt1 - 1)
    if clean is not None:
        err, edt = err[clean], edt[clean]
        assert clean.sum() > 0.5*len(clean)
    assert np.isfinite(r1).all() and err.max() < TOL and edt.max() < TOL, (err.max(), edt.max())


def cells_untouched_by_h1(om, a):
    """Cells that are neither at a physical boundary nor next to a cell that is."""
    n = om.nelem
    bcell = np.zeros(n, dtype=bool)
    bcell[a["intfac"][:om.nbface, 0]] = True
    near = bcell.copy()
    interior = a["intfac"][om.nbface:, :2]
    near[interior[bcell[interior[:, 1]], 0]] = True
    near[interior[bcell[interior[:, 0]], 1]] = True
    return ~near


@pytest.mark.parametrize("flux", ["LLF", "VANLEER", "AUSM", "AUSMPLUS", "ROE", "HLL", "HLLC"])
@pytest.mark.parametrize("order2", [False, True])
def test_all_fluxes_first_and_second_order(flux, order2):
    om, a, r0, dt0, r1, dt1 = run("2dcylinderhybrid.msh", flux=flux, order2=order2)
    compare(a, r0, dt0, r1, dt1)


@pytest.mark.parametrize("mesh,shock", [("naca0012luo.msh", True), ("bump", False)])
@pytest.mark.parametrize("gradient", ["LEASTSQUARES", "GREENGAUSS"])
@pytest.mark.parametrize("recon,lp", [("NONE", 1.0), ("WENO", 1.0), ("WENO", 20.0), ("VANALBADA", 1.0)])
def test_second_order_reconstructions(mesh, shock, gradient, recon, lp):
    om, a, r0, dt0, r1, dt1 = run(mesh, flux="HLLC", gradient=gradient, recon=recon, lp=lp, shock=shock)
    compare(a, r0, dt0, r1, dt1)


@pytest.mark.parametrize("mesh,shock", [("naca0012luo.msh", True), ("bump", True)])
@pytest.mark.parametrize("gradient", ["LEASTSQUARES", "GREENGAUSS"])
@pytest.mark.parametrize("recon,lp", [("BARTHJESPERSEN", 1.0), ("VENKATAKRISHNAN", 0.5), ("VENKATAKRISHNAN", 6.0)])
def test_limiters_away_from_the_h1_cells(mesh, shock, gradient, recon, lp):
    om, a, r0, dt0, r1, dt1 = run(mesh, flux="ROE", gradient=gradient, recon=recon, lp=lp, shock=shock)
    compare(a, r0, dt0, r1, dt1, cells_untouched_by_h1(om, a))


@pytest.mark.parametrize("const_visc", [False, True])
@pytest.mark.parametrize("order2,recon", [(False, "NONE"), (True, "NONE"), (True, "VANALBADA"), (True, "WENO")])
def test_viscous_residual(const_visc, order2, recon):
    om, a, r0, dt0, r1, dt1 = run("NACA0012_lam_hybrid_1.msh", flux="ROE", recon=recon, order2=order2, viscous=True, const_visc=const_visc)
    compare(a, r0, dt0, r1, dt1)


def test_persistent_reference_flow_serial_and_openmp_builds():
    """The handle-based entry points bench.py's reference arm times (ref_flow_create / ref_flow_residual), in the serial
    build and in the build with the reference's OpenMP pragmas on: the same numbers as the one-shot call (weighted
    least squares: the Green-Gauss boundary loop is the one that races under OpenMP)."""
    om = orc.Mesh.read(mesh_path("naca0012luo.msh"))
    a = om.arrays()
    phys = lib.make_physics(1.4, 0.6, 288.15, 800.0, 0.72, 0.03)
    bcs = [(t, lib.BC[ty], v) for (t, ty, v) in INVISCID_BCS if t in set(a["btags"].tolist())]
    of = orc.Flow(om, phys, lib.FLUX["ROE"], lib.GRAD["LEASTSQUARES"], lib.RECON["VANALBADA"], 1.0, True, 0, bcs)
    rc, _, _ = of.geometry()
    u = synth.perturbed_state(rc, 1.4, 0.6, 0.03, amp=0.08, shock=True)
    r1, dt1 = orc.ref_residual(a, phys, "ROE", "LEASTSQUARES", "VANALBADA", 1.0, True, bcs, u)
    for omp in ([False, True] if orc.have_ref_c_omp() else [False]):
        rf = orc.RefFlow(a, phys, "ROE", "LEASTSQUARES", "VANALBADA", 1.0, True, bcs, omp=omp)
        assert rf.threads() >= 1
        for _ in range(2):       # repeated evaluations on one object
            r, dt = rf.residual(u)
            # atomics reorder the sums under OpenMP: round-off, not bits
            assert np.abs(r - r1).max() < 1e-13*np.abs(r1).max() and np.abs(dt/dt1 - 1).max() < 1e-13
    r0, dt0, _, _ = of.residual(u)
    compare(a, r0, dt0, r1, dt1)
