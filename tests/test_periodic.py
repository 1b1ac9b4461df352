"""Periodic boundaries (BASELINE configs[4] / SURVEY 8d config 5: the isentropic-vortex box is periodic in x and y).

The reference pairs periodic faces in UMesh::compute_periodic_map (src/mesh/mesh.cpp:369-424) but cannot run them: it has
no periodic FlowBC (abc.cpp:493-494 throws) and its face loop would count each periodic edge twice, so
there is no reference result to match. The product turns every periodic pair into interior faces whose far cell is a ghost
copy of the partner's cell displaced by the period, filled by the same exchange that serves a partition. The oracle for
it is an ORDINARY evaluation on the 3 x 3 unfolded mesh (synth.unfold_periodic): in the centre copy every cell has its true
periodic neighbourhood (two layers deep, enough for every second-order scheme here), so the oracle's residual there is what
the periodic evaluation must give - to 1e-12; the unfolded node coordinates differ from base + period by round-off."""
import os

import numpy as np
import pytest

import orc
from common import rel_err_by_component
from fvens_b200 import lib, synth

L = 10.0
G, M = 1.4, 0.5
PBCS = [(3, "periodic", ()), (4, "periodic", ())]


def periodic_mesh(n, hybrid):
    arrs = synth.periodic_square(n, tri_fraction=0.3 if hybrid else 0.0, jitter=0.15 if hybrid else 0.0)
    um = lib.UMesh.from_arrays(*arrs)
    assert um.compute_periodic_map(3, 0) == n and um.compute_periodic_map(4, 1) == n
    return arrs, um


def smooth_periodic_state(rc, amp=0.05):
    x, y = rc[:, 0], rc[:, 1]
    d = amp*np.sin(2*np.pi*x/L)*np.cos(4*np.pi*y/L)
    pinf = 1.0/(G*M*M)
    rho = 1.0 + d
    vx = 1.0 + d
    vy = 0.3 + 0.5*d
    p = pinf*(1.0 + d)
    return np.ascontiguousarray(np.stack([rho, rho*vx, rho*vy, p/(G - 1.0) + 0.5*rho*(vx*vx + vy*vy)], axis=1))


def test_periodic_map_pairs_opposite_sides():
    arrs, um = periodic_mesh(8, True)
    pp = um.periodic_partners()
    a = um.arrays()
    mid = 0.5*(a["coords"][a["intfac"][:um.nbface, 2]] + a["coords"][a["intfac"][:um.nbface, 3]])
    assert (pp >= 0).all() and (pp[pp] == np.arange(um.nbface)).all()
    d = np.abs(mid - mid[pp])
    # partners sit one period apart in exactly one coordinate
    assert np.allclose(np.sort(d, axis=1), [[0.0, L]]*um.nbface, atol=1e-9)


def test_device_mesh_has_one_ghost_per_periodic_face():
    arrs, um = periodic_mesh(8, True)
    dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=32, device=-2)      # host-only build
    sc, rc, idx = dm.halo_lists()
    assert dm.nghost == um.nbface == 32 and list(sc) == [32] and list(rc) == [32]
    # ghost rows are copies of own cells: the permutation lists the partner's global cell
    perm = dm.permutation()
    assert (np.sort(perm[:dm.ncell]) == np.arange(um.nelem)).all() and (perm[dm.ncell:] < um.nelem).all()


def test_unfolded_mesh_is_consistent():
    arrs, um = periodic_mesh(6, True)
    un = synth.unfold_periodic(arrs[0], arrs[1], arrs[2], L)
    om = orc.Mesh.from_arrays(*un)
    assert om.nelem == 9*um.nelem and om.nbface == 4*3*6
    assert abs(om.arrays()["area"].sum() - 9*L*L) < 1e-9


CASES = [
    ("inviscid", dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="NONE")),
    ("inviscid", dict(flux="HLLC", gradient="GREENGAUSS", reconstruction="VENKATAKRISHNAN", limiter_param=2.0)),
    ("inviscid", dict(flux="AUSM", gradient="LEASTSQUARES", reconstruction="BARTHJESPERSEN")),
    ("inviscid", dict(flux="VANLEER", gradient="LEASTSQUARES", reconstruction="WENO", limiter_param=5.0)),
    ("inviscid", dict(flux="HLL", gradient="GREENGAUSS", reconstruction="VANALBADA")),
    ("inviscid", dict(flux="LLF", gradient="NONE", reconstruction="NONE", order2=False)),
    ("viscous", dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="NONE")),
    ("viscous", dict(flux="AUSMPLUS", gradient="GREENGAUSS", reconstruction="VENKATAKRISHNAN", limiter_param=2.0)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("reorder", ["none", "hilbert"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_periodic_residual_equals_the_unfolded_mesh(case, reorder):
    import torch
    kind, num = CASES[case]
    n = 12
    arrs, um = periodic_mesh(n, True)
    phys = lib.make_physics(G, M, 288.15, 500.0, 0.72, 0.0, kind == "viscous", False)
    dm = lib.DeviceMesh(um, reorder=reorder, tile_cells=32)
    fl = lib.FlowFV(dm, phys, bcs=PBCS, **num)
    rc = synth.cell_centres(*arrs[:3])
    u = smooth_periodic_state(rc)
    du = torch.from_numpy(u).cuda()
    res = torch.full_like(du, 7.0)
    dt = torch.zeros(len(u), dtype=torch.float64, device="cuda")
    fl.compute_residual(du, res, True, dt, accumulate=False)
    # the oracle on the unfolded mesh, centre copy
    un = synth.unfold_periodic(arrs[0], arrs[1], arrs[2], L)
    om = orc.Mesh.from_arrays(*un)
    rcu = synth.cell_centres(*un[:3])
    of = orc.Flow(om, phys, lib.FLUX[num["flux"]], lib.GRAD.get(num["gradient"], 0), lib.RECON[num["reconstruction"]],
                  num.get("limiter_param", 1.0), num.get("order2", True), 0, [(9, lib.BC["farfield"], (0.0, 0.0))])
    r0, dt0, _, _ = of.residual(smooth_periodic_state(rcu))
    ne = um.nelem
    assert rel_err_by_component(res.cpu().numpy(), r0[:ne]) < 1e-12
    assert np.abs(dt.cpu().numpy()/dt0[:ne] - 1).max() < 1e-12
    # accumulate contract and a second call (pushed rows, evaluation counter) give the same
    res2 = torch.ones_like(du)
    fl.compute_residual(du, res2, True, dt, accumulate=True)
    assert rel_err_by_component((res2 - 1.0).cpu().numpy(), r0[:ne]) < 1e-12


@pytest.mark.gpu
def test_periodic_free_stream_and_conservation():
    """A uniform state has zero residual on the periodic box (no boundary at all), and for any state the residuals sum to
    zero: every edge, periodic ones included, gives and takes the same flux up to round-off."""
    import torch
    arrs, um = periodic_mesh(24, True)
    phys = lib.make_physics(G, M, 288.15, 5000.0, 0.72, 0.3)
    dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=64)
    fl = lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "VENKATAKRISHNAN", 2.0, True, 0, PBCS)
    uinf = lib.freestream(phys)
    du = torch.from_numpy(np.tile(uinf, (um.nelem, 1))).cuda()
    res = torch.zeros_like(du); dt = torch.zeros(um.nelem, dtype=torch.float64, device="cuda")
    fl.compute_residual(du, res, True, dt, accumulate=False)
    assert float(res.abs().max()) < 1e-12
    du = torch.from_numpy(smooth_periodic_state(synth.cell_centres(*arrs[:3]), amp=0.2)).cuda()
    fl.compute_residual(du, res, True, dt, accumulate=False)
    assert float(res.sum(dim=0).abs().max()) < 1e-11*float(res.abs().sum(dim=0).max())


@pytest.mark.gpu
def test_periodic_forward_euler_steps_match_stepping_the_unfolded_oracle():
    """fvg_euler_step / fvg_forward_euler_solve on the periodic box against explicit steps with the oracle's residual on the
    unfolded mesh (whose 9 copies stay identical only as long as its outer boundary has not reached the centre: 3 steps)."""
    import torch
    arrs, um = periodic_mesh(12, False)
    phys = lib.make_physics(G, M, 288.15, 5000.0, 0.72, 0.0)
    dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=32)
    fl = lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "NONE", 1.0, True, 0, PBCS)
    rc = synth.cell_centres(*arrs[:3])
    u0 = smooth_periodic_state(rc)
    du = torch.from_numpy(u0).cuda()
    code, steps, hist = fl.solve_forward_euler(du, 0.4, 1e-30, 3)
    assert code == 5 and steps == 3
    un = synth.unfold_periodic(arrs[0], arrs[1], arrs[2], L)
    om = orc.Mesh.from_arrays(*un)
    of = orc.Flow(om, phys, lib.FLUX["ROE"], lib.GRAD["LEASTSQUARES"], lib.RECON["NONE"], 1.0, True, 0, [(9, lib.BC["farfield"], (0.0, 0.0))])
    uu = smooth_periodic_state(synth.cell_centres(*un[:3]))
    area = om.arrays()["area"]
    ne = um.nelem
    h0 = []
    for _ in range(3):
        r, dtm, _, _ = of.residual(uu)
        h0.append(np.sqrt((r[:ne, 3]**2*area[:ne]).sum()))
        uu = uu + (0.4*dtm/area)[:, None]*r
    assert rel_err_by_component(du.cpu().numpy(), uu[:ne]) < 1e-11
    assert np.abs(np.array(hist)/np.array(h0) - 1).max() < 1e-10
    du2 = torch.from_numpy(u0).cuda()
    for _ in range(3):
        fl.euler_step(du2, 0.4)
    assert torch.equal(du2, du)


def density_error(u, rc, area, t):
    # the vortex has crossed the periodic box t/L times: compare with the exact solution wrapped into the box
    x = rc.copy()
    x[:, 0] = (x[:, 0] - t + L/2) % L - L/2 + t
    ex = synth.isentropic_vortex(x, G, M, t)
    return float(np.sqrt(((u[:, 0] - ex[:, 0])**2*area).sum()))


@pytest.mark.gpu
@pytest.mark.parametrize("hybrid,rk", [(False, 2), (True, 3)])
def test_vortex_order_of_accuracy_with_true_periodic_boundaries(hybrid, rk):
    """BASELINE configs[4]: the vortex convected once around the periodic box (t = 10: it leaves through the right side and
    comes back through the left); density error in the area-weighted L2 norm on 32^2, 64^2, 128^2 cells (jittered hybrid
    meshes too): second order, as tests/flow_conv.cpp:62-89 judges it."""
    import torch
    errs, ncell = [], []
    phys = lib.make_physics(G, M, 288.15, 5000.0, 0.72, 0.0)
    for n in (32, 64, 128):
        arrs, um = periodic_mesh(n, hybrid)
        dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=256)
        fl = lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "NONE", 1.0, True, 0, PBCS)
        rc = synth.cell_centres(*arrs[:3])
        area = um.arrays()["area"]
        du = torch.from_numpy(synth.isentropic_vortex(rc, G, M, 0.0)).cuda()
        code, steps, time = fl.solve_tvdrk(du, rk, 0.4, 10.0)
        assert code == 0 and 10.0 <= time < 10.05
        errs.append(density_error(du.cpu().numpy(), rc, area, time)); ncell.append(len(area))
    p = [np.log(errs[k+1]/errs[k])/np.log(np.sqrt(ncell[k]/ncell[k+1])) for k in range(2)]
    print("density errors", errs, "orders", p)
    assert 1.7 < p[0] < 2.8 and 1.7 < p[1] < 2.8 and errs[2] < 1e-2


@pytest.mark.gpu
def test_periodic_marker_and_boundary_condition_must_agree():
    arrs, um = periodic_mesh(8, False)
    phys = lib.make_physics(G, M, 288.15, 5000.0, 0.72, 0.0)
    dm = lib.DeviceMesh(um, reorder="none", tile_cells=32)
    with pytest.raises(lib.FvgError) as e:
        lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "NONE", 1.0, True, 0, [(3, "farfield", ()), (4, "periodic", ())])
    assert e.value.code == 4
    um2 = lib.UMesh.from_arrays(*arrs)        # no pairing computed
    dm2 = lib.DeviceMesh(um2, reorder="none", tile_cells=32)
    with pytest.raises(lib.FvgError) as e:
        lib.FlowFV(dm2, phys, "ROE", "LEASTSQUARES", "NONE", 1.0, True, 0, PBCS)
    assert e.value.code == 4
