"""Subdomain meshes on the GPU. The exchange is emulated inside one process (ghost rows are filled from a
global array), so this runs on a single B200: it checks everything the multi-GPU path does on the device -
own/ghost numbering, cut faces with a ghost on either side, split passes - against the single-mesh result
(bitwise: cut faces keep their global orientation, so every rank evaluates the identical expression) and
against the oracle."""
import numpy as np
import pytest
import torch
import orc
from common import rel_err_by_component, INVISCID_BCS, VISCOUS_BCS
from gpu_common import load_mesh
from fvens_b200 import lib, synth

pytestmark = pytest.mark.gpu


def run_partitioned(um, phys, bcs, nranks, u, numerics, reorder="hilbert", tile=64):
    part = lib.partition_sfc(um, nranks)
    n = um.nelem
    ranks = []
    for r in range(nranks):
        dm = lib.DeviceMesh(um, reorder=reorder, tile_cells=tile, cell_rank=part, rank=r, nranks=nranks)
        fl = lib.FlowFV(dm, phys, bcs=bcs, **numerics)
        ids = torch.from_numpy(dm.permutation().astype(np.int64)).cuda()
        ntot = dm.ncell + dm.nghost
        lg = torch.zeros((ntot, 8), dtype=torch.float64, device="cuda")
        gu = torch.zeros((ntot, 8), dtype=torch.float64, device="cuda")
        fl.use_buffers(lg, gu)
        ranks.append(dict(dm=dm, fl=fl, ids=ids, lg=lg, gu=gu))
    ug = torch.from_numpy(u).cuda()
    glg = torch.zeros((n, 8), dtype=torch.float64, device="cuda"); ggu = torch.zeros_like(glg)
    weno = numerics.get("reconstruction", "NONE") == "WENO"
    order2 = numerics.get("order2", True)
    for R in ranks:
        R["u"] = ug[R["ids"]].contiguous()          # own rows + ghost rows (= the state exchange)
        if order2:
            R["fl"].gradient_pass(R["u"], 0)
    if order2:
        def share(name, glob):
            for R in ranks:
                glob[R["ids"][:R["dm"].ncell]] = R[name][:R["dm"].ncell]
            for R in ranks:
                R[name][R["dm"].ncell:] = glob[R["ids"][R["dm"].ncell:]]
        share("gu", ggu)
        if weno:
            for R in ranks:
                R["fl"].gradient_pass(R["u"], 1)
        share("lg", glg)
    res = torch.zeros((n, 4), dtype=torch.float64, device="cuda"); dt = torch.zeros(n, dtype=torch.float64, device="cuda")
    for R in ranks:
        nc = R["dm"].ncell
        r = torch.zeros((nc, 4), dtype=torch.float64, device="cuda"); d = torch.zeros(nc, dtype=torch.float64, device="cuda")
        R["fl"].face_pass(R["u"], r, True, d)
        res[R["ids"][:nc]] = r; dt[R["ids"][:nc]] = d
    torch.cuda.synchronize()
    return res.cpu().numpy(), dt.cpu().numpy()


CASES = [
    dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="VENKATAKRISHNAN", limiter_param=2.0),
    dict(flux="HLLC", gradient="GREENGAUSS", reconstruction="BARTHJESPERSEN"),
    dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="WENO", limiter_param=3.0),
    dict(flux="VANLEER", gradient="LEASTSQUARES", reconstruction="VANALBADA"),
    dict(flux="AUSM", order2=False),
]


@pytest.mark.parametrize("nranks", [2, 3])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_partitioned_residual_equals_single_mesh(case, nranks):
    numerics = CASES[case]
    um, om, rc = load_mesh("bump:40:15")
    phys = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.02)
    bcs = [b for b in INVISCID_BCS if b[0] in (2, 3, 4)]
    u = synth.perturbed_state(rc, 1.4, 0.5, 0.02)
    res, dt = run_partitioned(um, phys, bcs, nranks, u, numerics)
    dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=64)
    fl = lib.FlowFV(dm, phys, bcs=bcs, **numerics)
    du = torch.from_numpy(u).cuda(); r1 = torch.zeros_like(du); d1 = torch.zeros(um.nelem, dtype=torch.float64, device="cuda")
    fl.compute_residual(du, r1, True, d1, accumulate=False)
    torch.cuda.synchronize()
    assert np.array_equal(res, r1.cpu().numpy())          # bitwise: the 1/2/3-"GPU" answers are identical
    assert np.array_equal(dt, d1.cpu().numpy())
    of = orc.Flow(om, phys, lib.FLUX[numerics["flux"]], lib.GRAD.get(numerics.get("gradient", "NONE"), 0),
                  lib.RECON[numerics.get("reconstruction", "NONE")], numerics.get("limiter_param", 1.0),
                  numerics.get("order2", True), 0, [(t, lib.BC[ty], v) for (t, ty, v) in bcs])
    r0, dt0, _, _ = of.residual(u)
    assert rel_err_by_component(res, r0) < 1e-12 and np.abs(dt/dt0 - 1).max() < 1e-12


def test_partitioned_viscous_with_limiter():
    um, om, rc = load_mesh("2dcylinderhybrid.msh")
    phys = lib.make_physics(1.4, 0.5, 288.15, 200.0, 0.72, 0.0, True, False)
    bcs = [b for b in VISCOUS_BCS if b[0] in (2, 4)]
    u = synth.perturbed_state(rc, 1.4, 0.5, 0.0)
    numerics = dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="VENKATAKRISHNAN", limiter_param=2.0)
    res, dt = run_partitioned(um, phys, bcs, 2, u, numerics, tile=32)
    of = orc.Flow(om, phys, 4, 2, 4, 2.0, True, 0, [(t, lib.BC[ty], v) for (t, ty, v) in bcs])
    r0, dt0, _, _ = of.residual(u)
    assert rel_err_by_component(res, r0) < 1e-12 and np.abs(dt/dt0 - 1).max() < 1e-12


def test_single_call_entry_points_refuse_subdomain_meshes():
    um, _, rc = load_mesh("bump:40:15")
    part = lib.partition_sfc(um, 2)
    dm = lib.DeviceMesh(um, cell_rank=part, rank=0, nranks=2, tile_cells=64)
    fl = lib.FlowFV(dm, lib.make_physics(), bcs=[b for b in INVISCID_BCS if b[0] in (2, 3, 4)])
    n = dm.ncell + dm.nghost
    u = torch.ones((n, 4), dtype=torch.float64, device="cuda"); r = torch.zeros((dm.ncell, 4), dtype=torch.float64, device="cuda")
    with pytest.raises(lib.FvgError) as e:
        fl.compute_residual(u, r, False, None)
    assert e.value.code == 4
