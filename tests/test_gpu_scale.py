"""The benchmark configuration itself under test: the synthetic hybrid bump channel of BASELINE.json configs[2]
(SURVEY 8d generator), host mesh in Hilbert order, 256-cell tiles, device numbering = host numbering - at 1M cells
for both numerics bench.py knows, and at the full 10M cells for the headline numerics. The oracle (restated reference
loops, OpenMP) finishes a 10M-cell evaluation in a couple of seconds, so full-size parity is checked directly and
not only through size-independent properties; those are checked as well (free-stream preservation, bitwise
repeatability, bitwise independence of the tile size, host-buffer pipeline = device-resident path)."""
import os
import sys
import numpy as np
import pytest
import torch

import orc
from common import ROOT, rel_err_by_component
from fvens_b200 import lib

sys.path.insert(0, ROOT)
import bench     # noqa: E402  (the workload builder is the one the benchmark times)

pytestmark = pytest.mark.gpu
TOL = 1e-12


_cases = {}


def build(cells, numerics, tile=256, bcs=None):
    if cells not in _cases:
        _cases.clear()                        # one mesh at a time: the 10M-cell arrays are large
        _cases[cells] = bench.build_case(cells, numerics, tile)
    um, arrs, u, _ = _cases[cells]
    flux, grad, recon, lp = bench.NUMERICS[numerics][:4]
    phys = lib.make_physics(1.4, bench.MINF, 288.15, 5000.0, 0.72, 0.0)
    bcs = bcs or bench.BCS
    dm = lib.DeviceMesh(um, reorder="none", tile_cells=tile)
    fl = lib.FlowFV(dm, phys, flux, grad, recon, lp, True, 0, bcs)
    return um, arrs, u, phys, dm, fl, (flux, grad, recon, lp), bcs


def oracle_residual(arrs, u, phys, num, bcs):
    flux, grad, recon, lp = num
    orc.set_threads(os.cpu_count() or 1)
    om = orc.Mesh.from_arrays(*arrs)
    of = orc.Flow(om, phys, lib.FLUX[flux], lib.GRAD[grad], lib.RECON[recon], lp, True, 0,
                  [(t, lib.BC[ty], v) for (t, ty, v) in bcs])
    r0, dt0, _, _ = of.residual(u)
    return r0, dt0


def gpu_eval(fl, u):
    du = torch.from_numpy(u).cuda()
    res = torch.empty_like(du)
    dt = torch.empty(len(u), dtype=torch.float64, device="cuda")
    fl.compute_residual(du, res, True, dt, accumulate=False)
    torch.cuda.synchronize()
    return res.cpu().numpy(), dt.cpu().numpy()


@pytest.mark.parametrize("numerics", ["roe-wls-venkat", "hllc-gg-bj"])
def test_bench_configuration_1m_cells_against_the_oracle(numerics):
    um, arrs, u, phys, dm, fl, num, bcs = build(1.0e6, numerics)
    assert dm.info.tile_cells == 256 and dm.info.bank_conflict_groups < 0.02*dm.info.bank_groups
    r, dt = gpu_eval(fl, u)
    r0, dt0 = oracle_residual(arrs, u, phys, num, bcs)
    assert rel_err_by_component(r, r0) < TOL and np.abs(dt/dt0 - 1).max() < TOL
    # the same mesh cut into smaller tiles gives the same bits (per-cell sums in local-face order, identical face
    # expressions in both tiles of a cut face)
    dm2 = lib.DeviceMesh(um, reorder="none", tile_cells=128)
    fl2 = lib.FlowFV(dm2, phys, num[0], num[1], num[2], num[3], True, 0, bcs)
    r2, dt2 = gpu_eval(fl2, u)
    assert np.array_equal(r, r2) and np.array_equal(dt, dt2)


def test_full_size_10m_cells():
    um, arrs, u, phys, dm, fl, num, bcs = build(10.0e6, "roe-wls-venkat")
    assert um.nelem > 9.9e6
    r, dt = gpu_eval(fl, u)
    r0, dt0 = oracle_residual(arrs, u, phys, num, bcs)
    assert rel_err_by_component(r, r0) < TOL and np.abs(dt/dt0 - 1).max() < TOL
    del r0, dt0
    # run-to-run: same bits
    r1, dt1 = gpu_eval(fl, u)
    assert np.array_equal(r, r1) and np.array_equal(dt, dt1)
    del r1, dt1
    # drop-in mode (host buffers, chunked pipeline): same bits
    hr = np.empty_like(u); hdt = np.empty(len(u))
    fl.compute_residual_host(u, hr, True, hdt, accumulate=False)
    assert np.array_equal(hr, r) and np.array_equal(hdt, dt)
    # a fused forward-Euler step = residual + dt + update computed separately (same expression order), and its norm
    du = torch.from_numpy(u).cuda()
    n2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    fl.euler_step(du, 0.5, n2)
    torch.cuda.synchronize()
    area = um.arrays()["area"]
    expect = u + (0.5*dt/area)[:, None]*r
    assert np.abs(du.cpu().numpy() - expect).max() < 1e-13*np.abs(u).max()
    assert abs(float(n2.item())/float((r[:, 3]**2*area).sum()) - 1) < 1e-11


def test_full_size_freestream_preservation():
    # uniform free stream with far-field states on every boundary: each cell's fluxes are F(u_inf).n summed over a
    # closed polygon, i.e. zero to round-off whatever the mesh (10M cells, limiter active on a constant field)
    bcs = [(t, "farfield", (0.0, 0.0)) for t in (2, 3, 4)]
    um, arrs, u, phys, dm, fl, num, _ = build(10.0e6, "roe-wls-venkat", bcs=bcs)
    uinf = np.tile(lib.freestream(phys), (um.nelem, 1))
    r, dt = gpu_eval(fl, uinf)
    a = um.arrays()
    flux_scale = np.abs(orc.flux("orc", 0, phys, uinf[:1], uinf[:1], np.array([[0.6, 0.8]]))).max()*a["facemetric"][:, 2].max()
    assert np.abs(r).max() < 4e-13*flux_scale
    assert np.isfinite(dt).all() and (dt > 0).all()
