"""The CUDA path held DIRECTLY against the reference's own object code: the reference's mesh reader and UMesh, its
FlowFV::compute_residual and everything under it, compiled from its unmodified sources (oracle/ref_tier_e.cpp ->
oracle/_ref/libfvens_ref_e.so, built where /root/reference exists and shipped with the tree), given the same mesh FILE
and state as fvg_residual - no oracle and no stand-in for the mesh in between.
Tolerance 1e-12 relative per component (BASELINE.json north_star). Barth-Jespersen / Venkatakrishnan: cells at a
physical boundary and their neighbours are excluded, the reference reads undefined memory there."""
import numpy as np
import pytest
import torch

import orc
from common import mesh_path, rel_err_by_component, INVISCID_BCS, VISCOUS_BCS
from fvens_b200 import lib, synth
from test_oracle_ref_c import cells_untouched_by_h1

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not orc.have_ref_e(), reason="oracle/_ref/libfvens_ref_e.so not shipped")]
TOL = 1e-12


@pytest.mark.parametrize("cfg", [
    dict(mesh="NACA0012_inv.su2", flux="ROE", gradient="LEASTSQUARES", recon="VENKATAKRISHNAN", lp=2.0, shock=True),     # config 1 numerics
    dict(mesh="naca0012luo.msh", flux="HLLC", gradient="GREENGAUSS", recon="BARTHJESPERSEN", shock=True),               # config 3 numerics
    dict(mesh="NACA0012_lam_hybrid_1.msh", flux="ROE", gradient="LEASTSQUARES", recon="NONE", viscous=True),             # config 2
    dict(mesh="2dcylinderhybrid.msh", flux="AUSM", gradient="LEASTSQUARES", recon="WENO", lp=20.0),                      # config 4 numerics
    dict(mesh="2dcylinderhybrid.msh", flux="VANLEER", gradient="GREENGAUSS", recon="VANALBADA"),
    dict(mesh="naca0012luo.msh", flux="HLL", order2=False),
    dict(mesh="NACA0012_lam_hybrid_1.msh", flux="LLF", gradient="GREENGAUSS", recon="VANALBADA", viscous=True, const_visc=True),
])
def test_cuda_residual_against_reference_object_code(cfg):
    mesh = cfg["mesh"]; flux = cfg["flux"]; gradient = cfg.get("gradient", "NONE"); recon = cfg.get("recon", "NONE")
    lp = cfg.get("lp", 1.0); order2 = cfg.get("order2", True); viscous = cfg.get("viscous", False); cv = cfg.get("const_visc", False)
    um = lib.UMesh.read(mesh_path(mesh))
    om = orc.Mesh.read(mesh_path(mesh))
    a = om.arrays()
    phys = lib.make_physics(1.4, 0.6, 288.15, 800.0, 0.72, 0.03, viscous, cv)
    tags = set(a["btags"].tolist())
    named = [b for b in (VISCOUS_BCS if viscous else INVISCID_BCS) if b[0] in tags]
    rc = synth.cell_centres(a["coords"], a["nnode"], np.where(np.arange(4)[None, :] < a["nnode"][:, None], a["inpoel"], -1))
    u = synth.perturbed_state(rc, 1.4, 0.6, 0.03, amp=0.08, shock=cfg.get("shock", False))
    rf = orc.RefCase.read(mesh_path(mesh)).flow(phys, flux, gradient if order2 else "NONE", recon if order2 else "NONE", lp, order2,
                                                [(t, lib.BC[ty], v) for (t, ty, v) in named])
    r1, dt1 = rf.residual(u)
    dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=128)
    fl = lib.FlowFV(dm, phys, flux, gradient, recon, lp, order2, 0, named)
    du = torch.from_numpy(u).cuda()
    res = torch.zeros_like(du); dt = torch.zeros(len(u), dtype=torch.float64, device="cuda")
    fl.compute_residual(du, res, True, dt, accumulate=False)
    torch.cuda.synchronize()
    r, d = res.cpu().numpy(), dt.cpu().numpy()
    keep = cells_untouched_by_h1(om, a) if recon in ("BARTHJESPERSEN", "VENKATAKRISHNAN") else np.ones(om.nelem, dtype=bool)
    assert keep.sum() > 0.5*om.nelem
    scale = np.abs(r1).max(axis=0)
    assert (np.abs(r - r1)[keep]/scale).max() < TOL and np.abs(d/dt1 - 1)[keep].max() < TOL
