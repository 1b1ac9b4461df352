"""Helpers for the GPU parity tests: build the same case in the CUDA engine (through the C ABI) and
in the CPU oracle, on the same seeded state."""
import numpy as np
import torch
import orc
from common import mesh_path, INVISCID_BCS, VISCOUS_BCS
from fvens_b200 import lib, synth

_cache = {}


def load_mesh(name):
    """name: fixture file, or 'bump:nx:ny' / 'ogrid:nt:nr' / 'square:n'. Returns (UMesh, oracle Mesh, centres)."""
    if name in _cache:
        return _cache[name]
    if ":" in name:
        kind, *args = name.split(":")
        args = [int(a) for a in args]
        if kind == "bump":
            arrs = synth.bump_channel(*args)
        elif kind == "ogrid":
            arrs = synth.ogrid_cylinder(*args, tri_fraction=0.3)
        else:
            arrs = synth.square(*args, tri_fraction=0.4, jitter=0.15)
        um = lib.UMesh.from_arrays(*arrs)
        om = orc.Mesh.from_arrays(*arrs)
    else:
        um = lib.UMesh.read(mesh_path(name))
        om = orc.Mesh.read(mesh_path(name))
    a = um.arrays()
    rc = synth.cell_centres(a["coords"], a["nnode"], np.pad(a["inpoel"], ((0, 0), (0, 4-a["inpoel"].shape[1])), constant_values=-1))
    _cache[name] = (um, om, rc)
    return _cache[name]


def make_case(mesh, flux="ROE", gradient="LEASTSQUARES", recon="NONE", order2=True, viscous=False,
              const_visc=False, limiter_param=2.0, bnd_policy=0, reorder="hilbert", tile=128,
              Minf=0.5, aoa=0.02, Reinf=5000.0, shock=False, amp=0.05):
    um, om, rc = load_mesh(mesh)
    phys = lib.make_physics(1.4, Minf, 288.15, Reinf, 0.72, aoa, viscous, const_visc)
    bcs = VISCOUS_BCS if viscous else INVISCID_BCS
    tags = set(um.arrays()["btags"][:, 0].tolist())
    bcs = [b for b in bcs if b[0] in tags]
    dm = lib.DeviceMesh(um, reorder=reorder, tile_cells=tile)
    fl = lib.FlowFV(dm, phys, flux, gradient, recon, limiter_param, order2, bnd_policy, bcs)
    of = orc.Flow(om, phys, lib.FLUX[flux], lib.GRAD[gradient], lib.RECON[recon], limiter_param, order2, bnd_policy,
                  [(t, lib.BC[ty], v) for (t, ty, v) in bcs])
    u = synth.perturbed_state(rc, 1.4, Minf, aoa, amp=amp, shock=shock)
    return fl, of, u, um


def gpu_residual(fl, u, gettimesteps=True):
    du = torch.from_numpy(u).cuda()
    res = torch.zeros_like(du)
    dt = torch.zeros(len(u), dtype=torch.float64, device="cuda")
    fl.compute_residual(du, res, gettimesteps, dt)
    torch.cuda.synchronize()
    return res.cpu().numpy(), dt.cpu().numpy()
