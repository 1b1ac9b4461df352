"""The reference-side binding of INTEGRATION.md section 2 (fvens_b200/host/reference_binding/flow_spatial_b200.hpp):
FlowFV_B200 derives from the reference's own FlowFV and overrides compute_residual with fvg_residual_host. It is
compiled against the reference's unmodified headers and sources (oracle/ref_binding.cpp -> oracle/_ref/
libfvens_ref_binding.so) and linked with libfvens_b200.so.

CPU: the library builds, needs exactly six ABI symbols, and constructing the binding without a GPU fails loudly with
the library's error (no silent fallback to the reference's CPU residual).
GPU (first run and green on a B200 in round 2, profiles/r02_pytest_gpu_formerly_unrun.log): the reference's mesh reader + the binding give the
reference's own residual to 1e-12, and the reference's own SteadyForwardEulerSolver object code, driving the CUDA
residual through the binding, reproduces its CPU run; so does the binding's device-resident driver
(ode_b200.hpp: SteadyForwardEulerSolver_B200, a SteadySolver of the reference around fvg_forward_euler_solve)."""
import subprocess

import numpy as np
import pytest

import orc
from common import mesh_path, rel_err_by_component, INVISCID_BCS, VISCOUS_BCS
from fvens_b200 import lib, synth

def _gpus():
    try:
        return lib.device_count()
    except Exception:
        return 0


pytestmark = pytest.mark.skipif(not orc.have_ref_binding(), reason="oracle/_ref/libfvens_ref_binding.so not built (needs /root/reference)")


def case(cls, mesh, viscous=False):
    rc = cls.read(mesh_path(mesh))
    phys = lib.make_physics(1.4, 0.5, 288.15, 100.0, 0.72, 0.02, viscous, False)
    tags = set(np.asarray(rc.arrays()["btags"]).reshape(rc.nbface, -1)[:, 0].tolist())
    bcs = [(t, lib.BC[ty], v) for (t, ty, v) in (VISCOUS_BCS if viscous else INVISCID_BCS) if t in tags]
    a = rc.arrays()
    inpoel = np.asarray(a["inpoel"]).reshape(rc.nelem, -1)
    cen = synth.cell_centres(a["coords"], a["nnode"], np.where(np.arange(inpoel.shape[1])[None, :] < a["nnode"][:, None], inpoel, -1))
    u = synth.perturbed_state(cen, 1.4, 0.5, 0.02, amp=0.05)
    return rc, phys, bcs, u


def test_binding_library_needs_only_the_residual_entry_points():
    out = subprocess.run(["nm", "-D", "--undefined-only", orc.REFBIND_PATH], capture_output=True, text=True).stdout
    used = sorted(ln.split()[-1] for ln in out.splitlines() if " fvg_" in ln)
    # flow_spatial_b200.hpp: create/destroy + fvg_residual_host; ode_b200.hpp: device buffer + fvg_forward_euler_solve
    assert used == ["fvg_flow_create", "fvg_flow_destroy", "fvg_forward_euler_solve", "fvg_free", "fvg_last_error", "fvg_malloc",
                    "fvg_memcpy", "fvg_mesh_create", "fvg_mesh_destroy", "fvg_residual_host"]


@pytest.mark.skipif(_gpus() > 0, reason="needs a machine WITHOUT a GPU")
def test_binding_fails_loudly_without_a_gpu():
    rc, phys, bcs, u = case(orc.RefBindingCase, "2dcylinderhybrid.msh")
    with pytest.raises(RuntimeError, match="FlowFV_B200"):
        rc.flow_b200(phys, "ROE", "LEASTSQUARES", "VANALBADA", 1.0, True, bcs)
    # the same case with the reference's own Spatial object still works in this library
    r, dt = rc.flow(phys, "ROE", "LEASTSQUARES", "VANALBADA", 1.0, True, bcs).residual(u)
    assert np.isfinite(r).all() and (dt > 0).all()
    # the device-resident driver refuses a Spatial that is not the binding's (it could only run it on the CPU)
    with pytest.raises(RuntimeError, match="needs a FlowFV_B200"):
        rc.forward_euler_b200(u, 0.4, 1e-30, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [("2dcylinderhybrid.msh", "ROE", "LEASTSQUARES", "VANALBADA", True, False),
                                 ("naca0012luo.msh", "HLLC", "GREENGAUSS", "NONE", True, False),
                                 ("NACA0012_inv.su2", "AUSM", "LEASTSQUARES", "WENO", True, False),
                                 ("2dcylinderhybrid.msh", "HLL", "NONE", "NONE", False, False),
                                 ("2dcylinderhybrid.msh", "ROE", "LEASTSQUARES", "NONE", True, True)])
def test_reference_code_on_top_of_the_cuda_residual(cfg):
    mesh, flux, gradient, recon, order2, viscous = cfg
    g, r = gradient if order2 else "NONE", recon if order2 else "NONE"
    ref, phys, bcs, u = case(orc.RefCase, mesh, viscous)
    ref.flow(phys, flux, g, r, 1.0, order2, bcs)
    gpu, _, _, _ = case(orc.RefBindingCase, mesh, viscous)
    gpu.flow_b200(phys, flux, g, r, 1.0, order2, bcs)
    r0, dt0 = ref.residual(u)
    r1, dt1 = gpu.residual(u)
    assert rel_err_by_component(r1, r0) < 1e-12 and np.abs(dt1/dt0 - 1).max() < 1e-12
    # the reference's explicit solver (its object code) stepping the GPU residual
    nsteps = 40
    c0, s0, rel0, abs0, u0 = ref.forward_euler(u, 0.4, 1e-30, nsteps)
    c1, s1, rel1, abs1, u1 = gpu.forward_euler(u, 0.4, 1e-30, nsteps)
    assert (c0, s0) == (c1, s1) == (1, nsteps)
    assert np.abs(abs1/abs0 - 1).max() < 1e-6          # the reference stores its history in single precision
    assert rel_err_by_component(u1, u0) < 1e-10
    # the binding's device-resident driver (one upload, fused steps on the GPU, one download): same contract
    c2, s2, rel2, abs2, u2 = gpu.forward_euler_b200(u, 0.4, 1e-30, nsteps)
    assert (c2, s2) == (1, nsteps)
    assert np.abs(abs2/abs0 - 1).max() < 1e-6 and np.abs(rel2/rel0 - 1).max() < 1e-6
    assert rel_err_by_component(u2, u0) < 1e-10
