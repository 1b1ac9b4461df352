"""Tier B: the oracle's restatement of the gradient schemes and reconstructions (oracle/orc_spatial.hpp) against the
REFERENCE'S OWN OBJECT CODE for them - spatial/agradientschemes.cpp, areconstruction.cpp,
limitedlinearreconstruction.cpp and musclreconstruction.cpp compiled unmodified, in place, against the stand-ins of
oracle/ref_shim_b (oracle/ref_tier_b.cpp -> oracle/_ref/libfvens_ref_b.so). This pins Green-Gauss, weighted least
squares, linear, WENO, Van Albada MUSCL, Barth-Jespersen and Venkatakrishnan to the reference itself, including the
limiters that no reference test or control file exercises. At boundary cells the reference's limiters
read one row past the cell states per boundary face (H1); the harness hands them a matrix that has those rows,
filled with the boundary ghost states, which is the oracle's (and the kernels') bnd_policy 0."""
import numpy as np
import pytest

import orc
from common import mesh_path, INVISCID_BCS
from fvens_b200 import lib, synth

pytestmark = pytest.mark.skipif(not orc.have_ref_b(), reason="oracle/_ref/libfvens_ref_b.so not built (needs /root/reference)")
TOL = 1e-13


def case(mesh, shock=False):
    if mesh.startswith("bump"):
        om = orc.Mesh.from_arrays(*synth.bump_channel(36, 14))
    else:
        om = orc.Mesh.read(mesh_path(mesh))
    a = om.arrays()
    phys = lib.make_physics(1.4, 0.6, 288.15, 5000.0, 0.72, 0.03)
    tags = set(a["btags"].tolist())
    bcs = [(t, lib.BC[ty], v) for (t, ty, v) in INVISCID_BCS if t in tags]
    of = orc.Flow(om, phys, 4, 2, 0, 1.0, True, 0, bcs)
    rc, gr, rcbp = of.geometry()
    u = synth.perturbed_state(rc, 1.4, 0.6, 0.03, amp=0.08, shock=shock)
    up = orc.cons2prim("orc", phys, u)
    ug = orc.cons2prim("orc", phys, of.boundary_states(u[a["intfac"][:om.nbface, 0]]))
    return om, a, phys, bcs, rc, gr, rcbp, up, ug


def relerr(x, y):
    return np.abs(x - y).max()/max(np.abs(y).max(), 1e-300)


@pytest.mark.parametrize("mesh", ["2dcylinderhybrid.msh", "naca0012luo.msh", "NACA0012_lam_hybrid_1.msh", "bump"])
@pytest.mark.parametrize("gradient", ["GREENGAUSS", "LEASTSQUARES", "NONE"])
def test_gradient_schemes_against_reference_object_code(mesh, gradient):
    om, a, phys, bcs, rc, gr, rcbp, up, ug = case(mesh)
    of = orc.Flow(om, phys, 4, lib.GRAD[gradient], 0, 1.0, True, 0, bcs)
    g_orc = of.gradients(up, ug)
    g_ref = orc.ref_gradients(a, lib.GRAD[gradient], rc, rcbp, up, ug)
    assert np.isfinite(g_ref).all()
    if gradient == "NONE":
        assert not g_ref.any() and not g_orc.any()
    else:
        assert np.abs(g_ref).max() > 1e-3 and relerr(g_orc, g_ref) < TOL


@pytest.mark.parametrize("mesh,shock", [("2dcylinderhybrid.msh", False), ("naca0012luo.msh", True), ("bump", True)])
@pytest.mark.parametrize("recon,param", [("NONE", 0.0), ("WENO", 1.0), ("WENO", 20.0), ("VANALBADA", 0.0), ("BARTHJESPERSEN", 0.0),
                                         ("VENKATAKRISHNAN", 0.5), ("VENKATAKRISHNAN", 6.0)])
@pytest.mark.parametrize("gradient", ["LEASTSQUARES", "GREENGAUSS"])
def test_reconstructions_against_reference_object_code(mesh, shock, recon, param, gradient):
    om, a, phys, bcs, rc, gr, rcbp, up, ug = case(mesh, shock)
    grad = orc.ref_gradients(a, lib.GRAD[gradient], rc, rcbp, up, ug)
    of = orc.Flow(om, phys, 4, lib.GRAD[gradient], lib.RECON[recon], param, True, 0, bcs)
    ufl, ufr = of.face_values(up, ug, grad)
    rl, rr = orc.ref_face_values(a, lib.RECON[recon], param, rc, rcbp, gr, up, ug, grad)
    nb = om.nbface
    # the reference writes the left value of every face and the right value of every interior face
    assert np.isfinite(rl).all() and np.isfinite(rr[nb:]).all() and np.isnan(rr[:nb]).all()
    assert relerr(ufl, rl) < TOL and relerr(ufr[nb:], rr[nb:]) < TOL
    if recon in ("BARTHJESPERSEN", "VENKATAKRISHNAN", "VANALBADA") and shock:
        # the limiter is active somewhere: the limited values differ from the plain linear extrapolation
        ll, _ = orc.ref_face_values(a, 0, 0.0, rc, rcbp, gr, up, ug, grad)
        assert np.abs(ll - rl).max() > 1e-6
