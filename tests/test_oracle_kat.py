"""The reference's own known-answer tests, run against the CPU oracle:
 - tests/finite-volume/testgradientschemes.cpp:36-90  WLS gradients + linear reconstruction are
   exact for a linear field (RMS < 10 eps) on testperiodic, 2dcylinderhybrid, squareunsquad0
 - tests/flow-general/testwallbcs.cpp:14-79  zero normal mass/energy flux through walls for six fluxes
 - tests/mesh/mesh.cpp  intfac / esuel / elemface mutual consistency
"""
import numpy as np
import pytest
import orc
from common import mesh_path
from fvens_b200 import lib

EPS = np.finfo(np.float64).eps
KAT_MESHES = ["testperiodic.msh", "2dcylinderhybrid.msh", "squareunsquad0.msh"]


def linear_fields(x):
    # the reference's function 2x + 0.5y + 2.5 in variable 0, other slopes in the other variables
    c = np.array([[2.0, 0.5, 2.5], [-1.0, 3.0, 0.25], [0.75, -0.5, -1.0], [4.0, 1.5, 10.0]])
    return np.stack([c[k, 0]*x[:, 0] + c[k, 1]*x[:, 1] + c[k, 2] for k in range(4)], axis=1)


@pytest.mark.parametrize("mesh", KAT_MESHES)
def test_wls_one_exact(mesh):
    m = orc.Mesh.read(mesh_path(mesh))
    p = lib.make_physics()
    tags = sorted(set(m.arrays()["btags"].tolist()))
    fl = orc.Flow(m, p, flux=4, gradient=2, recon=0, bcs=[(t, 1, (0, 0)) for t in tags])
    rc, gr, rcbp = fl.geometry()
    u = linear_fields(rc); ug = linear_fields(rcbp)
    grad = fl.gradients(u, ug)
    ufl, ufr = fl.face_values(u, ug, grad)
    exact = linear_fields(gr)
    nb = m.nbface
    for k in range(4):
        err = np.sqrt(((ufl[:, k]-exact[:, k])**2).sum()/m.naface)
        lr = np.sqrt(((ufl[nb:, k]-ufr[nb:, k])**2).sum()/m.naface)
        scale = max(1.0, np.abs(exact[:, k]).max()/3.0)   # the reference's field is O(3)
        assert err < 10*EPS*scale and lr < 10*EPS*scale


def wall_test_state():
    return np.array([1.0, 0.5, 0.5, 10.0/(1.4-1.0) + 0.5*0.5])


@pytest.mark.parametrize("flux", [0, 2, 3, 4, 5, 6])   # LLF AUSM AUSMPLUS ROE HLL HLLC (the reference's matrix)
@pytest.mark.parametrize("wall", [7, 0])                # adiabatic wall (v_t = 0), slip wall
def test_wall_bcs_zero_flux(flux, wall):
    m = orc.Mesh.read(mesh_path("testperiodic.msh"))
    a = m.arrays()
    p = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.0, viscous=True)
    fl = orc.Flow(m, p, flux=flux, gradient=2, recon=0, bcs=[(4, 1, (0, 0)), (2, wall, (0.0, 0)), (3, 6, (0.0, 290.0))])
    ins = np.tile(wall_test_state(), (m.nbface, 1))
    gs = fl.boundary_states(ins)
    nrm = np.ascontiguousarray(a["facemetric"][: m.nbface, :2])
    f = orc.flux("orc", flux, p, ins, gs, nrm)
    sel = a["btags"] == 2
    assert sel.sum() == 8
    assert np.abs(f[sel, 0]).max() < 10*2.2e-16
    assert np.abs(f[sel, 3]).max() < (10 if wall == 7 else 100)*2.2e-16


@pytest.mark.parametrize("mesh", ["2dcylinderhybrid.msh", "testperiodic.msh", "squarecoarse.msh", "testhybrid.msh"])
def test_mesh_topology_consistency(mesh):
    m = orc.Mesh.read(mesh_path(mesh))
    a = m.arrays()
    ne, nb = m.nelem, m.nbface
    intfac, esuel, elemface, nnode, inpoel = a["intfac"], a["esuel"], a["elemface"], a["nnode"], a["inpoel"]
    assert m.naface == (nnode.sum() + nb)//2
    for f in range(m.naface):
        L, R, n0, n1 = intfac[f]
        assert 0 <= L < ne
        jl = list(elemface[L, :nnode[L]]).index(f)
        assert esuel[L, jl] == R
        # the face's nodes are consecutive nodes of the left cell, in its orientation
        assert inpoel[L, jl] == n0 and inpoel[L, (jl+1) % nnode[L]] == n1
        if f < nb:
            assert R == ne + f
        else:
            assert L < R < ne
            jr = list(elemface[R, :nnode[R]]).index(f)
            assert esuel[R, jr] == L
    # normals point out of the left cell and have unit length
    rc = np.array([a["coords"][inpoel[i, :nnode[i]]].mean(axis=0) for i in range(ne)])
    mid = 0.5*(a["coords"][intfac[:, 2]] + a["coords"][intfac[:, 3]])
    d = mid - rc[intfac[:, 0]]
    assert ((d*a["facemetric"][:, :2]).sum(axis=1) > 0).all()
    np.testing.assert_allclose(np.linalg.norm(a["facemetric"][:, :2], axis=1), 1.0, rtol=1e-14)
    assert (a["area"] > 0).all()
