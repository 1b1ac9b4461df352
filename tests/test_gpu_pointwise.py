"""CUDA device functions (through the C-ABI pointwise hooks) against the golden vectors produced by the
reference's own object code, and against the oracle for the viscous flux. Tolerance 1e-12 relative to
the magnitude of the result vector (FP64; the device code shares reciprocals and contracts FMAs)."""
import numpy as np
import pytest
import orc
from common import golden
from fvens_b200 import lib

pytestmark = pytest.mark.gpu
TOL = 1e-12


def phys_from(g, key):
    return lib.make_physics(*g[f"phys_{key}"][:6])


def vec_rel(a, b):
    scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-300)
    return (np.abs(a-b)/scale).max()


@pytest.mark.parametrize("key", ["a", "b"])
@pytest.mark.parametrize("flux", list(lib.FLUX))
def test_flux_matches_reference_goldens(key, flux):
    g = golden()
    p = phys_from(g, key)
    ul, ur, n = g[f"ul_{key}"], g[f"ur_{key}"], g[f"n_{key}"]
    out = lib.flux_pointwise(flux, p, ul, ur, n)
    ref = g[f"flux{lib.FLUX[flux]}_{key}"]
    # scale: the analytical flux magnitudes of the two states (a numerical flux is a difference of those)
    fl = np.abs(orc.flux("orc", 0, p, ul, ul, n)).max(axis=1) + np.abs(orc.flux("orc", 0, p, ur, ur, n)).max(axis=1)
    err = (np.abs(out-ref).max(axis=1)/fl).max()
    assert err < TOL, err


@pytest.mark.parametrize("key", ["a", "b"])
@pytest.mark.parametrize("bt", [0, 1, 2, 3, 4, 6, 7])
def test_bc_matches_reference_goldens(key, bt):
    g = golden()
    p = phys_from(g, key)
    out = lib.bc_pointwise(bt, g[f"bcvals{bt}_{key}"], p, g[f"ul_{key}"], g[f"n_{key}"])
    ref = g[f"bc{bt}_{key}"]
    ok = np.isfinite(ref).all(axis=1)
    assert ok.sum() > 0.5*len(ok)
    assert vec_rel(out[ok], ref[ok]) < TOL


def test_periodic_bc_is_rejected():
    p = lib.make_physics()
    with pytest.raises(lib.FvgError) as e:
        lib.bc_pointwise(5, (0, 0), p, np.ones((1, 4)), np.array([[1.0, 0.0]]))
    assert e.value.code == 4


@pytest.mark.parametrize("order2", [True, False])
@pytest.mark.parametrize("const_visc", [True, False])
def test_viscous_flux_matches_oracle(order2, const_visc):
    rng = np.random.default_rng(11)
    n = 1500
    p = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.0, True, const_visc)

    def states():
        rho = rng.uniform(0.5, 2, n); vx = rng.uniform(-1, 1, n); vy = rng.uniform(-1, 1, n); pr = rng.uniform(1, 4, n)
        return np.stack([rho, rho*vx, rho*vy, pr/0.4 + 0.5*rho*(vx*vx+vy*vy)], axis=1)
    ucl, ucr = states(), states()
    ul = ucl*(1 + 0.01*rng.standard_normal((n, 4))); ur = ucr*(1 + 0.01*rng.standard_normal((n, 4)))
    th = rng.uniform(0, 2*np.pi, n)
    nrm = np.stack([np.cos(th), np.sin(th)], axis=1)
    rcl = rng.uniform(-1, 1, (n, 2)); rcr = rcl + rng.uniform(0.05, 0.2, (n, 1))*nrm + 0.02*rng.standard_normal((n, 2))
    gl = rng.standard_normal((n, 8)); gr = rng.standard_normal((n, 8))
    ref = orc.viscous_flux(p, order2, nrm, rcl, rcr, ucl, ucr, gl, gr, ul, ur)
    out = lib.viscous_flux_pointwise(p, order2, nrm, rcl, rcr, ucl, ucr, gl, gr, ul, ur)
    assert vec_rel(out, ref) < TOL


def test_freestream_state():
    p = lib.make_physics(1.4, 0.8, 288.15, 5000.0, 0.72, 1.25*np.pi/180)
    assert np.array_equal(lib.freestream(p), orc.freestream("orc", p))


def test_empty_batches_are_fine():
    p = lib.make_physics()
    z4 = np.zeros((0, 4)); z2 = np.zeros((0, 2))
    assert lib.flux_pointwise("ROE", p, z4, z4, z2).shape == (0, 4)
    assert lib.bc_pointwise(0, (0, 0), p, z4, z2).shape == (0, 4)
