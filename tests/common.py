"""Shared helpers for the test-suite: fixture paths, standard cases, comparison metrics."""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MESHDIR = os.path.join(ROOT, "tests", "golden", "meshes")


def mesh_path(name):
    return os.path.join(MESHDIR, name)


def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "tier_a.npz"))


def rel_err_by_component(a, b):
    """max |a-b| per component, normalised by that component's infinity norm in b (the flux scale;
    SURVEY section 7: entries of a residual are differences of O(1) fluxes)."""
    a = np.asarray(a); b = np.asarray(b)
    a2 = a.reshape(-1, a.shape[-1]); b2 = b.reshape(-1, b.shape[-1])
    scale = np.maximum(np.abs(b2).max(axis=0), 1e-300)
    return (np.abs(a2-b2).max(axis=0)/scale).max()


def rel_err(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return np.abs(a-b).max()/max(np.abs(b).max(), 1e-300)


# marker -> boundary condition for the fixture meshes (reference control files):
#   naca0012 / 2dcylinder: 2 = slip wall, 4 = far field (testcases/naca0012/*.ctrl)
#   visc-naca0012: 2 = adiabatic wall, 4 = far field
#   tests/flow-general/test.ctrl: 4 far field, 2 adiabatic wall, 3 isothermal wall
INVISCID_BCS = [(2, "slipwall", (0.0, 0.0)), (3, "inflowoutflow", (0.0, 0.0)), (4, "farfield", (0.0, 0.0)),
                (1, "extrapolation", (0.0, 0.0))]
VISCOUS_BCS = [(2, "adiabaticwall", (0.0, 0.0)), (3, "isothermalwall", (0.1, 1.02)), (4, "farfield", (0.0, 0.0)),
               (1, "slipwall", (0.0, 0.0))]
