"""Generates the committed fixtures under tests/golden/. Run in the build container, where
/root/reference exists:   python tests/golden/make_fixtures.py

1. meshes/: input meshes of the reference's own tests and test cases (data files, copied verbatim;
   they are inputs, not source code) - the GPU box has no /root/reference.
2. tier_a.npz: known-answer vectors produced by the REFERENCE'S OWN OBJECT CODE
   (oracle/_ref/libfvens_ref_a.so = the reference's ens_gasdynamics sources compiled unmodified):
   every inviscid flux, every boundary condition, cons<->prim and the free stream, on seeded random
   admissible states. The oracle restatement and the CUDA device functions are both pinned to these.
"""
import os
import shutil
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

MESHES = [
    "tests/common-input/testperiodic.msh", "tests/common-input/2dcylinderhybrid.msh",
    "tests/common-input/testhybrid.msh", "tests/common-input/testhybrid_part1.msh",
    "tests/common-input/testhybrid_part2.msh", "tests/common-input/testhybrid_part3.msh",
    "tests/common-input/testhybrid-distb.dat", "tests/common-input/testhybrid-distb_part1.dat",
    "tests/common-input/testhybrid-distb_part2.dat", "tests/common-input/testhybrid-distb_part3.dat",
    "tests/common-input/squarecoarse.msh", "tests/common-input/squarecoarselevels.dat",
    "tests/heat/grids/squareunsquad0.msh",
    "testcases/2dcylinder/grids/2dcylinder0.msh", "testcases/2dcylinder/grids/2dcylinder1.msh",
    "testcases/2dcylinder/grids/2dcylinder2.msh",
    "testcases/naca0012/grids/naca0012luo.msh", "testcases/naca0012/grids/NACA0012_inv.su2",
    "testcases/visc-naca0012/grids/NACA0012_lam_hybrid_1.msh",
]


def random_states(rng, n, gamma):
    rho = rng.uniform(0.3, 2.5, n)
    vx = rng.uniform(-2.0, 2.0, n)
    vy = rng.uniform(-2.0, 2.0, n)
    p = rng.uniform(0.2, 4.0, n)
    E = p/(gamma-1.0) + 0.5*rho*(vx*vx+vy*vy)
    return np.stack([rho, rho*vx, rho*vy, E], axis=1)


def main():
    import orc
    from fvens_b200 import lib
    os.makedirs(os.path.join(HERE, "meshes"), exist_ok=True)
    for m in MESHES:
        shutil.copyfile(os.path.join(REF, m), os.path.join(HERE, "meshes", os.path.basename(m)))

    rng = np.random.default_rng(20261017)
    n = 1024
    out = {}
    physs = {"a": lib.make_physics(1.4, 0.8, 288.15, 5000.0, 0.72, 1.25*np.pi/180),
             "b": lib.make_physics(1.33, 0.3, 300.0, 1.0e5, 0.7, -0.2)}
    for key, p in physs.items():
        out[f"phys_{key}"] = np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr, p.aoa])
        ul = random_states(rng, n, p.gamma)
        # half of the pairs are near-identical states (smooth flow), the rest arbitrary jumps
        ur = random_states(rng, n, p.gamma)
        ur[: n//2] = ul[: n//2]*(1.0 + 0.01*rng.standard_normal((n//2, 4)))
        th = rng.uniform(0, 2*np.pi, n)
        nrm = np.stack([np.cos(th), np.sin(th)], axis=1)
        out[f"ul_{key}"], out[f"ur_{key}"], out[f"n_{key}"] = ul, ur, nrm
        for fid in range(7):
            out[f"flux{fid}_{key}"] = orc.flux("ref", fid, p, ul, ur, nrm)
        for bt, vals in [(0, (0, 0)), (1, (0, 0)), (2, (0, 0)), (3, (1.2*1.0/(p.gamma*p.Minf**2), 1.1)), (4, (0, 0)),
                         (6, (0.3, 1.05)), (7, (0.2, 0))]:
            out[f"bc{bt}_{key}"] = orc.ghost_state("ref", bt, vals, p, ul, nrm)
            out[f"bcvals{bt}_{key}"] = np.array(vals, dtype=np.float64)
        out[f"prim_{key}"] = orc.cons2prim("ref", p, ul)
        out[f"cons_{key}"] = orc.prim2cons("ref", p, out[f"prim_{key}"])
        out[f"uinf_{key}"] = orc.freestream("ref", p)
    np.savez_compressed(os.path.join(HERE, "tier_a.npz"), **out)
    print("wrote", len(MESHES), "mesh fixtures and tier_a.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
