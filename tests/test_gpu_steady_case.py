"""The fvens_steady executable (fvens_b200/host/fvens_steady.cpp: control file -> mesh -> first-order starter -> main
explicit solve -> residual history, surface files, VTU) end to end on the GPU, against the same run driven through the
Python mirror of the class surface: Cl, Cd, entropy norm, history, file formats."""
import math
import os
import re
import subprocess
import xml.dom.minidom

import numpy as np
import pytest

from common import ROOT, MESHDIR

CTRL = os.path.join(ROOT, "tests", "golden", "ctrl")
STEADY = os.path.join(ROOT, "tests", "cpp", "fvens_steady")


@pytest.mark.gpu
def test_fvens_steady_end_to_end(tmp_path):
    import torch
    from fvens_b200 import lib
    nsteps = 40
    mesh = os.path.join(MESHDIR, "2dcylinder1.msh")
    prefix = str(tmp_path / "cyl")
    # the control file's output names are relative: run in the temporary directory
    r = subprocess.run([STEADY, os.path.join(CTRL, "expl-cyl-ls-hllc.ctrl"), "--source_dir", CTRL, "--mesh_file", mesh,
                        "--log_file_prefix", prefix, "--max_timesteps", str(nsteps)], capture_output=True, text=True, timeout=600,
                       cwd=str(tmp_path))
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "--------------- End" in r.stdout
    m = re.search(r"Functionals: h (\S+) entropy (\S+) CL (\S+) CDp (\S+) CDf (\S+)", r.stdout)
    h, entropy, cl, cdp, cdf = (float(x) for x in m.groups())

    # the same case through the Python mirror: first-order starter, then the main solve, both capped at nsteps
    um = lib.UMesh.read(mesh)
    phys = lib.make_physics(1.4, 0.38, 298.0, float("inf"), float("nan"), 0.0)
    bcs = [(2, "slipwall", (0, 0)), (4, "farfield", (0, 0))]
    dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=256)
    u = torch.from_numpy(np.tile(lib.freestream(phys), (um.nelem, 1))).cuda()
    f1 = lib.FlowFV(dm, phys, "HLLC", "NONE", "NONE", 1.0, False, 0, bcs)
    f1.solve_forward_euler(u, 0.5, 1e-1, nsteps)
    f2 = lib.FlowFV(dm, phys, "HLLC", "LEASTSQUARES", "NONE", 1.0, True, 0, bcs)
    code, steps, hist = f2.solve_forward_euler(u, 0.25, 1e-4, nsteps)
    g = torch.zeros(um.nelem, 8, dtype=torch.float64, device="cuda")
    f2.getGradients(u, g)
    cl0, cdp0, cdf0 = f2.computeSurfaceData(u, g, 2)
    assert abs(cl - cl0) < 1e-12 + 1e-10*abs(cl0) and abs(cdp - cdp0) < 1e-12 + 1e-10*abs(cdp0) and cdf == 0.0 == cdf0
    assert abs(entropy/f2.entropy_error(u) - 1) < 1e-10 and abs(h - 1.0/math.sqrt(um.nelem)) < 1e-12

    # residual history in the reference's format (spatial/aoutput.cpp:617-636): header, rule, one line per step
    lines = open(prefix + "-residual_history.log").read().splitlines()
    assert lines[0].split() == ["#", "NStep", "Log", "rel", "resi", "Log", "abs", "resi", "Tot.Wtime", "Lin.Wtime", "Lin.iters", "CFL"]
    assert lines[1].startswith("#---") and len(lines) == 2 + steps
    rows = np.array([[float(x) for x in ln.split()] for ln in lines[2:]])
    assert (rows[:, 0] == np.arange(1, steps+1)).all() and rows[0, 1] == 0.0 and (rows[:, 6] == 0.25).all()
    assert np.abs(rows[:, 2] - np.log10(hist)).max() < 1e-5          # printed with 6 significant digits, norm stored as float
    # surface file: one line per wall face + the coefficient line; VTU: well-formed, one value per point
    surf = open(os.path.join(str(tmp_path), "2dcyl-surf_w2.out")).read().splitlines()
    a = um.arrays()
    nwall = int((a["btags"][:, 0] == 2).sum())
    assert surf[0].startswith("#  x") and len(surf) == nwall + 3 and surf[-2].startswith("# Cl")
    assert abs(float(surf[-1].split()[1]) - cl) < 1e-5*max(abs(cl), 1e-3)
    pts = np.array([[float(x) for x in ln.split()] for ln in surf[1:1+nwall]])
    assert np.abs(np.hypot(pts[:, 0], pts[:, 1]) - 0.5).max() < 0.02      # face midpoints of the cylinder of radius 0.5
    doc = xml.dom.minidom.parse(os.path.join(str(tmp_path), "2dcyl.vtu"))
    piece = doc.getElementsByTagName("Piece")[0]
    assert int(piece.getAttribute("NumberOfPoints")) == a["coords"].shape[0] and int(piece.getAttribute("NumberOfCells")) == um.nelem
    names = [d.getAttribute("Name") for d in doc.getElementsByTagName("PointData")[0].getElementsByTagName("DataArray")]
    assert names == ["density", "mach-number", "pressure", "temperature", "velocity"]
    dens = np.array(doc.getElementsByTagName("PointData")[0].getElementsByTagName("DataArray")[0].firstChild.data.split(), dtype=float)
    assert len(dens) == a["coords"].shape[0] and (dens > 0.5).all() and (dens < 1.5).all()
