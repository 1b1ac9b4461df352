#!/usr/bin/env python
"""What the locality renumbering buys: the same hybrid bump-channel mesh evaluated with
  native     the generator's row-major numbering, engine reorder = none
  hilbert    host mesh renumbered along the Hilbert curve (what bench.py does), engine reorder = none
  shuffled+h a random numbering given to the engine, which renumbers it itself (reorder = hilbert; every call
             gathers the state into device order and scatters the result back)
  shuffled+r the same with reverse Cuthill-McKee
  shuffled   a random numbering used as it is (tiles shrink until their halo fits: the no-locality floor)
usage: python tools/order_sweep.py [cells]     prints one line per variant"""
import os
import sys
import time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                           # noqa: E402
from fvens_b200 import lib, synth      # noqa: E402

cells = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0e6
nx, ny = bench.lattice_for(cells)
coords, nnode, inpoel, bface = synth.bump_channel(nx, ny, seed=12345)
phys = lib.make_physics(1.4, bench.MINF, 288.15, 5000.0, 0.72, 0.0)
rng = np.random.default_rng(7)
shuffle = rng.permutation(len(nnode))


def run(name, perm, reorder):
    nn, ip = (nnode, inpoel) if perm is None else (nnode[perm], inpoel[perm])
    um = lib.UMesh.from_arrays(coords, nn, ip, bface)
    if name == "hilbert":
        p = um.hilbert_ordering(); um.reorder_cells(p); nn, ip = nn[p], ip[p]
    t0 = time.time()
    dm = lib.DeviceMesh(um, reorder=reorder, tile_cells=256)
    tb = time.time() - t0
    fl = lib.FlowFV(dm, phys, "ROE", "LEASTSQUARES", "VENKATAKRISHNAN", 2.0, True, 0, bench.BCS)
    u = torch.from_numpy(synth.perturbed_state(synth.cell_centres(coords, nn, ip), 1.4, bench.MINF)).cuda()
    res = torch.empty_like(u); dt = torch.empty(len(u), dtype=torch.float64, device="cuda")
    for _ in range(5):
        fl.compute_residual(u, res, True, dt, accumulate=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for _ in range(n):
        fl.compute_residual(u, res, True, dt, accumulate=False)
    e1.record(); torch.cuda.synchronize()
    i = dm.info
    print(f"{name:11s} reorder={reorder:8s} cells {um.nelem} tiles {i.ntile} (mean {um.nelem/i.ntile:6.1f} cells) cut-face copies {i.ncut_dup:8d} "
          f"mean |i-j| {i.mean_neighbour_distance:10.1f}  device-mesh build {tb:5.1f} s  {e0.elapsed_time(e1)/n:7.3f} ms/eval "
          f"{um.naface/(e0.elapsed_time(e1)/n*1e-3)/1e9:6.2f} Gfaces/s", flush=True)


run("native", None, "none")
run("hilbert", None, "none")
run("shuffled+h", shuffle, "hilbert")
run("shuffled+r", shuffle, "rcm")
run("shuffled", shuffle, "none")
