import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from test_gpu_partition import run_partitioned, CASES
from gpu_common import load_mesh
from common import INVISCID_BCS
from fvens_b200 import lib, synth
um, om, rc = load_mesh("bump:40:15")
phys = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.02)
bcs = [b for b in INVISCID_BCS if b[0] in (2, 3, 4)]
u = synth.perturbed_state(rc, 1.4, 0.5, 0.02)
for case in (0,4):
    numerics = CASES[case]
    res, dt = run_partitioned(um, phys, bcs, 2, u, numerics)
    outs = {}
    for tile in (64, 32, 256):
        dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=tile)
        fl = lib.FlowFV(dm, phys, bcs=bcs, **numerics)
        du = torch.from_numpy(u).cuda(); r1 = torch.zeros_like(du); d1 = torch.zeros(um.nelem, dtype=torch.float64, device="cuda")
        fl.compute_residual(du, r1, True, d1, accumulate=False); torch.cuda.synchronize()
        outs[tile] = r1.cpu().numpy()
    part = lib.partition_sfc(um, 2)
    d = np.abs(res - outs[64])
    bad = np.nonzero(d.max(axis=1) > 0)[0]
    print("case", case, "max diff part vs single", d.max(), "ncells differing", len(bad), "scale", np.abs(res).max())
    print(" tile64 vs tile32 diff", np.abs(outs[64]-outs[32]).max(), (np.abs(outs[64]-outs[32]).max(axis=1)>0).sum(), " vs 256:", np.abs(outs[64]-outs[256]).max())
    a = um.arrays()
    # are differing cells adjacent to the partition boundary?
    nb = um.nbface
    cut = set()
    for L,R in a["intfac"][nb:,:2]:
        if part[L]!=part[R]: cut.add(L); cut.add(R)
    print(" differing cells on the cut:", sum(1 for c in bad if c in cut), "of", len(bad), "; cut cells", len(cut))
