"""Debug helper: two or three ranks on one GPU, fused engine, prints where the residual differs from the single-mesh one."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fvens_b200 import lib, synth
from fvens_b200.dist import DistFlow

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
dist.init_process_group("gloo")
arrs = synth.bump_channel(120, 45)
um = lib.UMesh.from_arrays(*arrs)
rc = synth.cell_centres(arrs[0], arrs[1], arrs[2])
phys = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.0)
bcs = [(2, "slipwall", (0, 0)), (3, "inflowoutflow", (0, 0)), (4, "inflowoutflow", (0, 0))]
u0 = synth.perturbed_state(rc, 1.4, 0.5)
part = lib.partition_sfc(um, world)
order2 = os.environ.get("DBG_ORDER2", "1") == "1"
num = dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="VENKATAKRISHNAN", limiter_param=2.0, order2=order2)
df = DistFlow(um, part, rank, world, phys, dev, tile_cells=128, bcs=bcs, **num)
ids = torch.from_numpy(df.global_ids.astype(np.int64)).to(dev)
n = df.ncell + df.nghost
u = torch.zeros((n, 4), dtype=torch.float64, device=dev)
u[:df.ncell] = torch.from_numpy(u0).to(dev)[ids[:df.ncell]]
res = torch.zeros((df.ncell, 4), dtype=torch.float64, device=dev); dt = torch.zeros(df.ncell, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    for rep in range(2):
        df.residual(u, res, dt)
        stream.synchronize()
        print(f"[{rank}] rep {rep} counters {df.engine.counters()} nan rows {int(torch.isnan(res).any(dim=1).sum())} of {df.ncell}, "
              f"zero rows {int((res == 0).all(dim=1).sum())}, ghost {df.nghost}", flush=True)
import ctypes as C
rows = np.zeros((max(df.nghost, 1), 4)); flags = np.zeros(48, dtype=np.uint64)
for par in (0, 1):
    lib.check(lib.load().fvg_dist_debug_window(df.engine._h, 0, par, rows.ctypes.data_as(C.POINTER(C.c_double)), flags.ctypes.data_as(C.POINTER(C.c_ulonglong))))
    expect = u0[df.global_ids[df.ncell:]]
    print(f"[{rank}] U area parity {par}: zero rows {int((rows == 0).all(axis=1).sum())} of {df.nghost}, matching rows {int((rows == expect).all(axis=1).sum())}; flags U {flags[:world]} GU {flags[16:16+world]} LG {flags[32:32+world]}", flush=True)
df.check()
full = torch.zeros((um.nelem, 4), dtype=torch.float64, device=dev); full[ids[:df.ncell]] = res
c = full.cpu(); dist.all_reduce(c)
if rank == 0:
    dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=128, device=0)
    fl = lib.FlowFV(dm, phys, bcs=bcs, **num)
    du = torch.from_numpy(u0).to(dev); r1 = torch.zeros_like(du); d1 = torch.zeros(um.nelem, dtype=torch.float64, device=dev)
    fl.compute_residual(du, r1, True, d1, accumulate=False)
    r1 = r1.cpu()
    bad = (c != r1).any(dim=1)
    print("differing cells", int(bad.sum()), "of", um.nelem, "max abs diff", float(torch.nan_to_num(c - r1, nan=1e300).abs().max()))
    # are the differing cells next to the partition boundary?
    pr = torch.from_numpy(part)
    print("differing cells per owner", [int((bad & (pr == r)).sum()) for r in range(world)])
dist.destroy_process_group()
