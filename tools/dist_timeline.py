#!/usr/bin/env python
"""Where a multi-GPU evaluation spends its time: CUDA events around every stage of DistFlow.residual (send of the state
rows, gradient/limiter pass, send of the gradient rows, face pass) on every rank, for the bench mesh.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_timeline.py [cells] [evals]
Prints, per rank, mean / min / max microseconds per stage, the sum of the stages and the measured time per evaluation
(the difference is launch gaps), plus the same evaluation timed without any event in between. Stage times include the
in-kernel wait for the neighbours when FVG_FUSED_RECV=1 (default); FVG_FUSED_RECV=0 shows the exchange kernels instead."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                              # noqa: E402
from fvens_b200 import lib                # noqa: E402
from fvens_b200.dist import DistFlow      # noqa: E402


def main():
    cells = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0e6
    nev = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    lib.load()
    um, arrs, u, _ = bench.build_case(cells, "roe-wls-venkat", 256)
    phys = lib.make_physics(1.4, bench.MINF, 288.15, 5000.0, 0.72, 0.0)
    part = lib.partition_sfc(um, world)
    df = DistFlow(um, part, rank, world, phys, dev, reorder="none", tile_cells=256, bcs=bench.BCS, flux="ROE",
                  gradient="LEASTSQUARES", reconstruction="VENKATAKRISHNAN", limiter_param=2.0, order2=True)
    ids = torch.from_numpy(df.global_ids.astype(np.int64))
    du = torch.zeros((df.ncell + df.nghost, 4), dtype=torch.float64, device=dev)
    du[:df.ncell] = torch.from_numpy(u[ids[:df.ncell].numpy()]).to(dev)
    res = torch.empty((df.ncell, 4), dtype=torch.float64, device=dev)
    dtm = torch.empty(df.ncell, dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    fl, win = df.flow, getattr(df.halo, "win", None)
    fused = df.fused_recv

    def stages():
        """The stage sequence of DistFlow.residual, as closures"""
        if fused:
            return [("send u", lambda: fl.ghost_source(0, win, win.post(du, 4, stream=s))),
                    ("gradient pass", lambda: fl.gradient_pass(du, 0, stream=s)),
                    ("send gradients", lambda: fl.ghost_source(1, win, win.post(df.lg, 8, stream=s))),
                    ("face pass", lambda: fl.face_pass(du, res, True, dtm, accumulate=False, stream=s))]
        return [("exchange u", lambda: df.halo.exchange(du)),
                ("gradient pass", lambda: fl.gradient_pass(du, 0, stream=s)),
                ("exchange gradients", lambda: df.halo.exchange(df.lg)),
                ("face pass", lambda: fl.face_pass(du, res, True, dtm, accumulate=False, stream=s))]

    st = stages()
    for _ in range(5):
        df.residual(du, res, dtm)
    dist.barrier(); torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(st) + 1)] for _ in range(nev)]
    for k in range(nev):
        ev[k][0].record()
        for j, (_, f) in enumerate(st):
            f()
            ev[k][j+1].record()
    if fused:
        fl.ghost_source(0); fl.ghost_source(1)
    dist.barrier(); torch.cuda.synchronize()
    t = np.array([[ev[k][j].elapsed_time(ev[k][j+1])*1e3 for j in range(len(st))] for k in range(nev)])
    total = ev[0][0].elapsed_time(ev[-1][-1])*1e3/nev
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(nev):
        df.residual(du, res, dtm)
    e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1)*1e3/nev
    lines = [f"rank {rank}: own cells {df.ncell} ghosts {df.nghost} tiles {df.dmesh.info.ntile} "
             f"{'in-kernel receive' if fused else 'exchange kernels'}"]
    for j, (name, _) in enumerate(st):
        lines.append(f"    {name:20s} mean {t[:, j].mean():8.1f} us  min {t[:, j].min():8.1f}  max {t[:, j].max():8.1f}")
    lines.append(f"    sum of stages {t.sum(axis=1).mean():8.1f} us, per evaluation with events {total:8.1f} us, without {plain:8.1f} us")
    out = [None]*world
    dist.all_gather_object(out, "\n".join(lines))
    if rank == 0:
        print("\n".join(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
