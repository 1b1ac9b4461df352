"""Multi-GPU correctness check, launched by torchrun (one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_check.py
Every rank builds its subdomain of the same mesh, runs residual + a few forward-Euler steps with real NCCL
halo exchanges, and rank 0 compares with the single-GPU engine (bitwise) and prints OK/FAIL."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fvens_b200 import lib, synth          # noqa: E402
from fvens_b200.dist import DistFlow       # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    # MGPU_SAME_DEVICE=1: every rank uses cuda:0 and gloo carries the set-up traffic and the scalar reductions, so
    # the whole multi-rank path (subdomain meshes, peer-memory windows over CUDA IPC, split passes) runs on ONE GPU
    same = os.environ.get("MGPU_SAME_DEVICE", "0") == "1"
    if same:
        lr = 0
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if same:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)

    def allreduce(t):
        if same:
            c = t.cpu(); dist.all_reduce(c); t.copy_(c)
        else:
            dist.all_reduce(t)
    arrs = synth.bump_channel(120, 45)
    um = lib.UMesh.from_arrays(*arrs)
    rc = synth.cell_centres(arrs[0], arrs[1], arrs[2])
    phys = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.0)
    bcs = [(2, "slipwall", (0, 0)), (3, "inflowoutflow", (0, 0)), (4, "inflowoutflow", (0, 0))]
    u0 = synth.perturbed_state(rc, 1.4, 0.5)
    part = (lib.partition_rcb if os.environ.get("MGPU_PARTITION", "sfc") == "rcb" else lib.partition_sfc)(um, world)
    ok = True
    for numerics in (dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="VENKATAKRISHNAN", limiter_param=2.0),
                     dict(flux="HLLC", gradient="GREENGAUSS", reconstruction="WENO", limiter_param=2.0)):
        df = DistFlow(um, part, rank, world, phys, dev, tile_cells=128, bcs=bcs, **numerics)
        ids = torch.from_numpy(df.global_ids.astype(np.int64)).to(dev)
        n = df.ncell + df.nghost
        u = torch.zeros((n, 4), dtype=torch.float64, device=dev)
        u[:df.ncell] = torch.from_numpy(u0).to(dev)[ids[:df.ncell]]
        res = torch.zeros((df.ncell, 4), dtype=torch.float64, device=dev); dt = torch.zeros(df.ncell, dtype=torch.float64, device=dev)
        df.residual(u, res, dt)
        # a few fused steps with exchanges, ping-pong buffers
        unew = torch.zeros_like(u); n2 = torch.zeros(1, dtype=torch.float64, device=dev)
        hist = []
        cur, nxt = u.clone(), unew
        for _ in range(5):
            df.euler_step(cur, nxt, 0.4, n2)
            t = n2.clone(); allreduce(t)
            hist.append(float(t.sqrt().item()))
            cur, nxt = nxt, cur
        # gather to rank 0
        full_r = torch.zeros((um.nelem, 4), dtype=torch.float64, device=dev); full_u = torch.zeros_like(full_r)
        full_r[ids[:df.ncell]] = res; full_u[ids[:df.ncell]] = cur[:df.ncell]
        allreduce(full_r); allreduce(full_u)
        if rank == 0:
            dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=128, device=lr)
            fl = lib.FlowFV(dm, phys, bcs=bcs, **numerics)
            du = torch.from_numpy(u0).to(dev); r1 = torch.zeros_like(du); d1 = torch.zeros(um.nelem, dtype=torch.float64, device=dev)
            fl.compute_residual(du, r1, True, d1, accumulate=False)
            code, steps, h1 = fl.solve_forward_euler(du, 0.4, 1e-30, 5)
            same_r = torch.equal(full_r, r1)
            same_u = torch.equal(full_u, du)
            hrel = max(abs(a/b - 1) for a, b in zip(hist, h1))
            print(f"{numerics['reconstruction']}: residual bitwise {same_r}, state after 5 steps bitwise {same_u}, norm history rel diff {hrel:.2e}")
            ok = ok and same_r and same_u and hrel < 1e-13
    if rank == 0:
        print("MGPU_CHECK", "OK" if ok else "FAIL", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
