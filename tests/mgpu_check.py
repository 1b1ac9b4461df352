"""Multi-GPU correctness check, launched by torchrun (one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_check.py
Every rank builds its subdomain of the same mesh, runs residual evaluations, a few forward-Euler steps and the
multi-GPU solver with real halo traffic, and rank 0 compares with the single-GPU engine (bitwise) and prints OK/FAIL.
MGPU_SAME_DEVICE=1: all ranks share cuda:0 (CUDA IPC works between processes on one device), gloo does the set-up.
MGPU_WITHHOLD=1: rank 0 skips one evaluation; the others must report FVG_ERR_COMM instead of a result."""
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

CASES = [
    ("inviscid", dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="VENKATAKRISHNAN", limiter_param=2.0)),
    ("inviscid", dict(flux="HLLC", gradient="GREENGAUSS", reconstruction="WENO", limiter_param=2.0)),
    ("inviscid", dict(flux="AUSM", gradient="LEASTSQUARES", reconstruction="VANALBADA")),
    ("inviscid", dict(flux="LLF", gradient="NONE", reconstruction="NONE", order2=False)),
    ("viscous", dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="NONE")),
    ("viscous", dict(flux="HLL", gradient="GREENGAUSS", reconstruction="BARTHJESPERSEN")),
    # doubly periodic box: periodic rows travel with the partition's rows (some of them from a rank to itself)
    ("periodic", dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="VENKATAKRISHNAN", limiter_param=2.0)),
]


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
    def all_agree(flag):
        """True on every rank iff `flag` is true on all of them"""
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cpu" if same else dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())
    arrs = synth.bump_channel(120, 45)
    um = lib.UMesh.from_arrays(*arrs)
    rc = synth.cell_centres(arrs[0], arrs[1], arrs[2])
    inviscid = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.0)
    viscous = lib.make_physics(1.4, 0.5, 288.15, 5000.0, 0.72, 0.0, viscous=True, const_visc=False)
    bcs_i = [(2, "slipwall", (0, 0)), (3, "inflowoutflow", (0, 0)), (4, "inflowoutflow", (0, 0))]
    bcs_v = [(2, "adiabaticwall", (0, 0)), (3, "farfield", (0, 0)), (4, "farfield", (0, 0))]
    u0 = synth.perturbed_state(rc, 1.4, 0.5)
    partition = lib.partition_rcb if os.environ.get("MGPU_PARTITION", "sfc") == "rcb" else lib.partition_sfc
    part = partition(um, world)
    base = (um, u0, part)
    parrs = synth.periodic_square(40, tri_fraction=0.3, jitter=0.15)
    pum = lib.UMesh.from_arrays(*parrs)
    pum.compute_periodic_map(3, 0); pum.compute_periodic_map(4, 1)
    prc = synth.cell_centres(parrs[0], parrs[1], parrs[2])
    pu0 = synth.perturbed_state(prc/5.0, 1.4, 0.5)          # sin(2 pi x/5) cos(3 pi y/5): periodic on [-5,5]^2
    periodic = (pum, pu0, partition(pum, world))
    bcs_p = [(3, "periodic", (0, 0)), (4, "periodic", (0, 0))]
    stream = torch.cuda.Stream(device=dev)          # graphs cannot be captured on the legacy default stream
    ok = True

    if os.environ.get("MGPU_WITHHOLD", "0") == "1":
        # a rank that stops delivering rows must surface as FVG_ERR_COMM on its neighbours, not as a wrong result
        df = DistFlow(um, part, rank, world, inviscid, dev, tile_cells=128, bcs=bcs_i, **CASES[0][1])
        n = df.ncell + df.nghost
        u = torch.zeros((n, 4), dtype=torch.float64, device=dev)
        ids = torch.from_numpy(df.global_ids.astype(np.int64)).to(dev)
        u[:df.ncell] = torch.from_numpy(u0).to(dev)[ids[:df.ncell]]
        res = torch.zeros((df.ncell, 4), dtype=torch.float64, device=dev); dt = torch.zeros(df.ncell, dtype=torch.float64, device=dev)
        df.residual(u, res, dt)
        torch.cuda.synchronize()
        df.check()
        dist.barrier()
        caught = False
        if rank != 0:
            df.residual(u, res, dt)          # rank 0 never runs this evaluation
            torch.cuda.synchronize()
            try:
                df.check()
            except lib.FvgError as e:
                caught = e.code == 7
        else:
            caught = True
        good = all_agree(caught)
        if rank == 0:
            print("MGPU_CHECK", "OK" if good else "FAIL", "withhold world", world, flush=True)
        dist.destroy_process_group()
        os._exit(0 if good else 1)       # the windows of a broken exchange are not worth an orderly teardown

    for kind, numerics in CASES:
        phys, bcs = (viscous, bcs_v) if kind == "viscous" else ((inviscid, bcs_p) if kind == "periodic" else (inviscid, bcs_i))
        if kind == "periodic" and os.environ.get("FVG_DIST", "fused") != "fused":
            continue          # only the fused engine delivers a rank's periodic rows to itself
        um, u0, part = periodic if kind == "periodic" else base
        df = DistFlow(um, part, rank, world, phys, dev, tile_cells=128, bcs=bcs, **numerics)
        ids = torch.from_numpy(df.global_ids.astype(np.int64)).to(dev)
        n = df.ncell + df.nghost
        u = torch.zeros((n, 4), dtype=torch.float64, device=dev)
        u[:df.ncell] = torch.from_numpy(u0).to(dev)[ids[:df.ncell]]
        res = torch.zeros((df.ncell, 4), dtype=torch.float64, device=dev); dt = torch.zeros(df.ncell, dtype=torch.float64, device=dev)
        hist = []
        torch.cuda.synchronize()          # the fills above ran on the default stream, the evaluations run on `stream`
        with torch.cuda.stream(stream):
            # three evaluations: the second and third replay the captured graph and must give the same bits
            df.residual(u, res, dt)
            r_first = res.clone()
            df.residual(u, res, dt)
            df.residual(u, res, dt)
            stream.synchronize()
            replay_same = torch.equal(r_first, res)
            # a few fused steps with exchanges, ping-pong buffers
            unew = torch.zeros_like(u); n2 = torch.zeros(1, dtype=torch.float64, device=dev)
            cur, nxt = u.clone(), unew
            for _ in range(5):
                df.euler_step(cur, nxt, 0.4, n2)
                t = n2.clone()
                if not df.norm_is_global:
                    stream.synchronize(); allreduce(t)
                hist.append(float(t.sqrt().item()))
                cur, nxt = nxt, cur
            # residual of the stepped state (its rows were pushed by the step epilogue), then of the initial one again
            res5 = torch.zeros_like(res)
            df.residual(cur, res5, dt)
            stream.synchronize()
        df.check()
        # the product's multi-GPU solver: 40 steps from the initial state
        hs = None
        usolve = u.clone()
        if df.engine is not None:
            code, steps, hs = df.solve_forward_euler(usolve, 0.4, 1e-30, 40, check_every=7)
            ok = ok and code == 5 and steps == 40
        # gather to rank 0
        def gather(x):
            full = torch.zeros((um.nelem, x.shape[1]), dtype=torch.float64, device=dev)
            full[ids[:df.ncell]] = x[:df.ncell]
            allreduce(full)
            return full
        full_r, full_u, full_r5, full_us = gather(res), gather(cur), gather(res5), gather(usolve)
        replay_all = all_agree(replay_same)
        if rank == 0:
            dm = lib.DeviceMesh(um, reorder="hilbert", tile_cells=128, device=lr)
            fl = lib.FlowFV(dm, phys, bcs=bcs, **numerics)
            du = torch.from_numpy(u0).to(dev); r1 = torch.zeros_like(du); d1 = torch.zeros(um.nelem, dtype=torch.float64, device=dev)
            fl.compute_residual(du, r1, True, d1, accumulate=False)
            code, steps, h1 = fl.solve_forward_euler(du, 0.4, 1e-30, 5)
            r5 = torch.zeros_like(du)
            fl.compute_residual(du, r5, True, d1, accumulate=False)
            same_r = torch.equal(full_r, r1)
            same_u = torch.equal(full_u, du)
            same_r5 = torch.equal(full_r5, r5)
            hrel = max(abs(a/b - 1) for a, b in zip(hist, h1))
            msg = (f"{numerics['flux']}+{numerics['reconstruction']}{'+viscous' if kind == 'viscous' else ''}: residual bitwise {same_r}, "
                   f"graph replay bitwise {replay_all}, state after 5 steps bitwise {same_u}, residual of it bitwise {same_r5}, "
                   f"norm history rel diff {hrel:.2e}")
            good = same_r and same_u and same_r5 and hrel < 1e-13 and replay_all
            if hs is not None:
                du2 = torch.from_numpy(u0).to(dev)
                code, steps, h40 = fl.solve_forward_euler(du2, 0.4, 1e-30, 40)
                same_us = torch.equal(full_us, du2)
                hrel40 = max(abs(a/b - 1) for a, b in zip(hs, h40))
                msg += f"; solver: state after 40 steps bitwise {same_us}, history rel diff {hrel40:.2e}"
                good = good and same_us and hrel40 < 1e-12
            print(msg)
            ok = ok and good
        del df
    ok = all_agree(ok)
    if rank == 0:
        print("MGPU_CHECK", "OK" if ok else "FAIL", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
