#!/usr/bin/env python
"""Benchmark of the residual hot path (BASELINE.json metric: residual evals/s and Gfaces/s on a
10M-cell hybrid mesh; % of HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one evaluation of FlowFV::compute_residual(u, r, gettimesteps=true, dt) on the synthetic
10M-cell hybrid tri/quad Gaussian-bump channel with the north-star headline numerics
Roe + weighted least squares + Venkatakrishnan. `value` is device-resident throughput; `e2e` goes
through the host-buffer C-ABI entry (H2D of u, D2H of r and dt inside the timed region).
The reference arm (--impl reference) times the reference's own compute_residual (oracle/_ref: its sources compiled
unmodified, OpenMP, all host threads; the oracle's restatement if that build is absent) on a bounded sample of the same
mesh family.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NUMERICS = {
    # name: (flux, gradient, reconstruction, limiter_param, per-cell algorithmic bytes of pass A, of pass B)
    "roe-wls-venkat": ("ROE", "LEASTSQUARES", "VENKATAKRISHNAN", 2.0, 144 + 32 + 8, 160),
    "hllc-gg-bj": ("HLLC", "GREENGAUSS", "BARTHJESPERSEN", 0.0, 144 + 8, 160),
}
FLUXES = ["LLF", "VANLEER", "AUSM", "HLL", "HLLC", "ROE"]
# the default workload's free-stream Mach number and boundary conditions (also used by tests/test_gpu_scale.py and tools/)
MINF = 0.2
BCS = [(2, "slipwall", (0.0, 0.0)), (3, "inflowoutflow", (0.0, 0.0)), (4, "inflowoutflow", (0.0, 0.0))]


def lattice_for(cells, aspect=2.6667, split=1.0/3.0):
    """Base lattice nx x ny (nx/ny = aspect) whose mesh has ~`cells` cells when a fraction `split` of the quads is cut in two."""
    nbase = cells/(1.0 + split)
    ny = int(round((nbase/aspect)**0.5))
    nx = int(round(aspect*ny))
    return nx, ny


class Workload:
    """One benchmark configuration: mesh arrays (numpy only, Hilbert-ordered), state, physics, boundary conditions,
    numerics and the algorithmic bytes of SURVEY 8(d). Shared by both arms and by the cpu_baseline leg."""

    def __init__(self, args, world=1):
        self.name = args.workload
        self.args = args
        w = args.workload
        self.periodic = []
        self.viscous = False
        self.reference_can_run = True
        if w == "bump":          # BASELINE configs[2] mesh with the north-star headline numerics (or --numerics hllc-gg-bj)
            flux, grad, recon, lp, cA, cB = NUMERICS[args.numerics]
            self.minf, self.bcs = MINF, BCS
            self.num = dict(flux=flux, gradient=grad, reconstruction=recon, limiter_param=lp, order2=True)
            self.cA, self.cB, self.cX = cA, cB, 0
            self.what = (f"synthetic hybrid tri/quad Gaussian-bump channel, {args.cells/1e6:g}M cells, {flux}+{grad}+{recon} second-order "
                         f"residual with local time steps (BASELINE configs[2] mesh, north-star headline numerics)")
            self.tag = args.numerics
        elif w == "viscous":     # BASELINE configs[1] numerics at benchmark size: laminar Navier-Stokes, Roe + WLS + MAG viscous flux
            self.minf, self.viscous = 0.5, True
            self.bcs = [(2, "adiabaticwall", (0.0, 0.0)), (3, "farfield", (0.0, 0.0)), (4, "farfield", (0.0, 0.0))]
            self.num = dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="NONE", limiter_param=1.0, order2=True)
            self.cA, self.cB, self.cX = 144 + 32, 160, 0
            self.what = (f"synthetic hybrid tri/quad bump channel with no-slip adiabatic walls, {args.cells/1e6:g}M cells, laminar "
                         f"Navier-Stokes (Re 5000, Sutherland), ROE + LEASTSQUARES + modified-average-gradient viscous flux, second order "
                         f"(BASELINE configs[1] numerics at benchmark size)")
            self.tag = "roe-wls-none-viscous"
        elif w == "ogrid-weno":  # BASELINE configs[3]: O-grid around a cylinder, quads, WLS + WENO, one of the six fluxes
            self.minf = 0.38
            self.bcs = [(2, "slipwall", (0.0, 0.0)), (4, "farfield", (0.0, 0.0))]
            self.num = dict(flux=args.flux.upper(), gradient="LEASTSQUARES", reconstruction="WENO", limiter_param=args.weno_lambda, order2=True)
            self.cA, self.cB, self.cX = 144 + 32 + 0, 160, 192
            self.what = (f"synthetic O-grid around a cylinder (2dcylstruct.geo scaled up), quads, {args.cells/1e6:g}M cells, "
                         f"{args.flux.upper()} + LEASTSQUARES + WENO (lambda = {args.weno_lambda:g}) second-order residual with local time steps "
                         f"(BASELINE configs[3]; 50M cells at 8 GPUs is {50/8:g}M per GPU)")
            self.tag = f"{args.flux.lower()}-wls-weno"
        elif w == "vortex":      # BASELINE configs[4]: doubly periodic box of the isentropic vortex, n x n quads
            self.minf = 0.5
            self.bcs = [(3, "periodic", (0.0, 0.0)), (4, "periodic", (0.0, 0.0))]
            self.periodic = [(3, 0), (4, 1)]
            self.num = dict(flux="ROE", gradient="LEASTSQUARES", reconstruction="NONE", limiter_param=1.0, order2=True)
            self.cA, self.cB, self.cX = 144 + 32, 160, 0
            self.n = int(args.n if args.n else round(args.cells**0.5))
            if args.scaling == "weak":
                self.n = int(round(self.n*world**0.5))
            self.what = (f"isentropic-vortex box [-5,5]^2, {self.n} x {self.n} quads, periodic in x and y, ROE + LEASTSQUARES + linear "
                         f"reconstruction, second-order residual with local time steps (BASELINE configs[4]; "
                         f"{'weak' if args.scaling == 'weak' else 'strong'} scaling of the halo exchange)")
            self.tag = "roe-wls-none-periodic"
            self.reference_can_run = False       # no periodic FlowBC in the reference
        else:
            raise SystemExit(f"unknown workload {w}")

    def physics(self):
        from fvens_b200 import lib
        return lib.make_physics(1.4, self.minf, 288.15, 5000.0, 0.72, 0.0, self.viscous, False)

    def arrays(self):
        """(coords, nnode, inpoel, bface), state, lattice - plain numpy, cells renumbered along a Hilbert curve (the
        reference's `-mesh_reorder` step with a locality order; synth.hilbert_order is the numpy twin of
        fvg_umesh_hilbert_ordering). Both arms are built from this, so they time the same mesh in the same numbering."""
        from fvens_b200 import synth
        a = self.args
        if self.name in ("bump", "viscous"):
            lat = lattice_for(a.cells)
            coords, nnode, inpoel, bface = synth.bump_channel(lat[0], lat[1], seed=12345)
        elif self.name == "ogrid-weno":
            lat = lattice_for(a.cells, aspect=12500.0/4000.0, split=0.0)
            coords, nnode, inpoel, bface = synth.ogrid_cylinder(lat[0], lat[1])
        else:
            lat = (self.n, self.n)
            coords, nnode, inpoel, bface = synth.periodic_square(self.n)
        rc = synth.cell_centres(coords, nnode, inpoel)
        if getattr(a, "cell_order", "hilbert") == "hilbert":
            perm = synth.hilbert_order(rc)
            nnode, inpoel, rc = np.ascontiguousarray(nnode[perm]), np.ascontiguousarray(inpoel[perm]), rc[perm]
        if self.name == "vortex":
            u = synth.isentropic_vortex(rc, 1.4, self.minf, 0.0)
        else:
            u = synth.perturbed_state(rc, 1.4, self.minf)
        return (coords, nnode, inpoel, bface), np.ascontiguousarray(u), lat

    def host_mesh(self, arrs):
        from fvens_b200 import lib
        um = lib.UMesh.from_arrays(*arrs)
        for marker, axis in self.periodic:
            um.compute_periodic_map(marker, axis)
        return um

    def algorithmic_bytes(self, nc, nf):
        """pass A = c_A*N_c + 16*N_f (face midpoints), pass B = 160*N_c + 48*N_f, WENO pass 192*N_c."""
        return self.cA*nc + 16*nf + self.cX*nc, self.cB*nc + 48*nf

    def config(self, nc, nf, lat):
        """`config` of the JSON line: the workload only, identical in both arms (the GPU arm's layout goes under `layout`)."""
        return {"workload": self.what, "cells": int(nc), "faces": int(nf), "lattice": [int(lat[0]), int(lat[1])],
                "numerics": self.tag,
                "cell_order": "Hilbert curve (host renumbering, identical in both arms)" if getattr(self.args, "cell_order", "hilbert") == "hilbert"
                else "the generator's row-major numbering as the caller's; the engine renumbers internally (reorder = hilbert) and "
                     "gathers / scatters through the permutation inside its two kernels (the drop-in path of FlowFV_B200)",
                "l2": "inputs (32 B state + 64 B gradients + ~300 B mesh per cell: 4 GB at 10M cells) exceed the 126 MB L2 from "
                      "0.4M cells per GPU up; no explicit flush"}


def build_case(cells, numerics, tile):
    """Host mesh of the CUDA arm for the default workload (used by tests/test_gpu_scale.py and tools/)."""
    ns = argparse.Namespace(workload="bump", numerics=numerics, cells=cells)
    wl = Workload(ns)
    arrs, u, lat = wl.arrays()
    return wl.host_mesh(arrs), arrs, u, lat


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(",")])

    def count(self, t0=None):
        return sum(1 for r in self.rows if len(r) >= 10 and (t0 is None or r[0] >= t0))

    def stop(self, windows):
        """Summary over the samples that arrived inside one of the (t0, t1) load windows."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r[1:] for r in self.rows if len(r) >= 10 and any(a <= r[0] <= b for a, b in windows)]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for k, nm in enumerate(names):
                if r[5+k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(wl, steps, warmup, extras=False, prebuilt=None):
    """The reference's CPU path on the host cores, on the SAME mesh as the CUDA arm. When
    oracle/_ref/libfvens_ref_c_omp.so exists (the reference's own flow_spatial.cpp and everything it calls, compiled
    unmodified with its OpenMP pragmas on - oracle/ref_tier_c.cpp) that is what is timed (kind "reference"); otherwise
    the oracle's restatement of the same loops (kind "port"). Only numpy, tests/orc.py and oracle/ are used here.
    Returns (Gfaces/s, ms/step, info)."""
    # all host cores for the CPU arm, also under torchrun (which sets OMP_NUM_THREADS=1 for its workers unless the caller
    # has set it): the OpenMP runtime of the oracle libraries reads the variable when it is first loaded, i.e. below
    if os.environ.get("OMP_NUM_THREADS", "1") == "1" and "FVG_CPU_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    elif "FVG_CPU_THREADS" in os.environ:
        os.environ["OMP_NUM_THREADS"] = os.environ["FVG_CPU_THREADS"]
    os.environ.setdefault("OMP_PROC_BIND", "close")
    import orc
    from fvens_b200 import lib      # constants only (name tables, the physics struct); the shared library is not loaded
    arrs, u, (nx, ny) = prebuilt if prebuilt is not None else wl.arrays()
    om = orc.Mesh.from_arrays(*arrs)
    flux, grad, recon, lp = wl.num["flux"], wl.num["gradient"], wl.num["reconstruction"], wl.num["limiter_param"]
    phys = wl.physics()
    bcs = [(t, lib.BC[ty], v) for (t, ty, v) in wl.bcs]
    build = "-O3 -msse4.2 (its default release flags)"
    extra = {}

    def once(flow):
        t = time.perf_counter()
        flow.residual(u, True, want=False)
        return time.perf_counter() - t
    if orc.have_ref_c_omp():
        rf = orc.RefFlow(om.arrays(), phys, flux, grad, recon, lp, True, bcs, omp=True)
        kind, cores = "reference", rf.threads()
        once(rf)
        ta = [once(rf) for _ in range(2)]
        extra["ms_per_eval_sse42_build"] = min(ta)*1e3
        if os.path.exists(orc.REFC_OMP_AVX2_PATH) and orc.host_has_avx2_fma():
            # the reference's -DAVX_2 build option: keep whichever build is faster on this host (best of two single
            # evaluations each after one untimed evaluation)
            rf2 = orc.RefFlow(om.arrays(), phys, flux, grad, recon, lp, True, bcs, omp=True, path=orc.REFC_OMP_AVX2_PATH)
            once(rf2)
            tb = [once(rf2) for _ in range(2)]
            extra["ms_per_eval_avx2_build"] = min(tb)*1e3
            if min(tb) < min(ta):
                rf, build = rf2, "-O3 -mavx2 -mfma (its AVX_2 build option)"
            else:
                del rf2
        if extras:
            # one evaluation on a single thread (the libraries share one OpenMP runtime; orc_set_num_threads sets its ICV)
            orc.set_threads(1)
            extra["ms_per_eval_1_thread"] = once(rf)*1e3
            orc.set_threads(cores)

        def evaluate():
            rf.residual(u, True, want=False)
    else:
        orc.set_threads(os.cpu_count() or 1)
        of = orc.Flow(om, phys, lib.FLUX[flux], lib.GRAD[grad], lib.RECON[recon], lp, True, 0, bcs)
        kind, cores = "port", orc.num_threads()

        def evaluate():
            of.residual(u)
    for _ in range(warmup):
        evaluate()
    t0 = time.perf_counter()
    for _ in range(steps):
        evaluate()
    dt = (time.perf_counter() - t0)/steps
    what = ("the reference's own FlowFV::compute_residual (flow_spatial.cpp and its callees compiled unmodified, OpenMP, " + build + ")"
            if kind == "reference" else "the oracle's restatement of the reference's loops (OpenMP)")
    info = {"cells": om.nelem, "faces": om.naface, "cores": cores, "kind": kind, "lattice": (nx, ny), "extra": extra,
            "sample": f"{what} on the whole workload mesh: {nx}x{ny} base lattice = {om.nelem} cells / {om.naface} "
                      f"faces, Hilbert-ordered (same arrays, numbering, state and numerics as the CUDA arm), {steps} evaluations "
                      f"after {warmup} warm-up, {cores} threads"}
    return om.naface/dt/1e9, dt*1e3, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=float, default=10.0e6)
    ap.add_argument("--workload", default="bump", choices=["bump", "viscous", "ogrid-weno", "vortex"],
                    help="bump: the headline (BASELINE configs[2] mesh, north-star numerics); viscous: configs[1] numerics at size; "
                         "ogrid-weno: configs[3]; vortex: configs[4] periodic box")
    ap.add_argument("--numerics", default="roe-wls-venkat", choices=list(NUMERICS), help="bump workload only")
    ap.add_argument("--flux", default="roe", choices=[f.lower() for f in FLUXES], help="ogrid-weno workload: the inviscid flux")
    ap.add_argument("--weno-lambda", type=float, default=1.0, help="ogrid-weno workload: central weight of the WENO average (1 or 20)")
    ap.add_argument("--vortex-n", dest="n", type=int, default=0, help="vortex workload: cells per side (default sqrt(--cells))")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="vortex workload with N > 1: fixed mesh, or n*sqrt(N)")
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--cpu-cells", type=float, default=0.0,
                    help="cells of the CPU arms' mesh; 0 (default) = the workload's own mesh (same config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cell-order", default="hilbert", choices=["hilbert", "caller"],
                    help="hilbert: the host mesh is renumbered along a Hilbert curve first (default, both arms); caller: arrays stay in "
                         "the generator's row-major order and the engine renumbers internally (single GPU)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--time-passes", action="store_true", help="N > 1: time the two passes on every rank (turns graph replay off)")
    ap.add_argument("--partition", default="sfc", choices=["sfc", "rcb"],
                    help="N > 1: Hilbert-curve chunks (default, the measured configuration) or recursive coordinate bisection")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = Workload(args, world)
    if args.impl == "reference":
        if rank != 0:
            return
        if not wl.reference_can_run:
            print(json.dumps({"impl": "reference", "unavailable": "the reference has no periodic FlowBC (abc.cpp:493-494 throws) and its "
                              "face loop would count periodic edges twice (SURVEY H8/H8b): it cannot evaluate this workload"}))
            return
        import __graft_entry__ as g
        g.build_oracle(quiet=True)        # the checker libraries only: this process never maps libfvens_b200.so
        # same mesh, same --steps / --warmup as the CUDA arm (a 10M-cell evaluation takes a few tenths of a second on a
        # 16-thread host; --cpu-cells bounds the sample if a host is too slow for that)
        if args.cpu_cells:
            args.cells = args.cpu_cells
            wl = Workload(args, world)
        gf, ms, info = cpu_reference(wl, max(1, args.steps), max(1, args.warmup), extras=True)
        line = {"impl": "reference", "metric": "Gfaces/s", "value": gf, "unit": "Gfaces/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": args.scaling if args.workload == "vortex" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": wl.config(info["cells"], info["faces"], info["lattice"]),
                "residual_evals_per_s": 1e3/ms,
                "cpu_baseline": {"value": gf, "unit": "Gfaces/s", "cores": info["cores"], "kind": info["kind"],
                                 "sample": info["sample"], **info["extra"]},
                "e2e": {"value": gf, "unit": "Gfaces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        try:    # evidence that this arm is process-clean: the shared objects of this repo mapped by this process
            maps = sorted({ln.split()[-1][len(ROOT)+1:] for ln in open("/proc/self/maps") if ROOT in ln and ".so" in ln})
            line["repo_libraries_mapped"] = maps
            assert not any("libfvens_b200" in m for m in maps), "the reference arm must not load the product library"
        except OSError:
            pass
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from fvens_b200 import lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    if args.gpus != world:
        sys.exit(f"bench.py: --gpus {args.gpus} needs one rank per GPU but WORLD_SIZE is {world}; launch it as "
                 f"python -m torch.distributed.run --nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus} ...")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # everything runs on one non-default stream (CUDA graphs of the multi-GPU evaluation cannot be captured on the legacy
    # default stream; torch.cuda.Event and the library's timing events are recorded on this same stream)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        g.build(quiet=True)
    if world > 1:
        dist.barrier()
    lib.load()

    # every rank builds the same global mesh; with N > 1 it is partitioned (strong scaling of the 10M-cell case)
    arrs, u, (nx, ny) = wl.arrays()
    um = wl.host_mesh(arrs)
    phys = wl.physics()
    BCS = wl.bcs
    nc_glob, nf_glob = um.nelem, um.naface
    stream = torch.cuda.current_stream().cuda_stream
    num = wl.num
    if world == 1:
        dm = lib.DeviceMesh(um, reorder="none" if args.cell_order == "hilbert" else "hilbert", tile_cells=args.tile, device=local_rank)
        fl = lib.FlowFV(dm, phys, bcs=BCS, **num)
        nc = nc_glob
        du = torch.from_numpy(u).cuda()
        halo_launches = 0

        def evaluate():
            fl.compute_residual(du, res, True, dtm, accumulate=False, stream=stream)
        if os.environ.get("FVG_SELF_DIST") == "1":
            # diagnostic: the multi-GPU kernels (exchange engine with no neighbour) on one GPU, to price their extra code
            eng = lib.DistEngine(fl)
            eng.connect([eng.handle()], np.zeros((1, 1), dtype=np.int32))

            def evaluate():
                eng.residual(du, res, True, dtm, accumulate=False, stream=stream)
    else:
        from fvens_b200.dist import DistFlow
        part = (lib.partition_rcb if args.partition == "rcb" else lib.partition_sfc)(um, world)
        df = DistFlow(um, part, rank, world, phys, dev, reorder="none", tile_cells=args.tile, bcs=BCS, **num)
        dm, fl = df.dmesh, df.flow
        nc = df.ncell
        ids = torch.from_numpy(df.global_ids.astype(np.int64))
        du = torch.zeros((df.ncell + df.nghost, 4), dtype=torch.float64, device=dev)
        du[:nc] = torch.from_numpy(u[ids[:nc].numpy()]).to(dev)
        halo_launches = 0 if df.engine is not None else 2      # split schedule: send kernels per evaluation (state rows, gradient rows)

        def evaluate():
            df.residual(du, res, dtm, gettimesteps=True, exchange_state=True)
    res = torch.empty((nc, 4), dtype=torch.float64, device=dev)
    dtm = torch.empty(nc, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi needs a few hundred ms to deliver its first row: start it before the warm-up and wait for it
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        t_wait = time.perf_counter()
        while sampler.proc and sampler.count() == 0 and time.perf_counter() - t_wait < 5.0:
            time.sleep(0.02)
    for _ in range(args.warmup):
        evaluate()
    barrier()
    if world == 1:
        fl.timing(True)
    elif df.engine is not None and args.time_passes:
        # per-pass times on every rank (events between the kernels: no graph replay, no launch overlap while this is on)
        fl.timing(True)
    launches0 = fl.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        evaluate()
    e1.record()
    barrier()
    windows = [(w0, time.perf_counter())]
    ms_total = e0.elapsed_time(e1)
    launches = fl.launch_count() - launches0 + halo_launches*args.steps
    ms_cell, ms_face, ntimed = fl.timing(False) if (world == 1 or (df.engine is not None and args.time_passes)) else (0.0, 0.0, 0)
    if world > 1 and ntimed:
        tk = torch.tensor([ms_cell/ntimed, ms_face/ntimed], dtype=torch.float64, device=dev)
        tmin = tk.clone(); dist.all_reduce(tk, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        pass_times = {"gradient_limiter_pass_max": tk[0].item(), "face_pass_max": tk[1].item(),
                      "gradient_limiter_pass_min": tmin[0].item(), "face_pass_min": tmin[1].item(), "timed_evals": ntimed,
                      "note": "per-rank CUDA events between the two kernels (direct launches, no graph replay): max / min over the ranks"}
    else:
        pass_times = None
    # a short timed region can end between two nvidia-smi rows (100 ms apart): keep the same kernels running,
    # untimed, until at least three rows have been taken under this load
    probe = torch.tensor([0], dtype=torch.int32, device=dev)
    p0 = time.perf_counter()
    while True:
        if rank == 0:
            have = sum(1 for r in sampler.rows if len(r) >= 10 and r[0] >= w0 + 0.02)
            probe[0] = 1 if (sampler.proc is None or have >= 3 or time.perf_counter() - p0 > 3.0) else 0
        if world > 1:
            dist.broadcast(probe, 0)
        if int(probe.item()) == 1:
            break
        for _ in range(20):
            evaluate()
        barrier()
    windows.append((p0, time.perf_counter()))
    clocks = sampler.stop(windows) if rank == 0 else None
    if clocks is not None:
        clocks["sampled"] = "nvidia-smi every 100 ms during the timed region and an untimed continuation of the same kernels"
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = tmax.item()/args.steps

    # sanity: the timed output is a real residual (finite, non-trivial)
    chk = float(res.abs().max().item())
    assert np.isfinite(chk) and chk > 0.0

    # end to end through host buffers (pinned): H2D of the state rows, kernels (+ halo), D2H of residual and dt
    if world == 1:
        hu = torch.from_numpy(u).pin_memory()
    else:
        hu = torch.from_numpy(u[ids[:nc].numpy()]).pin_memory()
    hres = torch.empty((nc, 4), dtype=torch.float64).pin_memory()
    hdt = torch.empty(nc, dtype=torch.float64).pin_memory()

    def evaluate_e2e():
        if world == 1:
            fl.compute_residual_host(hu.data_ptr(), hres.data_ptr(), True, hdt.data_ptr(), accumulate=False)
        else:
            du[:nc].copy_(hu, non_blocking=True)
            evaluate()
            hres.copy_(res, non_blocking=True)
            hdt.copy_(dtm, non_blocking=True)
            torch.cuda.synchronize()
    for _ in range(2):
        evaluate_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        evaluate_e2e()
    barrier()
    t_e2e = torch.tensor([(time.perf_counter()-t0)/args.e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    assert torch.equal(hres, res.cpu())

    # fused pseudo-time step (residual + dt + update + norm), device resident
    n2 = torch.zeros(1, dtype=torch.float64, device=dev)
    nstep = max(5, args.steps//3)
    if world == 1:
        u2 = du.clone()

        def step():
            fl.euler_step(u2, 0.5, n2, stream=stream)
    else:
        cur = [du.clone(), torch.zeros_like(du)]

        def step():
            df.euler_step(cur[0], cur[1], 0.5, n2)
            if not df.norm_is_global:
                dist.all_reduce(n2)
            cur.reverse()
    for _ in range(3):
        step()
    barrier()
    s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(nstep):
        step()
    s1.record()
    barrier()
    teu = torch.tensor([s0.elapsed_time(s1)/nstep], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(teu, op=dist.ReduceOp.MAX)
    ms_euler = teu.item()
    # the product's multi-GPU pseudo-time loop (fvg_dist_forward_euler_solve: norm reduced through the peer windows one
    # step behind, so no rank waits for another inside a step); wall clock over the whole call, copies in and out included
    ms_solver = None
    if world > 1 and df.engine is not None:
        usol = du.clone()
        df.solve_forward_euler(usol, 0.5, 1e-300, 5, check_every=5)
        barrier()
        t0s = time.perf_counter()
        code, nst, _ = df.solve_forward_euler(usol, 0.5, 1e-300, nstep, check_every=nstep)
        torch.cuda.synchronize()
        tsol = torch.tensor([(time.perf_counter() - t0s)/max(nst, 1)*1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(tsol, op=dist.ReduceOp.MAX)
        ms_solver = tsol.item()
        df.check()
        del usol

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bA, bB = wl.algorithmic_bytes(nc_glob, nf_glob)
    gfaces = nf_glob/(ms_step*1e-3)/1e9
    info = dm.info
    traffic, traffic_src = None, None
    try:        # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this command
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(f"face_kernel:{wl.tag}:{args.tile}:{int(args.cells)}") if args.workload == "bump" else None
        if traffic is not None:
            traffic_src = ("not measured by this run: dram__bytes_read.sum + dram__bytes_write.sum of one face_kernel launch from the "
                           "committed `ncu --set full` capture of this command, " + str(tj.get("source", "profiles/")))
    except Exception:
        pass
    line = {
        "metric": "Gfaces/s", "value": gfaces, "unit": "Gfaces/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": args.scaling if args.workload == "vortex" else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl.config(nc_glob, nf_glob, (nx, ny)),
        "layout": {"cells_on_rank0": nc, "ghost_cells_on_rank0": int(info.nghost),
                   "tile_cells": info.tile_cells, "cut_face_duplicates_rank0": info.ncut_dup,
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} GPUs, {'coordinate-bisection' if args.partition == 'rcb' else 'Hilbert-curve'} partition, one ghost layer; "
                   + ("per evaluation: gradient pass + face pass, one C call replayed from a CUDA graph; the producing kernels store the "
                      "neighbours' state / gradient rows into their peer-mapped windows over NVLink (CUDA IPC), the consuming kernels wait "
                      "on arrival flags in the partition-boundary tiles, which run last (fvg_dist_residual)" if df.halo_kind == "fused" else
                      "per evaluation: state halo, gradient pass, gradient halo, face pass; halo transport: "
                      + (("peer-mapped windows over NVLink (CUDA IPC, direct stores + flags)"
                          + (", received inside the consuming kernels" if df.fused_recv else ", one send+receive kernel per exchange"))
                         if df.halo_kind == "peer" else "NCCL all-to-all with row splits"))},
        "residual_evals_per_s": 1e3/ms_step,
        "residual_roofline_frac": (bA + bB)/(ms_step*1e-3)/1e9/(peak*world),
        "euler_step": {"ms_per_step": ms_euler, "Gfaces/s": nf_glob/(ms_euler*1e-3)/1e9,
                       "solver_loop_ms_per_step": ms_solver,
                       "note": "fused residual + local dt + forward-Euler update + energy-residual norm"
                               + ((" + norm reduced over the ranks through the peer windows (no NCCL call)" if df.norm_is_global
                                   else " + all-reduce of the norm") if world > 1 else "")},
        "e2e": {"value": nf_glob/t_e2e.item()/1e9, "unit": "Gfaces/s", "ms_per_step": t_e2e.item()*1e3,
                "h2d_bytes_per_step": 32*nc, "d2h_bytes_per_step": 40*nc,
                "note": "pinned host buffers: H2D u, kernels" + (" + halos" if world > 1 else "") + ", D2H residual + dt"
                        + (" (bytes are per rank)" if world > 1 else " (fvg_residual_host: chunked upload / compute / download pipeline)")},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world == 1:
        t_face = ms_face/max(ntimed, 1)*1e-3
        t_cell = ms_cell/max(ntimed, 1)*1e-3
        ach_face = bB/t_face/1e9
        line["kernels_ms"] = {"gradient_limiter_pass": t_cell*1e3, "face_pass": t_face*1e3, "timed_evals": ntimed}
        line["roofline"] = {"bound": "hbm", "kernel": "face_kernel (reconstruct + flux + spectral radius + accumulate)",
                            "achieved": ach_face, "peak": peak, "unit": "GB/s", "frac": ach_face/peak, "frac_of_nominal_8000_GBs": ach_face/8000.0, "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes_per_launch": bB, "peak_source": peak_src,
                            "cell_pass": {"achieved": bA/t_cell/1e9, "frac": bA/t_cell/1e9/peak,
                                          "algorithmic_bytes_per_launch": bA}}
    else:
        if pass_times:
            line["kernels_ms"] = pass_times
        ach = (bA + bB)/world/(ms_step*1e-3)/1e9
        line["roofline"] = {"bound": "hbm", "kernel": "whole evaluation per GPU (cell pass + face pass + halos)",
                            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach/peak, "frac_of_nominal_8000_GBs": ach/8000.0, "traffic": None,
                            "algorithmic_bytes_per_launch": (bA + bB)/world, "peak_source": peak_src}
    if not args.no_cpu_baseline and world == 1 and wl.reference_can_run:
        wlc = wl
        if args.cpu_cells:
            args.cells = args.cpu_cells
            wlc = Workload(args, world)
        gf, ms, ci = cpu_reference(wlc, 6, 2, prebuilt=None if args.cpu_cells else (arrs, u, (nx, ny)))
        line["cpu_baseline"] = {"value": gf, "unit": "Gfaces/s", "cores": ci["cores"], "kind": ci["kind"],
                                "sample": ci["sample"], "ms_per_eval": ms, **ci["extra"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
