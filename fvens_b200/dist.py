"""Multi-GPU driver: one process per GPU, a subdomain device mesh per rank, ghost rows exchanged between
the passes. Plumbing only (torch.distributed for the transport); every kernel is in libfvens_b200.so.

Schedule per residual (reference: the three exchanges of SURVEY 2.1 reduced to two, ghost-cell based):
  1. ghost rows of the state u          (after the update; the reference's VecGhostUpdate of u)
  2. gradient/limiter pass on own cells
  3. ghost rows of the limited gradients (+ unlimited ones for WENO / viscous-with-limiter)
  4. face pass on own cells (cut faces evaluated identically on both ranks)
The ghost block of a device-ordered array is ordered by source rank, and each rank packs the rows a peer
needs in that peer's ghost order, so one all-to-all with row splits moves a whole exchange and the receive
buffer IS the ghost block (no unpack kernel).
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import lib


class HaloExchange:
    def __init__(self, dmesh, device, group=None):
        self.dmesh = dmesh
        self.group = group
        self.device = device
        sc, rc, idx = dmesh.halo_lists()
        self.send_counts = [int(x) for x in sc]
        self.recv_counts = [int(x) for x in rc]
        self.send_idx_host = idx.copy()
        self.nsend = int(sc.sum())
        self.ncell, self.nghost = dmesh.ncell, dmesh.nghost
        self._bufs = {}
        self.on_gpu = torch.device(device).type == "cuda"
        if not self.on_gpu:
            self.send_idx_t = torch.from_numpy(idx.astype(np.int64))

    def _sendbuf(self, width):
        if width not in self._bufs:
            self._bufs[width] = torch.empty((max(self.nsend, 1), width), dtype=torch.float64, device=self.device)
        return self._bufs[width]

    def exchange(self, arr):
        """arr: [ncell+nghost, width] device-ordered; fills the ghost rows from their owners."""
        width = arr.shape[1]
        sb = self._sendbuf(width)
        if self.on_gpu:
            self.dmesh.halo_pack(arr, width, sb, stream=torch.cuda.current_stream().cuda_stream)
        else:   # host-side test path (gloo): same pattern, numpy gather instead of the pack kernel
            if self.nsend:
                sb[:self.nsend] = arr[self.send_idx_t]
        ghost = arr[self.ncell:]
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            try:
                dist.all_to_all_single(ghost, sb[:self.nsend], self.recv_counts, self.send_counts, group=self.group)
            except (RuntimeError, NotImplementedError):
                self._exchange_p2p(ghost, sb)
        else:
            assert self.nghost == 0


    def _exchange_p2p(self, ghost, sb):
        ops, so, ro = [], 0, 0
        rank = dist.get_rank(self.group)
        for r in range(len(self.send_counts)):
            if r != rank and self.send_counts[r]:
                ops.append(dist.P2POp(dist.isend, sb[so:so+self.send_counts[r]], r, group=self.group))
            if r != rank and self.recv_counts[r]:
                ops.append(dist.P2POp(dist.irecv, ghost[ro:ro+self.recv_counts[r]], r, group=self.group))
            so += self.send_counts[r]; ro += self.recv_counts[r]
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class PeerHaloExchange:
    """Halo exchange through peer-mapped windows over NVLink (fvg_halo_*): two small kernels per exchange, no
    collective library and no host synchronisation. torch.distributed is used once, at set-up, to all-gather the
    64-byte IPC handles and the receive counts."""

    def __init__(self, dmesh, device, group=None, max_width=8):
        self.dmesh = dmesh
        self.win = lib.PeerHaloWindow(dmesh, max_width)
        world = dist.get_world_size(group)
        _, rc, _ = dmesh.halo_lists()
        mine = torch.zeros(64 + 4*world, dtype=torch.uint8)
        mine[:64] = torch.frombuffer(bytearray(self.win.handle()), dtype=torch.uint8)
        mine[64:] = torch.from_numpy(np.ascontiguousarray(rc, dtype=np.int32).view(np.uint8))
        if dist.get_backend(group) == "nccl":      # NCCL moves device tensors; gloo (single-GPU tests) host tensors
            mine = mine.to(device)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine, group=group)
        every = [e.cpu().numpy() for e in every]
        handles = [e[:64].tobytes() for e in every]
        counts = np.stack([e[64:].view(np.int32) for e in every])
        self.win.connect(handles, counts)
        dist.barrier(group=group)

    def exchange(self, arr):
        """arr: [ncell+nghost, width] device-ordered; fills the ghost rows from their owners."""
        self.win.exchange(arr, arr.shape[1], stream=torch.cuda.current_stream().cuda_stream)

    def send(self, arr):
        self.win.send(arr, arr.shape[1], stream=torch.cuda.current_stream().cuda_stream)

    def recv(self, arr):
        self.win.recv(arr, arr.shape[1], stream=torch.cuda.current_stream().cuda_stream)


def _allgather_handles(handle, recv_counts, device, group):
    """All-gather of the 64-byte IPC handles and the receive-count rows (set-up only)."""
    world = dist.get_world_size(group)
    mine = torch.zeros(64 + 4*world, dtype=torch.uint8)
    mine[:64] = torch.frombuffer(bytearray(handle), dtype=torch.uint8)
    mine[64:] = torch.from_numpy(np.ascontiguousarray(recv_counts, dtype=np.int32).view(np.uint8))
    if dist.get_backend(group) == "nccl":      # NCCL moves device tensors; gloo (single-GPU tests) host tensors
        mine = mine.to(device)
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine, group=group)
    every = [e.cpu().numpy() for e in every]
    return [e[:64].tobytes() for e in every], np.stack([e[64:].view(np.int32) for e in every])


class DistFlow:
    """FlowFV on one rank's subdomain + the halo exchanges. Arrays are device-ordered [ncell+nghost, .].

    Default on GPUs (FVG_DIST=fused): the fused evaluation of libfvens_b200.so (fvg_dist_*, csrc/dist.cu) - two kernels per
    residual, rows pushed to the neighbours by the producing kernels, one C call per evaluation replayed from a CUDA
    graph. FVG_DIST=split selects the round-1 schedule driven from here (send kernels + split passes), kept for A/B
    comparisons and for the CPU (gloo) tests of the exchange pattern."""

    def __init__(self, umesh, cell_rank, rank, nranks, phys, device, reorder="hilbert", tile_cells=256, group=None,
                 **numerics):
        self.dmesh = lib.DeviceMesh(umesh, reorder=reorder, tile_cells=tile_cells, device=torch.device(device).index or 0,
                                    cell_rank=cell_rank, rank=rank, nranks=nranks)
        self.flow = lib.FlowFV(self.dmesh, phys, **numerics)
        # transport: peer-mapped windows over NVLink on GPUs (FVG_HALO=nccl selects the all-to-all), gloo on the CPU
        fused = torch.device(device).type == "cuda" and nranks > 1 and os.environ.get("FVG_DIST", "fused") == "fused"
        use_peer = (torch.device(device).type == "cuda" and nranks > 1 and os.environ.get("FVG_HALO", "peer") == "peer"
                    and not fused)
        self.halo = PeerHaloExchange(self.dmesh, device, group) if use_peer else HaloExchange(self.dmesh, device, group)
        self.halo_kind = "peer" if use_peer else "collective"
        self.ncell, self.nghost = self.dmesh.ncell, self.dmesh.nghost
        n = self.ncell + self.nghost
        num = self.flow.num
        self.order2 = bool(num.order2)
        self.weno = num.reconstruction == lib.RECON["WENO"]
        limited = num.reconstruction in (lib.RECON["BARTHJESPERSEN"], lib.RECON["VENKATAKRISHNAN"])
        self.need_lg = self.order2 and num.reconstruction != lib.RECON["VANALBADA"]
        self.need_gu = self.order2 and (self.weno or num.reconstruction == lib.RECON["VANALBADA"] or
                                        (phys.viscous_sim and limited))
        self.lg = torch.zeros((n, 8), dtype=torch.float64, device=device) if self.need_lg else None
        self.gu = torch.zeros((n, 8), dtype=torch.float64, device=device) if self.need_gu else None
        if self.order2:
            self.flow.use_buffers(self.lg, self.gu)
        self.global_ids = self.dmesh.permutation()
        # FVG_OVERLAP=1: exchanges overlapped with the tiles that see no ghost cell. Off by default: on B200 boxes the
        # two extra launches per pass cost what the hidden exchanges save (10M cells: 0.600 vs 0.586 ms at 2 GPUs, 0.193
        # vs 0.189 ms at 8; profiles/r01_bench_v9_*). WENO always uses the plain sequence (its second stage needs every
        # neighbour's unlimited gradient first).
        self.overlap = (torch.device(device).type == "cuda" and nranks > 1 and not self.weno
                        and os.environ.get("FVG_OVERLAP", "0") == "1")
        # in-kernel receive (default with the peer windows; FVG_FUSED_RECV=0 turns it off): inviscid flows with linear
        # reconstruction or first order - the other passes read ghost rows straight from the arrays
        self.fused_recv = (use_peer and not self.overlap and not self.need_gu and not self.weno and not phys.viscous_sim
                           and os.environ.get("FVG_FUSED_RECV", "1") != "0")
        # the fused C-level evaluation
        self.engine = None
        if torch.device(device).type == "cuda" and nranks > 1 and os.environ.get("FVG_DIST", "fused") == "fused":
            self.engine = lib.DistEngine(self.flow)
            _, rc, _ = self.dmesh.halo_lists()
            handles, counts = _allgather_handles(self.engine.handle(), rc, device, group)
            self.engine.connect(handles, counts)
            dist.barrier(group=group)
            self.halo_kind = "fused"
        if self.overlap:
            self._halo_stream = torch.cuda.Stream(device=device, priority=-1)
            self._ev_u, self._ev_g, self._ev_l = (torch.cuda.Event() for _ in range(3))

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    @property
    def norm_is_global(self):
        """True if euler_step's norm2 is already reduced over the ranks (fused engine)."""
        return self.engine is not None

    def solve_forward_euler(self, u, cfl, tol, maxiter, check_every=1):
        """SteadyForwardEulerSolver::solve on the partitioned mesh (fvg_dist_forward_euler_solve): u [>= ncell, 4], own rows
        updated in place. Returns (status, steps, global history)."""
        if self.engine is None:
            raise RuntimeError("the multi-GPU solver needs the fused engine (FVG_DIST=fused on CUDA devices)")
        return self.engine.solve_forward_euler(u, cfl, tol, maxiter, check_every)

    def check(self):
        """Raises FvgError(COMM) if a neighbour stopped delivering rows (call at the caller's own sync points)."""
        if self.engine is not None:
            self.engine.status()
        elif self.halo_kind == "peer":
            seq = self.halo.win.status()
            if seq:
                raise lib.FvgError(7, f"a halo receive timed out waiting for exchange {seq}")

    def _gradients(self, u):
        if not self.order2:
            return
        s = self._stream()
        self.flow.gradient_pass(u, 0, stream=s)
        if self.weno:
            self.halo.exchange(self.gu)
            self.flow.gradient_pass(u, 1, stream=s)
        elif self.need_gu:
            self.halo.exchange(self.gu)
        if self.need_lg:
            self.halo.exchange(self.lg)

    def _overlapped(self, u, exchange_state, face):
        """Both exchanges hidden behind the tiles that need no ghost row:
             halo stream:    exchange(u) .......................... exchange(gradients) ..........
             compute stream: gradients, interior tiles | boundary tiles | face pass, interior | boundary
        The exchange kernels are enqueued on a high-priority stream BEFORE the compute kernel they overlap with (the
        face kernel is persistent and would otherwise hold every SM until it ends)."""
        cur = torch.cuda.current_stream()
        hs = self._halo_stream
        fl = self.flow
        if exchange_state:
            hs.wait_stream(cur)                 # u is final (and the previous evaluation has consumed the ghost rows)
            with torch.cuda.stream(hs):
                self.halo.exchange(u)
                self._ev_u.record(hs)
        if self.order2:
            fl.select_tiles(1)
            fl.gradient_pass(u, 0, stream=cur.cuda_stream)
            if exchange_state:
                cur.wait_event(self._ev_u)
            fl.select_tiles(2)
            fl.gradient_pass(u, 0, stream=cur.cuda_stream)
            self._ev_g.record(cur)
            hs.wait_event(self._ev_g)
            with torch.cuda.stream(hs):
                if self.need_gu:
                    self.halo.exchange(self.gu)
                if self.need_lg:
                    self.halo.exchange(self.lg)
                self._ev_l.record(hs)
            fl.select_tiles(1)
            face(cur.cuda_stream)
            cur.wait_event(self._ev_l)
        else:
            fl.select_tiles(1)
            face(cur.cuda_stream)
            if exchange_state:
                cur.wait_event(self._ev_u)
        fl.select_tiles(2)
        face(cur.cuda_stream)
        fl.select_tiles(0)

    def _in_kernel_receive(self, u, face):
        """Two sends and two passes, no receive kernel: each pass waits for the neighbours' rows itself, only in the CTAs
        that reach a tile with ghost cells in its halo, and reads the ghost rows from the halo window."""
        s = self._stream()
        fl, win = self.flow, self.halo.win
        fl.ghost_source(0, win, win.post(u, 4, stream=s))
        if self.order2:
            fl.gradient_pass(u, 0, stream=s)
            fl.ghost_source(1, win, win.post(self.lg, 8, stream=s))
        face(s)
        fl.ghost_source(0); fl.ghost_source(1)

    def residual(self, u, res, dtm, gettimesteps=True, exchange_state=True):
        """u [ncell+nghost,4]; res [ncell,4] (overwritten); dtm [ncell]."""
        if self.engine is not None:
            if not exchange_state:
                raise ValueError("the fused evaluation always brings the neighbours' state rows")
            return self.engine.residual(u, res, gettimesteps, dtm, accumulate=False, stream=self._stream())
        if self.fused_recv and exchange_state:
            return self._in_kernel_receive(u, lambda s: self.flow.face_pass(u, res, gettimesteps, dtm, accumulate=False, stream=s))
        if self.overlap:
            return self._overlapped(u, exchange_state,
                                    lambda s: self.flow.face_pass(u, res, gettimesteps, dtm, accumulate=False, stream=s))
        if exchange_state:
            self.halo.exchange(u)
        self._gradients(u)
        self.flow.face_pass(u, res, gettimesteps, dtm, accumulate=False, stream=self._stream())

    def euler_step(self, u, unew, cfl, norm2, exchange_state=True):
        """One forward-Euler step: unew (own rows) from u; norm2 (device scalar) = this rank's sum of r_E^2*area - or, with
        the fused engine, already the sum over ALL ranks (see `norm_is_global`)."""
        if self.engine is not None:
            return self.engine.euler_step(u, unew, cfl, norm2, stream=self._stream())
        if self.fused_recv and exchange_state:
            return self._in_kernel_receive(u, lambda s: self.flow.euler_face_pass(u, unew, cfl, norm2, stream=s))
        if self.overlap:
            return self._overlapped(u, exchange_state,
                                    lambda s: self.flow.euler_face_pass(u, unew, cfl, norm2, stream=s))
        if exchange_state:
            self.halo.exchange(u)
        self._gradients(u)
        self.flow.euler_face_pass(u, unew, cfl, norm2, stream=self._stream())
