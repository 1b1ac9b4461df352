"""Host-side logic of the multi-GPU path, without a GPU: the subdomain builder (own cells, one ghost layer,
send/receive lists) and the exchange pattern over a real 2-process gloo group - the analogue of the
reference's tests/solvers/testtracevector.cpp (every ghost must receive a value made of its owner's rank and
the global cell id)."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import mesh_path
from fvens_b200 import lib, synth


def build_parts(um, nranks, reorder="hilbert", tile=64, partitioner="sfc"):
    part = (lib.partition_rcb if partitioner == "rcb" else lib.partition_sfc)(um, nranks)
    return part, [lib.DeviceMesh(um, reorder=reorder, tile_cells=tile, device=-2, cell_rank=part, rank=r, nranks=nranks)
                  for r in range(nranks)]


@pytest.mark.parametrize("partitioner", ["sfc", "rcb"])
@pytest.mark.parametrize("nranks", [2, 3, 8])
@pytest.mark.parametrize("mesh", ["2dcylinderhybrid.msh", "bump"])
def test_subdomains_are_consistent(mesh, nranks, partitioner):
    um = lib.UMesh.from_arrays(*synth.bump_channel(40, 15)) if mesh == "bump" else lib.UMesh.read(mesh_path(mesh))
    a = um.arrays()
    part, dms = build_parts(um, nranks, partitioner=partitioner)
    counts = np.bincount(part, minlength=nranks)
    assert counts.min() >= um.nelem//nranks - 1 and counts.max() <= um.nelem//nranks + 1      # balanced
    owned = np.concatenate([dm.permutation()[:dm.ncell] for dm in dms])
    assert sorted(owned.tolist()) == list(range(um.nelem))                        # every cell owned exactly once
    nb = um.nbface
    for r, dm in enumerate(dms):
        ids = dm.permutation()
        own, ghost = ids[:dm.ncell], ids[dm.ncell:]
        assert (part[own] == r).all() and (part[ghost] != r).all()
        # ghosts = exactly the cells of other ranks across a face of an own cell
        expect = set()
        ownset = set(own.tolist())
        for L, R in a["intfac"][nb:, :2]:
            if (L in ownset) != (R in ownset):
                expect.add(int(R if L in ownset else L))
        assert set(ghost.tolist()) == expect and len(ghost) == len(expect)
        # ghosts grouped by owner rank, ascending
        assert (np.diff(part[ghost]) >= 0).all()
        sc, rc, idx = dm.halo_lists()
        assert rc.sum() == dm.nghost and np.array_equal(rc, np.bincount(part[ghost], minlength=nranks))
        assert sc[r] == 0 and rc[r] == 0 and (idx < dm.ncell).all()
    # what rank s sends to rank r is, in order, rank r's ghost block from s
    for r, dr in enumerate(dms):
        gr = dr.permutation()[dr.ncell:]
        _, rc, _ = dr.halo_lists()
        off = np.concatenate(([0], np.cumsum(rc)))
        for s, ds in enumerate(dms):
            sc, _, idx = ds.halo_lists()
            so = np.concatenate(([0], np.cumsum(sc)))
            sent = ds.permutation()[idx[so[r]:so[r+1]]]
            assert np.array_equal(sent, gr[off[s]:off[s+1]])
    # every face of the mesh is evaluated by the ranks owning its cells, and by no one else
    for r, dm in enumerate(dms):
        face, _, _ = dm.stream()
        faces = set(face[face >= 0].tolist())
        ownset = set(dm.permutation()[:dm.ncell].tolist())
        expect = {f for f in range(um.naface) if a["intfac"][f, 0] in ownset or (f >= nb and a["intfac"][f, 1] in ownset)}
        assert faces == expect


def _worker(rank, world, port, meshname, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fvens_b200 import lib as L, synth as S
        from fvens_b200.dist import HaloExchange
        um = L.UMesh.from_arrays(*S.bump_channel(40, 15))
        part = L.partition_sfc(um, world)
        dm = L.DeviceMesh(um, reorder="hilbert", tile_cells=64, device=-2, cell_rank=part, rank=rank, nranks=world)
        ids = dm.permutation()
        hx = HaloExchange(dm, "cpu")
        arr = torch.zeros((dm.ncell + dm.nghost, 4), dtype=torch.float64)
        own = torch.from_numpy(ids[:dm.ncell].astype(np.float64))
        for v in range(4):
            arr[:dm.ncell, v] = rank*1.0e7 + own*10 + v
        hx.exchange(arr)
        g = ids[dm.ncell:]
        expect = np.stack([part[g]*1.0e7 + g*10 + v for v in range(4)], axis=1)
        ok = bool(np.array_equal(arr[dm.ncell:].numpy(), expect)) and dm.nghost > 0
        # a 1-double all-reduce like the residual norm
        t = torch.tensor([float(dm.ncell)], dtype=torch.float64)
        dist.all_reduce(t)
        ok = ok and int(t.item()) == um.nelem
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, "bump", q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    results = dict(q.get(timeout=5) for _ in range(world))
    assert all(p.exitcode == 0 for p in procs)
    assert results == {r: True for r in range(world)}


def test_scotch_graph_export_and_mapping_import(tmp_path):
    """The cell adjacency graph the reference hands to Scotch (getCellAdjLists, mesh/meshpartitioning.cpp:376-430)
    as a Scotch source-graph file, and a Scotch mapping file read back as the cell -> rank map of
    fvg_mesh_create_part (the drop-in point for a real Scotch partition)."""
    um = lib.UMesh.read(mesh_path("2dcylinderhybrid.msh"))
    a = um.arrays()
    ptrs, store = um.cell_adjacency()
    n = um.nelem
    interior = a["intfac"][um.nbface:]
    assert ptrs[-1] == 2*len(interior)                       # every interior face is an arc in both directions
    adj = [set(store[ptrs[i]:ptrs[i+1]].tolist()) for i in range(n)]
    assert all(len(adj[i]) == ptrs[i+1] - ptrs[i] for i in range(n))
    for L, R in interior[:, :2]:
        assert R in adj[L] and L in adj[R]
    # neighbours come in local-face order, as the reference's list does
    for i in (0, n//2, n-1):
        want = [e for e in a["esuel"][i, :a["nnode"][i]] if 0 <= e < n]
        assert store[ptrs[i]:ptrs[i+1]].tolist() == want
    g = tmp_path / "mesh.grf"
    um.write_scotch_graph(g)
    lines = g.read_text().split("\n")
    assert lines[0] == "0" and lines[1].split() == [str(n), str(ptrs[-1])] and lines[2] == "0 000"
    for i in (0, 17, n-1):
        row = [int(x) for x in lines[3+i].split()]
        assert row[0] == len(row) - 1 and row[1:] == store[ptrs[i]:ptrs[i+1]].tolist()
    assert len([ln for ln in lines if ln]) == 3 + n
    # a mapping file (here: the SFC partition written in Scotch's format, lines in shuffled order)
    part = lib.partition_sfc(um, 3)
    order = np.random.default_rng(0).permutation(n)
    mp = tmp_path / "mesh.map"
    mp.write_text(f"{n}\n" + "".join(f"{v}\t{part[v]}\n" for v in order))
    cr, nparts = um.read_scotch_map(mp)
    assert nparts == 3 and np.array_equal(cr, part)
    mp.write_text(f"{n-1}\n")
    with pytest.raises(lib.FvgError):
        um.read_scotch_map(mp)


@pytest.mark.parametrize("nranks", [2, 5])
def test_tile_send_lists_are_the_halo_send_lists_grouped_by_tile(nranks):
    """fvg_mesh_tile_send_lists: the rows a tile's cells contribute to the neighbours' ghost blocks (for kernels that push
    their output to the peers as they produce it) are exactly the per-rank send lists, regrouped by tile."""
    um = lib.UMesh.from_arrays(*synth.bump_channel(60, 24))
    part, meshes = build_parts(um, nranks, tile=64)
    for r, dm in enumerate(meshes):
        sc, rc, idx = dm.halo_lists()
        off, tr = dm.tile_send_lists()
        t0 = dm.tile_offsets()
        assert off[0] == 0 and off[-1] == len(idx) == len(tr) and (np.diff(off) >= 0).all()
        tile = np.repeat(np.arange(dm.info.ntile), np.diff(off))
        cell = t0[tile] + tr[:, 0]
        assert (tr[:, 0] >= 0).all() and (cell < t0[tile + 1]).all()
        so = np.concatenate(([0], np.cumsum(sc)))
        want = {(int(idx[k]), p, k - int(so[p])) for p in range(nranks) for k in range(so[p], so[p+1])}
        got = {(int(c), int(p), int(q)) for c, (_, p, q) in zip(cell, tr)}
        assert got == want and len(got) == len(tr)
        # rows of one peer are visited in ascending order inside a tile (ascending cells)
        for t in range(dm.info.ntile):
            seg = tr[off[t]:off[t+1]]
            for p in set(seg[:, 1].tolist()):
                rows = seg[seg[:, 1] == p, 2]
                assert (np.diff(rows) > 0).all()


def edge_cut(um, part):
    a = um.arrays()
    f = a["intfac"][um.nbface:, :2]
    return int((part[f[:, 0]] != part[f[:, 1]]).sum())


@pytest.mark.parametrize("nranks", [2, 4, 5, 8])
def test_partition_quality_edge_cut_and_balance(nranks):
    """the in-tree partitioners are judged by edge cut and balance. On the benchmark geometry (bump
    channel, aspect 2.67) both are balanced to one cell; recursive coordinate bisection cuts fewer faces than the
    Hilbert-curve partition (except where the curve's own cut happens to be straight), and both stay within a small factor of the straight-cut estimate
    (nranks - 1 cuts across the short side for strips)."""
    nx, ny = 400, 150
    um = lib.UMesh.from_arrays(*synth.bump_channel(nx, ny))
    sfc, rcb = lib.partition_sfc(um, nranks), lib.partition_rcb(um, nranks)
    for part in (sfc, rcb):
        c = np.bincount(part, minlength=nranks)
        assert c.max() - c.min() <= 1 and part.min() == 0 and part.max() == nranks - 1
    cs, cr = edge_cut(um, sfc), edge_cut(um, rcb)
    strips = (nranks - 1)*ny*4/3            # a straight cut crosses ny lattice rows; a split quad adds a diagonal
    assert cr <= 1.15*cs and cr < 1.6*strips and cs < 3.0*strips
    # deterministic
    assert np.array_equal(rcb, lib.partition_rcb(um, nranks))


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("nranks,partition", [(2, "sfc"), (3, "rcb"), (5, "sfc")])
def test_send_lists_and_ghost_blocks_of_all_ranks_agree(nranks, partition, periodic):
    """The exchange pattern of the fused evaluation, checked on the CPU with host-only device meshes of EVERY rank: the
    rows rank q sends to rank r - in q's send order - are exactly the cells of r's ghost block from q, in r's ghost
    order. Holds with the late numbering of the partition-boundary tiles (ghost order is keyed on the global locality
    order, not on the owner's device numbering) and with periodic pairs, whose ghost copies may come from the rank
    itself; every own cell appears once, every tile that sends is in one late block."""
    if periodic:
        arrs = synth.periodic_square(36, tri_fraction=0.3, jitter=0.1)
        um = lib.UMesh.from_arrays(*arrs)
        um.compute_periodic_map(3, 0); um.compute_periodic_map(4, 1)
    else:
        arrs = synth.bump_channel(60, 24)
        um = lib.UMesh.from_arrays(*arrs)
    part = (lib.partition_sfc if partition == "sfc" else lib.partition_rcb)(um, nranks)
    meshes = [lib.DeviceMesh(um, reorder="hilbert", tile_cells=64, device=-2, cell_rank=part, rank=r, nranks=nranks) for r in range(nranks)]
    perms = [m.permutation() for m in meshes]
    lists = [m.halo_lists() for m in meshes]
    owned = np.concatenate([p[:m.ncell] for p, m in zip(perms, meshes)])
    assert np.array_equal(np.sort(owned), np.arange(um.nelem))
    for r in range(nranks):
        sc, rc, idx = lists[r]
        assert rc.sum() == meshes[r].nghost and sc.sum() == len(idx)
        goff = np.concatenate(([0], np.cumsum(rc)))
        for q in range(nranks):
            scq, _, idxq = lists[q]
            soff = np.concatenate(([0], np.cumsum(scq)))
            sent = perms[q][idxq[soff[r]:soff[r+1]]]                     # global cells q pushes to r, in q's order
            ghosts = perms[r][meshes[r].ncell + goff[q]: meshes[r].ncell + goff[q+1]]
            assert np.array_equal(sent, ghosts), (r, q)
            assert (part[ghosts] == q).all()
            if q == r and not periodic:
                assert len(ghosts) == 0
        toff, _ = meshes[r].tile_send_lists()
        sends = np.diff(toff) > 0
        first, last = np.argmax(sends), len(sends) - 1 - np.argmax(sends[::-1])
        assert sends[first:last+1].all()
