"""Product host mesh (fvens::UMesh in libfvens_b200.so, edge-bucket algorithms) against the oracle's
restatement of the reference algorithms (esup searches): identical numbering, orientation, metrics.
Also the C-ABI surface: the library loads without a GPU and exports every symbol of include/fvens_b200.h,
and the host-only device-mesh build yields a valid tiling and edge colouring."""
import os
import re
import numpy as np
import pytest
import orc
from common import mesh_path, ROOT
from fvens_b200 import lib, synth

ALL_MESHES = ["testperiodic.msh", "2dcylinderhybrid.msh", "testhybrid.msh", "squarecoarse.msh",
              "squareunsquad0.msh", "2dcylinder0.msh", "2dcylinder1.msh", "naca0012luo.msh",
              "NACA0012_inv.su2", "NACA0012_lam_hybrid_1.msh"]


def compare(um, om):
    a, b = um.arrays(), om.arrays()
    assert (um.npoin, um.nelem, um.nbface, um.naface, um.ninface) == (om.npoin, om.nelem, om.nbface, om.naface, om.ninface)
    mw = um.maxnnode
    assert np.array_equal(a["coords"], b["coords"])
    assert np.array_equal(a["nnode"], b["nnode"])
    for k in ("inpoel", "esuel", "elemface"):
        assert np.array_equal(a[k], b[k][:, :mw]), k
    assert np.array_equal(a["intfac"], b["intfac"])
    assert np.array_equal(a["btags"][:, 0], b["btags"])
    assert np.array_equal(a["facemetric"], b["facemetric"])   # bit-exact
    assert np.array_equal(a["area"], b["area"])


@pytest.mark.parametrize("mesh", ALL_MESHES)
def test_umesh_matches_oracle_on_reference_meshes(mesh):
    compare(lib.UMesh.read(mesh_path(mesh)), orc.Mesh.read(mesh_path(mesh)))


def test_mesh_sizes_of_the_configs():
    # SURVEY section 8 size table
    m = lib.UMesh.read(mesh_path("NACA0012_inv.su2"))
    assert (m.nelem, m.nbface, m.naface) == (10216, 250, 15449)
    m = lib.UMesh.read(mesh_path("naca0012luo.msh"))
    assert (m.nelem, m.nbface, m.naface) == (2857, 113, 4342)
    m = lib.UMesh.read(mesh_path("NACA0012_lam_hybrid_1.msh"))
    assert (m.nelem, m.nbface, m.naface) == (13156, 250, 21897)


@pytest.mark.parametrize("gen", ["bump", "ogrid", "square"])
def test_umesh_matches_oracle_on_synthetic_meshes(gen):
    if gen == "bump":
        arrs = synth.bump_channel(40, 15)
    elif gen == "ogrid":
        arrs = synth.ogrid_cylinder(48, 12, tri_fraction=0.3)
    else:
        arrs = synth.square(16, tri_fraction=0.5, jitter=0.2)
    um = lib.UMesh.from_arrays(*arrs)
    compare(um, orc.Mesh.from_arrays(*arrs))
    assert (um.arrays()["area"] > 0).all()


def test_bump_channel_is_half_triangles():
    _, nnode, _, _ = synth.bump_channel(300, 110)
    frac = (nnode == 3).mean()
    assert 0.45 < frac < 0.55


def test_bad_inputs_are_reported():
    with pytest.raises(lib.FvgError) as e:
        lib.UMesh.read("/nonexistent/mesh.msh")
    assert e.value.code == 3
    coords, nnode, inpoel, bface = synth.square(4)
    bad = bface.copy(); bad[0, :2] = [0, 12]     # not an edge of any cell
    with pytest.raises(lib.FvgError):
        lib.UMesh.from_arrays(coords, nnode, inpoel, bad)
    with pytest.raises(lib.FvgError):            # a hole: boundary face missing
        lib.UMesh.from_arrays(coords, nnode, inpoel, bface[1:])


def test_reorder_cells_and_orderings():
    arrs = synth.bump_channel(32, 12)
    um = lib.UMesh.from_arrays(*arrs)
    for perm in (um.rcm_ordering(), um.hilbert_ordering()):
        assert sorted(perm.tolist()) == list(range(um.nelem))
    a0 = um.arrays()
    perm = um.hilbert_ordering()
    um.reorder_cells(perm)
    a1 = um.arrays()
    assert np.array_equal(a1["inpoel"], a0["inpoel"][perm])
    assert np.array_equal(a1["area"], a0["area"][perm])
    # the reordered mesh is again what the reference algorithms give for that cell order
    coords, nnode, inpoel, bface = arrs
    compare(um, orc.Mesh.from_arrays(coords, nnode[perm], inpoel[perm], bface))


def bandwidth(intfac, nb):
    return np.abs(intfac[nb:, 0] - intfac[nb:, 1]).max()


def test_rcm_reduces_bandwidth():
    rng = np.random.default_rng(3)
    coords, nnode, inpoel, bface = synth.square(24, tri_fraction=0.4)
    shuffle = rng.permutation(len(nnode))
    um = lib.UMesh.from_arrays(coords, nnode[shuffle], inpoel[shuffle], bface)
    before = bandwidth(um.arrays()["intfac"], um.nbface)
    um.reorder_cells(um.rcm_ordering())
    after = bandwidth(um.arrays()["intfac"], um.nbface)
    assert after < before/5


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fvens_b200.h")).read()
    names = sorted(set(re.findall(r"\b(fvg_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 30
    L = lib.load()
    for n in names:
        assert hasattr(L, n), n


@pytest.mark.parametrize("reorder", ["none", "hilbert", "rcm"])
@pytest.mark.parametrize("tile", [32, 128])
def test_tiling_and_colouring_are_valid(reorder, tile):
    um = lib.UMesh.read(mesh_path("2dcylinderhybrid.msh"))
    dm = lib.DeviceMesh(um, reorder=reorder, tile_cells=tile, device=-2)     # host-only build
    a = um.arrays()
    new2old = dm.permutation()
    assert sorted(new2old.tolist()) == list(range(um.nelem))
    old2new = np.empty_like(new2old); old2new[new2old] = np.arange(um.nelem)
    face, colour, tile_of = dm.stream()
    info = dm.info
    t0 = dm.tile_offsets()
    assert t0[0] == 0 and t0[-1] == um.nelem and (np.diff(t0) > 0).all() and np.diff(t0).max() <= tile
    cell_tile = np.repeat(np.arange(info.ntile), np.diff(t0))
    real = face >= 0
    assert real.sum() == um.naface + info.ncut_dup and info.max_colours <= 8
    assert (np.diff(np.concatenate(([0], np.cumsum(np.bincount(tile_of, minlength=info.ntile))))) % 4 == 0).all()

    def tiles_of_face(f):
        L = old2new[a["intfac"][f, 0]]
        ts = {int(cell_tile[L])}
        if f >= um.nbface:
            ts.add(int(cell_tile[old2new[a["intfac"][f, 1]]]))
        return ts
    # every face appears exactly once in every tile it touches
    seen = {}
    for e in np.nonzero(real)[0]:
        f = int(face[e])
        assert int(tile_of[e]) in tiles_of_face(f)
        assert int(tile_of[e]) not in seen.setdefault(f, set())
        seen[f].add(int(tile_of[e]))
    assert len(seen) == um.naface
    for f, ts in seen.items():
        assert ts == tiles_of_face(f)
    # no two entries of one (tile, colour) touch the same tile-owned cell
    used = set()
    for e in np.nonzero(real)[0]:
        f = int(face[e]); t = int(tile_of[e])
        cells = [old2new[a["intfac"][f, 0]]] + ([old2new[a["intfac"][f, 1]]] if f >= um.nbface else [])
        for c in cells:
            if cell_tile[c] == t:
                key = (t, int(colour[e]), int(c))
                assert key not in used
                used.add(key)
    # entries of a tile are ordered by kind: both cells in the tile, cut by the tile boundary, physical boundary
    # (the flux phase of the face kernel relies on warps being uniform in kind); padding entries may sit inside the
    # first two kinds (unused positions of a bank residue class) and at the end
    nconfl = ngroups = 0
    for t in range(info.ntile):
        fs = face[tile_of == t]
        kinds = []
        groups = {}
        for p, f in enumerate(fs):
            if f < 0:
                continue
            f = int(f)
            if f < um.nbface:
                kinds.append(2)
                continue
            kinds.append(0 if len(tiles_of_face(f)) == 1 else 1)
            # shared-memory bank residues: the entries that the 8 consecutive cells of a quarter warp reach through their
            # local face j should sit in distinct residue classes mod 8
            for side in (0, 1):
                old = int(a["intfac"][f, side])
                c = int(old2new[old])
                if cell_tile[c] != t:
                    continue
                j = int(np.nonzero(a["elemface"][old] == f)[0][0])
                groups.setdefault(((c - t0[t]) >> 3, j), {})[p] = p % 8     # keyed by entry: both cells of an entry in one group = one address
        assert (np.diff(kinds) >= 0).all()
        ngroups += len(groups)
        nconfl += sum(1 for g in groups.values() if len(set(g.values())) < len(g))
    assert ngroups == info.bank_groups and nconfl == info.bank_conflict_groups
    # (tiny tiles fall back to dense packing when the padding would exceed the staging capacity)
    assert nconfl <= (0.03 if tile >= 128 else 0.15)*ngroups


def test_tiles_shrink_to_fit_the_halo_capacity():
    # a shuffled numbering makes almost every neighbour a halo cell: tiles must shrink, not fail
    rng = np.random.default_rng(5)
    coords, nnode, inpoel, bface = synth.square(40, tri_fraction=0.3)
    sh = rng.permutation(len(nnode))
    um = lib.UMesh.from_arrays(coords, nnode[sh], inpoel[sh], bface)
    dm = lib.DeviceMesh(um, reorder="none", tile_cells=256, device=-2)
    assert np.diff(dm.tile_offsets()).max() < 100
    dm2 = lib.DeviceMesh(um, reorder="hilbert", tile_cells=256, device=-2)
    assert np.diff(dm2.tile_offsets()).mean() > 200


def test_hilbert_order_makes_compact_tiles():
    um = lib.UMesh.from_arrays(*synth.bump_channel(96, 36))
    none = lib.DeviceMesh(um, reorder="none", tile_cells=128, device=-2).info
    hil = lib.DeviceMesh(um, reorder="hilbert", tile_cells=128, device=-2).info
    # row-major strips cut far more faces than Hilbert patches
    assert hil.ncut_dup < 0.5*none.ncut_dup


def test_tiles_of_a_quad_mesh_are_cut_at_whole_rounds_of_the_face_kernel(monkeypatch):
    """All-quad meshes: a 256-cell tile has 2 entries per cell plus half its perimeter, just over two rounds of the face
    kernel's 256 threads (and over the staging capacity unless it is a perfect square, which made the earlier cut fall
    back to 204 cells). The cut ends a tile at 512 entries instead: more cells per tile, both rounds full."""
    from fvens_b200 import synth
    n = 160
    coords, nnode, inpoel, bface = synth.periodic_square(n)
    rc = synth.cell_centres(coords, nnode, inpoel)
    perm = synth.hilbert_order(rc)
    um = lib.UMesh.from_arrays(coords, np.ascontiguousarray(nnode[perm]), np.ascontiguousarray(inpoel[perm]), bface)

    def stats():
        dm = lib.DeviceMesh(um, reorder="none", device=-2)
        face, _, tile_of = dm.stream()
        real = np.bincount(tile_of[face >= 0], minlength=dm.info.ntile)
        return np.diff(dm.tile_offsets()), real
    cells, real = stats()
    assert real.max() <= 512 and np.median(real) >= 500
    assert cells.mean() > 225
    monkeypatch.setenv("FVG_TILE_ENTRY_CAP", "0")            # the capacity alone: the earlier rounds' cut
    cells_old, real_old = stats()
    assert cells_old.mean() < cells.mean() - 15 and real_old.max() <= 544
