"""Tier E: the whole chain in the reference's own object code - its mesh readers and UMesh (topology, orientation,
areas, face metrics), its FlowFV::compute_residual, its explicit solver, its surface functionals and entropy norm,
compiled from its unmodified sources without any stand-in for the mesh (oracle/ref_tier_e.cpp). Held against the
oracle (oracle/orc_mesh.hpp, orc_spatial.hpp), the checker of every GPU parity test:
  mesh arrays: integers bit for bit, metrics to round-off, for every fixture mesh (Gmsh 2 and SU2) and a synthetic one;
  residual + time steps from the mesh FILE: 1e-12; 40 forward-Euler steps: 1e-10; Cl / Cd / entropy norm: 1e-10."""
import os
import subprocess

import numpy as np
import pytest

import orc
from common import ROOT, mesh_path, rel_err_by_component, INVISCID_BCS, VISCOUS_BCS
from fvens_b200 import lib, synth

pytestmark = pytest.mark.skipif(not orc.have_ref_e(), reason="oracle/_ref/libfvens_ref_e.so not built (needs /root/reference)")
MESHES = ["2dcylinder0.msh", "2dcylinder1.msh", "2dcylinder2.msh", "2dcylinderhybrid.msh", "naca0012luo.msh", "NACA0012_inv.su2",
          "NACA0012_lam_hybrid_1.msh", "squarecoarse.msh", "squareunsquad0.msh", "testhybrid.msh", "testperiodic.msh"]


def same_mesh(a, b):
    for k in ("inpoel", "nnode", "bface", "esuel", "elemface", "intfac", "btags"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["coords"], b["coords"])
    assert np.abs(a["area"]/b["area"] - 1).max() < 1e-14
    assert np.abs(a["facemetric"] - b["facemetric"]).max() < 1e-14*np.abs(b["facemetric"]).max()


@pytest.mark.parametrize("mesh", MESHES)
def test_mesh_topology_and_metrics_against_the_reference_mesh_class(mesh):
    om = orc.Mesh.read(mesh_path(mesh))
    rm = orc.RefCase.read(mesh_path(mesh))
    assert (om.npoin, om.nelem, om.nbface, om.naface) == (rm.npoin, rm.nelem, rm.nbface, rm.naface)
    same_mesh(om.arrays(), rm.arrays())


def test_mesh_from_arrays():
    arrs = synth.bump_channel(30, 12)
    same_mesh(orc.Mesh.from_arrays(*arrs).arrays(), orc.RefCase.from_arrays(*arrs).arrays())
    arrs = synth.ogrid_cylinder(24, 10, tri_fraction=0.3)
    same_mesh(orc.Mesh.from_arrays(*arrs).arrays(), orc.RefCase.from_arrays(*arrs).arrays())


@pytest.mark.parametrize("cfg", [
    dict(mesh="NACA0012_inv.su2", flux="ROE", gradient="LEASTSQUARES", recon="VANALBADA"),
    dict(mesh="naca0012luo.msh", flux="HLLC", gradient="GREENGAUSS", recon="WENO", lp=20.0),
    dict(mesh="2dcylinderhybrid.msh", flux="AUSMPLUS", order2=False),
    dict(mesh="NACA0012_lam_hybrid_1.msh", flux="ROE", gradient="LEASTSQUARES", recon="NONE", viscous=True),
])
def test_whole_chain_from_the_mesh_file(cfg):
    mesh = cfg["mesh"]; flux = cfg["flux"]; gradient = cfg.get("gradient", "NONE"); recon = cfg.get("recon", "NONE")
    lp = cfg.get("lp", 1.0); order2 = cfg.get("order2", True); viscous = cfg.get("viscous", False)
    om = orc.Mesh.read(mesh_path(mesh))
    a = om.arrays()
    phys = lib.make_physics(1.4, 0.5, 288.15, 200.0, 0.72, 0.03, viscous, False)
    tags = set(a["btags"].tolist())
    bcs = [(t, lib.BC[ty], v) for (t, ty, v) in (VISCOUS_BCS if viscous else INVISCID_BCS) if t in tags]
    of = orc.Flow(om, phys, lib.FLUX[flux], lib.GRAD[gradient], lib.RECON[recon], lp, order2, 0, bcs)
    rf = orc.RefCase.read(mesh_path(mesh)).flow(phys, flux, gradient, recon, lp, order2, bcs)
    rc, _, _ = of.geometry()
    u = synth.perturbed_state(rc, 1.4, 0.5, 0.03, amp=0.06)
    r0, dt0, _, _ = of.residual(u)
    r1, dt1 = rf.residual(u)
    assert rel_err_by_component(r0, r1) < 1e-12 and np.abs(dt0/dt1 - 1).max() < 1e-12
    # forward Euler from this state
    code0, steps0, hist0, u0 = of.forward_euler(u, 0.3, 1e-30, 40)
    code1, steps1, rel1, abs1, u1 = rf.forward_euler(u, 0.3, 1e-30, 40)
    assert (code0, steps0, code1, steps1) == (1, 40, 1, 40) and rel_err_by_component(u0, u1) < 1e-10
    assert np.abs(hist0/abs1 - 1).max() < 3e-7
    # functionals on the wall marker of the advanced state
    wall = 2
    cl1, cdp1, cdf1, ent1 = rf.surface_and_entropy(u1, wall)
    cl0, cdp0, cdf0 = of.surface_data(u0, of.get_gradients(u0), wall)
    ent0 = of.entropy_error(u0)
    scale = max(abs(cl1), abs(cdp1), 1e-3)
    assert abs(cl0 - cl1) < 1e-9*scale and abs(cdp0 - cdp1) < 1e-9*scale and abs(cdf0 - cdf1) < 1e-9*max(abs(cdf1), 1e-3)
    assert abs(ent0/ent1 - 1) < 1e-9


@pytest.mark.parametrize("mesh", ["2dcylinderhybrid.msh", "NACA0012_inv.su2", "testhybrid.msh"])
def test_product_host_mesh_against_the_reference_mesh_class(mesh):
    """The PRODUCT's host mesh (fvens_b200/host/mesh.hpp + csrc/umesh.cpp, what the device mesh is built from) against
    the reference's UMesh object code directly, also after UMesh::reorder_cells with the same permutation."""
    um = lib.UMesh.read(mesh_path(mesh))
    rm = orc.RefCase.read(mesh_path(mesh))

    def same(a, b):
        mw = a["inpoel"].shape[1]
        for k in ("inpoel", "esuel", "elemface"):
            assert np.array_equal(a[k], b[k][:, :mw]), k
        assert np.array_equal(a["nnode"], b["nnode"]) and np.array_equal(a["intfac"], b["intfac"])
        assert np.array_equal(a["btags"][:, 0], b["btags"]) and np.array_equal(a["coords"], b["coords"])
        assert np.abs(a["area"]/b["area"] - 1).max() < 1e-14
        assert np.abs(a["facemetric"] - b["facemetric"]).max() < 1e-14*np.abs(b["facemetric"]).max()
    same(um.arrays(), rm.arrays())
    perm = np.random.default_rng(11).permutation(um.nelem).astype(np.int32)
    um.reorder_cells(perm)
    rm.reorder_cells(perm)
    same(um.arrays(), rm.arrays())


def _vtu_arrays(path):
    import xml.dom.minidom
    doc = xml.dom.minidom.parse(str(path))
    out = {}
    for d in doc.getElementsByTagName("DataArray"):
        name = d.getAttribute("Name") or "points"
        out[name] = (d.getAttribute("type"), d.firstChild.data)
    piece = doc.getElementsByTagName("Piece")[0]
    return out, int(piece.getAttribute("NumberOfPoints")), int(piece.getAttribute("NumberOfCells"))


def test_output_files_of_the_host_layer_against_the_reference_writers(tmp_path):
    """The artefacts fvens_steady leaves behind - surface files, volume file, VTU - written by the host layer
    (fvens_b200/host/casesolvers.hpp, free functions on host arrays: no GPU) and by the reference's own FlowOutput /
    VTU writer (spatial/aoutput.cpp, tier E) for the same viscous state. Text files: identical, or equal to the printed
    precision where the gradients come from different (1e-13-equal) sources. VTU: identical blocks, except density and
    pressure - the reference divides its point sums of the conserved variables by NVARS times the area sum
    (aoutput.cpp:115-127), the host layer writes the true area-weighted averages."""
    import os
    import subprocess
    from common import ROOT
    mesh = mesh_path("2dcylinderhybrid.msh")
    om = orc.Mesh.read(mesh)
    a = om.arrays()
    phys = lib.make_physics(1.4, 0.5, 288.15, 150.0, 0.72, 0.05, True, False)
    bcs = [(2, lib.BC["adiabaticwall"], (0.0, 0.0)), (4, lib.BC["farfield"], (0.0, 0.0))]
    of = orc.Flow(om, phys, lib.FLUX["ROE"], lib.GRAD["LEASTSQUARES"], lib.RECON["NONE"], 1.0, True, 0, bcs)
    rc, _, _ = of.geometry()
    u = synth.perturbed_state(rc, 1.4, 0.5, 0.05, amp=0.07)
    grads = of.get_gradients(u)
    rf = orc.RefCase.read(mesh).flow(phys, "ROE", "LEASTSQUARES", "NONE", 1.0, True, bcs)
    rdir = tmp_path / "ref"; hdir = tmp_path / "host"
    rdir.mkdir(); hdir.mkdir()
    rf.write_outputs(u, [2], [4], rdir / "case", rdir / "case.vtu", rdir / "case")
    (tmp_path / "u.bin").write_bytes(np.ascontiguousarray(u).tobytes())
    (tmp_path / "g.bin").write_bytes(np.ascontiguousarray(grads).tobytes())
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_controlparser"), "--outputs", mesh, str(tmp_path / "u.bin"),
                        str(tmp_path / "g.bin"), "1.4", "0.5", "288.15", "150.0", "0.72", "0.05", "1", "0", "2", "4", str(hdir / "case")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    # volume file and the "other boundary" surface file: pure functions of mesh and state -> the same text
    assert (hdir / "case-vol.out").read_text() == (rdir / "case-vol.out").read_text()
    assert (hdir / "case-surf_o4.out").read_text() == (rdir / "case-surf_o4.out").read_text()
    # wall surface file: same layout; numbers equal to the 6 printed digits (gradients from the oracle vs the reference)
    hw = (hdir / "case-surf_w2.out").read_text().splitlines(); rw = (rdir / "case-surf_w2.out").read_text().splitlines()
    assert len(hw) == len(rw) and hw[0] == rw[0] and hw[-2] == rw[-2] and len(hw) > 10
    for x, y in zip(hw[1:-2] + [hw[-1][1:]], rw[1:-2] + [rw[-1][1:]]):
        p, q = np.array(x.split(), dtype=float), np.array(y.split(), dtype=float)
        assert np.abs(p - q).max() <= 2e-5*max(np.abs(q).max(), 1e-3)
    assert abs(float(hw[-1].split()[3])) > 1e-4          # a viscous case: the skin-friction drag is there
    # VTU
    hv, hnp, hnc = _vtu_arrays(hdir / "case.vtu"); rv, rnp, rnc = _vtu_arrays(rdir / "case.vtu")
    assert (hnp, hnc) == (rnp, rnc) == (a["coords"].shape[0], om.nelem) and list(hv) == list(rv)
    for name in ("mach-number", "temperature", "velocity", "points", "connectivity", "offsets", "types"):
        assert hv[name] == rv[name], name
    for name in ("density", "pressure"):
        h = np.array(hv[name][1].split(), dtype=float); rr = np.array(rv[name][1].split(), dtype=float)
        assert hv[name][0] == rv[name][0] and np.abs(h/(4.0*rr) - 1).max() < 1e-9


def _read_dist_golden(path):
    lines = open(path).read().split("\n")
    i = [l.strip().lower() for l in lines].index("#connfaces")
    elems = np.array([int(x) for x in lines[1:i] if x.strip()], dtype=np.int32)
    conn = np.array([[int(y) for y in x.split()] for x in lines[i+1:] if x.strip()], dtype=np.int32).reshape(-1, 4)
    return elems, conn


def test_reference_partitioner_reproduces_its_goldens_and_the_product_agrees():
    """tests/mesh/distributedmesh.cpp of the reference, replayed on its own object code (mesh/meshpartitioning.cpp in
    tier E, one process playing the three ranks): the trivial 3-way partition of testhybrid.msh gives the golden
    subdomain meshes (testhybrid_part*.msh), global cell indices and connectivity faces (testhybrid-distb_part*.dat).
    Then the PRODUCT's subdomain construction (fvg_mesh_create_part, host-only build) for the same cell -> rank map: the
    same own cells in the same order, and ghost cells = the neighbours behind the reference's connectivity faces, owned
    by the ranks the reference names."""
    gm = orc.RefCase.read(mesh_path("testhybrid.msh"))
    dist = gm.trivial_partition(3)
    um = lib.UMesh.read(mesh_path("testhybrid.msh"))
    assert sorted(set(dist.tolist())) == [0, 1, 2]
    for r in range(3):
        lm = gm.restrict_to_rank(dist, 3, r)
        gold = orc.RefCase.read(mesh_path(f"testhybrid_part{r+1}.msh"))
        la, ga = lm.arrays(), gold.arrays()
        assert (lm.npoin, lm.nelem, lm.nbface) == (gold.npoin, gold.nelem, gold.nbface)
        assert np.array_equal(la["nnode"], ga["nnode"]) and np.array_equal(la["inpoel"], ga["inpoel"])
        assert np.abs(la["coords"] - ga["coords"]).max() < 1e-12
        elems, conn4 = _read_dist_golden(mesh_path(f"testhybrid-distb_part{r+1}.dat"))
        glob, conn = lm.connectivity()
        assert np.array_equal(glob, elems) and np.array_equal(conn[:, :4], conn4)
        assert (dist[glob] == r).all()
        # the product, for the same distribution
        dm = lib.DeviceMesh(um, reorder="none", tile_cells=32, device=-2, cell_rank=dist, rank=r, nranks=3)
        ids = dm.permutation()
        # same own cells (the device numbers the tiles that see a ghost cell last, so the order is its own)
        assert np.array_equal(np.sort(ids[:dm.ncell]), glob)
        assert set(ids[dm.ncell:].tolist()) == set(conn[:, 3].tolist())
        owner = {int(c): int(rk) for rk, c in zip(conn[:, 2], conn[:, 3])}
        assert all(dist[g] == owner[int(g)] for g in ids[dm.ncell:])


@pytest.mark.parametrize("partition", [lib.partition_sfc, lib.partition_rcb])
@pytest.mark.parametrize("nranks", [2, 5])
def test_product_subdomains_against_the_reference_partitioner_on_a_hilbert_partition(nranks, partition):
    """The space-filling-curve partition bench.py uses and the coordinate-bisection one, restricted to each rank by the reference's own
    restrictMeshToPartitions and by the product: own cells (ascending global index), ghost sets, owners."""
    arrs = synth.bump_channel(40, 16)
    um = lib.UMesh.from_arrays(*arrs)
    part = partition(um, nranks)
    gm = orc.RefCase.from_arrays(*arrs)
    for r in range(nranks):
        glob, conn = gm.restrict_to_rank(part, nranks, r).connectivity()
        dm = lib.DeviceMesh(um, reorder="none", tile_cells=64, device=-2, cell_rank=part, rank=r, nranks=nranks)
        ids = dm.permutation()
        assert np.array_equal(np.sort(ids[:dm.ncell]), glob) and (part[glob] == r).all()
        # tiles that send rows to a neighbour (= tiles that see a ghost cell) are one block late in the device numbering,
        # followed only by the interior tiles that hide the latency of their pushes
        toff, _ = dm.tile_send_lists()
        sends = np.diff(toff) > 0
        first, last = np.argmax(sends), len(sends) - 1 - np.argmax(sends[::-1])
        assert sends.any() and sends[first:last+1].all() and (len(sends) < 4 or first >= (len(sends) - sends.sum())//2)
        assert set(ids[dm.ncell:].tolist()) == set(conn[:, 3].tolist()) and len(conn) >= dm.nghost > 0
        assert all(part[c] == rk for rk, c in zip(conn[:, 2], conn[:, 3]))


@pytest.mark.parametrize("mesh", ["testhybrid.msh", "2dcylinderhybrid.msh", "NACA0012_inv.su2", "NACA0012_lam_hybrid_1.msh", "squarecoarse.msh"])
def test_gmsh_writer_against_the_reference_writer(mesh, tmp_path):
    """UMesh::writeGmsh2 (mesh/mesh.cpp:205-286; what utilities/convertformat.cpp produces, e.g. SU2 -> msh): the host
    layer's file equals the reference's byte for byte, and reading it back gives the same mesh."""
    a, b = str(tmp_path / "ref.msh"), str(tmp_path / "own.msh")
    orc.RefCase.read(mesh_path(mesh)).write_gmsh2(a)
    um = lib.UMesh.read(mesh_path(mesh))
    um.write_gmsh2(b)
    assert open(a, "rb").read() == open(b, "rb").read()
    back = lib.UMesh.read(b).arrays()
    orig = um.arrays()
    for k in ("nnode", "inpoel", "esuel", "intfac"):
        assert np.array_equal(back[k], orig[k]), k
    # an SU2 mesh has one tag per boundary face; the file carries the second one Gmsh requires
    assert np.array_equal(back["btags"][:, 0], orig["btags"][:, 0])
    assert np.array_equal(back["coords"], orig["coords"])           # 20 significant digits: exact round trip
    with pytest.raises(lib.FvgError):
        um.write_gmsh2(str(tmp_path / "no" / "such" / "dir.msh"))


CONVERT = os.path.join(ROOT, "tests", "cpp", "convertformat")


@pytest.mark.parametrize("mesh", ["NACA0012_inv.su2", "squarecoarse.msh", "2dcylinderhybrid.msh"])
def test_convertformat_program(mesh, tmp_path):
    """utilities/convertformat.cpp: <in> <out> msh|vtu. msh: the reference writer's bytes. vtu (writeMeshToVtu,
    spatial/aoutput.cpp:557-615): the reference's bytes on single-cell-type meshes; on a hybrid mesh the reference's
    `offsets` array is nnode(i)*(i+1) instead of the running node count, every other line is the same."""
    a, b = str(tmp_path / "ref"), str(tmp_path / "own")
    r = subprocess.run([CONVERT, mesh_path(mesh), b + ".msh", "msh"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert orc.ref_convertformat(mesh_path(mesh), a + ".msh", "msh") == 0
    assert open(a + ".msh", "rb").read() == open(b + ".msh", "rb").read()
    r = subprocess.run([CONVERT, mesh_path(mesh), b + ".vtu", "vtu"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert orc.ref_convertformat(mesh_path(mesh), a + ".vtu", "vtu") == 0
    la, lb = open(a + ".vtu").read().split("\n"), open(b + ".vtu").read().split("\n")
    nnode = lib.UMesh.read(mesh_path(mesh)).arrays()["nnode"]
    if len(set(nnode.tolist())) == 1:
        assert la == lb
    else:
        assert len(la) == len(lb)
        i0 = lb.index('\t\t\t<DataArray type="UInt32" Name="offsets" Format="ascii">') + 1
        assert la[:i0] == lb[:i0] and la[i0+len(nnode):] == lb[i0+len(nnode):]
        assert [int(x) for x in lb[i0:i0+len(nnode)]] == np.cumsum(nnode).tolist()
        assert [int(x) for x in la[i0:i0+len(nnode)]] == (nnode*np.arange(1, len(nnode)+1)).tolist()
    # usage / bad format
    assert subprocess.run([CONVERT, mesh_path(mesh)], capture_output=True).returncode == 2
    assert subprocess.run([CONVERT, mesh_path(mesh), b + ".x", "stl"], capture_output=True).returncode != 0


@pytest.mark.parametrize("mesh", ["testhybrid.msh", "2dcylinderhybrid.msh", "NACA0012_inv.su2"])
def test_cell_adjacency_graph_against_the_reference(mesh):
    """fvg_umesh_cell_adjacency = getCellAdjLists (mesh/meshpartitioning.cpp:376-430), the graph the reference would
    hand to SCOTCH_graphBuild: same CSR, entry for entry."""
    p0, s0 = orc.RefCase.read(mesh_path(mesh)).cell_adjacency()
    p1, s1 = lib.UMesh.read(mesh_path(mesh)).cell_adjacency()
    assert np.array_equal(p0, p1) and np.array_equal(s0, s1)
