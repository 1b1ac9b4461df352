"""ctypes wrapper of the CPU oracle (oracle/liborc.so) and of the reference builds (oracle/_ref/libfvens_ref_a.so:
the reference's gas-dynamics sources; libfvens_ref_b.so: its gradient and reconstruction sources).
TEST INFRASTRUCTURE ONLY - never imported by fvens_b200/."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_PATH = os.path.join(ROOT, "oracle", "liborc.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libfvens_ref_a.so")
REFB_PATH = os.path.join(ROOT, "oracle", "_ref", "libfvens_ref_b.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_orc = None
_ref = None


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def orc():
    global _orc
    if _orc is None:
        _orc = C.CDLL(ORC_PATH)
        _orc.orc_mesh_read.restype = C.c_void_p
        _orc.orc_mesh_from_arrays.restype = C.c_void_p
        _orc.orc_flow_create.restype = C.c_void_p
        _orc.orc_flow_entropy.restype = C.c_double
    return _orc


def have_ref():
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_PATH)
    return _ref


_refb = None


def have_ref_b():
    return os.path.exists(REFB_PATH)


def ref_b():
    global _refb
    if _refb is None:
        _refb = C.CDLL(REFB_PATH)
    return _refb


def _ref_mesh_args(a):
    """Mesh arrays (Mesh.arrays()) in the argument order of oracle/ref_tier_b.cpp; keeps the converted arrays alive."""
    sizes = np.array([a["coords"].shape[0], len(a["nnode"]), len(a["btags"]), a["intfac"].shape[0], 4], dtype=np.int32)
    keep = [sizes, np.ascontiguousarray(a["coords"]), np.ascontiguousarray(a["nnode"], dtype=np.int32),
            np.ascontiguousarray(a["inpoel"], dtype=np.int32), np.ascontiguousarray(a["esuel"], dtype=np.int32),
            np.ascontiguousarray(a["elemface"], dtype=np.int32), np.ascontiguousarray(a["intfac"], dtype=np.int32),
            np.ascontiguousarray(a["facemetric"]), np.ascontiguousarray(a["area"])]
    args = [_ip(keep[0]), _dp(keep[1]), _ip(keep[2]), _ip(keep[3]), _ip(keep[4]), _ip(keep[5]), _ip(keep[6]), _dp(keep[7]), _dp(keep[8])]
    return args, keep


def ref_gradients(a, gradient, rc, rcbp, u, ug):
    """The reference's own GradientScheme classes (spatial/agradientschemes.cpp compiled in place): [nelem][8]."""
    args, keep = _ref_mesh_args(a)
    rc, rcbp, u, ug = (np.ascontiguousarray(x, dtype=np.float64) for x in (rc, rcbp, u, ug))
    grad = np.zeros((len(a["nnode"]), 8))
    ref_b().ref_gradients(int(gradient), *args, _dp(rc), _dp(rcbp), _dp(u), _dp(ug), _dp(grad))
    return grad


def ref_face_values(a, recon, param, rc, rcbp, gr, u, ug, grad, fill=np.nan):
    """The reference's own SolutionReconstruction classes (areconstruction.cpp, limitedlinearreconstruction.cpp,
    musclreconstruction.cpp compiled in place): ufl, ufr [naface][4]; entries a class does not write keep `fill`."""
    args, keep = _ref_mesh_args(a)
    rc, rcbp, gr, u, ug, grad = (np.ascontiguousarray(x, dtype=np.float64) for x in (rc, rcbp, gr, u, ug, grad))
    nf = a["intfac"].shape[0]
    ufl = np.full((nf, 4), fill); ufr = np.full((nf, 4), fill)
    ref_b().ref_face_values(int(recon), C.c_double(param), *args, _dp(rc), _dp(rcbp), _dp(gr), _dp(u), _dp(ug), _dp(grad),
                            _dp(ufl), _dp(ufr))
    return ufl, ufr


REFC_PATH = os.path.join(ROOT, "oracle", "_ref", "libfvens_ref_c.so")
_refc = None


def have_ref_c():
    return os.path.exists(REFC_PATH)


def ref_residual(a, p, flux, gradient, recon, limiter_param, order2, bcs, u, gettimesteps=True):
    """The reference's own FlowFV::compute_residual (oracle/ref_tier_c.cpp). flux / gradient / recon: the control
    file keys (upper case); bcs: (tag, type id, (v0, v1)). Returns (residual [nelem][4], dt [nelem])."""
    global _refc
    if _refc is None:
        _refc = C.CDLL(REFC_PATH)
    args, keep = _ref_mesh_args(a)
    btags = np.ascontiguousarray(a["btags"], dtype=np.int32).reshape(-1)
    ph = np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr, p.aoa], dtype=np.float64)
    tt = np.array([[t, ty] for (t, ty, v) in bcs], dtype=np.int32).reshape(-1)
    vv = np.array([[v[0], v[1]] for (t, ty, v) in bcs], dtype=np.float64).reshape(-1)
    u = np.ascontiguousarray(u, dtype=np.float64)
    n = len(a["nnode"])
    res = np.zeros((n, 4)); dtm = np.zeros(n)
    rc = _refc.ref_residual(*args, _ip(btags), _dp(ph), flux.encode(), gradient.encode(), recon.encode(), C.c_double(limiter_param),
                            int(order2), int(p.viscous_sim), int(p.const_visc), len(bcs), _ip(tt), _dp(vv), _dp(u),
                            int(gettimesteps), _dp(res), _dp(dtm))
    if rc != 0:
        raise RuntimeError(f"reference compute_residual returned {rc}")
    return res, dtm


REFC_OMP_PATH = os.path.join(ROOT, "oracle", "_ref", "libfvens_ref_c_omp.so")
REFC_OMP_AVX2_PATH = os.path.join(ROOT, "oracle", "_ref", "libfvens_ref_c_omp_avx2.so")


def host_has_avx2_fma():
    try:
        flags = next(ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")).split()
    except (OSError, StopIteration):
        return False
    return "avx2" in flags and "fma" in flags


class RefFlow:
    """The reference's own FlowFV object (oracle/ref_tier_c.cpp), kept alive for repeated evaluations: the CPU arm of
    bench.py. omp=True loads the build with the reference's OpenMP pragmas enabled."""

    def __init__(self, a, p, flux, gradient, recon, limiter_param, order2, bcs, omp=True, path=None):
        path = path or (REFC_OMP_PATH if omp else REFC_PATH)
        self.lib = C.CDLL(path)
        self.lib.ref_flow_create.restype = C.c_void_p
        args, keep = _ref_mesh_args(a)
        btags = np.ascontiguousarray(a["btags"], dtype=np.int32).reshape(-1)
        ph = np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr, p.aoa], dtype=np.float64)
        tt = np.array([[t, ty] for (t, ty, v) in bcs], dtype=np.int32).reshape(-1)
        vv = np.array([[v[0], v[1]] for (t, ty, v) in bcs], dtype=np.float64).reshape(-1)
        self.n = len(a["nnode"])
        self.h = C.c_void_p(self.lib.ref_flow_create(*args, _ip(btags), _dp(ph), flux.encode(), gradient.encode(), recon.encode(),
                                                     C.c_double(limiter_param), int(order2), int(p.viscous_sim), int(p.const_visc),
                                                     len(bcs), _ip(tt), _dp(vv)))

    def threads(self):
        return int(self.lib.ref_num_threads())

    def residual(self, u, gettimesteps=True, want=True):
        u = np.ascontiguousarray(u, dtype=np.float64)
        res = np.zeros((self.n, 4)) if want else None
        dtm = np.zeros(self.n) if want else None
        rc = self.lib.ref_flow_residual(self.h, _dp(u), int(gettimesteps), _dp(res) if want else None, _dp(dtm) if want else None)
        if rc != 0:
            raise RuntimeError(f"reference compute_residual returned {rc}")
        return res, dtm

    def __del__(self):
        try:
            self.lib.ref_flow_destroy(self.h)
        except Exception:
            pass


REFD_PATH = os.path.join(ROOT, "oracle", "_ref", "libfvens_ref_d.so")


def have_ref_d():
    return os.path.exists(REFD_PATH)


class RefSolver(RefFlow):
    """Tier D: the reference's own SteadyForwardEulerSolver::solve on the reference's own FlowFV (oracle/ref_tier_d.cpp)."""

    def __init__(self, a, p, flux, gradient, recon, limiter_param, order2, bcs):
        # the handle functions are the same symbols in the tier-D library
        RefFlow.__init__(self, a, p, flux, gradient, recon, limiter_param, order2, bcs, omp=False, path=REFD_PATH)

    def forward_euler(self, u, cfl, tol, maxiter):
        """Returns (code, steps, relative history, absolute history, final state); code 0 converged, 1 max iterations."""
        u = np.array(u, dtype=np.float64, copy=True)
        steps = C.c_int(0)
        hrel = np.zeros(max(maxiter, 1)); habs = np.zeros(max(maxiter, 1))
        code = self.lib.ref_flow_forward_euler(self.h, C.c_double(cfl), C.c_double(tol), int(maxiter), _dp(u), C.byref(steps),
                                               _dp(hrel), _dp(habs))
        return code, steps.value, hrel[:steps.value], habs[:steps.value], u

    def tvdrk(self, u, order, cfl, finaltime, logfile=os.devnull):
        """The reference's TVDRKSolver::solve as it is. Returns (code, final state)."""
        u = np.array(u, dtype=np.float64, copy=True)
        code = self.lib.ref_flow_tvdrk(self.h, int(order), C.c_double(cfl), C.c_double(finaltime), logfile.encode(), _dp(u))
        return code, u

    def history_text(self, steps, rel, abs_, wtime, cfl):
        f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
        st = np.ascontiguousarray(steps, dtype=np.int32)
        rel, abs_, wtime, cfl = f32(rel), f32(abs_), f32(wtime), f32(cfl)
        buf = C.create_string_buffer(200*(len(st) + 3))
        fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        n = self.lib.ref_convergence_history_text(len(st), _ip(st), fp(rel), fp(abs_), fp(wtime), fp(cfl), buf, len(buf))
        assert n >= 0
        return buf.value.decode()


REFE_PATH = os.path.join(ROOT, "oracle", "_ref", "libfvens_ref_e.so")
_refe = None


def have_ref_e():
    return os.path.exists(REFE_PATH)


def _load_ref_e():
    global _refe
    if _refe is None:
        _refe = C.CDLL(REFE_PATH)
        _refe.ref_e_mesh_read.restype = C.c_void_p
        _refe.ref_e_mesh_from_arrays.restype = C.c_void_p
        _refe.ref_e_restrict_to_rank.restype = C.c_void_p
    return _refe


def ref_convertformat(inmesh, outmesh, fmt):
    """The reference's utilities/convertformat.cpp (tier E)."""
    return _load_ref_e().ref_e_convertformat(str(inmesh).encode(), str(outmesh).encode(), fmt.encode())


class RefCase:
    """Tier E: the reference's own UMesh (readers, topology, metrics), FlowFV, explicit solver and output unit, all
    compiled from its unmodified sources (oracle/ref_tier_e.cpp). No stand-in for the mesh."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference mesh construction failed")
        self.h = C.c_void_p(handle)
        s = np.zeros(5, dtype=np.int32)
        self.lib.ref_e_mesh_sizes(self.h, _ip(s))
        self.npoin, self.nelem, self.nbface, self.naface, self.maxnnode = (int(x) for x in s)

    @staticmethod
    def _lib():
        return _load_ref_e()

    lib = property(lambda self: type(self)._lib())

    @classmethod
    def read(cls, path):
        return cls(cls._lib().ref_e_mesh_read(str(path).encode()))

    @classmethod
    def from_arrays(cls, coords, nnode, inpoel, bface):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        nnode = np.ascontiguousarray(nnode, dtype=np.int32)
        inpoel = np.ascontiguousarray(inpoel, dtype=np.int32)
        bface = np.ascontiguousarray(bface, dtype=np.int32)
        return cls(cls._lib().ref_e_mesh_from_arrays(len(coords), _dp(coords), len(nnode), inpoel.shape[1], _ip(nnode), _ip(inpoel),
                                                     len(bface), _ip(bface)))

    def arrays(self):
        d = dict(coords=np.zeros((self.npoin, 2)), inpoel=np.zeros((self.nelem, 4), dtype=np.int32),
                 nnode=np.zeros(self.nelem, dtype=np.int32), bface=np.zeros((self.nbface, 3), dtype=np.int32),
                 esuel=np.zeros((self.nelem, 4), dtype=np.int32), elemface=np.zeros((self.nelem, 4), dtype=np.int32),
                 intfac=np.zeros((self.naface, 4), dtype=np.int32), btags=np.zeros(self.nbface, dtype=np.int32),
                 facemetric=np.zeros((self.naface, 3)), area=np.zeros(self.nelem))
        self.lib.ref_e_mesh_get(self.h, _dp(d["coords"]), _ip(d["inpoel"]), _ip(d["nnode"]), _ip(d["bface"]), _ip(d["esuel"]),
                                _ip(d["elemface"]), _ip(d["intfac"]), _ip(d["btags"]), _dp(d["facemetric"]), _dp(d["area"]))
        return d

    def write_gmsh2(self, path):
        self.lib.ref_e_write_gmsh2(self.h, str(path).encode())

    def write_mesh_vtu(self, path):
        self.lib.ref_e_write_mesh_vtu(self.h, str(path).encode())

    def cell_adjacency(self):
        """getCellAdjLists of the reference: (ptrs [nelem+1], store)."""
        n = self.lib.ref_e_cell_adjacency(self.h, None, None)
        ptrs = np.zeros(self.nelem + 1, dtype=np.int32); store = np.zeros(max(n, 1), dtype=np.int32)
        self.lib.ref_e_cell_adjacency(self.h, _ip(ptrs), _ip(store))
        return ptrs, store[:n]

    def trivial_partition(self, nranks):
        """cell -> rank of the reference's TrivialReplicatedGlobalMeshPartitioner."""
        dist = np.zeros(self.nelem, dtype=np.int32)
        self.lib.ref_e_trivial_partition(self.h, int(nranks), _ip(dist))
        return dist

    def restrict_to_rank(self, dist, nranks, rank):
        """The subdomain mesh the reference's restrictMeshToPartitions builds for `rank` (a new RefCase, mesh only)."""
        dist = np.ascontiguousarray(dist, dtype=np.int32)
        self.lib.ref_e_restrict_to_rank.restype = C.c_void_p
        return RefCase(self.lib.ref_e_restrict_to_rank(self.h, _ip(dist), int(nranks), int(rank)))

    def connectivity(self):
        """(global cell ids [nelem], connectivity faces [nconn][5]: local cell, local face, neighbour rank, neighbour's global
        cell, global face id)."""
        n = self.lib.ref_e_connectivity(self.h, None, None)
        glob = np.zeros(self.nelem, dtype=np.int32); conn = np.zeros((max(n, 1), 5), dtype=np.int32)
        self.lib.ref_e_connectivity(self.h, _ip(glob), _ip(conn))
        return glob, conn[:n]

    def reorder_cells(self, perm):
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        assert len(perm) == self.nelem
        self.lib.ref_e_mesh_reorder(self.h, _ip(perm))

    def flow(self, p, flux, gradient, recon, limiter_param, order2, bcs):
        ph = np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr, p.aoa], dtype=np.float64)
        tt = np.array([[t, ty] for (t, ty, v) in bcs], dtype=np.int32).reshape(-1)
        vv = np.array([[v[0], v[1]] for (t, ty, v) in bcs], dtype=np.float64).reshape(-1)
        rc = self.lib.ref_e_flow_create(self.h, _dp(ph), flux.encode(), gradient.encode(), recon.encode(), C.c_double(limiter_param),
                                        int(order2), int(p.viscous_sim), int(p.const_visc), len(bcs), _ip(tt), _dp(vv))
        if rc != 0:
            raise RuntimeError("reference FlowFV construction failed")
        self.phys = p
        return self

    def flow_b200(self, p, flux, gradient, recon, limiter_param, order2, bcs):
        """The Spatial object becomes the reference-side binding FlowFV_B200 (RefBindingCase only): compute_residual goes
        to the GPU through libfvens_b200, everything else stays the reference's."""
        ph = np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr, p.aoa], dtype=np.float64)
        tt = np.array([[t, ty] for (t, ty, v) in bcs], dtype=np.int32).reshape(-1)
        vv = np.array([[v[0], v[1]] for (t, ty, v) in bcs], dtype=np.float64).reshape(-1)
        rc = self.lib.ref_e_flow_create_b200(self.h, _dp(ph), flux.encode(), gradient.encode(), recon.encode(),
                                             C.c_double(limiter_param), int(order2), int(p.viscous_sim), int(p.const_visc),
                                             len(bcs), _ip(tt), _dp(vv))
        if rc != 0:
            self.lib.ref_binding_error.restype = C.c_char_p
            raise RuntimeError("FlowFV_B200: " + self.lib.ref_binding_error().decode())
        self.phys = p
        return self

    def residual(self, u, gettimesteps=True):
        u = np.ascontiguousarray(u, dtype=np.float64)
        res = np.zeros((self.nelem, 4)); dtm = np.zeros(self.nelem)
        rc = self.lib.ref_e_flow_residual(self.h, _dp(u), int(gettimesteps), _dp(res), _dp(dtm))
        if rc != 0:
            raise RuntimeError(f"reference compute_residual returned {rc}")
        return res, dtm

    def forward_euler(self, u, cfl, tol, maxiter):
        u = np.array(u, dtype=np.float64, copy=True)
        steps = C.c_int(0)
        hrel = np.zeros(max(maxiter, 1)); habs = np.zeros(max(maxiter, 1))
        code = self.lib.ref_e_flow_forward_euler(self.h, C.c_double(cfl), C.c_double(tol), int(maxiter), _dp(u), C.byref(steps),
                                                 _dp(hrel), _dp(habs))
        return code, steps.value, hrel[:steps.value], habs[:steps.value], u

    def forward_euler_b200(self, u, cfl, tol, maxiter):
        """The binding's device-resident driver SteadyForwardEulerSolver_B200 (RefBindingCase after flow_b200)."""
        u = np.array(u, dtype=np.float64, copy=True)
        steps = C.c_int(0)
        hrel = np.zeros(max(maxiter, 1)); habs = np.zeros(max(maxiter, 1))
        code = self.lib.ref_e_flow_forward_euler_b200(self.h, C.c_double(cfl), C.c_double(tol), int(maxiter), _dp(u), C.byref(steps),
                                                      _dp(hrel), _dp(habs))
        if code == 3:
            self.lib.ref_binding_error.restype = C.c_char_p
            raise RuntimeError(self.lib.ref_binding_error().decode())
        return code, steps.value, hrel[:steps.value], habs[:steps.value], u

    def write_outputs(self, u, walls, others, basename, vtufile, volprefix):
        """The reference's own surface / VTU / volume files for the state u."""
        p = self.phys
        ph = np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr], dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        w = np.ascontiguousarray(walls, dtype=np.int32); ot = np.ascontiguousarray(others, dtype=np.int32)
        rc = self.lib.ref_e_write_outputs(self.h, _dp(u), C.c_double(p.aoa), _dp(ph), len(w), _ip(w), len(ot), _ip(ot),
                                          str(basename).encode(), str(vtufile).encode(), str(volprefix).encode())
        if rc != 0:
            raise RuntimeError(f"reference output failed ({rc})")

    def surface_and_entropy(self, u, marker):
        """(Cl, Cdp, Cdf, entropy norm) from the reference's computeSurfaceData / getGradients / FlowOutput."""
        p = self.phys
        ph = np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr], dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros(4)
        rc = self.lib.ref_e_surface_and_entropy(self.h, _dp(u), int(marker), C.c_double(p.aoa), _dp(ph), _dp(out))
        if rc != 0:
            raise RuntimeError("reference surface data failed")
        return tuple(out)

    def __del__(self):
        try:
            self.lib.ref_e_destroy(self.h)
        except Exception:
            pass


def have_ref_c_omp():
    return os.path.exists(REFC_OMP_PATH)


def set_threads(n):
    orc().orc_set_num_threads(int(n))


def num_threads():
    return orc().orc_num_threads()


def phys5(p):
    return np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr], dtype=np.float64)


def _pw(lib, prefix):
    return lambda name: getattr(lib, prefix + name)


def flux(lib_kind, flux_id, p, ul, ur, n):
    lib, pre = (orc(), "orc_") if lib_kind == "orc" else (ref(), "ref_")
    ul, ur, n = (np.ascontiguousarray(a, dtype=np.float64) for a in (ul, ur, n))
    out = np.zeros_like(ul)
    ph = phys5(p)
    getattr(lib, pre + "flux")(int(flux_id), _dp(ph), len(ul), _dp(ul), _dp(ur), _dp(n), _dp(out))
    return out


def ghost_state(lib_kind, bc_type, vals, p, ins, n):
    lib, pre = (orc(), "orc_") if lib_kind == "orc" else (ref(), "ref_")
    ins, n = (np.ascontiguousarray(a, dtype=np.float64) for a in (ins, n))
    out = np.zeros_like(ins)
    ph = phys5(p)
    v = np.array((list(vals) + [0.0, 0.0])[:2], dtype=np.float64)
    getattr(lib, pre + "ghost_state")(int(bc_type), _dp(v), _dp(ph), C.c_double(p.aoa), len(ins), _dp(ins), _dp(n), _dp(out))
    return out


def viscous_flux(p, order2, n, rcl, rcr, ucl, ucr, gl, gr, ul, ur):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (n, rcl, rcr, ucl, ucr, gl, gr, ul, ur)]
    out = np.zeros_like(arrs[3])
    ph = phys5(p)
    orc().orc_viscous_flux(_dp(ph), int(order2), int(p.const_visc), len(arrs[0]), *[_dp(a) for a in arrs], _dp(out))
    return out


def cons2prim(lib_kind, p, u):
    lib, pre = (orc(), "orc_") if lib_kind == "orc" else (ref(), "ref_")
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    ph = phys5(p)
    getattr(lib, pre + "cons2prim")(_dp(ph), len(u), _dp(u), _dp(out))
    return out


def prim2cons(lib_kind, p, u):
    lib, pre = (orc(), "orc_") if lib_kind == "orc" else (ref(), "ref_")
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    ph = phys5(p)
    getattr(lib, pre + "prim2cons")(_dp(ph), len(u), _dp(u), _dp(out))
    return out


def freestream(lib_kind, p):
    lib, pre = (orc(), "orc_") if lib_kind == "orc" else (ref(), "ref_")
    out = np.zeros(4)
    ph = phys5(p)
    getattr(lib, pre + "freestream")(_dp(ph), C.c_double(p.aoa), _dp(out))
    return out


class Mesh:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle mesh construction failed")
        self.h = C.c_void_p(handle)
        s = np.zeros(5, dtype=np.int32)
        orc().orc_mesh_sizes(self.h, _ip(s))
        self.npoin, self.nelem, self.nbface, self.naface, self.ninface = (int(x) for x in s)

    @classmethod
    def read(cls, path):
        return cls(orc().orc_mesh_read(str(path).encode()))

    @classmethod
    def from_arrays(cls, coords, nnode, inpoel, bface):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        nnode = np.ascontiguousarray(nnode, dtype=np.int32)
        inpoel = np.ascontiguousarray(inpoel, dtype=np.int32)
        bface = np.ascontiguousarray(bface, dtype=np.int32)
        return cls(orc().orc_mesh_from_arrays(len(coords), _dp(coords), len(nnode), _ip(nnode), _ip(inpoel),
                                              len(bface), _ip(bface)))

    def arrays(self):
        d = dict(coords=np.zeros((self.npoin, 2)), inpoel=np.zeros((self.nelem, 4), dtype=np.int32),
                 nnode=np.zeros(self.nelem, dtype=np.int32), bface=np.zeros((self.nbface, 3), dtype=np.int32),
                 esuel=np.zeros((self.nelem, 4), dtype=np.int32), elemface=np.zeros((self.nelem, 4), dtype=np.int32),
                 intfac=np.zeros((self.naface, 4), dtype=np.int32), btags=np.zeros(self.nbface, dtype=np.int32),
                 facemetric=np.zeros((self.naface, 3)), area=np.zeros(self.nelem))
        orc().orc_mesh_get(self.h, _dp(d["coords"]), _ip(d["inpoel"]), _ip(d["nnode"]), _ip(d["bface"]),
                           _ip(d["esuel"]), _ip(d["elemface"]), _ip(d["intfac"]), _ip(d["btags"]),
                           _dp(d["facemetric"]), _dp(d["area"]))
        return d

    def __del__(self):
        try:
            orc().orc_mesh_free(self.h)
        except Exception:
            pass


class Flow:
    """Oracle FlowFV. bcs: list of (tag, type_id, (v0, v1))."""

    def __init__(self, mesh, p, flux=4, gradient=2, recon=0, limiter_param=1.0, order2=True, bnd_policy=0, bcs=()):
        self.mesh = mesh
        ph = np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr, p.aoa], dtype=np.float64)
        io = np.array([p.viscous_sim, p.const_visc, flux, gradient, recon, int(order2), bnd_policy], dtype=np.int32)
        bt = np.array([[t, ty] for (t, ty, _) in bcs], dtype=np.int32).reshape(-1, 2)
        bv = np.array([(list(v) + [0.0, 0.0])[:2] for (_, _, v) in bcs], dtype=np.float64).reshape(-1, 2)
        self.h = C.c_void_p(orc().orc_flow_create(mesh.h, _dp(ph), _ip(io), C.c_double(limiter_param), len(bcs),
                                                  _ip(bt), _dp(bv)))

    def geometry(self):
        m = self.mesh
        rc = np.zeros((m.nelem, 2)); gr = np.zeros((m.naface, 2)); rcbp = np.zeros((m.nbface, 2))
        orc().orc_flow_geometry(self.h, _dp(rc), _dp(gr), _dp(rcbp), None, None)
        return rc, gr, rcbp

    def residual(self, u, gettimesteps=True, want_grad=False, want_faces=False):
        m = self.mesh
        u = np.ascontiguousarray(u, dtype=np.float64)
        res = np.zeros((m.nelem, 4)); dtm = np.zeros(m.nelem)
        grad = np.zeros((m.nelem, 8)) if want_grad else None
        faces = np.zeros((2, m.naface, 4)) if want_faces else None
        rc = orc().orc_flow_residual(self.h, _dp(u), _dp(res), int(gettimesteps), _dp(dtm),
                                     _dp(grad) if want_grad else None, _dp(faces) if want_faces else None)
        if rc != 0:
            raise RuntimeError("oracle residual failed")
        return res, dtm, grad, faces

    def gradients(self, uprim, ug):
        grad = np.zeros((self.mesh.nelem, 8))
        uprim = np.ascontiguousarray(uprim); ug = np.ascontiguousarray(ug)
        orc().orc_flow_gradients(self.h, _dp(uprim), _dp(ug), _dp(grad))
        return grad

    def face_values(self, uprim, ug, grad):
        m = self.mesh
        ufl = np.zeros((m.naface, 4)); ufr = np.zeros((m.naface, 4))
        uprim = np.ascontiguousarray(uprim); ug = np.ascontiguousarray(ug); grad = np.ascontiguousarray(grad)
        orc().orc_flow_face_values(self.h, _dp(uprim), _dp(ug), _dp(grad), _dp(ufl), _dp(ufr))
        return ufl, ufr

    def boundary_states(self, ins):
        ins = np.ascontiguousarray(ins)
        gs = np.zeros_like(ins)
        orc().orc_flow_boundary_states(self.h, _dp(ins), _dp(gs))
        return gs

    def get_gradients(self, u):
        u = np.ascontiguousarray(u)
        g = np.zeros((self.mesh.nelem, 8))
        orc().orc_flow_get_gradients(self.h, _dp(u), _dp(g))
        return g

    def surface_data(self, u, grads, marker):
        out = np.zeros(3)
        u = np.ascontiguousarray(u); grads = np.ascontiguousarray(grads)
        orc().orc_flow_surface_data(self.h, _dp(u), _dp(grads), int(marker), _dp(out))
        return out

    def entropy_error(self, u):
        u = np.ascontiguousarray(u)
        return orc().orc_flow_entropy(self.h, _dp(u))

    def forward_euler(self, u, cfl, tol, maxiter):
        u = np.ascontiguousarray(u, dtype=np.float64).copy()
        steps = C.c_int(0)
        hist = np.zeros(max(maxiter, 1))
        code = orc().orc_forward_euler(self.h, _dp(u), C.c_double(cfl), C.c_double(tol), int(maxiter),
                                       C.byref(steps), _dp(hist))
        return code, steps.value, hist[:steps.value].copy(), u

    def __del__(self):
        try:
            orc().orc_flow_free(self.h)
        except Exception:
            pass


REFBIND_PATH = os.path.join(ROOT, "oracle", "_ref", "libfvens_ref_binding.so")
_refbind = None


def have_ref_binding():
    return os.path.exists(REFBIND_PATH)


class RefBindingCase(RefCase):
    """Tier E plus the reference-side binding (oracle/ref_binding.cpp): the reference's mesh, solver and output code with
    FlowFV_B200 (fvens_b200/host/reference_binding/flow_spatial_b200.hpp) as the Spatial object."""

    @staticmethod
    def _lib():
        global _refbind
        if _refbind is None:
            _refbind = C.CDLL(REFBIND_PATH)
            _refbind.ref_e_mesh_read.restype = C.c_void_p
            _refbind.ref_e_mesh_from_arrays.restype = C.c_void_p
            _refbind.ref_e_restrict_to_rank.restype = C.c_void_p
        return _refbind
