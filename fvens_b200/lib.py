"""ctypes binding of libfvens_b200.so (C ABI declared in include/fvens_b200.h).

This module is plumbing for the Python test-suite and bench.py: it loads the in-tree shared library
and wraps the handles. There is no Python or CPU implementation of any operator here: if the
library is missing, or a device entry point fails, the call raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FVENS_B200_LIB") or os.path.join(_HERE, "libfvens_b200.so")

FLUX = {"LLF": 0, "VANLEER": 1, "AUSM": 2, "AUSMPLUS": 3, "ROE": 4, "HLL": 5, "HLLC": 6}
GRAD = {"NONE": 0, "ZERO": 0, "GREENGAUSS": 1, "LEASTSQUARES": 2}
RECON = {"NONE": 0, "WENO": 1, "VANALBADA": 2, "BARTHJESPERSEN": 3, "VENKATAKRISHNAN": 4}
# spatial/abctypes.hpp:13-22
BC = {"slipwall": 0, "farfield": 1, "inflowoutflow": 2, "subsonicinflow": 3, "extrapolation": 4,
      "periodic": 5, "isothermalwall": 6, "adiabaticwall": 7}
REORDER = {"none": 0, "hilbert": 1, "rcm": 2}

STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "IO", 4: "UNSUPPORTED", 5: "TOLERANCE", 6: "NUMERICAL", 7: "COMM"}


class FvgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fvens_b200 error {code} ({STATUS.get(code, '?')}): {msg}")
        self.code = code


class ToleranceError(FvgError):
    """Solver hit maxiter (reference: Tolerance_error)."""


class NumericalError(FvgError):
    """Non-finite residual (reference: Numerical_error)."""


class HostMeshView(C.Structure):
    _fields_ = [("npoin", C.c_int), ("nelem", C.c_int), ("nbface", C.c_int), ("naface", C.c_int),
                ("ninface", C.c_int), ("nconnface", C.c_int), ("maxnnode", C.c_int), ("nbtag", C.c_int),
                ("coords", C.POINTER(C.c_double)), ("inpoel", C.POINTER(C.c_int)),
                ("nnode", C.POINTER(C.c_int)), ("esuel", C.POINTER(C.c_int)),
                ("elemface", C.POINTER(C.c_int)), ("intfac", C.POINTER(C.c_int)),
                ("btags", C.POINTER(C.c_int)), ("facemetric", C.POINTER(C.c_double)),
                ("area", C.POINTER(C.c_double)), ("bpartner", C.POINTER(C.c_int))]


class MeshOpts(C.Structure):
    _fields_ = [("reorder", C.c_int), ("tile_cells", C.c_int), ("device", C.c_int)]


class MeshInfo(C.Structure):
    _fields_ = [("ncell", C.c_int), ("nbface", C.c_int), ("naface", C.c_int), ("ntile", C.c_int),
                ("tile_cells", C.c_int), ("nstream", C.c_int), ("ncut_dup", C.c_int),
                ("max_colours", C.c_int), ("reorder", C.c_int), ("mean_neighbour_distance", C.c_double),
                ("nghost", C.c_int), ("nsend", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
                ("entry_capacity", C.c_int), ("halo_capacity", C.c_int),
                ("bank_groups", C.c_longlong), ("bank_conflict_groups", C.c_longlong)]


class Physics(C.Structure):
    _fields_ = [("gamma", C.c_double), ("Minf", C.c_double), ("Tinf", C.c_double), ("Reinf", C.c_double),
                ("Pr", C.c_double), ("aoa", C.c_double), ("viscous_sim", C.c_int), ("const_visc", C.c_int)]


class Numerics(C.Structure):
    _fields_ = [("flux", C.c_int), ("gradient", C.c_int), ("reconstruction", C.c_int),
                ("limiter_param", C.c_double), ("order2", C.c_int), ("bnd_policy", C.c_int)]


class BCStruct(C.Structure):
    _fields_ = [("tag", C.c_int), ("type", C.c_int), ("vals", C.c_double*2)]


_lib = None


def load():
    """Loads the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C fvens_b200/csrc`. There is no fallback implementation.")
    lib = C.CDLL(LIB_PATH)
    lib.fvg_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(code):
    if code == 0:
        return
    msg = load().fvg_last_error().decode()
    if code == 5:
        raise ToleranceError(code, msg)
    if code == 6:
        raise NumericalError(code, msg)
    raise FvgError(code, msg)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _ptr(t):
    """Device (torch tensor) or raw integer address -> void*"""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, int):
        return C.c_void_p(t)
    assert t.is_cuda and t.is_contiguous() and str(t.dtype) == "torch.float64"
    return C.c_void_p(t.data_ptr())


def make_physics(gamma=1.4, Minf=0.5, Tinf=288.15, Reinf=5000.0, Pr=0.72, aoa=0.0, viscous=False,
                 const_visc=False):
    return Physics(gamma, Minf, Tinf, Reinf, Pr, aoa, int(viscous), int(const_visc))


def phys_array(p):
    """The oracle's parameter vector {gamma, Minf, Tinf, Reinf, Pr, aoa}"""
    return np.array([p.gamma, p.Minf, p.Tinf, p.Reinf, p.Pr, p.aoa])


class UMesh:
    """Host mesh handle = fvens::UMesh<double,2> after constructMesh/preprocessMesh."""

    def __init__(self, handle):
        self._h = handle
        v = HostMeshView()
        check(load().fvg_umesh_view(self._h, C.byref(v)))
        self.view = v
        self.npoin, self.nelem, self.nbface, self.naface = v.npoin, v.nelem, v.nbface, v.naface
        self.ninface, self.maxnnode = v.ninface, v.maxnnode

    @classmethod
    def read(cls, path):
        h = C.c_void_p()
        check(load().fvg_umesh_read(str(path).encode(), C.byref(h)))
        return cls(h)

    def write_gmsh2(self, path):
        """UMesh::writeGmsh2: Gmsh 2.2 ASCII (the conversion utilities/convertformat.cpp does)."""
        check(load().fvg_umesh_write_gmsh2(self._h, str(path).encode()))

    @classmethod
    def from_arrays(cls, coords, nnode, inpoel, bface):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        nnode = np.ascontiguousarray(nnode, dtype=np.int32)
        inpoel = np.ascontiguousarray(inpoel, dtype=np.int32)
        bface = np.ascontiguousarray(bface, dtype=np.int32)
        assert inpoel.shape == (len(nnode), 4) and coords.shape[1] == 2 and bface.shape[1] == 3
        h = C.c_void_p()
        check(load().fvg_umesh_from_arrays(len(coords), _dp(coords), len(nnode), _ip(nnode), _ip(inpoel),
                                           len(bface), _ip(bface), C.byref(h)))
        return cls(h)

    def _arr(self, ptr, shape, dtype):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).copy()

    def arrays(self):
        """Copies of the derived arrays, names as the reference's accessors."""
        v = self.view
        mw = v.maxnnode
        return dict(
            coords=self._arr(v.coords, (v.npoin, 2), np.float64),
            inpoel=self._arr(v.inpoel, (v.nelem, mw), np.int32),
            nnode=self._arr(v.nnode, (v.nelem,), np.int32),
            esuel=self._arr(v.esuel, (v.nelem, mw), np.int32),
            elemface=self._arr(v.elemface, (v.nelem, mw), np.int32),
            intfac=self._arr(v.intfac, (v.naface, 4), np.int32),
            btags=self._arr(v.btags, (v.nbface, v.nbtag), np.int32),
            facemetric=self._arr(v.facemetric, (v.naface, 3), np.float64),
            area=self._arr(v.area, (v.nelem,), np.float64))

    def compute_periodic_map(self, marker, axis):
        """UMesh::compute_periodic_map: pairs the boundary faces with this marker along `axis` (0: x-periodic, 1: y-periodic);
        returns the number of pairs. Device meshes built afterwards treat the pairs as interior faces."""
        n = C.c_int(0)
        check(load().fvg_umesh_compute_periodic_map(self._h, int(marker), int(axis), C.byref(n)))
        self.__init__(self._h)
        return n.value

    def periodic_partners(self):
        v = self.view
        if not v.bpartner:
            return np.full(v.nbface, -1, dtype=np.int32)
        return self._arr(v.bpartner, (v.nbface,), np.int32)

    def reorder_cells(self, perm):
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        check(load().fvg_umesh_reorder_cells(self._h, _ip(perm)))
        self.__init__(self._h)

    def cell_adjacency(self):
        """CSR adjacency of the cells across interior faces (the reference's getCellAdjLists)."""
        ptrs = np.zeros(self.nelem + 1, dtype=np.int32)
        check(load().fvg_umesh_cell_adjacency(self._h, _ip(ptrs), None))
        store = np.zeros(max(int(ptrs[-1]), 1), dtype=np.int32)
        check(load().fvg_umesh_cell_adjacency(self._h, _ip(ptrs), _ip(store)))
        return ptrs, store[:ptrs[-1]]

    def write_scotch_graph(self, path):
        check(load().fvg_umesh_write_scotch_graph(self._h, str(path).encode()))

    def read_scotch_map(self, path):
        """cell -> part from a Scotch mapping file; returns (cell_rank, nparts)."""
        cr = np.zeros(self.nelem, dtype=np.int32)
        npart = C.c_int(0)
        check(load().fvg_partition_read_scotch_map(self._h, str(path).encode(), _ip(cr), C.byref(npart)))
        return cr, npart.value

    def rcm_ordering(self):
        perm = np.zeros(self.nelem, dtype=np.int32)
        check(load().fvg_umesh_rcm_ordering(self._h, _ip(perm)))
        return perm

    def hilbert_ordering(self):
        perm = np.zeros(self.nelem, dtype=np.int32)
        check(load().fvg_umesh_hilbert_ordering(self._h, _ip(perm)))
        return perm

    def close(self):
        if self._h:
            load().fvg_umesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMesh:
    """Device mesh of the whole host mesh, or (cell_rank given) of one rank's subdomain with its ghost layer."""

    def __init__(self, umesh, reorder="hilbert", tile_cells=0, device=-1, cell_rank=None, rank=0, nranks=1):
        self.umesh = umesh
        opts = MeshOpts(REORDER[reorder] if isinstance(reorder, str) else int(reorder), tile_cells, device)
        h = C.c_void_p()
        if cell_rank is None:
            check(load().fvg_mesh_create(C.byref(umesh.view), C.byref(opts), C.byref(h)))
        else:
            cell_rank = np.ascontiguousarray(cell_rank, dtype=np.int32)
            assert len(cell_rank) == umesh.nelem
            check(load().fvg_mesh_create_part(C.byref(umesh.view), C.byref(opts), _ip(cell_rank), int(rank), int(nranks),
                                              C.byref(h)))
        self._h = h
        info = MeshInfo()
        check(load().fvg_mesh_get_info(self._h, C.byref(info)))
        self.info = info
        self.ncell, self.nbface, self.naface = info.ncell, info.nbface, info.naface
        self.nghost, self.nranks, self.rank = info.nghost, info.nranks, info.rank

    def permutation(self):
        """Global (reference) cell id of every device row: own cells, then ghosts."""
        p = np.zeros(self.ncell + self.nghost, dtype=np.int32)
        check(load().fvg_mesh_permutation(self._h, _ip(p)))
        return p

    def halo_lists(self):
        sc = np.zeros(self.nranks, dtype=np.int32); rc = np.zeros(self.nranks, dtype=np.int32)
        idx = np.zeros(max(self.info.nsend, 1), dtype=np.int32)
        check(load().fvg_mesh_halo_lists(self._h, _ip(sc), _ip(rc), _ip(idx)))
        return sc, rc, idx[:self.info.nsend]

    def halo_pack(self, src, width, sendbuf, stream=None):
        check(load().fvg_halo_pack(self._h, _ptr(src), int(width), _ptr(sendbuf), C.c_void_p(stream or 0)))

    def tile_send_lists(self):
        """(tile_off [ntile+1], triples [n][3] = tile-local cell, peer, row in my block of the peer's ghost range)."""
        off = np.zeros(self.info.ntile + 1, dtype=np.int32)
        check(load().fvg_mesh_tile_send_lists(self._h, _ip(off), None))
        tr = np.zeros((max(int(off[-1]), 1), 3), dtype=np.int32)
        check(load().fvg_mesh_tile_send_lists(self._h, _ip(off), _ip(tr)))
        return off, tr[:off[-1]]

    def tile_offsets(self):
        t = np.zeros(self.info.ntile+1, dtype=np.int32)
        check(load().fvg_mesh_tile_offsets(self._h, _ip(t)))
        return t

    def stream(self):
        n = self.info.nstream
        f, c, t = (np.zeros(n, dtype=np.int32) for _ in range(3))
        check(load().fvg_mesh_stream(self._h, _ip(f), _ip(c), _ip(t)))
        return f, c, t

    def close(self):
        if self._h:
            load().fvg_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FlowFV:
    """Device flow discretisation handle; method names follow the reference's FlowFV."""

    def __init__(self, dmesh, phys, flux="ROE", gradient="LEASTSQUARES", reconstruction="NONE",
                 limiter_param=1.0, order2=True, bnd_policy=0, bcs=()):
        self.dmesh = dmesh
        self.phys = phys
        num = Numerics(FLUX[flux.upper()], GRAD.get(gradient.upper(), 0), RECON[reconstruction.upper()],
                       limiter_param, int(order2), bnd_policy)
        self.num = num
        arr = (BCStruct*max(len(bcs), 1))()
        for i, (tag, typ, vals) in enumerate(bcs):
            arr[i].tag = tag
            arr[i].type = BC[typ] if isinstance(typ, str) else int(typ)
            arr[i].vals[0], arr[i].vals[1] = (list(vals) + [0.0, 0.0])[:2]
        self.bcs = [(a.tag, a.type, (a.vals[0], a.vals[1])) for a in arr[:len(bcs)]]
        h = C.c_void_p()
        check(load().fvg_flow_create(dmesh._h, C.byref(phys), C.byref(num), arr, len(bcs), C.byref(h)))
        self._h = h

    # device-pointer entry points (torch float64 CUDA tensors)
    def compute_residual(self, u, res, gettimesteps=True, dtm=None, accumulate=True, stream=None):
        check(load().fvg_residual(self._h, _ptr(u), _ptr(res), int(accumulate), int(gettimesteps), _ptr(dtm),
                                  C.c_void_p(stream or 0)))

    def compute_residual_host(self, u, res, gettimesteps=True, dtm=None, accumulate=True):
        """Drop-in mode: host arrays (numpy, or raw addresses of pinned buffers) in and out, H2D/D2H inside."""
        def hp(a):
            if a is None:
                return None
            if isinstance(a, int):
                return C.cast(C.c_void_p(a), C.POINTER(C.c_double))
            assert a.dtype == np.float64 and a.flags.c_contiguous
            return _dp(a)
        check(load().fvg_residual_host(self._h, hp(u), hp(res), int(accumulate), int(gettimesteps), hp(dtm)))

    def timing(self, enable=True):
        """Returns (ms gradient pass, ms face pass, evaluations) since the previous call; (re)arms the timers."""
        out = np.zeros(3)
        check(load().fvg_flow_timing(self._h, int(enable), _dp(out)))
        return out[0], out[1], int(out[2])

    def compute_gradients(self, uprim, ug, grad, stream=None):
        check(load().fvg_gradients(self._h, _ptr(uprim), _ptr(ug), _ptr(grad), C.c_void_p(stream or 0)))

    def compute_face_values(self, uprim, ug, grad, ufl, ufr, stream=None):
        check(load().fvg_face_values(self._h, _ptr(uprim), _ptr(ug), _ptr(grad), _ptr(ufl), _ptr(ufr),
                                     C.c_void_p(stream or 0)))

    def compute_boundary_states(self, ins, gs, stream=None):
        check(load().fvg_boundary_states(self._h, _ptr(ins), _ptr(gs), C.c_void_p(stream or 0)))

    def jacobian_vector_product(self, u, res, mdt, x, y, eps=1e-7, stream=None):
        """MatrixFreeSpatialJacobian::apply: y = mdt*x + (r(u + h x) - r(u))/h, h = eps/|x|; res = compute_residual(u)."""
        check(load().fvg_jacobian_vector_product(self._h, _ptr(u), _ptr(res), _ptr(mdt), _ptr(x), C.c_double(eps), _ptr(y),
                                                 C.c_void_p(stream or 0)))

    def getGradients(self, u, grads, stream=None):
        check(load().fvg_get_gradients(self._h, _ptr(u), _ptr(grads), C.c_void_p(stream or 0)))

    def computeSurfaceData(self, u, grads, marker):
        out = np.zeros(3)
        check(load().fvg_surface_data(self._h, _ptr(u), _ptr(grads), int(marker), _dp(out)))
        return tuple(out)

    def entropy_error(self, u):
        out = C.c_double(0)
        check(load().fvg_entropy_error(self._h, _ptr(u), C.byref(out)))
        return out.value

    def euler_step(self, u, cfl, resnorm2=None, stream=None):
        check(load().fvg_euler_step(self._h, _ptr(u), C.c_double(cfl), _ptr(resnorm2), C.c_void_p(stream or 0)))

    def solve_forward_euler(self, u, cfl, tol, maxiter, check_every=1):
        """Returns (status code, steps, history); raises only on hard errors."""
        steps = C.c_int(0)
        hist = np.zeros(max(maxiter, 1))
        code = load().fvg_forward_euler_solve(self._h, _ptr(u), C.c_double(cfl), C.c_double(tol), int(maxiter),
                                              int(check_every), C.byref(steps), _dp(hist))
        if code not in (0, 5, 6):
            check(code)
        return code, steps.value, hist[:steps.value].copy()

    def solve_tvdrk(self, u, order, cfl, finaltime, maxsteps=0):
        """TVDRKSolver::solve as an SSP Runge-Kutta scheme (fvg_tvdrk_solve). Returns (status code, steps, time)."""
        steps = C.c_int(0); time = C.c_double(0)
        code = load().fvg_tvdrk_solve(self._h, _ptr(u), int(order), C.c_double(cfl), C.c_double(finaltime), int(maxsteps),
                                      C.byref(steps), C.byref(time))
        if code not in (0, 6):
            check(code)
        return code, steps.value, time.value

    # split passes for multi-GPU drivers
    def use_buffers(self, lg, gu):
        self._bufs = (lg, gu)      # keep the tensors alive
        check(load().fvg_flow_use_buffers(self._h, _ptr(lg), _ptr(gu)))

    def gradient_pass(self, u, stage=0, stream=None):
        check(load().fvg_gradient_pass(self._h, _ptr(u), int(stage), C.c_void_p(stream or 0)))

    def face_pass(self, u, res, gettimesteps=True, dtm=None, accumulate=False, stream=None):
        check(load().fvg_face_pass(self._h, _ptr(u), _ptr(res), int(accumulate), int(gettimesteps), _ptr(dtm),
                                   C.c_void_p(stream or 0)))

    def euler_face_pass(self, u, unew, cfl, resnorm2=None, stream=None):
        check(load().fvg_euler_face_pass(self._h, _ptr(u), _ptr(unew), C.c_double(cfl), _ptr(resnorm2),
                                         C.c_void_p(stream or 0)))

    def ghost_source(self, which, window=None, token=0):
        """Ghost rows of the state (which=0) / the reconstruction gradients (which=1) come from the halo window of
        exchange `token` (PeerHaloWindow.post) instead of the array; None/0 switches back."""
        check(load().fvg_flow_ghost_source(self._h, int(which), window._h if window is not None else None,
                                           C.c_ulonglong(token)))

    def select_tiles(self, part):
        """Tiles covered by the split passes: 0 all, 1 interior (no ghost cell in sight), 2 partition boundary."""
        check(load().fvg_flow_select_tiles(self._h, int(part)))

    def launch_count(self):
        c = C.c_longlong(0)
        check(load().fvg_flow_launch_count(self._h, C.byref(c)))
        return c.value

    def close(self):
        if self._h:
            load().fvg_flow_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def flux_pointwise(flux, phys, ul, ur, n):
    ul, ur, n = (np.ascontiguousarray(a, dtype=np.float64) for a in (ul, ur, n))
    out = np.zeros_like(ul)
    fid = FLUX[flux.upper()] if isinstance(flux, str) else int(flux)
    check(load().fvg_flux_pointwise(fid, C.byref(phys), len(ul), _dp(ul), _dp(ur), _dp(n), _dp(out)))
    return out


def bc_pointwise(bctype, vals, phys, ins, n):
    ins, n = (np.ascontiguousarray(a, dtype=np.float64) for a in (ins, n))
    out = np.zeros_like(ins)
    bc = BCStruct(0, BC[bctype] if isinstance(bctype, str) else int(bctype))
    bc.vals[0], bc.vals[1] = (list(vals) + [0.0, 0.0])[:2]
    check(load().fvg_bc_pointwise(C.byref(bc), C.byref(phys), len(ins), _dp(ins), _dp(n), _dp(out)))
    return out


def viscous_flux_pointwise(phys, order2, n, rcl, rcr, ucl, ucr, gl, gr, ul, ur):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (n, rcl, rcr, ucl, ucr, gl, gr, ul, ur)]
    out = np.zeros_like(arrs[3])
    check(load().fvg_viscous_flux_pointwise(C.byref(phys), int(order2), len(arrs[0]), *[_dp(a) for a in arrs], _dp(out)))
    return out


def freestream(phys):
    out = np.zeros(4)
    check(load().fvg_freestream(C.byref(phys), _dp(out)))
    return out


class PeerHaloWindow:
    """fvg_halo: this rank's peer-memory halo window (see include/fvens_b200.h). The caller all-gathers the
    handles and receive counts (any transport) and passes them to connect()."""

    def __init__(self, dmesh, max_width=8):
        self.dmesh = dmesh
        h = C.c_void_p()
        check(load().fvg_halo_create(dmesh._h, int(max_width), C.byref(h)))
        self._h = h

    def handle(self):
        buf = (C.c_ubyte*64)()
        check(load().fvg_halo_ipc_handle(self._h, buf))
        return bytes(buf)

    def connect(self, handles, all_recv_counts):
        """handles: nranks x 64 bytes; all_recv_counts: [nranks, nranks] int array (row r = rank r's recv counts)."""
        hb = b"".join(handles)
        arc = np.ascontiguousarray(all_recv_counts, dtype=np.int32)
        check(load().fvg_halo_connect(self._h, C.c_char_p(hb), _ip(arc)))

    def send(self, arr, width, stream=None):
        check(load().fvg_halo_send(self._h, _ptr(arr), int(width), C.c_void_p(stream or 0)))

    def post(self, arr, width, stream=None):
        """Send only; returns the token of this exchange for FlowFV.ghost_source."""
        tok = C.c_ulonglong(0)
        check(load().fvg_halo_post(self._h, _ptr(arr), int(width), C.c_void_p(stream or 0), C.byref(tok)))
        return int(tok.value)

    def recv(self, arr, width, stream=None):
        check(load().fvg_halo_recv(self._h, _ptr(arr), int(width), C.c_void_p(stream or 0)))

    def exchange(self, arr, width, stream=None):
        check(load().fvg_halo_exchange(self._h, _ptr(arr), int(width), C.c_void_p(stream or 0)))

    def status(self):
        v = C.c_ulonglong(0)
        check(load().fvg_halo_status(self._h, C.byref(v)))
        return int(v.value)

    def close(self):
        if self._h:
            load().fvg_halo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistEngine:
    """fvg_dist: the fused multi-GPU evaluation of one rank (include/fvens_b200.h). The caller all-gathers the handles and
    receive counts (any transport) and passes them to connect()."""

    def __init__(self, flow):
        self.flow = flow
        h = C.c_void_p()
        check(load().fvg_dist_create(flow._h, C.byref(h)))
        self._h = h

    def handle(self):
        buf = (C.c_ubyte*64)()
        check(load().fvg_dist_ipc_handle(self._h, buf))
        return bytes(buf)

    def connect(self, handles, all_recv_counts):
        hb = b"".join(handles)
        arc = np.ascontiguousarray(all_recv_counts, dtype=np.int32)
        check(load().fvg_dist_connect(self._h, C.c_char_p(hb), _ip(arc)))

    def residual(self, u, res, gettimesteps=True, dtm=None, accumulate=False, stream=None):
        check(load().fvg_dist_residual(self._h, _ptr(u), _ptr(res), int(accumulate), int(gettimesteps), _ptr(dtm),
                                       C.c_void_p(stream or 0)))

    def euler_step(self, u, unew, cfl, resnorm2=None, stream=None):
        check(load().fvg_dist_euler_step(self._h, _ptr(u), _ptr(unew), C.c_double(cfl), _ptr(resnorm2),
                                         C.c_void_p(stream or 0)))

    def invalidate_state(self):
        check(load().fvg_dist_invalidate_state(self._h))

    def solve_forward_euler(self, u, cfl, tol, maxiter, check_every=1):
        """Returns (status code, steps, global history); raises on hard errors (FVG_ERR_COMM included)."""
        steps = C.c_int(0)
        hist = np.zeros(max(maxiter, 1))
        code = load().fvg_dist_forward_euler_solve(self._h, _ptr(u), C.c_double(cfl), C.c_double(tol), int(maxiter),
                                                   int(check_every), C.byref(steps), _dp(hist))
        if code not in (0, 5, 6):
            check(code)
        return code, steps.value, hist[:steps.value].copy()

    def status(self):
        """Raises FvgError (code 7, COMM) if a wait for a neighbour's rows has timed out."""
        v = C.c_ulonglong(0)
        check(load().fvg_dist_status(self._h, C.byref(v)))
        return int(v.value)

    def counters(self):
        ev = C.c_longlong(0); gr = C.c_longlong(0); dk = C.c_ulonglong(0)
        check(load().fvg_dist_counters(self._h, C.byref(ev), C.byref(gr), C.byref(dk)))
        return ev.value, gr.value, int(dk.value)

    def close(self):
        if self._h:
            load().fvg_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def partition_sfc(umesh, nranks):
    """cell -> rank map: the Hilbert order of the cells cut into nranks equal chunks."""
    part = np.zeros(umesh.nelem, dtype=np.int32)
    check(load().fvg_partition_sfc(umesh._h, int(nranks), _ip(part)))
    return part


def partition_rcb(umesh, nranks):
    """cell -> rank map by recursive coordinate bisection of the cell centres (balanced to one cell for any nranks)."""
    part = np.zeros(umesh.nelem, dtype=np.int32)
    check(load().fvg_partition_rcb(umesh._h, int(nranks), _ip(part)))
    return part


def device_count():
    c = C.c_int(0)
    check(load().fvg_device_count(C.byref(c)))
    return c.value


def tvdrk_coefficients(order):
    """initialize_TVDRK_Coeffs (ode/aodesolver.cpp:45-67): [order][3] weights of u^n, the stage state and the update."""
    c = np.zeros((int(order), 3))
    check(load().fvg_tvdrk_coefficients(int(order), _dp(c)))
    return c
