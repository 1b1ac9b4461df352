/* fvens_b200 C ABI: the B200 (sm_100a) residual engine behind the FVENS class surface.
 *
 * FVENS itself has no C API or FFI: its operator boundary is a set of C++ abstract classes plus
 * string-keyed factories. Each entry point below replaces one of those C++
 * interfaces 1:1 and cites it (paths relative to the reference's src/). The C++ host classes in
 * fvens_b200/host/ (same names and signatures as the reference) are thin wrappers over this ABI.
 *
 * Conventions: every function returns 0 on success or an fvg_status code, never throws, never
 * takes ownership of caller memory. `d_` pointers are device memory on the device the mesh was
 * created on, `h_` pointers are host memory. `stream` is a cudaStream_t passed as void* (NULL =
 * the default stream); device-pointer calls are asynchronous on that stream unless stated.
 * State, residual, gradient and face arrays use the reference's layouts:
 *   u, res     [nelem][4]  conserved (rho, rho vx, rho vy, rho E)     (PETSc Vec with block size 4)
 *   dtm        [nelem]
 *   grad       [nelem][8]  GradBlock_t col-major 2x4: grad[8*i + idim + 2*ivar]  (aconstants.hpp:80-90)
 *   ufl, ufr   [naface][4] face states in the reference's face numbering (intfac order)
 *   ug         [nbface][4] boundary ghost states
 */
#ifndef FVENS_B200_H
#define FVENS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
	FVG_OK = 0,
	FVG_ERR_INVALID = 1,      /* bad argument / inconsistent mesh / unknown BC tag */
	FVG_ERR_CUDA = 2,         /* CUDA runtime error (see fvg_last_error) */
	FVG_ERR_IO = 3,           /* mesh file could not be read */
	FVG_ERR_UNSUPPORTED = 4,  /* option the engine does not implement */
	FVG_ERR_TOLERANCE = 5,    /* solver hit maxiter   (reference: Tolerance_error, aerrorhandling.hpp) */
	FVG_ERR_NUMERICAL = 6,    /* non-finite residual  (reference: Numerical_error) */
	FVG_ERR_COMM = 7          /* peer-memory / halo set-up failure */
} fvg_status;

/* Message of the last error on the calling thread ("" if none). */
const char *fvg_last_error(void);
/* Number of visible CUDA devices; fails (FVG_ERR_CUDA) when there is no usable GPU. */
int fvg_device_count(int *count);

/* Device memory for host code that does not link the CUDA runtime itself (the C++ class surface in
 * fvens_b200/host/ holds its device-resident Vecs through these; the reference's counterpart is PETSc's
 * VecCreate / VecGetArray / VecDestroy, linalg/alinalg.cpp:9-52, linalg/petscutils.hpp). kind: 0 = host to
 * device, 1 = device to host, 2 = device to device. Copies are synchronous. */
int fvg_malloc(void **d_ptr, unsigned long long bytes);
int fvg_free(void *d_ptr);
int fvg_memcpy(void *dst, const void *src, unsigned long long bytes, int kind);
int fvg_memset(void *d_ptr, int value, unsigned long long bytes);

/* ------------------------------------------------------------------------------------------------
 * Host mesh = fvens::UMesh<double,2> (mesh/mesh.hpp:25-500) built by constructMesh + preprocessMesh
 * (mesh/ameshutils.cpp:39-153): read -> correctBoundaryFaceOrientation -> compute_topological ->
 * compute_areas -> compute_face_data. CPU only, no CUDA calls. */
typedef struct fvg_umesh fvg_umesh;

/* readMesh (mesh/meshreaders.cpp:35-64): .msh = Gmsh 2.2 ASCII, .su2 = SU2 with integer markers */
int fvg_umesh_read(const char *path, fvg_umesh **out);
/* From arrays (MeshData, mesh/meshreaders.hpp:29-56): inpoel [nelem][4] (-1 padded),
 * bface [nbface][3] = node0, node1, marker. */
int fvg_umesh_from_arrays(int npoin, const double *coords, int nelem, const int *nnode,
                          const int *inpoel, int nbface, const int *bface, fvg_umesh **out);
/* UMesh::writeGmsh2 (mesh/mesh.cpp:205-286): Gmsh 2.2 ASCII, what utilities/convertformat.cpp produces from any
 * readable format (e.g. SU2 -> msh). */
int fvg_umesh_write_gmsh2(const fvg_umesh *m, const char *path);
void fvg_umesh_destroy(fvg_umesh *m);
/* UMesh::reorder_cells (mesh/mesh.cpp:85-99) followed by the topology/metric rebuild of
 * preprocessMesh (mesh/ameshutils.cpp:60-100): new cell i = old cell perm[i]. */
int fvg_umesh_reorder_cells(fvg_umesh *m, const int *perm);
/* Reverse Cuthill-McKee ordering of the cell adjacency graph, the stand-in for `-mesh_reorder rcm`
 * (mesh/ameshutils.cpp:246-288, which calls PETSc MatGetOrdering). perm[new] = old. */
int fvg_umesh_rcm_ordering(const fvg_umesh *m, int *perm);
/* getCellAdjLists (mesh/meshpartitioning.cpp:376-430): CSR adjacency of the cells across interior faces, the graph
 * the reference hands to Scotch (ptrs [nelem+1]; store [ptrs[nelem]], may be NULL to get the sizes only). */
int fvg_umesh_cell_adjacency(const fvg_umesh *m, int *ptrs, int *store);
/* The same graph as a Scotch source-graph file (.grf), and the reader of a Scotch mapping file (.map: number of
 * vertices, then "vertex part" lines) into a cell->rank map for fvg_mesh_create_part: where Scotch is available
 * (`gpart N mesh.grf mesh.map`) its partition replaces fvg_partition_sfc without touching anything else. */
int fvg_umesh_write_scotch_graph(const fvg_umesh *m, const char *path);
int fvg_partition_read_scotch_map(const fvg_umesh *m, const char *path, int *cell_rank, int *nparts);
/* Hilbert space-filling-curve ordering of the cell centres (the locality order the device mesh uses;
 * contiguous index ranges are compact patches, which is what both the tiles and the multi-GPU
 * partitions want). perm[new] = old. */
int fvg_umesh_hilbert_ordering(const fvg_umesh *m, int *perm);

/* Non-owning view of the reference's mesh arrays (what UMesh's g*() accessors return). */
typedef struct {
	int npoin, nelem, nbface, naface, ninface, nconnface;
	int maxnnode, nbtag;
	const double *coords;       /* [npoin][2]              gcoords   */
	const int *inpoel;          /* [nelem][maxnnode]       ginpoel   */
	const int *nnode;           /* [nelem]                 gnnode (== gnfael for linear cells) */
	const int *esuel;           /* [nelem][maxnnode]       gesuel: >= nelem+nconnface means physical boundary */
	const int *elemface;        /* [nelem][maxnnode]       gelemface */
	const int *intfac;          /* [naface][4] L,R,n0,n1   gintfac   */
	const int *btags;           /* [nbface][nbtag]         gbtags    */
	const double *facemetric;   /* [naface][3] nx,ny,len   gfacemetric */
	const double *area;         /* [nelem]                 garea     */
	const int *bpartner;        /* [nbface] partner boundary face of a periodic face, -1 otherwise; may be NULL
	                               (UMesh::compute_periodic_map, mesh.cpp:369-424). The device mesh turns a periodic
	                               pair into interior faces: each of the two boundary faces gets the partner's cell,
	                               displaced by the period, as its right cell (one flux per geometric edge
	                               and cell, unlike the reference's dead periodic path, which would count it twice) */
} fvg_host_mesh;
int fvg_umesh_view(const fvg_umesh *m, fvg_host_mesh *view);
/* UMesh::compute_periodic_map(bcm, axis) (mesh/mesh.cpp:369-424): pairs the boundary faces with marker `marker` whose
 * midpoints agree in the coordinate other than `axis` (0: periodic in x, 1: periodic in y). One call per periodic
 * direction, each with its own marker; take the view (fvg_umesh_view) afterwards. Returns the number of pairs. */
int fvg_umesh_compute_periodic_map(fvg_umesh *m, int marker, int axis, int *npairs);

/* ------------------------------------------------------------------------------------------------
 * Device mesh: SoA copy of the arrays above, cells renumbered for locality, faces regrouped into
 * per-tile streams and edge-coloured inside each tile so that the face->cell accumulation needs no
 * atomics and is deterministic. Invisible to callers: all API arrays stay in reference numbering. */
typedef struct fvg_mesh fvg_mesh;

enum { FVG_REORDER_NONE = 0, FVG_REORDER_HILBERT = 1, FVG_REORDER_RCM = 2 };
typedef struct {
	int reorder;       /* FVG_REORDER_*; HILBERT = space-filling-curve order of the cell centres */
	int tile_cells;    /* maximum own cells per tile (0 = default 256); multiple of 32, <= 1024. Tiles are
	                      consecutive cell ranges, shortened where needed so that their halo (distinct
	                      out-of-tile neighbours) and face stream fit the shared-memory staging areas */
	int device;        /* CUDA device ordinal, -1 = current; -2 = host-only build (renumbering, tiling and
	                      colouring can be inspected without a GPU; no flow can be created on it) */
} fvg_mesh_opts;

typedef struct {
	int ncell, nbface, naface, ntile, tile_cells, nstream, ncut_dup, max_colours, reorder;
	double mean_neighbour_distance;   /* mean |i-j| over interior faces in device numbering */
	int nghost, nsend, rank, nranks;  /* subdomain meshes: ghost cells, cells sent per exchange */
	int entry_capacity, halo_capacity; /* per-tile staging capacities the kernels size their shared memory with */
	long long bank_groups, bank_conflict_groups; /* (quarter-warp, local-face) groups of stream entries; groups in which two
	                                     entries share a shared-memory bank residue (0 = conflict-free scatter/gather) */
} fvg_mesh_info;

int fvg_mesh_create(const fvg_host_mesh *hm, const fvg_mesh_opts *opts, fvg_mesh **out);
/* One subdomain of a mesh distributed over `nranks` GPUs (one process per GPU). Every rank passes the same
 * GLOBAL host mesh and cell->rank map; the engine restricts it (the job of the reference's
 * ReplicatedGlobalMeshPartitioner::restrictMeshToPartitions, mesh/meshpartitioning.cpp:24-159): own cells,
 * one ghost cell per cell across a cut face, faces. Arrays of such a mesh are in DEVICE order: rows
 * [0, ncell) are the own cells, rows [ncell, ncell+nghost) the ghosts grouped by owner rank (see
 * fvg_mesh_permutation for the global ids and fvg_mesh_halo_lists for the exchange pattern). */
int fvg_mesh_create_part(const fvg_host_mesh *hm, const fvg_mesh_opts *opts, const int *cell_rank, int rank,
                         int nranks, fvg_mesh **out);
/* Space-filling-curve partition: the Hilbert order of the cells cut into nranks equal chunks (the stand-in
 * for the Scotch partition of mesh/meshpartitioning.cpp:432-480; Scotch is not available offline). */
int fvg_partition_sfc(const fvg_umesh *m, int nranks, int *cell_rank);
/* Recursive coordinate bisection of the cell centres (longer side of the bounding box first, cell counts in proportion
 * to the ranks on either side, so any nranks is balanced to one cell): the geometric partitioner SURVEY.md 8e asks for
 * in Scotch's absence; 30-40 % fewer cut faces than the curve partition on the benchmark mesh. Any cell->rank map is
 * valid input to fvg_mesh_create_part and gives bitwise the same residual. */
int fvg_partition_rcb(const fvg_umesh *m, int nranks, int *cell_rank);
void fvg_mesh_destroy(fvg_mesh *m);
/* Halo pattern of a subdomain mesh: send_counts[r] own cells go to rank r, recv_counts[r] ghost rows come
 * from rank r (ghost rows are ordered by source rank, so a recv buffer IS the ghost block); send_idx
 * (may be NULL) lists the own device cells to pack, grouped by destination rank. */
int fvg_mesh_halo_lists(const fvg_mesh *m, int *send_counts, int *recv_counts, int *send_idx);
/* Packs rows of a device-ordered array for the peers: sendbuf[k][:] = src[send_idx[k]][:] (width doubles). */
int fvg_halo_pack(const fvg_mesh *m, const double *d_src, int width, double *d_sendbuf, void *stream);
/* The same send pattern per tile: tile_off [ntile+1]; cell_peer_row (may be NULL) holds, for the entries of tile t,
 * triples {tile-local cell, peer rank, row inside this rank's block of the peer's ghost range}, i.e. where a kernel
 * that has just produced the row of that cell would store it in the peer's ghost block. */
int fvg_mesh_tile_send_lists(const fvg_mesh *m, int *tile_off, int *cell_peer_row);
int fvg_mesh_get_info(const fvg_mesh *m, fvg_mesh_info *info);

/* Peer-memory halo exchange (one process per GPU, all GPUs in one NVLink/NVSwitch box). Each rank owns a window
 * (receive areas + arrival flags in one device allocation) that its neighbours map through CUDA IPC; an exchange
 * is a send kernel (pack + direct stores into the neighbours' windows + release of a sequence flag) and a receive
 * kernel (acquire the flags, copy the rows into the ghost block): no host round trip and no collective library.
 * The GPU counterpart of the reference's L2TraceVector::updateSharedFacesBegin/End (linalg/tracevector.cpp:214-325)
 * and of its VecGhostUpdateBegin/End calls (ode/aodesolver.cpp:212,247; spatial/flow_spatial.cpp:711-729).
 * Set-up: every rank creates its window, the 64-byte handles and the recv_counts rows of fvg_mesh_halo_lists are
 * all-gathered by the caller (any transport), then fvg_halo_connect maps the neighbours. Every rank must issue the
 * same sequence of send/recv pairs. max_width and width are even numbers of doubles per row (4 = state, 8 = gradients). */
typedef struct fvg_halo fvg_halo;
int fvg_halo_create(fvg_mesh *mesh, int max_width, fvg_halo **out);
int fvg_halo_ipc_handle(fvg_halo *h, void *handle64);
/* handles: [nranks][64] bytes; all_recv_counts: [nranks][nranks], row r = rank r's recv_counts */
int fvg_halo_connect(fvg_halo *h, const void *handles, const int *all_recv_counts);
/* d_arr: device-ordered array [ncell + nghost][width]; send reads its own rows, recv fills its ghost rows */
int fvg_halo_send(fvg_halo *h, const double *d_arr, int width, void *stream);
int fvg_halo_recv(fvg_halo *h, double *d_arr, int width, void *stream);
/* send + recv in one launch (the pushing and the waiting CTAs are co-resident) */
int fvg_halo_exchange(fvg_halo *h, double *d_arr, int width, void *stream);
/* In-kernel receive: fvg_halo_post only sends (this rank's rows into the neighbours' windows, flag released) and
 * returns a token; fvg_flow_ghost_source(flow, which, halo, token) tells the flow that the ghost rows of the state
 * (which = 0) or of the reconstruction gradients (which = 1) are NOT in the array but in the halo window of that
 * exchange. The next split passes, run with fvg_flow_select_tiles(flow, 3), then wait for the neighbours' flags
 * inside the kernel - only the CTAs that reach a partition-boundary tile, only once, and after all their interior
 * tiles - and gather the ghost rows straight from the window: no receive kernel, no copy, and the wait hides behind
 * the interior work. Valid for the two most recent exchanges of a window. token 0 / halo NULL switches back to the
 * array. Inviscid linear-reconstruction or first-order flows only (other passes read ghost rows from the arrays). */
int fvg_halo_post(fvg_halo *h, const double *d_arr, int width, void *stream, unsigned long long *token);
/* 0 if every receive so far saw its neighbours arrive; else the sequence number of a receive that gave up waiting */
int fvg_halo_status(fvg_halo *h, unsigned long long *h_timed_out_seq);
void fvg_halo_destroy(fvg_halo *h);
/* ------------------------------------------------------------------------------------------------
 * Fused multi-GPU evaluation (one process per GPU of one NVLink/NVSwitch box): the product form of the split passes
 * above. FlowFV::compute_residual on a partitioned mesh (spatial/flow_spatial.cpp:637-816 with its ghost updates and
 * trace exchanges, :711-788; linalg/tracevector.cpp:214-325) is the SAME two kernels as on one GPU (three with WENO):
 * the kernel that produces a row a neighbour needs - the gradient pass for gradient rows, the step epilogue (or the
 * first wave of the evaluation's first kernel) for state rows - stores it straight into the neighbour's peer-mapped
 * window over NVLink, and the consuming kernel waits for the neighbours' rows only in the CTAs that reach a
 * partition-boundary tile, after the interior tiles. The evaluation number lives on the device, so an evaluation is
 * captured once into a CUDA graph and replayed (any stream but the legacy default stream; FVG_GRAPH=0 disables).
 * Set-up as for fvg_halo: create, all-gather the 64-byte handles and the recv_counts rows, connect.
 * Arrays are device-ordered: [ncell + nghost][.] for the state (the ghost rows of the ARRAY are never read: they
 * live in the window), [ncell][.] for residual / time steps. Every rank must issue the same sequence of calls. */
typedef struct fvg_dist fvg_dist;
typedef struct fvg_flow fvg_flow;    /* declared below */
int fvg_dist_create(fvg_flow *flow, fvg_dist **out);
int fvg_dist_ipc_handle(fvg_dist *d, void *handle64);
int fvg_dist_connect(fvg_dist *d, const void *handles, const int *all_recv_counts);
/* FlowFV::compute_residual on this rank's subdomain; accumulate as for fvg_residual */
int fvg_dist_residual(fvg_dist *d, const double *d_u, double *d_res, int accumulate, int gettimesteps, double *d_dtm,
                      void *stream);
/* One forward-Euler pseudo-time step (ode/aodesolver.cpp:189-247): d_unew's own rows from d_u (two distinct arrays);
 * the step epilogue pushes the new state rows to the neighbours, so the next evaluation of d_unew needs no state
 * exchange (pass another array, or call fvg_dist_invalidate_state after modifying it, and the rows are pushed again).
 * d_resnorm2 (device, may be NULL on ALL ranks): the sum over ALL ranks of r_E^2 * area for this step (:218-229), reduced
 * through the windows in rank order - bitwise the same on every rank. */
int fvg_dist_euler_step(fvg_dist *d, const double *d_u, double *d_unew, double cfl, double *d_resnorm2, void *stream);
int fvg_dist_invalidate_state(fvg_dist *d);
/* SteadyForwardEulerSolver::solve (ode/aodesolver.cpp:136-282) on the partitioned mesh: same contract as
 * fvg_forward_euler_solve; d_u holds this rank's own rows [ncell][4] (ghost rows, if present, are ignored), h_hist the
 * GLOBAL residual norms (identical on all ranks). FVG_ERR_COMM if a neighbour stopped delivering rows. */
int fvg_dist_forward_euler_solve(fvg_dist *d, double *d_u, double cfl, double tol, int maxiter, int check_every,
                                 int *h_steps, double *h_hist);
/* FVG_OK, or FVG_ERR_COMM once a wait for a neighbour's rows has timed out (FVG_HALO_TIMEOUT_MS, default 20 s); the
 * flag value that never arrived is returned. Synchronises the device. */
int fvg_dist_status(fvg_dist *d, unsigned long long *h_timed_out);
int fvg_dist_counters(fvg_dist *d, long long *evaluations, long long *graph_replays, unsigned long long *device_evaluation);
/* test hook: the receive area of (row type 0 state / 1 unlimited gradients / 2 reconstruction gradients, evaluation parity)
 * as [nghost][4 or 8] doubles, and the 3 x 16 arrival flags */
int fvg_dist_debug_window(fvg_dist *d, int type, int parity, double *h_rows, unsigned long long *h_flags);
void fvg_dist_destroy(fvg_dist *d);

/* cell_new2old[ncell + nghost]: device cell i holds reference (global) cell cell_new2old[i]. */
int fvg_mesh_permutation(const fvg_mesh *m, int *cell_new2old);
/* tile_cell0[ntile+1]: device cells [tile_cell0[t], tile_cell0[t+1]) are tile t's own cells. */
int fvg_mesh_tile_offsets(const fvg_mesh *m, int *tile_cell0);
/* Stream entries: reference face id (-1 for padding entries), colour and tile of every entry (test hook for the colouring
 * validity check: no two entries of one colour in a tile may touch the same tile-owned cell). */
int fvg_mesh_stream(const fvg_mesh *m, int *entry_face, int *entry_colour, int *entry_tile);

/* ------------------------------------------------------------------------------------------------
 * Flow discretisation = FlowFV<double,order2,constVisc> (spatial/flow_spatial.hpp:33-55,150-262)
 * built by create_const_flowSpatialDiscretization (utilities/afactory.cpp:217-270). */
typedef struct fvg_flow fvg_flow;

/* FlowPhysicsConfig (spatial/flow_spatial.hpp:33-44) minus the BC list */
typedef struct {
	double gamma, Minf, Tinf, Reinf, Pr, aoa;
	int viscous_sim, const_visc;
} fvg_physics;

/* Factory keys of utilities/afactory.cpp:38-81 (fluxes), 111-127 (gradients), 178-211 (reconstruction) */
enum { FVG_FLUX_LLF = 0, FVG_FLUX_VANLEER = 1, FVG_FLUX_AUSM = 2, FVG_FLUX_AUSMPLUS = 3,
       FVG_FLUX_ROE = 4, FVG_FLUX_HLL = 5, FVG_FLUX_HLLC = 6 };
enum { FVG_GRAD_ZERO = 0, FVG_GRAD_GREENGAUSS = 1, FVG_GRAD_LEASTSQUARES = 2 };
enum { FVG_RECON_NONE = 0, FVG_RECON_WENO = 1, FVG_RECON_VANALBADA = 2, FVG_RECON_BARTHJESPERSEN = 3,
       FVG_RECON_VENKATAKRISHNAN = 4 };
/* Boundary-neighbour policy of the Barth-Jespersen / Venkatakrishnan limiters. The reference reads
 * out of bounds there; 0 = use the boundary ghost state, 1 = skip (as its WENO does). */
enum { FVG_BND_GHOST = 0, FVG_BND_SKIP = 1 };

/* FlowNumericsConfig (spatial/flow_spatial.hpp:47-55) with integer keys */
typedef struct {
	int flux, gradient, reconstruction;
	double limiter_param;     /* Venkatakrishnan K / WENO central weight (never parsed by the reference, H2) */
	int order2;
	int bnd_policy;
} fvg_numerics;

/* FlowBCConfig (spatial/abc.hpp:34-40); type = BCType of spatial/abctypes.hpp:13-22:
 * 0 slip wall, 1 far field, 2 inflow-outflow, 3 subsonic inflow (vals = p0, T0), 4 extrapolation,
 * 5 periodic (accepted for a marker whose faces the mesh has all paired, see fvg_host_mesh::bpartner; no ghost state is
 * ever computed for it), 6 isothermal wall (vals = tangential velocity, T), 7 adiabatic wall (vals[0] = v_t) */
typedef struct {
	int tag, type;
	double vals[2];
} fvg_bc;

int fvg_flow_create(fvg_mesh *mesh, const fvg_physics *phys, const fvg_numerics *num,
                    const fvg_bc *bcs, int nbc, fvg_flow **out);
void fvg_flow_destroy(fvg_flow *f);

/* Spatial::compute_residual (spatial/aspatial.hpp:62-63, flow_spatial.cpp:637-816): adds -r(u) into
 * d_res when accumulate != 0 (the reference's contract: caller zeroes), else overwrites it.
 * d_dtm may be NULL when gettimesteps == 0. */
int fvg_residual(fvg_flow *f, const double *d_u, double *d_res, int accumulate, int gettimesteps,
                 double *d_dtm, void *stream);
/* Same through host buffers (drop-in mode): uploads u (and h_res when accumulate != 0, the reference's
 * add-into contract), runs the kernels, downloads the residual and the time steps. Synchronous.
 * Pinned host memory makes the copies run at full PCIe rate but is not required. */
int fvg_residual_host(fvg_flow *f, const double *h_u, double *h_res, int accumulate, int gettimesteps,
                      double *h_dtm);

/* The two passes of fvg_residual as separate calls, so that a multi-GPU driver can exchange ghost rows in
 * between (the reference's gradient halo, spatial/flow_spatial.cpp:711-729, and trace exchange :738,:782):
 *   fvg_gradient_pass(stage 0)  -> limited gradients of the own cells in the flow's d_lg buffer
 *                                  (WENO: unlimited gradients in d_gu; exchange them, then stage 1 -> d_lg)
 *   [exchange the ghost rows of d_lg (and d_gu for viscous runs with a limiter)]
 *   fvg_face_pass / fvg_euler_face_pass
 * The state must be device-ordered with valid ghost rows. fvg_flow_buffers returns the gradient buffers
 * ([ncell+nghost][8] each; NULL if the numerics do not use one). */
int fvg_gradient_pass(fvg_flow *f, const double *d_u, int stage, void *stream);
int fvg_face_pass(fvg_flow *f, const double *d_u, double *d_res, int accumulate, int gettimesteps,
                  double *d_dtm, void *stream);
int fvg_euler_face_pass(fvg_flow *f, const double *d_u, double *d_unew, double cfl, double *d_resnorm2,
                        void *stream);
/* Which tiles the split passes above cover from now on: 0 = all (default), 1 = tiles that see no ghost cell,
 * 2 = tiles on the partition boundary. A multi-GPU driver runs part 1 while the ghost rows are in flight and
 * part 2 once they have arrived (the reference overlaps its trace exchange with the interior faces the same way,
 * spatial/flow_spatial.cpp:738-782). On an unpartitioned mesh every tile is in part 1 and part 2 is empty;
 * fvg_euler_face_pass sums the norm after part 2 (or 0). WENO's stage-1 pass always covers all tiles.
 * 3 = all tiles in ONE launch, interior tiles first: the order the in-kernel receive below relies on. */
int fvg_flow_select_tiles(fvg_flow *f, int part);
/* see fvg_halo_post */
int fvg_flow_ghost_source(fvg_flow *f, int which, fvg_halo *h, unsigned long long token);
int fvg_flow_buffers(fvg_flow *f, double **d_lg, double **d_gu);
/* Makes the flow use caller-owned gradient buffers ([ncell+nghost][8] doubles each, device memory that must
 * outlive the flow's use of them), e.g. tensors that a communication library can address. */
int fvg_flow_use_buffers(fvg_flow *f, double *d_lg, double *d_gu);

/* GradientScheme::compute_gradients (spatial/agradientschemes.hpp:44-48) on primitive cell states */
int fvg_gradients(fvg_flow *f, const double *d_uprim, const double *d_ug, double *d_grad, void *stream);
/* SolutionReconstruction::compute_face_values (spatial/areconstruction.hpp:37-41) */
int fvg_face_values(fvg_flow *f, const double *d_uprim, const double *d_ug, const double *d_grad,
                    double *d_ufl, double *d_ufr, void *stream);
/* FlowFV_base::compute_boundary_states (spatial/flow_spatial.cpp:74-93): ins, gs [nbface][4] conserved */
int fvg_boundary_states(fvg_flow *f, const double *d_ins, double *d_gs, void *stream);
/* MatrixFreeSpatialJacobian::apply (linalg/alinalg.cpp:143-230): y = mdt x + (r(u + h x) - r(u))/h with h = eps/|x|_2,
 * the product of the pseudo-time-shifted residual Jacobian with x by one extra residual evaluation (a Krylov solver
 * can sit on top of the GPU residual without any matrix). d_res is what fvg_residual left for the state d_u
 * (accumulate = 0), d_mdt the diagonal shift per cell (the reference passes area/dt), all device arrays in the
 * caller's cell order; the reference's default eps is 1e-7. Asynchronous on `stream`, no host synchronisation
 * (|x| is reduced and consumed on the device). */
int fvg_jacobian_vector_product(fvg_flow *f, const double *d_u, const double *d_res, const double *d_mdt, const double *d_x,
                                double eps, double *d_y, void *stream);
/* FlowFV_base::getGradients (spatial/flow_spatial.cpp:96-112): gradients of the CONSERVED variables */
int fvg_get_gradients(fvg_flow *f, const double *d_u, double *d_grads, void *stream);
/* FlowFV_base::computeSurfaceData (spatial/flow_spatial.cpp:131-310): out3 = Cl, Cdp, Cdf over the
 * faces carrying `marker`. Synchronous; h_out3 is host memory. */
int fvg_surface_data(fvg_flow *f, const double *d_u, const double *d_grads, int marker, double *h_out3);
/* compute_entropy_cell (spatial/aoutput.cpp:28-63). Synchronous. */
int fvg_entropy_error(fvg_flow *f, const double *d_u, double *h_out);

/* One pseudo-time step of SteadyForwardEulerSolver::solve (ode/aodesolver.cpp:177-251):
 * r = -r(u); dt; u += cfl*dt/area*r; *d_resnorm2 = sum_cells r_E^2 * area. Fused: no residual array
 * is written. d_resnorm2 is one double of device memory (may be NULL). */
int fvg_euler_step(fvg_flow *f, double *d_u, double cfl, double *d_resnorm2, void *stream);
/* Whole solve loop: steps until resi/initres <= tol (returns 0), maxiter reached (FVG_ERR_TOLERANCE)
 * or a non-finite norm (FVG_ERR_NUMERICAL). h_hist (may be NULL) receives the absolute residual norm
 * of every step (capacity maxiter). check_every >= 1: how often the norm is read back (1 = the
 * reference's behaviour). Synchronous. */
int fvg_forward_euler_solve(fvg_flow *f, double *d_u, double cfl, double tol, int maxiter,
                            int check_every, int *h_steps, double *h_hist);

/* TVDRKSolver (ode/aodesolver.hpp:229-258; solve: ode/aodesolver.cpp:672-785), the explicit time-accurate driver,
 * as a strong-stability-preserving Runge-Kutta scheme of order 1, 2 or 3 in Shu-Osher form with the reference's
 * coefficient table (initialize_TVDRK_Coeffs, ode/aodesolver.cpp:45-67):
 *     u(0) = u;  u(i+1) = a_i u + b_i u(i) + c_i dt/area * (-r(u(i)));  u <- u(order);  time += dt
 * dt = cfl * min over cells of the local time step at u (first stage), as the reference. Unlike the reference's loop
 * the residual is evaluated at the STAGE state and enters with the sign the forward-Euler loop uses (the reference
 * evaluates every stage at u and subtracts -r). Loop: while time <= finaltime - 1e-12 (the last step
 * is not clipped, as the reference) and, if maxsteps > 0, step < maxsteps. Returns FVG_ERR_NUMERICAL when dt is not
 * finite ("TVDRK solver diverged - dtmin is Nan or inf!"); d_u then holds the last completed step. Synchronous;
 * d_u [ncell][4] in the caller's cell order. */
int fvg_tvdrk_solve(fvg_flow *f, double *d_u, int order, double cfl, double finaltime, int maxsteps,
                    int *h_steps, double *h_time);
/* the coefficient table: h_coeffs [order][3] = (a_i, b_i, c_i) */
int fvg_tvdrk_coefficients(int order, double *h_coeffs);

/* Pointwise test hooks (host buffers; each launches a kernel - there is no CPU path):
 * InviscidFlux::get_flux (spatial/anumericalflux.hpp:32-34), FlowBC::computeGhostState
 * (spatial/abc.hpp:72-73), FlowFV::compute_viscous_flux (spatial/flow_spatial.cpp:349-395). */
int fvg_flux_pointwise(int flux_id, const fvg_physics *phys, int n, const double *h_ul,
                       const double *h_ur, const double *h_n, double *h_out);
int fvg_bc_pointwise(const fvg_bc *bc, const fvg_physics *phys, int n, const double *h_ins,
                     const double *h_n, double *h_out);
int fvg_viscous_flux_pointwise(const fvg_physics *phys, int order2, int n, const double *h_n,
                               const double *h_rcl, const double *h_rcr, const double *h_ucl,
                               const double *h_ucr, const double *h_gl, const double *h_gr,
                               const double *h_ul, const double *h_ur, double *h_out);
/* IdealGasPhysics::compute_freestream_state (physics/aphysics.cpp:44-58) */
int fvg_freestream(const fvg_physics *phys, double *h_uinf4);

/* Per-pass device timing of fvg_residual / fvg_euler_step with CUDA events recorded on the launching
 * stream. Call with enable = 1 to start, later with h_out3 to read {ms in the gradient/limiter pass,
 * ms in the face pass, number of evaluations timed} accumulated since the previous call (synchronises). */
int fvg_flow_timing(fvg_flow *f, int enable, double *h_out3);
/* Work counters since flow creation: kernels launched by this library for that flow. */
int fvg_flow_launch_count(const fvg_flow *f, long long *count);

#ifdef __cplusplus
}
#endif
#endif
