/* Internal declarations shared by the translation units of libfvens_b200.so. Not part of the ABI. */
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>              // CUtensorMap (type only: the encoder is fetched through the runtime, libcuda is not linked)
#include <string>
#include <vector>
#include <cstdint>
#include "gas.cuh"
#include "dist_dev.cuh"
#include "../../include/fvens_b200.h"

namespace fvg {

constexpr int MAXCOL = 8;          ///< greedy edge colouring of a degree-4 graph needs at most 7
#ifndef FVG_FACE_BLOCK
#define FVG_FACE_BLOCK 256
#endif
#ifndef FVG_FACE_MINB
#define FVG_FACE_MINB 2
#endif
#ifndef FVG_CELL_BLOCK
#define FVG_CELL_BLOCK 256
#endif
#ifndef FVG_CELL_MINB
#define FVG_CELL_MINB 3
#endif
constexpr int FACE_BLOCK = FVG_FACE_BLOCK;    ///< threads per CTA of the face kernel
constexpr int CELL_BLOCK = FVG_CELL_BLOCK;    ///< threads per CTA of the cell kernels

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FVG_CUDA(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
	return ::fvg::cuda_fail(e_, #call, __FILE__, __LINE__); } while(0)

/// Device-resident mesh in device (renumbered) cell order. All pointers are device memory.
///
/// Cells are grouped into TILES of consecutive cells (at most TC own cells, at most HMAX distinct
/// out-of-tile neighbour cells = the tile's HALO, at most EMAX faces). Inside a tile everything is
/// addressed with 16-bit LOCAL indices: own cell k -> k, halo cell h -> ncell_of_tile + h. Each tile
/// has a contiguous FACE STREAM segment (every face touching one of its own cells; a face cut by a
/// tile boundary has a copy in both tiles), ordered by kind (both cells in the tile / cut by the tile
/// boundary / physical boundary / padding), start and length padded to 4 entries.
struct DMesh {
	int ncell;              ///< own cells (tiles, residual rows); state-like arrays have ncell + nghost rows
	int nghost, nsend;      ///< ghost cells received from / own cells sent to other ranks per exchange
	int nbface, naface, ntile, TC, HMAX, EMAX, nstream;
	// per cell
	const uint4 *cloc;      ///< x,y: local index of the neighbour across local face 0..3 (4 x u16; NB_NONE / NB_BND);
	                        ///< z,w: local stream entry of local face 0..3 (4 x u16, bit 15 set if the cell is the entry's right cell)
	const double2 *rc;      ///< cell centres
	const double *area;
	const double4 *wlsV;    ///< inverse least-squares matrices (V00, V01, V10, V11)
	const double *clength;  ///< Venkatakrishnan length scale (longest edge)
	// per tile
	const int *tcell0;      ///< [ntile+1] first own cell of each tile
	const int *thoff;       ///< [ntile+1] offsets into thalo
	const int *thalo;       ///< halo cell ids (device numbering), ascending within a tile
	const int *fsoff;       ///< [ntile+1] stream segment of each tile (multiples of 4)
	const int4 *tbnd;       ///< [ntile] x: tile-local index of the first cut entry (one side in the halo), y: of the first
	                        ///< physical-boundary entry, z: number of boundary entries (halo + these <= HMAX), w: padding entries
	                        ///< (low 16 bits) | 0x10000 if the tile's halo contains a ghost cell of another rank
	// per stream entry
	const unsigned *fLR;    ///< local left | local right << 16; right >= LR_BND: boundary face with BC table index (right & 15);
	                        ///< LR_PAD: padding entry
	const double2 *fn;      ///< unit normal, left -> right
	const double *flen;
	const double2 *fgr;     ///< face midpoint
	const double2 *fgw;     ///< Green-Gauss face data: inverse-distance weights of the left and right cell (they add up to 1)
	const double2 *fgln;    ///< ... and len*n_x, len*n_y (agradientschemes.cpp:62-214 with the geometry folded in once)
	const unsigned short *ford; ///< per tile: tile-local positions of its real (non-padding) entries, ascending; length = stream
	                        ///< segment, the first (segment - tbnd.w) values are meaningful
	const int *fref;        ///< reference face id (intfac index); duplicate copy of a cut face: -1-id; padding: INT_MIN
	// per boundary face (reference order)
	const int *bcell;       ///< device index of the interior cell
	const int *bentry;      ///< stream entry
	const int *bslot;       ///< slot of the face's marker in the flow's BC table (sorted distinct markers)
	const double2 *rcbp;    ///< ghost cell centre
	// permutation (null when identity)
	const int *new2old;
	const int *old2new;
	const int *halo_src;    ///< new2old of every entry of thalo (null when identity): caller-ordered halo gathers without a dependent load
	const int *send_idx;    ///< [nsend] own cells packed for the peers, grouped by peer rank
	// tiles that see no ghost cell first, then the tiles on the partition boundary (null on an unpartitioned mesh):
	// the first group can run while the ghost rows are still being exchanged
	const int *tile_order;
	int ntile_interior;
	/// one record of three int4 per tile: {tile, c0, nc, h0} {nh, e0, ne, tbnd.w} {tbnd.x, tbnd.y, tbnd.z, 0}; tdesc in
	/// natural tile order, tdesc_ord in the order of tile_order (null when there is none)
	const int4 *tdesc, *tdesc_ord;
};

/// Where a pass finds the ghost rows of an array when they are NOT copied into the array: in the halo window of
/// exchange `seq`, valid once every sending neighbour's flag has reached `seq` (rows == nullptr: in the array itself)
struct GhostSrc {
	const double *rows = nullptr;              ///< window buffer of that exchange: [nghost][width]
	const unsigned long long *flags = nullptr; ///< [nranks] arrival flags, then the error word
	const int *recv_off = nullptr;             ///< [nranks+1] ghost-row offsets per source rank
	unsigned long long seq = 0;
	int nranks = 0;
};

/// Role of one kernel launch in the fused multi-GPU evaluation (see dist_dev.cuh). d == nullptr: not part of one.
struct DistRole {
	const DistDev *d = nullptr;
	const DistCtl *ctl = nullptr;                 ///< d->ctl (spares the kernels a dependent load)
	const double *ghost[X_COUNT][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   ///< local receive areas by type and evaluation parity
	unsigned wait = 0;         ///< bit mask of row types whose ghost rows this kernel reads from the window
	unsigned push = 0;         ///< bit mask of row types this kernel produces and pushes per tile (X_U: the step epilogue)
	int first = 0;             ///< first kernel of the evaluation: its prologue pushes the state rows
	int last = 0;              ///< last kernel of the evaluation: its last CTA advances the evaluation counter
	int force_push = 0;        ///< push the state rows in the prologue even if a step epilogue is on record as having done so
	int visc_type = X_LG;      ///< row type of the gradients the viscous flux reads (X_LG when no limiter alters them)
};

constexpr unsigned NB_NONE = 0xFFFFu;   ///< no such local face (4th slot of a triangle)
constexpr unsigned NB_BND = 0xFFFEu;    ///< neighbour is a physical boundary ghost
constexpr unsigned LR_BND = 0xFFE0u;    ///< right field of a boundary entry = LR_BND | bc index
constexpr unsigned LR_PAD = 0xFFFFFFFFu;

} // namespace fvg

/// Opaque handles of the ABI
struct fvg_umesh;

struct fvg_mesh {
	fvg::DMesh d;
	int device = 0;
	int reorder = 0;
	bool identity_perm = true;
	int ncut_dup = 0, max_colours = 0;
	long long bank_groups = 0, bank_conflict_groups = 0;   ///< (quarter-warp, local face) groups of stream entries; those with a repeated bank residue
	double mean_nbr_dist = 0;
	std::vector<void*> allocs;                 ///< everything cudaMalloc'ed for this mesh
	// host copies kept for the test hooks and for flow set-up
	std::vector<int> h_new2old, h_old2new;
	std::vector<int> h_fref, h_fcolour, h_ftile;
	std::vector<int> h_btag;
	std::vector<unsigned> h_fLR;               ///< kept so that flow creation can patch in the BC indices
	std::vector<int> h_bentry;
	std::vector<int> h_markers;                ///< sorted distinct boundary markers = slots of the BC table
	std::vector<int> h_marker_periodic;        ///< per slot: 1 if the marker's faces are periodic pairs (interior faces on the device)
	std::vector<int> h_tcell0, h_thoff, h_thalo;
	std::vector<int> h_tsoff, h_tsend;         ///< per-tile send lists: offsets [ntile+1]; triples (tile-local cell, peer, row in my block)
	std::vector<double> h_rc;                  ///< cell centres, device order (own cells)
	int nghost = 0, rank = 0, nranks = 1;
	int h_periodic_faces = 0;                  ///< own boundary faces that a periodic pairing turned into interior faces
	std::vector<int> send_counts, recv_counts, h_send_idx;
};

namespace fvg {

/// Which scratch arrays a flow needs
struct FlowPlan {
	int flux, gradient, recon, order2, bnd_policy, visc;   // visc = ViscMode
	bool need_lg;     ///< limited (or plain) gradients consumed by the face kernel's linear reconstruction
	bool need_gu;     ///< unlimited gradients (MUSCL, WENO input, viscous flux)
};

} // namespace fvg

struct fvg_flow {
	fvg_mesh *mesh = nullptr;
	fvg::GasParams gas;
	fvg::FlowPlan plan;
	fvg_physics phys;
	double *d_lg = nullptr;        ///< [ncell][8] limited gradients
	double *d_gu = nullptr;        ///< [ncell][8] unlimited gradients
	double *d_uperm = nullptr;     ///< [ncell][4] scratch state in device order (non-identity permutations)
	double *d_rperm = nullptr;     ///< [ncell][4] scratch residual in device order
	double *d_dtperm = nullptr;    ///< [ncell]
	double *d_u2 = nullptr;        ///< [ncell][4] second state buffer for the fused step
	double *d_rk = nullptr;        ///< scratch of fvg_tvdrk_solve: stage state, residual, local steps, areas, reduction blocks
	double *d_partial = nullptr;   ///< [ntile][FACE_BLOCK/32] partial norms per tile and warp
	double *d_norm = nullptr;      ///< [1]
	double *h_norm = nullptr;      ///< pinned [1]
	double *d_hu = nullptr, *d_hr = nullptr, *d_hdt = nullptr;   ///< staging for the host-buffer entry point
	double *d_jaux = nullptr, *d_jyg = nullptr, *d_jpart = nullptr, *d_jnorm = nullptr;   ///< scratch of fvg_jacobian_vector_product
	// chunked host-buffer pipeline (fvg_residual_host): tile ranges, their upload order and dependencies
	struct HostPipe {
		bool planned = false;
		int K = 0;
		std::vector<int> tile0;                 ///< [K+1] chunk -> first tile
		std::vector<int> order;                 ///< upload order of the chunks
		std::vector<unsigned long long> deps;   ///< chunk -> chunks owning its tiles' cells and halo cells (bit mask)
		cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
		std::vector<cudaEvent_t> ev_up, ev_face;
		cudaEvent_t ev_start = nullptr;         ///< orders the pipeline after earlier work on the default stream
	} pipe;
	std::vector<void*> allocs;
	long long launches = 0;
	int prefetch_distance = 0;
	int part = 0;                  ///< tiles the split passes cover: 0 all, 1 interior, 2 partition boundary, 3 all, interior first
	fvg::GhostSrc gs_u, gs_g;      ///< set by fvg_flow_ghost_source
	/// fused multi-GPU evaluation (dist.cu): while `active`, the passes run over all tiles, interior tiles first, with these roles
	struct DistRoles { bool active = false; fvg::DistRole cell, weno, face; } roles;
	struct fvg_dist *self_dist = nullptr;      ///< single-rank mesh with periodic ghost cells: the exchange engine with this rank as its own peer
	// optional per-pass timing (CUDA events on the launching stream)
	bool timing = false;
	std::vector<cudaEvent_t> ev;   ///< triples: before pass A, between A and B, after B
};

namespace fvg {

// ---- launchers (kernels_cell.cu) ---------------------------------------------------------------
struct CellArgs {
	DMesh m;
	GasParams gas;
	const double *u;       ///< [ncell][4] conserved (or primitive when prim_in)
	const double *ug;      ///< [nbface][4] primitive ghost states (prim_in only)
	const double *gin;     ///< given gradients (GRAD_GIVEN only)
	double *lg;            ///< out: limited gradients (may be null)
	double *gu;            ///< out: unlimited gradients (may be null)
	int bnd_policy;
	int prefetch_distance; ///< tiles ahead whose operands are pulled into L2 (0 = off)
	int tile0 = 0, tile1 = -1;   ///< tile range of this launch (tile1 < 0: all tiles)
	const int4 *tdesc = nullptr; ///< tile sequence the range indexes (DMesh::tdesc or tdesc_ord; set by the launcher from `ordered`)
	bool ordered = false;        ///< walk the tiles in the order of DMesh::tile_order (interior tiles first)
	GhostSrc gs_u;               ///< ghost rows of u (partition-boundary tiles wait for them inside the kernel)
	DistRole dist;               ///< fused multi-GPU evaluation (dist.cu): what this launch pushes and waits for
	int pdl = 0;                     ///< launched with the programmatic-dependent-launch attribute: griddepcontrol.wait before the first read
	const int *src_idx = nullptr;    ///< caller-ordered state: row of `u` holding device cell i (null: u is in device order)
	const int *halo_src = nullptr;   ///< ... and the same for the entries of thalo (precomputed: no dependent index load)
	double *ucopy = nullptr;         ///< ... the own rows are also written here in device order (read by the face pass)
};
int launch_cell_kernel(int grad, int lim, bool prim_in, const CellArgs &a, cudaStream_t s);
struct WenoArgs {
	DMesh m;
	double lambda;
	const double *gu;      ///< [ncell(+nghost)][8] unlimited gradients
	double *lg;            ///< out: [ncell][8]
	const int4 *tdesc = nullptr;
	bool ordered = false;
	GhostSrc gs_gu;        ///< ghost rows of gu
	DistRole dist;
};
int launch_weno_kernel(const WenoArgs &a, cudaStream_t s);
/// persistent-grid size for a kernel: SMs x resident CTAs, cached per (device, kernel, shared-memory size)
int resident_ctas(const void *kernel, int block, size_t smem);
/// launches with the programmatic-dependent-launch attribute when FVG_PDL is not "0"
bool pdl_enabled();

struct FaceValArgs {
	DMesh m;
	const double *up;      ///< primitive cell states, device order
	const double *ug;      ///< primitive ghost states [nbface][4]
	const double *g;       ///< gradients used for extrapolation (limited, or unlimited for MUSCL)
	double *ufl, *ufr;     ///< [naface][4], reference face order
	int muscl;
};
int launch_face_values(const FaceValArgs &a, cudaStream_t s);

int launch_permute_rows(const double *src, double *dst, const int *idx, int n, int width, bool gather,
                        bool accumulate, cudaStream_t s);
int launch_boundary_states(const DMesh &m, const GasParams &g, const double *ins,
                           double *gs, cudaStream_t s);
int launch_cons2prim(const GasParams &g, const double *u, double *p, int n, cudaStream_t s);
int launch_boundary_prim_ghosts(const DMesh &m, const GasParams &g, const double *u,
                                double *ug, bool prim_out, cudaStream_t s);
int launch_halo_pack(const DMesh &m, const double *src, int width, double *dst, cudaStream_t s);
int launch_final_norm(const double *partial, int n, double *out, cudaStream_t s);
int launch_sumsq(const double *x, long long n, double *partial, int nblk, double *out, cudaStream_t s);
int launch_perturb(const double *u, const double *x, const double *xnorm2, double eps, long long n, double *aux, cudaStream_t s);
int launch_jvp_combine(const double *x, const double *res, const double *yg, const double *mdt, const double *xnorm2,
                       double eps, int ncell, int nvars, double *y, cudaStream_t s);
int launch_surface_data(const DMesh &m, const GasParams &g, double aoa, const double *u, const double *grads,
                        int slot, double *out4, cudaStream_t s);
int launch_entropy(const DMesh &m, const GasParams &g, const double *u, double *out, cudaStream_t s);
int launch_pointwise_flux(int flux, const GasParams &g, int n, const double *ul, const double *ur,
                          const double *nrm, double *out, cudaStream_t s);
int launch_pointwise_bc(const GasParams &g, int n, const double *ins, const double *nrm, double *out,
                        cudaStream_t s);
int launch_pointwise_visc(const GasParams &g, bool order2, bool constvisc, int n, const double *nrm,
                          const double *rcl, const double *rcr, const double *ucl, const double *ucr,
                          const double *gl, const double *gr, const double *ul, const double *ur,
                          double *out, cudaStream_t s);

// ---- face kernel (face_kernel.cuh, one translation unit per flux) -------------------------------
enum FaceRecon { FR_FIRST = 0, FR_LINEAR = 1, FR_MUSCL = 2 };
enum FaceEpilogue { EP_RESIDUAL = 0, EP_STEP = 1 };

struct FaceArgs {
	DMesh m;
	GasParams gas;
	const double *u;       ///< [ncell][4] conserved, device order
	const double *lg;      ///< gradients for the linear reconstruction
	const double *gu;      ///< unlimited gradients (MUSCL, viscous)
	// epilogue
	int epilogue;          ///< FaceEpilogue
	int accumulate;        ///< EP_RESIDUAL: add into res instead of overwriting
	int gettimesteps;
	double *res;           ///< [ncell][4]
	double *dtm;           ///< [ncell]
	double cfl;            ///< EP_STEP
	double *unew;          ///< EP_STEP: [ncell][4]
	double *partial;       ///< EP_STEP: [ntile][FACE_BLOCK/32] sums of r_E^2*area per tile and warp
	int prefetch_distance; ///< tiles ahead whose operands are pulled into L2 (0 = off)
	// TMA descriptors of the per-cell row arrays staged per tile: box = TC rows, hardware-swizzled so that
	// "thread k reads row k" is free of shared-memory bank conflicts (32-byte rows: SWIZZLE_32B, 64-byte: SWIZZLE_64B)
	CUtensorMap tm_u;      ///< u as [ncell][4]
	CUtensorMap tm_g;      ///< lg as [ncell][8] (linear reconstruction only)
	int tile0 = 0, tile1 = -1;   ///< tile range of this launch (tile1 < 0: all tiles)
	const int4 *tdesc = nullptr; ///< tile sequence the range indexes (DMesh::tdesc or tdesc_ord; set by the launcher from `ordered`)
	bool ordered = false;        ///< walk the tiles in the order of DMesh::tile_order (interior tiles first)
	GhostSrc gs_u, gs_g;         ///< ghost rows of u and of the reconstruction gradients, see GhostSrc
	GhostSrc gs_v;               ///< ghost rows of the viscous flux's gradients (fused multi-GPU evaluation only)
	DistRole dist;               ///< fused multi-GPU evaluation (dist.cu)
	const int *dst_idx = nullptr;    ///< caller-ordered outputs: row of res / dtm / unew that receives device cell i (null: device order)
};
/// Rows per TMA box of the per-cell row arrays (a box has at most 256 rows and must tile TC exactly)
__host__ __device__ inline int tile_box_rows(int TC) {
	return TC <= 256 ? TC : (TC % 256 == 0 ? 256 : (TC % 128 == 0 ? 128 : (TC % 64 == 0 ? 64 : 32)));
}
/// Tensor map over a row-major [nrows][width] FP64 array (width 4 or 8), box = box_rows full rows, swizzle = row size
int make_row_tensor_map(CUtensorMap *tm, const double *base, size_t nrows, int width, int box_rows);
typedef int (*FaceLauncher)(int recon, int visc, const FaceArgs &a, cudaStream_t s);
int launch_face_llf(int recon, int visc, const FaceArgs &a, cudaStream_t s);
int launch_face_vanleer(int recon, int visc, const FaceArgs &a, cudaStream_t s);
int launch_face_ausm(int recon, int visc, const FaceArgs &a, cudaStream_t s);
int launch_face_ausmplus(int recon, int visc, const FaceArgs &a, cudaStream_t s);
int launch_face_roe(int recon, int visc, const FaceArgs &a, cudaStream_t s);
int launch_face_hll(int recon, int visc, const FaceArgs &a, cudaStream_t s);
int launch_face_hllc(int recon, int visc, const FaceArgs &a, cudaStream_t s);

GasParams make_gas(const fvg_physics &p, double limiter_param);
/// Pass A / pass B on a device-ordered conserved state (capi.cu); used by the C ABI entry points and by dist.cu
int run_gradient_pass(fvg_flow *f, const double *u, cudaStream_t s, int tile0 = 0, int tile1 = -1,
                      const int *src_idx = nullptr, double *ucopy = nullptr);
int run_face_pass(fvg_flow *f, const double *u, int epilogue, int accumulate, int gettimesteps,
                  double *res, double *dtm, double cfl, double *unew, cudaStream_t s, int tile0 = 0, int tile1 = -1,
                  const int *dst_idx = nullptr);
template <typename T> int flow_dev_alloc(fvg_flow *f, T **p, size_t count) {
	void *q = nullptr;
	FVG_CUDA(cudaMalloc(&q, (count > 0 ? count : 1)*sizeof(T)));
	f->allocs.push_back(q);
	*p = static_cast<T*>(q);
	return 0;
}
} // namespace fvg
struct fvg_halo;
extern "C" int fvg_halo_ghost_source(fvg_halo *h, unsigned long long token, fvg::GhostSrc *out);   // halo.cu (internal)
namespace fvg {

} // namespace fvg
