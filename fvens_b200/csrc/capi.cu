/* extern "C" entry points of libfvens_b200.so other than the device-mesh ones (device_mesh.cu):
 * host mesh wrappers, flow creation, the residual / pseudo-time orchestration and the test hooks.
 * Orchestration restates FlowFV::compute_residual (reference src/spatial/flow_spatial.cpp:637-816)
 * and SteadyForwardEulerSolver::solve (src/ode/aodesolver.cpp:136-282) as a two-pass schedule:
 *   pass A  cell kernel   cons->prim, boundary ghosts, gradient, limiter  -> limited gradients
 *   pass B  face kernel   reconstruct, flux, spectral radius, accumulate  -> residual + dt (or u_new)
 */
#include "engine.hpp"
#include "../host/mesh.hpp"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <memory>

struct fvg_umesh { fvens::UMesh<double,2> m; };

namespace fvg {

void rcm_order(int n, int mw, const int *esuel, const int *nnode, std::vector<int> &new2old);
void hilbert_order(int n, const double *rc, std::vector<int> &new2old);

GasParams make_gas(const fvg_physics &p, double limiter_param)
{
	GasParams G;
	std::memset(&G, 0, sizeof(G));
	G.g = p.gamma; G.Minf = p.Minf; G.Tinf = p.Tinf; G.Reinf = p.Reinf; G.Pr = p.Pr;
	G.sCT = 110.5/p.Tinf;
	G.gm1 = p.gamma - 1.0;
	G.igm1 = 1.0/(p.gamma - 1.0);
	G.gM2 = p.gamma*p.Minf*p.Minf;
	G.pinf = 1.0/(p.gamma*p.Minf*p.Minf);
	G.uinf[0] = 1.0;
	G.uinf[1] = std::cos(p.aoa)*std::cos(0.0);
	G.uinf[2] = std::sin(p.aoa)*std::cos(0.0);
	G.uinf[3] = G.pinf/(p.gamma - 1.0) + 0.5*1.0*1.0;
	G.limiter_param = limiter_param;
	G.nbc = 0;
	return G;
}

static FaceLauncher face_launcher(int flux)
{
	switch(flux) {
	case FLUX_LLF: return launch_face_llf;
	case FLUX_VANLEER: return launch_face_vanleer;
	case FLUX_AUSM: return launch_face_ausm;
	case FLUX_AUSMPLUS: return launch_face_ausmplus;
	case FLUX_ROE: return launch_face_roe;
	case FLUX_HLL: return launch_face_hll;
	case FLUX_HLLC: return launch_face_hllc;
	}
	return nullptr;
}

template <typename T>
static int dev_alloc(fvg_flow *f, T **p, size_t count) { return flow_dev_alloc(f, p, count); }

static int limiter_mode(int recon)
{
	if(recon == FVG_RECON_BARTHJESPERSEN) return 1;
	if(recon == FVG_RECON_VENKATAKRISHNAN) return 2;
	return 0;
}

/// Tile range and order of a pass: explicit range (natural order) if given, else the flow's selected part; the fused
/// multi-GPU evaluation walks all tiles, interior tiles first
static void resolve_tiles(const fvg_flow *f, int &tile0, int &tile1, bool &ordered)
{
	ordered = false;
	const DMesh &D = f->mesh->d;
	// (subdomain meshes are numbered interior tiles first; a single-rank periodic mesh keeps the caller's numbering and
	// walks the tile list instead)
	if(f->roles.active && D.tile_order && tile1 < 0) { ordered = f->mesh->nranks == 1; tile0 = 0; tile1 = D.ntile; return; }
	if(tile1 >= 0 || f->part == 0 || !D.tile_order) return;
	ordered = true;
	if(f->part == 1) { tile0 = 0; tile1 = D.ntile_interior; }
	else if(f->part == 2) { tile0 = D.ntile_interior; tile1 = D.ntile; }
	else { tile0 = 0; tile1 = D.ntile; }
}

/// Pass A on a conserved state (device order, or caller order with src_idx / ucopy): fills f->d_lg and/or f->d_gu
int run_gradient_pass(fvg_flow *f, const double *u, cudaStream_t s, int tile0, int tile1, const int *src_idx, double *ucopy)
{
	const FlowPlan &P = f->plan;
	if(!P.order2) return 0;
	CellArgs a;
	a.m = f->mesh->d; a.gas = f->gas; a.u = u; a.ug = nullptr; a.gin = nullptr;
	a.bnd_policy = P.bnd_policy;
	a.prefetch_distance = f->prefetch_distance;
	resolve_tiles(f, tile0, tile1, a.ordered);
	a.tile0 = tile0; a.tile1 = tile1;
	a.gs_u = f->gs_u;
	a.src_idx = src_idx; a.halo_src = src_idx ? f->mesh->d.halo_src : nullptr; a.ucopy = ucopy;
	if(f->roles.active) a.dist = f->roles.cell;
	int rc;
	if(P.recon == FVG_RECON_WENO) {
		a.lg = nullptr; a.gu = f->d_gu;
		if((rc = launch_cell_kernel(P.gradient, 0, false, a, s)) != 0) return rc;
		f->launches++;
		WenoArgs w;
		w.m = f->mesh->d; w.lambda = f->gas.limiter_param; w.gu = f->d_gu; w.lg = f->d_lg; w.ordered = a.ordered;
		if(f->roles.active) w.dist = f->roles.weno;
		if((rc = launch_weno_kernel(w, s)) != 0) return rc;
		f->launches++;
	}
	else if(P.recon == FVG_RECON_VANALBADA) {
		a.lg = nullptr; a.gu = f->d_gu;
		if((rc = launch_cell_kernel(P.gradient, 0, false, a, s)) != 0) return rc;
		f->launches++;
	}
	else {
		const int lim = limiter_mode(P.recon);
		a.lg = f->d_lg;
		// unlimited gradients are needed separately only when a limiter changes them and the viscous flux wants them
		a.gu = (P.visc != VISC_NONE && lim != 0) ? f->d_gu : nullptr;
		if((rc = launch_cell_kernel(P.gradient, lim, false, a, s)) != 0) return rc;
		f->launches++;
	}
	return 0;
}

static const double *viscous_gradients(const fvg_flow *f)
{
	const FlowPlan &P = f->plan;
	if(P.recon == FVG_RECON_WENO || P.recon == FVG_RECON_VANALBADA) return f->d_gu;
	if(limiter_mode(P.recon) != 0) return f->d_gu;
	return f->d_lg;   // no limiter: the stored gradients are the unlimited ones
}

int make_row_tensor_map(CUtensorMap *tm, const double *base, size_t nrows, int width, int box_rows)
{
	typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
	                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static EncodeTiled encode = nullptr;
	if(!encode) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qr;
		const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
		if(e != cudaSuccess || !fn || qr != cudaDriverEntryPointSuccess) {
			set_error("cuTensorMapEncodeTiled is not available from this driver");
			return FVG_ERR_CUDA;
		}
		encode = reinterpret_cast<EncodeTiled>(fn);
	}
	if((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (width != 4 && width != 8) || box_rows < 1 || box_rows > 256) {
		set_error("tensor map: the array must be 16-byte aligned, rows of 4 or 8 doubles, at most 256 rows per tile");
		return FVG_ERR_INVALID;
	}
	const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)std::max<size_t>(nrows, 1)};
	const cuuint64_t gstride[1] = {(cuuint64_t)width*8};
	const cuuint32_t box[2] = {(cuuint32_t)width, (cuuint32_t)box_rows};
	const cuuint32_t estride[2] = {1, 1};
	const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estride,
	                          CU_TENSOR_MAP_INTERLEAVE_NONE, width == 4 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B,
	                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if(r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return FVG_ERR_CUDA; }
	return 0;
}

int run_face_pass(fvg_flow *f, const double *u, int epilogue, int accumulate, int gettimesteps,
                  double *res, double *dtm, double cfl, double *unew, cudaStream_t s, int tile0, int tile1, const int *dst_idx)
{
	const FlowPlan &P = f->plan;
	FaceArgs a;
	a.m = f->mesh->d; a.gas = f->gas; a.u = u;
	a.lg = f->d_lg; a.gu = viscous_gradients(f);
	a.epilogue = epilogue; a.accumulate = accumulate; a.gettimesteps = gettimesteps;
	a.res = res; a.dtm = dtm; a.cfl = cfl; a.unew = unew; a.partial = f->d_partial;
	a.prefetch_distance = f->prefetch_distance;
	resolve_tiles(f, tile0, tile1, a.ordered);
	a.tile0 = tile0; a.tile1 = tile1;
	a.gs_u = f->gs_u; a.gs_g = f->gs_g;
	a.dst_idx = dst_idx;
	if(f->roles.active) a.dist = f->roles.face;
	const int recon = !P.order2 ? FR_FIRST : (P.recon == FVG_RECON_VANALBADA ? FR_MUSCL : FR_LINEAR);
	int rt = make_row_tensor_map(&a.tm_u, u, (size_t)a.m.ncell, 4, tile_box_rows(a.m.TC));
	if(rt == 0 && recon == FR_LINEAR) rt = make_row_tensor_map(&a.tm_g, a.lg, (size_t)a.m.ncell, 8, tile_box_rows(a.m.TC));
	if(rt != 0) return rt;
	FaceLauncher L = face_launcher(P.flux);
	if(!L) { set_error("unknown flux id"); return FVG_ERR_INVALID; }
	const int rc = L(recon, P.visc, a, s);
	if(rc == 0) f->launches++;
	return rc;
}

} // namespace fvg

using namespace fvg;

extern "C" {

// ------------------------------------------------------------------------------------ device memory

int fvg_malloc(void **d_ptr, unsigned long long bytes)
{
	if(!d_ptr) { set_error("fvg_malloc: null argument"); return FVG_ERR_INVALID; }
	*d_ptr = nullptr;
	FVG_CUDA(cudaMalloc(d_ptr, std::max<size_t>((size_t)bytes, 8)));
	return 0;
}

int fvg_free(void *d_ptr)
{
	if(d_ptr) FVG_CUDA(cudaFree(d_ptr));
	return 0;
}

int fvg_memcpy(void *dst, const void *src, unsigned long long bytes, int kind)
{
	if(bytes == 0) return 0;
	if(!dst || !src || kind < 0 || kind > 2) { set_error("fvg_memcpy: bad argument"); return FVG_ERR_INVALID; }
	const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
	FVG_CUDA(cudaMemcpy(dst, src, (size_t)bytes, k));
	return 0;
}

int fvg_memset(void *d_ptr, int value, unsigned long long bytes)
{
	if(bytes == 0) return 0;
	if(!d_ptr) { set_error("fvg_memset: null argument"); return FVG_ERR_INVALID; }
	FVG_CUDA(cudaMemset(d_ptr, value, (size_t)bytes));
	return 0;
}

// ------------------------------------------------------------------------------------ host mesh

static int finish_umesh(std::unique_ptr<fvg_umesh> &h, fvg_umesh **out)
{
	h->m.correctBoundaryFaceOrientation();
	h->m.compute_topological();
	h->m.compute_areas();
	h->m.compute_face_data();
	*out = h.release();
	return 0;
}

int fvg_umesh_read(const char *path, fvg_umesh **out)
{
	if(!path || !out) { set_error("fvg_umesh_read: null argument"); return FVG_ERR_INVALID; }
	*out = nullptr;
	try {
		std::unique_ptr<fvg_umesh> h(new fvg_umesh{fvens::UMesh<double,2>(fvens::readMesh(path))});
		return finish_umesh(h, out);
	} catch(std::exception &e) { set_error(e.what()); return FVG_ERR_IO; }
}

int fvg_umesh_write_gmsh2(const fvg_umesh *m, const char *path)
{
	if(!m || !path) { set_error("fvg_umesh_write_gmsh2: null argument"); return FVG_ERR_INVALID; }
	try { m->m.writeGmsh2(path); }
	catch(std::exception &e) { set_error(e.what()); return FVG_ERR_IO; }
	return 0;
}

int fvg_umesh_from_arrays(int npoin, const double *coords, int nelem, const int *nnode,
                          const int *inpoel, int nbface, const int *bface, fvg_umesh **out)
{
	if(!coords || !nnode || !inpoel || !out || (nbface > 0 && !bface) || npoin <= 0 || nelem <= 0) {
		set_error("fvg_umesh_from_arrays: bad argument"); return FVG_ERR_INVALID;
	}
	*out = nullptr;
	try {
		fvens::MeshData md;
		md.npoin = npoin; md.nelem = nelem; md.nbface = nbface;
		md.nnode.assign(nnode, nnode+nelem); md.nfael = md.nnode;
		md.maxnnode = 3;
		for(int i = 0; i < nelem; i++) {
			if(nnode[i] != 3 && nnode[i] != 4) { set_error("fvg_umesh_from_arrays: cells must have 3 or 4 nodes"); return FVG_ERR_INVALID; }
			if(nnode[i] == 4) md.maxnnode = 4;
		}
		md.maxnfael = md.maxnnode; md.nnofa = 2; md.nbtag = 1; md.ndtag = 0;
		md.coords.assign(coords, coords + 2*(size_t)npoin);
		md.inpoel.assign((size_t)nelem*md.maxnnode, -1);
		for(int i = 0; i < nelem; i++)
			for(int j = 0; j < nnode[i]; j++) {
				const int p = inpoel[4*(size_t)i+j];
				if(p < 0 || p >= npoin) { set_error("fvg_umesh_from_arrays: node index out of range"); return FVG_ERR_INVALID; }
				md.inpoel[(size_t)i*md.maxnnode+j] = p;
			}
		md.bface.assign(bface, bface + 3*(size_t)nbface);
		std::unique_ptr<fvg_umesh> h(new fvg_umesh{fvens::UMesh<double,2>(md)});
		return finish_umesh(h, out);
	} catch(std::exception &e) { set_error(e.what()); return FVG_ERR_INVALID; }
}

void fvg_umesh_destroy(fvg_umesh *m) { delete m; }

int fvg_umesh_reorder_cells(fvg_umesh *m, const int *perm)
{
	if(!m || !perm) { set_error("fvg_umesh_reorder_cells: null argument"); return FVG_ERR_INVALID; }
	try {
		const int n = m->m.gnelem();
		std::vector<char> seen(n, 0);
		for(int i = 0; i < n; i++) {
			if(perm[i] < 0 || perm[i] >= n || seen[perm[i]]) { set_error("fvg_umesh_reorder_cells: not a permutation"); return FVG_ERR_INVALID; }
			seen[perm[i]] = 1;
		}
		m->m.reorder_cells(perm);
		m->m.compute_topological();
		m->m.compute_areas();
		m->m.compute_face_data();
	} catch(std::exception &e) { set_error(e.what()); return FVG_ERR_INVALID; }
	return 0;
}

int fvg_umesh_rcm_ordering(const fvg_umesh *m, int *perm)
{
	if(!m || !perm) { set_error("fvg_umesh_rcm_ordering: null argument"); return FVG_ERR_INVALID; }
	std::vector<int> p;
	rcm_order(m->m.gnelem(), m->m.gmaxnfael(), m->m.esuelData(), m->m.nnodeData(), p);
	std::memcpy(perm, p.data(), sizeof(int)*p.size());
	return 0;
}

int fvg_umesh_cell_adjacency(const fvg_umesh *m, int *ptrs, int *store)
{
	if(!m || !ptrs) { set_error("fvg_umesh_cell_adjacency: null argument"); return FVG_ERR_INVALID; }
	const auto &M = m->m;
	const int n = M.gnelem(), mw = M.gmaxnfael();
	const int *const esuel = M.esuelData();
	ptrs[0] = 0;
	for(int i = 0; i < n; i++) {
		int k = 0;
		for(int j = 0; j < M.gnfael(i); j++) { const int e = esuel[(size_t)i*mw+j]; if(e >= 0 && e < n) k++; }
		ptrs[i+1] = ptrs[i] + k;
	}
	if(store)
		for(int i = 0; i < n; i++) {
			int p = ptrs[i];
			for(int j = 0; j < M.gnfael(i); j++) { const int e = esuel[(size_t)i*mw+j]; if(e >= 0 && e < n) store[p++] = e; }
		}
	return 0;
}

int fvg_umesh_write_scotch_graph(const fvg_umesh *m, const char *path)
{
	if(!m || !path) { set_error("fvg_umesh_write_scotch_graph: null argument"); return FVG_ERR_INVALID; }
	const int n = m->m.gnelem();
	std::vector<int> ptrs((size_t)n + 1);
	fvg_umesh_cell_adjacency(m, ptrs.data(), nullptr);
	std::vector<int> store((size_t)std::max(ptrs[n], 1));
	fvg_umesh_cell_adjacency(m, ptrs.data(), store.data());
	FILE *f = std::fopen(path, "w");
	if(!f) { set_error(std::string("fvg_umesh_write_scotch_graph: cannot open ") + path); return FVG_ERR_IO; }
	// Scotch source graph (.grf): version, vertex and arc counts, base value and flags (no weights, no labels),
	// then per vertex its degree and neighbours
	std::fprintf(f, "0\n%d %d\n0 000\n", n, ptrs[n]);
	for(int i = 0; i < n; i++) {
		std::fprintf(f, "%d", ptrs[i+1] - ptrs[i]);
		for(int p = ptrs[i]; p < ptrs[i+1]; p++) std::fprintf(f, " %d", store[p]);
		std::fprintf(f, "\n");
	}
	std::fclose(f);
	return 0;
}

int fvg_partition_read_scotch_map(const fvg_umesh *m, const char *path, int *cell_rank, int *nparts)
{
	if(!m || !path || !cell_rank) { set_error("fvg_partition_read_scotch_map: null argument"); return FVG_ERR_INVALID; }
	FILE *f = std::fopen(path, "r");
	if(!f) { set_error(std::string("fvg_partition_read_scotch_map: cannot open ") + path); return FVG_ERR_IO; }
	const int n = m->m.gnelem();
	int cnt = 0, rc = 0, maxp = -1;
	if(std::fscanf(f, "%d", &cnt) != 1 || cnt != n) { set_error("fvg_partition_read_scotch_map: the mapping does not have one line per cell"); rc = FVG_ERR_INVALID; }
	std::vector<char> seen((size_t)n, 0);
	for(int k = 0; k < n && rc == 0; k++) {
		int v = -1, p = -1;
		if(std::fscanf(f, "%d %d", &v, &p) != 2 || v < 0 || v >= n || p < 0 || seen[v]) { set_error("fvg_partition_read_scotch_map: bad or repeated entry"); rc = FVG_ERR_INVALID; break; }
		seen[v] = 1; cell_rank[v] = p; maxp = std::max(maxp, p);
	}
	std::fclose(f);
	if(rc == 0 && nparts) *nparts = maxp + 1;
	return rc;
}

int fvg_umesh_hilbert_ordering(const fvg_umesh *m, int *perm)
{
	if(!m || !perm) { set_error("fvg_umesh_hilbert_ordering: null argument"); return FVG_ERR_INVALID; }
	std::vector<double> rc(2*(size_t)m->m.gnelem());
	m->m.compute_cell_centres(rc.data());
	std::vector<int> p;
	hilbert_order(m->m.gnelem(), rc.data(), p);
	std::memcpy(perm, p.data(), sizeof(int)*p.size());
	return 0;
}

int fvg_umesh_view(const fvg_umesh *h, fvg_host_mesh *v)
{
	if(!h || !v) { set_error("fvg_umesh_view: null argument"); return FVG_ERR_INVALID; }
	const fvens::UMesh<double,2> &m = h->m;
	v->npoin = m.gnpoin(); v->nelem = m.gnelem(); v->nbface = m.gnbface(); v->naface = m.gnaface();
	v->ninface = m.gninface(); v->nconnface = m.gnConnFace(); v->maxnnode = m.gmaxnnode(); v->nbtag = m.gnbtag();
	v->coords = m.coordsData(); v->inpoel = m.inpoelData(); v->nnode = m.nnodeData();
	v->esuel = m.esuelData(); v->elemface = m.elemfaceData(); v->intfac = m.intfacData();
	v->btags = m.btagsData(); v->facemetric = m.facemetricData(); v->area = m.areaData();
	v->bpartner = m.periodicmapData();
	return 0;
}

int fvg_umesh_compute_periodic_map(fvg_umesh *h, int marker, int axis, int *npairs)
{
	if(!h || axis < 0 || axis > 1) { set_error("fvg_umesh_compute_periodic_map: bad argument"); return FVG_ERR_INVALID; }
	h->m.compute_periodic_map(marker, axis);
	int np = 0;
	for(int b = 0; b < h->m.gnbface(); b++) if(h->m.gbtags(b,0) == marker && h->m.gperiodicmap(b) >= 0) np++;
	if(npairs) *npairs = np/2;
	return 0;
}

// ------------------------------------------------------------------------------------ flow

int fvg_flow_create(fvg_mesh *mesh, const fvg_physics *phys, const fvg_numerics *num,
                    const fvg_bc *bcs, int nbc, fvg_flow **out)
{
	if(!mesh || !phys || !num || !out || (nbc > 0 && !bcs)) { set_error("fvg_flow_create: null argument"); return FVG_ERR_INVALID; }
	*out = nullptr;
	if(nbc > MAX_BC) { set_error("fvg_flow_create: at most 16 boundary conditions"); return FVG_ERR_UNSUPPORTED; }
	if(num->flux < 0 || num->flux >= FLUX_COUNT) { set_error("fvg_flow_create: unknown flux"); return FVG_ERR_INVALID; }
	if(num->gradient < 0 || num->gradient > 2) { set_error("fvg_flow_create: unknown gradient scheme"); return FVG_ERR_INVALID; }
	if(num->reconstruction < 0 || num->reconstruction > 4) { set_error("fvg_flow_create: unknown reconstruction"); return FVG_ERR_INVALID; }
	if(!(phys->gamma > 1.0) || !(phys->Minf > 0.0)) { set_error("fvg_flow_create: gamma must exceed 1 and Minf must be positive"); return FVG_ERR_INVALID; }
	if(mesh->device < 0) { set_error("fvg_flow_create: the mesh was built host-only (device = -2) and has no device arrays"); return FVG_ERR_INVALID; }
	FVG_CUDA(cudaSetDevice(mesh->device));

	std::unique_ptr<fvg_flow> f(new fvg_flow);
	f->mesh = mesh;
	f->phys = *phys;
	f->gas = make_gas(*phys, num->limiter_param);
	for(int i = 0; i < nbc; i++)
		if(bcs[i].type < 0 || bcs[i].type > 7) {
			// reference: create_const_flowBCs throws for types without a FlowBC class (abc.cpp:493-494)
			set_error("fvg_flow_create: boundary condition type " + std::to_string(bcs[i].type) + " is not available");
			return FVG_ERR_UNSUPPORTED;
		}
	// the BC table is indexed by the mesh's marker slots; every marker present in the mesh needs a BC
	// (reference: bcs.at(tag) throws std::out_of_range in compute_boundary_state, flow_spatial.cpp:92)
	f->gas.nbc = (int)mesh->h_markers.size();
	for(size_t sl = 0; sl < mesh->h_markers.size(); sl++) {
		int k = -1;
		for(int i = 0; i < nbc; i++) if(bcs[i].tag == mesh->h_markers[sl]) k = i;
		if(k < 0) { set_error("fvg_flow_create: no boundary condition for marker " + std::to_string(mesh->h_markers[sl])); return FVG_ERR_INVALID; }
		// a periodic marker is one whose faces the mesh has paired (fvg_host_mesh::bpartner): no ghost state exists for it.
		// The reference has no periodic FlowBC (abc.cpp:493-494 throws); here the pairing IS the boundary condition.
		const bool paired = sl < mesh->h_marker_periodic.size() && mesh->h_marker_periodic[sl] != 0;
		if((bcs[k].type == PERIODIC_BC) != paired) {
			set_error("fvg_flow_create: marker " + std::to_string(mesh->h_markers[sl]) + (paired ? " is paired as periodic in the mesh but its boundary condition is not 'periodic'"
			          : " has the periodic boundary condition but the mesh holds no pairing for it (UMesh::compute_periodic_map)"));
			return FVG_ERR_UNSUPPORTED;
		}
		f->gas.bc[sl].tag = bcs[k].tag; f->gas.bc[sl].type = bcs[k].type;
		f->gas.bc[sl].v0 = bcs[k].vals[0]; f->gas.bc[sl].v1 = bcs[k].vals[1];
	}
	FlowPlan &P = f->plan;
	P.flux = num->flux; P.gradient = num->gradient; P.recon = num->reconstruction;
	P.order2 = num->order2 ? 1 : 0; P.bnd_policy = num->bnd_policy;
	P.visc = !phys->viscous_sim ? VISC_NONE : (phys->const_visc ? VISC_CONST : VISC_SUTHERLAND);
	const bool limited = limiter_mode(P.recon) != 0;
	P.need_lg = P.order2 && P.recon != FVG_RECON_VANALBADA;
	P.need_gu = P.order2 && (P.recon == FVG_RECON_WENO || P.recon == FVG_RECON_VANALBADA ||
	                         (P.visc != VISC_NONE && limited));

	const int n = mesh->d.ncell + mesh->d.nghost;      // gradient rows exist for the ghosts too (filled by the halo exchange)
	int rc;
	if(P.need_lg && (rc = dev_alloc(f.get(), &f->d_lg, 8*(size_t)n)) != 0) return rc;
	if(P.need_gu && (rc = dev_alloc(f.get(), &f->d_gu, 8*(size_t)n)) != 0) return rc;
	if((rc = dev_alloc(f.get(), &f->d_partial, (size_t)mesh->d.ntile*(FACE_BLOCK/32))) != 0) return rc;
	if((rc = dev_alloc(f.get(), &f->d_norm, 1)) != 0) return rc;
	FVG_CUDA(cudaMallocHost((void**)&f->h_norm, sizeof(double)));
	{
		// L2 prefetch distance = about one wave of resident CTAs (overridable for experiments)
		int sms = 148;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, mesh->device);
		const char *env = getenv("FVG_PREFETCH_WAVES");
		const double waves = env ? atof(env) : 1.0;
		f->prefetch_distance = (int)(waves*3*sms);
	}
	*out = f.release();
	return 0;
}

/// Single-rank mesh with periodic ghost cells: the exchange engine (dist.cu) with this rank as its own peer, created
/// on first use. The periodic rows travel exactly like the rows of a partition - pushed by the producing kernels into
/// the (own) window, waited for in the tiles that see a ghost cell.
static int self_exchange(fvg_flow *f, fvg_dist **out)
{
	if(!f->self_dist) {
		fvg_dist *d = nullptr;
		int rc = fvg_dist_create(f, &d);
		if(rc != 0) return rc;
		const int rcv = f->mesh->recv_counts.empty() ? 0 : f->mesh->recv_counts[0];
		unsigned char handle[64] = {0};
		if((rc = fvg_dist_connect(d, handle, &rcv)) != 0) { fvg_dist_destroy(d); return rc; }
		f->self_dist = d;
	}
	*out = f->self_dist;
	return 0;
}
static bool has_periodic_ghosts(const fvg_flow *f) { return f->mesh->nranks == 1 && f->mesh->d.nghost > 0; }
static int reject_periodic(const fvg_flow *f, const char *who)
{
	if(f->mesh->d.nghost > 0 && f->mesh->nranks == 1) {
		set_error(std::string(who) + ": not available on a mesh with periodic boundaries (use fvg_residual and the drivers built on it)");
		return FVG_ERR_UNSUPPORTED;
	}
	return 0;
}

void fvg_flow_destroy(fvg_flow *f)
{
	if(!f) return;
	if(f->self_dist) fvg_dist_destroy(f->self_dist);
	for(void *p : f->allocs) cudaFree(p);
	if(f->h_norm) cudaFreeHost(f->h_norm);
	for(cudaEvent_t e : f->pipe.ev_up) cudaEventDestroy(e);
	for(cudaEvent_t e : f->pipe.ev_face) cudaEventDestroy(e);
	if(f->pipe.ev_start) cudaEventDestroy(f->pipe.ev_start);
	if(f->pipe.s_in) cudaStreamDestroy(f->pipe.s_in);
	if(f->pipe.s_run) cudaStreamDestroy(f->pipe.s_run);
	if(f->pipe.s_out) cudaStreamDestroy(f->pipe.s_out);
	delete f;
}

int fvg_flow_launch_count(const fvg_flow *f, long long *count)
{
	if(!f || !count) { set_error("fvg_flow_launch_count: null argument"); return FVG_ERR_INVALID; }
	*count = f->launches;
	return 0;
}

/// records a timing event on the stream when per-pass timing is enabled
static int mark(fvg_flow *f, cudaStream_t s)
{
	if(!f->timing) return 0;
	cudaEvent_t e;
	FVG_CUDA(cudaEventCreate(&e));
	f->ev.push_back(e);
	FVG_CUDA(cudaEventRecord(e, s));
	return 0;
}

int fvg_flow_timing(fvg_flow *f, int enable, double *h_out3)
{
	if(!f) { set_error("fvg_flow_timing: null argument"); return FVG_ERR_INVALID; }
	double cell = 0, face = 0; int n = 0;
	int rc = 0;
	if(!f->ev.empty()) {
		const cudaError_t e = cudaDeviceSynchronize();
		if(e != cudaSuccess) rc = cuda_fail(e, "cudaDeviceSynchronize", __FILE__, __LINE__);
		for(size_t k = 0; rc == 0 && k + 3 <= f->ev.size(); k += 3) {
			float a = 0, b = 0;
			cudaEventElapsedTime(&a, f->ev[k], f->ev[k+1]);
			cudaEventElapsedTime(&b, f->ev[k+1], f->ev[k+2]);
			cell += a; face += b; n++;
		}
		for(cudaEvent_t e2 : f->ev) cudaEventDestroy(e2);
		f->ev.clear();
	}
	if(h_out3) { h_out3[0] = cell; h_out3[1] = face; h_out3[2] = n; }
	f->timing = enable != 0;
	return rc;
}

/// lazily allocated scratch
static int ensure(fvg_flow *f, double **p, size_t count) { return *p ? 0 : dev_alloc(f, p, count); }

int fvg_residual(fvg_flow *f, const double *d_u, double *d_res, int accumulate, int gettimesteps,
                 double *d_dtm, void *stream)
{
	if(!f || !d_u || !d_res || (gettimesteps && !d_dtm)) { set_error("fvg_residual: null argument"); return FVG_ERR_INVALID; }
	if(f->mesh->nranks > 1) { set_error("fvg_residual: a subdomain mesh needs its neighbours' rows; use fvg_dist_residual"); return FVG_ERR_UNSUPPORTED; }
	if(has_periodic_ghosts(f)) {
		fvg_dist *sd = nullptr;
		const int rcd = self_exchange(f, &sd);
		return rcd != 0 ? rcd : fvg_dist_residual(sd, d_u, d_res, accumulate, gettimesteps, d_dtm, stream);
	}
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const DMesh &D = f->mesh->d;
	const int n = D.ncell;
	int rc;
	if(f->mesh->identity_perm) {
		if((rc = mark(f, s)) != 0) return rc;
		if((rc = run_gradient_pass(f, d_u, s)) != 0) return rc;
		if((rc = mark(f, s)) != 0) return rc;
		if((rc = run_face_pass(f, d_u, EP_RESIDUAL, accumulate, gettimesteps, d_res, d_dtm, 0.0, nullptr, s)) != 0) return rc;
		return mark(f, s);
	}
	// renumbered mesh, caller-ordered arrays. Second order: the permutation is fused into the passes - the gradient
	// pass gathers the state rows through new2old and leaves a device-ordered copy for the face pass, whose residual /
	// time-step stores go straight to the caller's rows (no permutation kernels, SURVEY 8b drop-in path)
	if((rc = ensure(f, &f->d_uperm, 4*(size_t)n)) != 0) return rc;
	if(f->plan.order2) {
		if((rc = mark(f, s)) != 0) return rc;
		if((rc = run_gradient_pass(f, d_u, s, 0, -1, D.new2old, f->d_uperm)) != 0) return rc;
		if((rc = mark(f, s)) != 0) return rc;
		if((rc = run_face_pass(f, f->d_uperm, EP_RESIDUAL, accumulate, gettimesteps, d_res, d_dtm, 0.0, nullptr, s, 0, -1, D.new2old)) != 0) return rc;
		return mark(f, s);
	}
	// first order (no gradient pass to gather in): one gather kernel, then the face pass scatters its stores
	if((rc = launch_permute_rows(d_u, f->d_uperm, D.new2old, n, 4, true, false, s)) != 0) return rc;
	f->launches++;
	return run_face_pass(f, f->d_uperm, EP_RESIDUAL, accumulate, gettimesteps, d_res, d_dtm, 0.0, nullptr, s, 0, -1, D.new2old);
}

/** Plan of the chunked host-buffer pipeline. The tiles are cut into K ranges of consecutive tiles (compact patches
 * of the mesh, since tiles follow the locality order). A chunk's gradient pass needs the state rows of its own and
 * its halo cells, i.e. the uploads of the chunks in deps[c]; its face pass needs the gradient passes of the same
 * chunks. Uploading the chunks in a sweep along the longer axis of the domain makes these sets complete early, so
 * the first residual rows travel back to the host while most of the state is still on its way in. */
static void plan_host_pipe(fvg_flow *f)
{
	fvg_flow::HostPipe &P = f->pipe;
	P.planned = true;
	P.K = 0;
	const fvg_mesh *m = f->mesh;
	const int ntile = m->d.ntile;
	int K = 48;
	if(const char *ev = getenv("FVG_HOST_CHUNKS")) K = atoi(ev);
	K = std::min(std::min(K, 64), ntile/8);
	if(K < 2 || !m->identity_perm || m->nranks > 1 || m->d.nghost > 0 || m->h_thoff.empty()) return;
	P.tile0.resize(K+1);
	for(int c = 0; c <= K; c++) P.tile0[c] = (int)((long long)ntile*c/K);
	std::vector<int> chunk_of_tile(ntile);
	for(int c = 0; c < K; c++) for(int t = P.tile0[c]; t < P.tile0[c+1]; t++) chunk_of_tile[t] = c;
	auto chunk_of_cell = [&](int i) {
		const int t = (int)(std::upper_bound(m->h_tcell0.begin(), m->h_tcell0.end(), i) - m->h_tcell0.begin()) - 1;
		return chunk_of_tile[t];
	};
	P.deps.assign(K, 0ull);
	std::vector<double> cx(K, 0.0), cy(K, 0.0);
	double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
	for(int c = 0; c < K; c++) {
		P.deps[c] |= 1ull << c;
		for(int t = P.tile0[c]; t < P.tile0[c+1]; t++)
			for(int h = m->h_thoff[t]; h < m->h_thoff[t+1]; h++)
				if(m->h_thalo[h] < m->d.ncell) P.deps[c] |= 1ull << chunk_of_cell(m->h_thalo[h]);
		const int i0 = m->h_tcell0[P.tile0[c]], i1 = m->h_tcell0[P.tile0[c+1]];
		for(int i = i0; i < i1; i++) {
			const double x = m->h_rc[2*(size_t)i], y = m->h_rc[2*(size_t)i+1];
			cx[c] += x; cy[c] += y;
			lo[0] = std::min(lo[0], x); hi[0] = std::max(hi[0], x); lo[1] = std::min(lo[1], y); hi[1] = std::max(hi[1], y);
		}
		cx[c] /= std::max(1, i1 - i0); cy[c] /= std::max(1, i1 - i0);
	}
	const std::vector<double> &key = (hi[0] - lo[0] >= hi[1] - lo[1]) ? cx : cy;
	P.order.resize(K);
	for(int c = 0; c < K; c++) P.order[c] = c;
	std::stable_sort(P.order.begin(), P.order.end(), [&](int a, int b) { return key[a] < key[b]; });
	bool ok = cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking) == cudaSuccess
	       && cudaStreamCreateWithFlags(&P.s_run, cudaStreamNonBlocking) == cudaSuccess
	       && cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking) == cudaSuccess;
	P.ev_up.resize(K); P.ev_face.resize(K);
	for(int c = 0; c < K && ok; c++)
		ok = cudaEventCreateWithFlags(&P.ev_up[c], cudaEventDisableTiming) == cudaSuccess
		  && cudaEventCreateWithFlags(&P.ev_face[c], cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&P.ev_start, cudaEventDisableTiming) == cudaSuccess;
	if(!ok) { cudaGetLastError(); return; }
	P.K = K;
}

int fvg_residual_host(fvg_flow *f, const double *h_u, double *h_res, int accumulate, int gettimesteps, double *h_dtm)
{
	if(!f || !h_u || !h_res || (gettimesteps && !h_dtm)) { set_error("fvg_residual_host: null argument"); return FVG_ERR_INVALID; }
	FVG_CUDA(cudaSetDevice(f->mesh->device));
	const size_t n = f->mesh->d.ncell;
	int rc;
	if((rc = ensure(f, &f->d_hu, 4*n)) != 0) return rc;
	if((rc = ensure(f, &f->d_hr, 4*n)) != 0) return rc;
	if((rc = ensure(f, &f->d_hdt, n)) != 0) return rc;
	if(!f->pipe.planned) plan_host_pipe(f);
	const fvg_flow::HostPipe &P = f->pipe;
	if(P.K >= 2 && f->plan.recon != FVG_RECON_WENO && !f->timing) {
		// chunked pipeline: uploads in sweep order on one stream, the two passes per chunk on a second one as soon
		// as their inputs are complete, downloads of finished chunks on a third (PCIe is full duplex)
		const std::vector<int> &tc0 = f->mesh->h_tcell0;
		std::vector<int> pos(P.K);
		for(int j = 0; j < P.K; j++) pos[P.order[j]] = j;
		unsigned long long uploaded = 0, cells_done = 0, faces_done = 0;
		// earlier calls on the default stream may still be using the flow's gradient buffers
		FVG_CUDA(cudaEventRecord(P.ev_start, nullptr));
		FVG_CUDA(cudaStreamWaitEvent(P.s_in, P.ev_start, 0));
		FVG_CUDA(cudaStreamWaitEvent(P.s_run, P.ev_start, 0));
		for(int j = 0; j < P.K; j++) {
			const int c = P.order[j];
			const size_t i0 = (size_t)tc0[P.tile0[c]], i1 = (size_t)tc0[P.tile0[c+1]];
			FVG_CUDA(cudaMemcpyAsync(f->d_hu + 4*i0, h_u + 4*i0, 4*(i1 - i0)*sizeof(double), cudaMemcpyHostToDevice, P.s_in));
			// the reference adds into the caller's residual: its rows ride up with the state rows and the face
			// pass of the chunk accumulates on the device
			if(accumulate) FVG_CUDA(cudaMemcpyAsync(f->d_hr + 4*i0, h_res + 4*i0, 4*(i1 - i0)*sizeof(double), cudaMemcpyHostToDevice, P.s_in));
			FVG_CUDA(cudaEventRecord(P.ev_up[j], P.s_in));
			uploaded |= 1ull << c;
			bool waited = false;
			// gradient passes whose inputs are now complete (in upload order), then the face passes they release
			for(int q = 0; q < P.K; q++) {
				const int d = P.order[q];
				if((cells_done >> d) & 1ull || (P.deps[d] & ~uploaded) != 0) continue;
				if(!waited) { FVG_CUDA(cudaStreamWaitEvent(P.s_run, P.ev_up[j], 0)); waited = true; }
				if((rc = run_gradient_pass(f, f->d_hu, P.s_run, P.tile0[d], P.tile0[d+1])) != 0) return rc;
				cells_done |= 1ull << d;
			}
			for(int q = 0; q < P.K; q++) {
				const int d = P.order[q];
				if((faces_done >> d) & 1ull || (P.deps[d] & ~cells_done) != 0) continue;
				if((rc = run_face_pass(f, f->d_hu, EP_RESIDUAL, accumulate, gettimesteps, f->d_hr, f->d_hdt, 0.0, nullptr, P.s_run,
				                       P.tile0[d], P.tile0[d+1])) != 0) return rc;
				faces_done |= 1ull << d;
				FVG_CUDA(cudaEventRecord(P.ev_face[d], P.s_run));
				FVG_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_face[d], 0));
				const size_t a0 = (size_t)tc0[P.tile0[d]], a1 = (size_t)tc0[P.tile0[d+1]];
				FVG_CUDA(cudaMemcpyAsync(h_res + 4*a0, f->d_hr + 4*a0, 4*(a1 - a0)*sizeof(double), cudaMemcpyDeviceToHost, P.s_out));
				if(gettimesteps) FVG_CUDA(cudaMemcpyAsync(h_dtm + a0, f->d_hdt + a0, (a1 - a0)*sizeof(double), cudaMemcpyDeviceToHost, P.s_out));
			}
		}
		FVG_CUDA(cudaStreamSynchronize(P.s_out));
		FVG_CUDA(cudaStreamSynchronize(P.s_run));
		if(faces_done != (P.K == 64 ? ~0ull : (1ull << P.K) - 1)) { set_error("fvg_residual_host: internal error, chunk schedule incomplete"); return FVG_ERR_INVALID; }
		return 0;
	}
	cudaStream_t s = nullptr;
	FVG_CUDA(cudaMemcpyAsync(f->d_hu, h_u, 4*n*sizeof(double), cudaMemcpyHostToDevice, s));
	// the reference adds into the caller's residual: upload it and accumulate on the device
	if(accumulate) FVG_CUDA(cudaMemcpyAsync(f->d_hr, h_res, 4*n*sizeof(double), cudaMemcpyHostToDevice, s));
	if((rc = fvg_residual(f, f->d_hu, f->d_hr, accumulate, gettimesteps, f->d_hdt, s)) != 0) return rc;
	FVG_CUDA(cudaMemcpyAsync(h_res, f->d_hr, 4*n*sizeof(double), cudaMemcpyDeviceToHost, s));
	if(gettimesteps) FVG_CUDA(cudaMemcpyAsync(h_dtm, f->d_hdt, n*sizeof(double), cudaMemcpyDeviceToHost, s));
	FVG_CUDA(cudaStreamSynchronize(s));
	return 0;
}

/// Scratch copy of a reference-ordered cell array in device order (or the array itself)
static int to_device_order(fvg_flow *f, const double *src, int width, double **scratch, const double **out, cudaStream_t s)
{
	if(f->mesh->identity_perm) { *out = src; return 0; }
	const int n = f->mesh->d.ncell;
	FVG_CUDA(cudaMallocAsync((void**)scratch, sizeof(double)*(size_t)n*width, s));
	const int rc = launch_permute_rows(src, *scratch, f->mesh->d.new2old, n, width, true, false, s);
	*out = *scratch;
	f->launches++;
	return rc;
}

int fvg_gradients(fvg_flow *f, const double *d_uprim, const double *d_ug, double *d_grad, void *stream)
{
	if(!f || !d_uprim || !d_grad || (f->mesh->d.nbface > 0 && !d_ug)) { set_error("fvg_gradients: null argument"); return FVG_ERR_INVALID; }
	{ const int rp = reject_periodic(f, "fvg_gradients"); if(rp != 0) return rp; }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const DMesh &D = f->mesh->d;
	double *su = nullptr, *sg = nullptr;
	const double *u = nullptr;
	int rc = to_device_order(f, d_uprim, 4, &su, &u, s);
	CellArgs a;
	a.m = D; a.gas = f->gas; a.u = u; a.ug = d_ug; a.gin = nullptr;
	a.lg = nullptr; a.bnd_policy = f->plan.bnd_policy; a.prefetch_distance = 0;
	if(rc == 0 && !f->mesh->identity_perm) {
		const cudaError_t e = cudaMallocAsync((void**)&sg, sizeof(double)*8*(size_t)D.ncell, s);
		if(e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
	}
	a.gu = f->mesh->identity_perm ? d_grad : sg;
	if(rc == 0) { rc = launch_cell_kernel(f->plan.gradient, 0, true, a, s); f->launches++; }
	if(rc == 0 && !f->mesh->identity_perm) { rc = launch_permute_rows(sg, d_grad, D.new2old, D.ncell, 8, false, false, s); f->launches++; }
	if(su) cudaFreeAsync(su, s);
	if(sg) cudaFreeAsync(sg, s);
	return rc;
}

int fvg_face_values(fvg_flow *f, const double *d_uprim, const double *d_ug, const double *d_grad,
                    double *d_ufl, double *d_ufr, void *stream)
{
	if(!f || !d_uprim || !d_grad || !d_ufl || !d_ufr || (f->mesh->d.nbface > 0 && !d_ug)) { set_error("fvg_face_values: null argument"); return FVG_ERR_INVALID; }
	{ const int rp = reject_periodic(f, "fvg_face_values"); if(rp != 0) return rp; }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const DMesh &D = f->mesh->d;
	const FlowPlan &P = f->plan;
	double *su = nullptr, *sg = nullptr, *slg = nullptr;
	const double *u = nullptr, *g = nullptr;
	int rc = to_device_order(f, d_uprim, 4, &su, &u, s);
	if(rc == 0) rc = to_device_order(f, d_grad, 8, &sg, &g, s);
	FaceValArgs fa;
	fa.m = D; fa.up = u; fa.ug = d_ug; fa.ufl = d_ufl; fa.ufr = d_ufr; fa.muscl = 0; fa.g = g;
	if(rc == 0 && P.recon != FVG_RECON_NONE && P.recon != FVG_RECON_VANALBADA) {
		const cudaError_t e = cudaMallocAsync((void**)&slg, sizeof(double)*8*(size_t)D.ncell, s);
		if(e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
		if(rc == 0 && P.recon == FVG_RECON_WENO) {
			WenoArgs w;
			w.m = D; w.lambda = f->gas.limiter_param; w.gu = g; w.lg = slg;
			rc = launch_weno_kernel(w, s); f->launches++;
		}
		else if(rc == 0) {
			CellArgs a;
			a.m = D; a.gas = f->gas; a.u = u; a.ug = d_ug; a.gin = g; a.lg = slg; a.gu = nullptr;
			a.bnd_policy = P.bnd_policy; a.prefetch_distance = 0;
			rc = launch_cell_kernel(3 /*given*/, limiter_mode(P.recon), true, a, s); f->launches++;
		}
		fa.g = slg;
	}
	if(P.recon == FVG_RECON_VANALBADA) fa.muscl = 1;
	if(rc == 0) { rc = launch_face_values(fa, s); f->launches++; }
	if(su) cudaFreeAsync(su, s);
	if(sg) cudaFreeAsync(sg, s);
	if(slg) cudaFreeAsync(slg, s);
	return rc;
}

int fvg_boundary_states(fvg_flow *f, const double *d_ins, double *d_gs, void *stream)
{
	if(!f || !d_ins || !d_gs) { set_error("fvg_boundary_states: null argument"); return FVG_ERR_INVALID; }
	const int rc = launch_boundary_states(f->mesh->d, f->gas, d_ins, d_gs, static_cast<cudaStream_t>(stream));
	if(rc == 0) f->launches++;
	return rc;
}

int fvg_jacobian_vector_product(fvg_flow *f, const double *d_u, const double *d_res, const double *d_mdt, const double *d_x,
                                double eps, double *d_y, void *stream)
{
	if(!f || !d_u || !d_res || !d_mdt || !d_x || !d_y || !(eps > 0.0)) { set_error("fvg_jacobian_vector_product: bad argument"); return FVG_ERR_INVALID; }
	if(f->mesh->nranks > 1) { set_error("fvg_jacobian_vector_product: single-GPU entry point"); return FVG_ERR_UNSUPPORTED; }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const int n = f->mesh->d.ncell;
	const int nblk = 1024;
	int rc;
	if((rc = ensure(f, &f->d_jaux, 4*(size_t)n)) != 0) return rc;
	if((rc = ensure(f, &f->d_jyg, 4*(size_t)n)) != 0) return rc;
	if((rc = ensure(f, &f->d_jpart, nblk)) != 0) return rc;
	if((rc = ensure(f, &f->d_jnorm, 1)) != 0) return rc;
	// |x|^2 stays on the device; aux = u + eps/|x| x; yg = -r(aux); y = mdt x + (res - yg)/(eps/|x|)
	if((rc = launch_sumsq(d_x, 4ll*n, f->d_jpart, nblk, f->d_jnorm, s)) != 0) return rc;
	if((rc = launch_perturb(d_u, d_x, f->d_jnorm, eps, 4ll*n, f->d_jaux, s)) != 0) return rc;
	if((rc = fvg_residual(f, f->d_jaux, f->d_jyg, 0, 0, nullptr, s)) != 0) return rc;
	if((rc = launch_jvp_combine(d_x, d_res, f->d_jyg, d_mdt, f->d_jnorm, eps, n, 4, d_y, s)) != 0) return rc;
	f->launches += 4;
	return 0;
}

int fvg_get_gradients(fvg_flow *f, const double *d_u, double *d_grads, void *stream)
{
	if(!f || !d_u || !d_grads) { set_error("fvg_get_gradients: null argument"); return FVG_ERR_INVALID; }
	{ const int rp = reject_periodic(f, "fvg_get_gradients"); if(rp != 0) return rp; }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const DMesh &D = f->mesh->d;
	double *su = nullptr, *sug = nullptr, *sg = nullptr;
	const double *u = nullptr;
	int rc = to_device_order(f, d_u, 4, &su, &u, s);
	if(rc == 0) {
		const cudaError_t e = cudaMallocAsync((void**)&sug, sizeof(double)*4*(size_t)std::max(D.nbface, 1), s);
		if(e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
	}
	// conserved ghost states from the conserved cell states; then the gradient scheme on conserved variables
	if(rc == 0) { rc = launch_boundary_prim_ghosts(D, f->gas, u, sug, false, s); f->launches++; }
	if(rc == 0 && !f->mesh->identity_perm) {
		const cudaError_t e = cudaMallocAsync((void**)&sg, sizeof(double)*8*(size_t)D.ncell, s);
		if(e != cudaSuccess) rc = cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
	}
	CellArgs a;
	a.m = D; a.gas = f->gas; a.u = u; a.ug = sug; a.gin = nullptr; a.lg = nullptr;
	a.gu = f->mesh->identity_perm ? d_grads : sg; a.bnd_policy = f->plan.bnd_policy; a.prefetch_distance = 0;
	if(rc == 0) { rc = launch_cell_kernel(f->plan.gradient, 0, true, a, s); f->launches++; }
	if(rc == 0 && !f->mesh->identity_perm) { rc = launch_permute_rows(sg, d_grads, D.new2old, D.ncell, 8, false, false, s); f->launches++; }
	if(su) cudaFreeAsync(su, s);
	if(sug) cudaFreeAsync(sug, s);
	if(sg) cudaFreeAsync(sg, s);
	return rc;
}

int fvg_surface_data(fvg_flow *f, const double *d_u, const double *d_grads, int marker, double *h_out3)
{
	if(!f || !d_u || !d_grads || !h_out3) { set_error("fvg_surface_data: null argument"); return FVG_ERR_INVALID; }
	cudaStream_t s = nullptr;
	double *su = nullptr, *sg = nullptr, *d4 = nullptr;
	const double *u = nullptr, *g = nullptr;
	int rc = to_device_order(f, d_u, 4, &su, &u, s);
	if(rc == 0) rc = to_device_order(f, d_grads, 8, &sg, &g, s);
	if(rc == 0) { const cudaError_t e = cudaMalloc((void**)&d4, 4*sizeof(double)); if(e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
	if(rc == 0) { int slot = -1;
		for(size_t q = 0; q < f->mesh->h_markers.size(); q++) if(f->mesh->h_markers[q] == marker) slot = (int)q;
		rc = launch_surface_data(f->mesh->d, f->gas, f->phys.aoa, u, g, slot, d4, s); f->launches++; }
	double h4[4] = {0,0,0,0};
	if(rc == 0) { const cudaError_t e = cudaMemcpy(h4, d4, sizeof(h4), cudaMemcpyDeviceToHost); if(e != cudaSuccess) rc = cuda_fail(e, "D2H", __FILE__, __LINE__); }
	h_out3[0] = h4[0]; h_out3[1] = h4[1]; h_out3[2] = h4[2];
	if(su) cudaFreeAsync(su, s);
	if(sg) cudaFreeAsync(sg, s);
	cudaFree(d4);
	return rc;
}

int fvg_entropy_error(fvg_flow *f, const double *d_u, double *h_out)
{
	if(!f || !d_u || !h_out) { set_error("fvg_entropy_error: null argument"); return FVG_ERR_INVALID; }
	double *su = nullptr;
	const double *u = nullptr;
	int rc = to_device_order(f, d_u, 4, &su, &u, nullptr);
	if(rc == 0) rc = launch_entropy(f->mesh->d, f->gas, u, f->d_norm, nullptr);
	if(su) cudaFreeAsync(su, nullptr);
	if(rc != 0) return rc;
	f->launches += 2;
	double v = 0;
	FVG_CUDA(cudaMemcpy(&v, f->d_norm, sizeof(double), cudaMemcpyDeviceToHost));
	*h_out = std::sqrt(v);
	return 0;
}

// ------------------------------------------------------------------------------------ multi-GPU pieces

int fvg_partition_sfc(const fvg_umesh *m, int nranks, int *cell_rank)
{
	if(!m || !cell_rank || nranks < 1) { set_error("fvg_partition_sfc: bad argument"); return FVG_ERR_INVALID; }
	const int n = m->m.gnelem();
	std::vector<double> rc(2*(size_t)n);
	m->m.compute_cell_centres(rc.data());
	std::vector<int> order;
	hilbert_order(n, rc.data(), order);
	for(int k = 0; k < n; k++) cell_rank[order[k]] = (int)(((long long)k*nranks)/n);
	return 0;
}

namespace {
/// Recursive coordinate bisection of the cells idx[lo, hi) into ranks [r0, r0 + nparts): split along the longer side
/// of the bounding box of the cell centres, cell counts in proportion to the number of ranks on either side
void rcb_split(const double *rc, std::vector<int> &idx, int lo, int hi, int r0, int nparts, int *cell_rank)
{
	if(nparts == 1) { for(int k = lo; k < hi; k++) cell_rank[idx[k]] = r0; return; }
	double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
	for(int k = lo; k < hi; k++)
		for(int d = 0; d < 2; d++) {
			const double x = rc[2*(size_t)idx[k]+d];
			mn[d] = std::min(mn[d], x); mx[d] = std::max(mx[d], x);
		}
	const int ax = (mx[1] - mn[1] > mx[0] - mn[0]) ? 1 : 0;
	const int nleft = nparts/2;
	const int mid = lo + (int)(((long long)(hi - lo)*nleft)/nparts);
	// ties in the coordinate are broken by the cell index, so the result does not depend on the library's nth_element
	std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [rc, ax](const int a, const int b) {
		const double xa = rc[2*(size_t)a+ax], xb = rc[2*(size_t)b+ax];
		return xa < xb || (xa == xb && a < b);
	});
	rcb_split(rc, idx, lo, mid, r0, nleft, cell_rank);
	rcb_split(rc, idx, mid, hi, r0 + nleft, nparts - nleft, cell_rank);
}
}

int fvg_partition_rcb(const fvg_umesh *m, int nranks, int *cell_rank)
{
	if(!m || !cell_rank || nranks < 1) { set_error("fvg_partition_rcb: bad argument"); return FVG_ERR_INVALID; }
	const int n = m->m.gnelem();
	std::vector<double> rc(2*(size_t)n);
	m->m.compute_cell_centres(rc.data());
	std::vector<int> idx(n);
	for(int i = 0; i < n; i++) idx[i] = i;
	rcb_split(rc.data(), idx, 0, n, 0, nranks, cell_rank);
	return 0;
}

int fvg_halo_pack(const fvg_mesh *m, const double *d_src, int width, double *d_sendbuf, void *stream)
{
	if(!m || !d_src || (m->d.nsend > 0 && !d_sendbuf) || width < 1) { set_error("fvg_halo_pack: bad argument"); return FVG_ERR_INVALID; }
	return launch_halo_pack(m->d, d_src, width, d_sendbuf, static_cast<cudaStream_t>(stream));
}

int fvg_flow_use_buffers(fvg_flow *f, double *d_lg, double *d_gu)
{
	if(!f) { set_error("fvg_flow_use_buffers: null argument"); return FVG_ERR_INVALID; }
	if((f->plan.need_lg && !d_lg) || (f->plan.need_gu && !d_gu)) { set_error("fvg_flow_use_buffers: this flow needs the buffer that was passed as NULL"); return FVG_ERR_INVALID; }
	if(d_lg) f->d_lg = d_lg;
	if(d_gu) f->d_gu = d_gu;
	return 0;
}

int fvg_flow_buffers(fvg_flow *f, double **d_lg, double **d_gu)
{
	if(!f) { set_error("fvg_flow_buffers: null argument"); return FVG_ERR_INVALID; }
	if(d_lg) *d_lg = f->d_lg;
	if(d_gu) *d_gu = f->d_gu;
	return 0;
}

int fvg_gradient_pass(fvg_flow *f, const double *d_u, int stage, void *stream)
{
	if(!f || !d_u) { set_error("fvg_gradient_pass: null argument"); return FVG_ERR_INVALID; }
	if(!f->mesh->identity_perm) { set_error("fvg_gradient_pass: split passes need a device-ordered state (reorder none or a subdomain mesh)"); return FVG_ERR_UNSUPPORTED; }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const FlowPlan &P = f->plan;
	if(!P.order2) return 0;
	if(P.recon != FVG_RECON_WENO) return stage == 0 ? run_gradient_pass(f, d_u, s) : 0;
	// WENO: stage 0 = unlimited gradients (exchange them), stage 1 = the weighted average
	CellArgs a;
	a.m = f->mesh->d; a.gas = f->gas; a.u = d_u; a.ug = nullptr; a.gin = nullptr; a.bnd_policy = P.bnd_policy;
	a.prefetch_distance = f->prefetch_distance;
	int rc;
	if(stage == 0) { a.lg = nullptr; a.gu = f->d_gu; rc = launch_cell_kernel(P.gradient, 0, false, a, s); }
	else {
		WenoArgs w;
		w.m = f->mesh->d; w.lambda = f->gas.limiter_param; w.gu = f->d_gu; w.lg = f->d_lg;
		rc = launch_weno_kernel(w, s);
	}
	if(rc == 0) f->launches++;
	return rc;
}

int fvg_face_pass(fvg_flow *f, const double *d_u, double *d_res, int accumulate, int gettimesteps, double *d_dtm, void *stream)
{
	if(!f || !d_u || !d_res || (gettimesteps && !d_dtm)) { set_error("fvg_face_pass: null argument"); return FVG_ERR_INVALID; }
	if(!f->mesh->identity_perm) { set_error("fvg_face_pass: split passes need a device-ordered state"); return FVG_ERR_UNSUPPORTED; }
	return run_face_pass(f, d_u, EP_RESIDUAL, accumulate, gettimesteps, d_res, d_dtm, 0.0, nullptr, static_cast<cudaStream_t>(stream));
}

int fvg_euler_face_pass(fvg_flow *f, const double *d_u, double *d_unew, double cfl, double *d_resnorm2, void *stream)
{
	if(!f || !d_u || !d_unew) { set_error("fvg_euler_face_pass: null argument"); return FVG_ERR_INVALID; }
	if(!f->mesh->identity_perm) { set_error("fvg_euler_face_pass: split passes need a device-ordered state"); return FVG_ERR_UNSUPPORTED; }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	int rc;
	if((rc = run_face_pass(f, d_u, EP_STEP, 0, 1, nullptr, nullptr, cfl, d_unew, s)) != 0) return rc;
	if(f->part == 1 && f->mesh->d.tile_order) return 0;      // the norm is summed once all tiles have their partial sums
	if((rc = launch_final_norm(f->d_partial, f->mesh->d.ntile*(FACE_BLOCK/32), f->d_norm, s)) != 0) return rc;
	f->launches++;
	if(d_resnorm2) FVG_CUDA(cudaMemcpyAsync(d_resnorm2, f->d_norm, sizeof(double), cudaMemcpyDeviceToDevice, s));
	return 0;
}

int fvg_flow_ghost_source(fvg_flow *f, int which, fvg_halo *h, unsigned long long token)
{
	if(!f || which < 0 || which > 1) { set_error("fvg_flow_ghost_source: bad argument"); return FVG_ERR_INVALID; }
	GhostSrc &g = which == 0 ? f->gs_u : f->gs_g;
	if(!h || token == 0) { g = GhostSrc(); return 0; }
	if(which == 1 && (f->plan.visc != VISC_NONE || f->plan.recon == FVG_RECON_VANALBADA || f->plan.recon == FVG_RECON_WENO)) {
		set_error("fvg_flow_ghost_source: viscous, MUSCL and WENO passes read ghost rows from the arrays; use fvg_halo_exchange");
		return FVG_ERR_UNSUPPORTED;
	}
	return fvg_halo_ghost_source(h, token, &g);
}

int fvg_flow_select_tiles(fvg_flow *f, int part)
{
	if(!f || part < 0 || part > 3) { set_error("fvg_flow_select_tiles: bad argument"); return FVG_ERR_INVALID; }
	f->part = part;
	return 0;
}

// ------------------------------------------------------------------------------------ pseudo-time

/// One fused step on device-ordered buffers: reads uin, writes uout, norm^2 -> f->d_norm
static int step_device_order(fvg_flow *f, const double *uin, double *uout, double cfl, cudaStream_t s)
{
	int rc;
	if((rc = mark(f, s)) != 0) return rc;
	if((rc = run_gradient_pass(f, uin, s)) != 0) return rc;
	if((rc = mark(f, s)) != 0) return rc;
	if((rc = run_face_pass(f, uin, EP_STEP, 0, 1, nullptr, nullptr, cfl, uout, s)) != 0) return rc;
	if((rc = mark(f, s)) != 0) return rc;
	if((rc = launch_final_norm(f->d_partial, f->mesh->d.ntile*(FACE_BLOCK/32), f->d_norm, s)) != 0) return rc;
	f->launches++;
	return 0;
}

int fvg_euler_step(fvg_flow *f, double *d_u, double cfl, double *d_resnorm2, void *stream)
{
	if(!f || !d_u) { set_error("fvg_euler_step: null argument"); return FVG_ERR_INVALID; }
	if(f->mesh->nranks > 1) { set_error("fvg_euler_step: use fvg_dist_euler_step on a subdomain mesh"); return FVG_ERR_UNSUPPORTED; }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const DMesh &D = f->mesh->d;
	const size_t n = D.ncell;
	int rc;
	if(has_periodic_ghosts(f)) {
		// periodic mesh: the exchange engine steps device-ordered ping-pong buffers
		fvg_dist *sd = nullptr;
		if((rc = self_exchange(f, &sd)) != 0) return rc;
		if((rc = ensure(f, &f->d_u2, 4*n)) != 0) return rc;
		if((rc = ensure(f, &f->d_uperm, 4*n)) != 0) return rc;
		if(f->mesh->identity_perm) FVG_CUDA(cudaMemcpyAsync(f->d_uperm, d_u, 4*n*sizeof(double), cudaMemcpyDeviceToDevice, s));
		else { if((rc = launch_permute_rows(d_u, f->d_uperm, D.new2old, (int)n, 4, true, false, s)) != 0) return rc; f->launches++; }
		if((rc = fvg_dist_invalidate_state(sd)) != 0) return rc;
		if((rc = fvg_dist_euler_step(sd, f->d_uperm, f->d_u2, cfl, f->d_norm, stream)) != 0) return rc;
		if(f->mesh->identity_perm) FVG_CUDA(cudaMemcpyAsync(d_u, f->d_u2, 4*n*sizeof(double), cudaMemcpyDeviceToDevice, s));
		else { if((rc = launch_permute_rows(f->d_u2, d_u, D.new2old, (int)n, 4, false, false, s)) != 0) return rc; f->launches++; }
		if(d_resnorm2) FVG_CUDA(cudaMemcpyAsync(d_resnorm2, f->d_norm, sizeof(double), cudaMemcpyDeviceToDevice, s));
		return 0;
	}
	if((rc = ensure(f, &f->d_u2, 4*n)) != 0) return rc;
	if(f->plan.order2) {
		// the gradient pass leaves a device-ordered copy of the state (gathered through the permutation when the mesh is
		// renumbered); the face pass reads only that copy, so it can store the new state straight into the caller's rows
		const int *perm = f->mesh->identity_perm ? nullptr : D.new2old;
		if((rc = mark(f, s)) != 0) return rc;
		if((rc = run_gradient_pass(f, d_u, s, 0, -1, perm, f->d_u2)) != 0) return rc;
		if((rc = mark(f, s)) != 0) return rc;
		if((rc = run_face_pass(f, f->d_u2, EP_STEP, 0, 1, nullptr, nullptr, cfl, d_u, s, 0, -1, perm)) != 0) return rc;
		if((rc = mark(f, s)) != 0) return rc;
		if((rc = launch_final_norm(f->d_partial, f->mesh->d.ntile*(FACE_BLOCK/32), f->d_norm, s)) != 0) return rc;
		f->launches++;
	} else if(f->mesh->identity_perm) {
		if((rc = step_device_order(f, d_u, f->d_u2, cfl, s)) != 0) return rc;
		FVG_CUDA(cudaMemcpyAsync(d_u, f->d_u2, 4*n*sizeof(double), cudaMemcpyDeviceToDevice, s));
	} else {
		if((rc = ensure(f, &f->d_uperm, 4*n)) != 0) return rc;
		if((rc = launch_permute_rows(d_u, f->d_uperm, D.new2old, (int)n, 4, true, false, s)) != 0) return rc;
		if((rc = step_device_order(f, f->d_uperm, f->d_u2, cfl, s)) != 0) return rc;
		if((rc = launch_permute_rows(f->d_u2, d_u, D.new2old, (int)n, 4, false, false, s)) != 0) return rc;
		f->launches += 2;
	}
	if(d_resnorm2) FVG_CUDA(cudaMemcpyAsync(d_resnorm2, f->d_norm, sizeof(double), cudaMemcpyDeviceToDevice, s));
	return 0;
}

int fvg_forward_euler_solve(fvg_flow *f, double *d_u, double cfl, double tol, int maxiter,
                            int check_every, int *h_steps, double *h_hist)
{
	if(!f || !d_u || !h_steps) { set_error("fvg_forward_euler_solve: null argument"); return FVG_ERR_INVALID; }
	if(f->mesh->nranks > 1) { set_error("fvg_forward_euler_solve: use fvg_dist_forward_euler_solve on a subdomain mesh"); return FVG_ERR_UNSUPPORTED; }
	if(has_periodic_ghosts(f)) {
		fvg_dist *sd = nullptr;
		const int rcd = self_exchange(f, &sd);
		return rcd != 0 ? rcd : fvg_dist_forward_euler_solve(sd, d_u, cfl, tol, maxiter, check_every, h_steps, h_hist);
	}
	if(check_every < 1) check_every = 1;
	*h_steps = 0;
	if(maxiter <= 0) return 0;
	FVG_CUDA(cudaSetDevice(f->mesh->device));
	cudaStream_t s = nullptr;
	const DMesh &D = f->mesh->d;
	const size_t n = D.ncell;
	int rc;
	if((rc = ensure(f, &f->d_u2, 4*n)) != 0) return rc;
	if((rc = ensure(f, &f->d_uperm, 4*n)) != 0) return rc;
	// state lives in device order in two ping-pong buffers for the whole loop
	double *cur = f->d_uperm, *nxt = f->d_u2;
	if(f->mesh->identity_perm) FVG_CUDA(cudaMemcpyAsync(cur, d_u, 4*n*sizeof(double), cudaMemcpyDeviceToDevice, s));
	else { if((rc = launch_permute_rows(d_u, cur, D.new2old, (int)n, 4, true, false, s)) != 0) return rc; f->launches++; }

	double *d_hist = nullptr;
	FVG_CUDA(cudaMalloc((void**)&d_hist, sizeof(double)*(size_t)maxiter));
	std::vector<double> hist((size_t)maxiter);
	int step = 0, status = FVG_OK;
	double initres = 1.0;
	// batches of check_every steps are enqueued back to back; their norms are read afterwards
	// (check_every = 1 is the reference's loop: one host read of the norm per step)
	while(rc == 0) {
		const int batch = std::min(check_every, maxiter - step);
		for(int k = 0; k < batch && rc == 0; k++) {
			rc = step_device_order(f, cur, nxt, cfl, s);
			if(rc == 0) {
				const cudaError_t e = cudaMemcpyAsync(d_hist + step + k, f->d_norm, sizeof(double), cudaMemcpyDeviceToDevice, s);
				if(e != cudaSuccess) rc = cuda_fail(e, "norm copy", __FILE__, __LINE__);
			}
			std::swap(cur, nxt);
		}
		if(rc != 0) break;
		cudaError_t e = cudaMemcpyAsync(hist.data() + step, d_hist + step, sizeof(double)*batch, cudaMemcpyDeviceToHost, s);
		if(e == cudaSuccess) e = cudaStreamSynchronize(s);
		if(e != cudaSuccess) { rc = cuda_fail(e, "norm read-back", __FILE__, __LINE__); break; }
		bool stop = false;
		for(int k = 0; k < batch; k++) {
			const double resi = std::sqrt(hist[step+k]);
			hist[step+k] = resi;
			if(step + k == 0) initres = resi;
			if(stop) continue;
			if(!std::isfinite(resi)) { status = FVG_ERR_NUMERICAL; stop = true; }
			else if(!(resi/initres > tol)) stop = true;
		}
		step += batch;
		if(step >= maxiter) { if(status == FVG_OK) status = FVG_ERR_TOLERANCE; stop = true; }
		if(stop) break;
	}
	if(rc == 0) {
		*h_steps = step;
		if(h_hist) std::memcpy(h_hist, hist.data(), sizeof(double)*step);
		if(f->mesh->identity_perm) {
			const cudaError_t e = cudaMemcpyAsync(d_u, cur, 4*n*sizeof(double), cudaMemcpyDeviceToDevice, s);
			if(e != cudaSuccess) rc = cuda_fail(e, "state copy", __FILE__, __LINE__);
		} else { rc = launch_permute_rows(cur, d_u, D.new2old, (int)n, 4, false, false, s); f->launches++; }
		const cudaError_t e2 = cudaStreamSynchronize(s);
		if(rc == 0 && e2 != cudaSuccess) rc = cuda_fail(e2, "stream sync", __FILE__, __LINE__);
	}
	cudaFree(d_hist);
	if(rc != 0) return rc;
	if(status == FVG_ERR_NUMERICAL) set_error("forward Euler: residual norm is not finite");
	if(status == FVG_ERR_TOLERANCE) set_error("forward Euler: exceeded max iterations");
	return status;
}

// ------------------------------------------------------------------------------------ pointwise hooks

namespace {
struct DevBuf {
	double *p = nullptr;
	~DevBuf() { if(p) cudaFree(p); }
	int put(const double *h, size_t n) {
		FVG_CUDA(cudaMalloc((void**)&p, std::max<size_t>(n,1)*sizeof(double)));
		if(h && n) FVG_CUDA(cudaMemcpy(p, h, n*sizeof(double), cudaMemcpyHostToDevice));
		return 0;
	}
	int get(double *h, size_t n) { if(n) FVG_CUDA(cudaMemcpy(h, p, n*sizeof(double), cudaMemcpyDeviceToHost)); return 0; }
};
}

int fvg_flux_pointwise(int flux_id, const fvg_physics *phys, int n, const double *h_ul,
                       const double *h_ur, const double *h_n, double *h_out)
{
	if(!phys || n < 0 || (n > 0 && (!h_ul || !h_ur || !h_n || !h_out))) { set_error("fvg_flux_pointwise: bad argument"); return FVG_ERR_INVALID; }
	const GasParams G = make_gas(*phys, 0.0);
	DevBuf a, b, c, o;
	int rc;
	if((rc = a.put(h_ul, 4*(size_t)n)) || (rc = b.put(h_ur, 4*(size_t)n)) || (rc = c.put(h_n, 2*(size_t)n)) || (rc = o.put(nullptr, 4*(size_t)n))) return rc;
	if((rc = launch_pointwise_flux(flux_id, G, n, a.p, b.p, c.p, o.p, nullptr)) != 0) return rc;
	FVG_CUDA(cudaDeviceSynchronize());
	return o.get(h_out, 4*(size_t)n);
}

int fvg_bc_pointwise(const fvg_bc *bc, const fvg_physics *phys, int n, const double *h_ins,
                     const double *h_n, double *h_out)
{
	if(!bc || !phys || n < 0 || (n > 0 && (!h_ins || !h_n || !h_out))) { set_error("fvg_bc_pointwise: bad argument"); return FVG_ERR_INVALID; }
	if(bc->type < 0 || bc->type > 7 || bc->type == PERIODIC_BC) { set_error("fvg_bc_pointwise: boundary condition type not available"); return FVG_ERR_UNSUPPORTED; }
	GasParams G = make_gas(*phys, 0.0);
	G.nbc = 1; G.bc[0].tag = bc->tag; G.bc[0].type = bc->type; G.bc[0].v0 = bc->vals[0]; G.bc[0].v1 = bc->vals[1];
	DevBuf a, c, o;
	int rc;
	if((rc = a.put(h_ins, 4*(size_t)n)) || (rc = c.put(h_n, 2*(size_t)n)) || (rc = o.put(nullptr, 4*(size_t)n))) return rc;
	if((rc = launch_pointwise_bc(G, n, a.p, c.p, o.p, nullptr)) != 0) return rc;
	FVG_CUDA(cudaDeviceSynchronize());
	return o.get(h_out, 4*(size_t)n);
}

int fvg_viscous_flux_pointwise(const fvg_physics *phys, int order2, int n, const double *h_n,
                               const double *h_rcl, const double *h_rcr, const double *h_ucl,
                               const double *h_ucr, const double *h_gl, const double *h_gr,
                               const double *h_ul, const double *h_ur, double *h_out)
{
	if(!phys || n < 0 || (n > 0 && (!h_n || !h_rcl || !h_rcr || !h_ucl || !h_ucr || !h_ul || !h_ur || !h_out)) ||
	   (order2 && n > 0 && (!h_gl || !h_gr))) { set_error("fvg_viscous_flux_pointwise: bad argument"); return FVG_ERR_INVALID; }
	const GasParams G = make_gas(*phys, 0.0);
	DevBuf nn, rl, rr, cl, cr, gl, gr, ul, ur, o;
	int rc;
	const size_t N = n;
	if((rc = nn.put(h_n, 2*N)) || (rc = rl.put(h_rcl, 2*N)) || (rc = rr.put(h_rcr, 2*N)) || (rc = cl.put(h_ucl, 4*N)) ||
	   (rc = cr.put(h_ucr, 4*N)) || (rc = gl.put(order2 ? h_gl : nullptr, 8*N)) || (rc = gr.put(order2 ? h_gr : nullptr, 8*N)) ||
	   (rc = ul.put(h_ul, 4*N)) || (rc = ur.put(h_ur, 4*N)) || (rc = o.put(nullptr, 4*N))) return rc;
	if((rc = launch_pointwise_visc(G, order2 != 0, phys->const_visc != 0, n, nn.p, rl.p, rr.p, cl.p, cr.p, gl.p, gr.p, ul.p, ur.p, o.p, nullptr)) != 0) return rc;
	FVG_CUDA(cudaDeviceSynchronize());
	return o.get(h_out, 4*N);
}

int fvg_freestream(const fvg_physics *phys, double *h_uinf4)
{
	if(!phys || !h_uinf4) { set_error("fvg_freestream: null argument"); return FVG_ERR_INVALID; }
	const GasParams G = make_gas(*phys, 0.0);
	for(int k = 0; k < 4; k++) h_uinf4[k] = G.uinf[k];
	return 0;
}

} // extern "C"
