/* The fused face kernel: reconstruct -> (boundary ghost) -> inviscid flux [+ viscous flux] ->
 * spectral radius -> coloured accumulation into the tile's cells -> residual/time-step or fused
 * forward-Euler epilogue. Restates P6-P11 of FlowFV::compute_residual (reference:
 * src/spatial/flow_spatial.cpp:722-812, compute_fluxes :489-563, compute_max_timestep :567-634)
 * and, with the EP_STEP epilogue, the update + norm of SteadyForwardEulerSolver::solve
 * (src/ode/aodesolver.cpp:204-223).
 *
 * One CTA per tile of consecutive cells; see the comment on face_kernel below for the three phases.
 * Faces cut by a tile boundary appear in both tiles' streams and are evaluated identically in both
 * (same left/right roles, same expression), so the scheme stays exactly conservative. Each cell's
 * residual is the sum of its faces' fluxes in local-face order: no atomics, bitwise reproducible.
 */
#pragma once
#include "engine.hpp"
#include "async_copy.cuh"

namespace fvg {

__device__ __forceinline__ void ld4(const double *p, double v[4]) {
	asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
	             : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
/// coherent variant for arrays the same kernel also writes
__device__ __forceinline__ void ld4c(const double *p, double v[4]) {
	asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
	             : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st4(double *p, const double v[4]) {
	asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
	             :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
/// 32-byte row from shared memory (two 16-byte loads)
__device__ __forceinline__ void lds4(const double *p, double v[4]) {
	const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p+2);
	v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

/// 32-byte row kept as two 16-byte planes `stride` elements apart
__device__ __forceinline__ void ldp4(const double2 *p, int stride, double v[4]) {
	const double2 a = p[0], b = p[stride];
	v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

/// Where the TMA swizzle modes put the 16-byte chunk c of row k (buffer at a multiple of 1024 bytes), in doubles from
/// the start of the buffer. 32-byte rows, SWIZZLE_32B: address bit 4 ^= bit 7; 64-byte rows, SWIZZLE_64B: bits 5:4 ^=
/// bits 8:7. With these, the 8 threads of a quarter warp reading chunk c of 8 consecutive rows hit 8 distinct bank groups.
__device__ __forceinline__ int rows32_chunk(int k, int c) { return 4*k + 2*(c ^ ((k >> 2) & 1)); }
__device__ __forceinline__ int rows64_chunk(int k, int c) { return 8*k + 2*(c ^ ((k >> 1) & 3)); }

/// Device cell index of a tile-local index
__device__ __forceinline__ int tile_global(const DMesh &M, int t, int c0, int nc, unsigned loc) {
	return loc < (unsigned)nc ? c0 + (int)loc : M.thalo[M.thoff[t] + (int)loc - nc];
}

/// ... with the tile's halo offset at hand
__device__ __forceinline__ int tile_global_h(const DMesh &M, int h0, int c0, int nc, unsigned loc) {
	return loc < (unsigned)nc ? c0 + (int)loc : M.thalo[h0 + (int)loc - nc];
}

/// u_face = u_cell + g . (gr - rc) for the four primitive variables (reconstruction_utils.hpp:17-32);
/// g = 8 gradients in GradBlock order, in shared memory
__device__ __forceinline__ void extrapolate4(const double pc[4], const double *g, double dx, double dy, double pf[4]) {
	double ga[4], gb[4];
	lds4(g, ga); lds4(g+4, gb);
	pf[0] = pc[0] + ga[0]*dx + ga[1]*dy;
	pf[1] = pc[1] + ga[2]*dx + ga[3]*dy;
	pf[2] = pc[2] + gb[0]*dx + gb[1]*dy;
	pf[3] = pc[3] + gb[2]*dx + gb[3]*dy;
}

/// Van Albada limited MUSCL increment (musclreconstruction.cpp:35-59), eps = 1e-8, k = 1/3
__device__ __forceinline__ double muscl_term(double delta, double dlr) {
	const double eps = 1e-8, k = 1.0/3.0;
	double phi = (2.0*delta*dlr + eps)/(delta*delta + dlr*dlr + eps);
	if(phi < 0.0) phi = 0.0;
	return phi*0.25*((1.0 - k*phi)*delta + (1.0 + k*phi)*dlr);
}

/// Shared-memory carve-up of the face kernel for the given capacities (host and device agree through this).
/// Per stream entry: the two face-state slots (later the flux and the spectral radii) and the midpoint;
/// per halo cell: its state, reconstruction gradient and centre, gathered asynchronously while phase A runs.
/// Everything indexed by stream entry is an array of 16-BYTE elements (a 32-byte face state is kept as two such
/// planes): entry e then sits in bank group e mod 8, consecutive entries (phase B) are conflict-free, and the
/// per-cell scatters / gathers of phases A and C are conflict-free by the placement of the entries (device_mesh.cu).
struct FaceSmem {
	int fsL, fsR, sgr, sn, slen, sLR, sord, hu, hg, hrc, su, sg, src, scl, sar, cbuf, bar, red, ring, gptr, total;   // byte offsets
	__host__ __device__ FaceSmem(int TC, int EMAX, int HMAX, bool mids, bool linear) {
		int o = 0;
		su = o; o += TC*32;                    // own cells: state, reconstruction gradient (both hardware-swizzled: keep them
		sg = o; o += linear ? TC*64 : 0;       // first, at multiples of 1024 bytes), centre
		src = o; o += linear ? TC*16 : 0;
		fsL = o; o += EMAX*32;                 // planes: [0, EMAX) first half of the state / flux, [EMAX, 2 EMAX) second half
		fsR = o; o += EMAX*32;
		sgr = o; o += mids ? EMAX*16 : 0;
		sn = o; o += EMAX*16;
		slen = o; o += EMAX*8;
		sLR = o; o += EMAX*4;
		sord = o; o += EMAX*2;
		hu = o; o += HMAX*32;
		hg = o; o += mids ? HMAX*64 : 0;
		hrc = o; o += mids ? HMAX*16 : 0;
		cbuf = TC*16 + (TC + 2)*8;             // stencil + area, two buffers of cbuf bytes
		scl = o; sar = o + TC*16; o += 2*cbuf;
		bar = o; o += 16;                      // two mbarriers
		red = o; o += 8*(FACE_BLOCK/32);       // warp partials of the norm
		ring = o; o += 3*48;                   // descriptor records of the current and the next two tiles
		gptr = o; o += 32;                     // fused multi-GPU evaluation: evaluation number, window areas of the ghost rows
		total = o;
	}
};

/// primitive face state of one side whose cell is NOT a tile cell (halo cell h): from the staged halo rows
template <int RECON>
__device__ __forceinline__ void halo_side_state(const FaceArgs &A, const double *hu, const double *hg, const double2 *hrc,
                                                int h, double2 gr, double pf[4])
{
	double uc[4];
	lds4(hu + 4*h, uc);
	if(RECON == FR_FIRST) { for(int k = 0; k < 4; k++) pf[k] = uc[k]; return; }
	double pc[4];
	cons2prim(A.gas, uc, pc);
	if(RECON == FR_MUSCL) { for(int k = 0; k < 4; k++) pf[k] = pc[k]; return; }
	const double2 rc = hrc[h];
	double ga[4], gb[4];
	lds4(hg + 8*h, ga); lds4(hg + 8*h + 4, gb);
	extrapolate_prim(pc, ga, gb, gr.x, gr.y, rc.x, rc.y, pf);
}

/// Called by a whole CTA before it touches ghost rows that arrive through a halo window: thread r waits until the
/// neighbour rank r has published exchange `seq` (bounded spin: on a timeout the window's error word is set and the
/// kernel goes on, the caller reads fvg_halo_status). The barrier makes the acquired rows visible to all threads.
__device__ __forceinline__ void ghost_wait(const GhostSrc &g, unsigned long long seq)
{
	if((int)threadIdx.x < g.nranks) {
		const int r = threadIdx.x;
		if(g.recv_off[r+1] > g.recv_off[r]) {
			long long spins = 0;
			while(ld_acquire_sys_u64(g.flags + r) < seq) {
				__nanosleep(64);
				if(++spins > 30000000ll) { atomicExch(const_cast<unsigned long long*>(g.flags) + g.nranks, seq); break; }
			}
		}
	}
	__syncthreads();
}

/// Row `g` of a device-ordered [.][width] array whose ghost rows (g >= ncell) may live in a halo window instead
__device__ __forceinline__ const double *ghost_aware_row(const double *arr, const double *win_rows, int ncell, int g, int width) {
	return (g >= ncell && win_rows) ? win_rows + (size_t)width*(size_t)(g - ncell) : arr + (size_t)width*(size_t)g;
}

/// Descriptor of one tile (all uniform across the CTA), unpacked from the tile's 48-byte record (DMesh::tdesc): the
/// records of the next two tiles of a CTA travel into a small shared-memory ring by cp.async while the current tile
/// computes, so no descriptor is held in registers across a tile and none is waited for
struct TileDesc { int t, c0, nc, h0, nh, e0, ne, nreal, ghost, cut0, bnd0; };
__device__ __forceinline__ TileDesc unpack_tile_desc(const int4 *rec) {
	const int4 a = rec[0], b = rec[1], c = rec[2];
	TileDesc D;
	D.t = a.x; D.c0 = a.y; D.nc = a.z; D.h0 = a.w;
	D.nh = b.x; D.e0 = b.y; D.ne = b.z;
	D.nreal = D.ne - (b.w & 0xFFFF); D.ghost = b.w >> 16;
	D.cut0 = c.x; D.bnd0 = c.y;
	return D;
}
/// threads 0..2 of a CTA: asynchronous copy of the record at position `pos` of the launch's tile sequence into `slot`
__device__ __forceinline__ void fetch_tile_desc(int4 *slot, const int4 *tdesc, int pos) {
	if(threadIdx.x < 3) cp_async16(slot + threadIdx.x, tdesc + 3*(size_t)pos + threadIdx.x);
}

/** Persistent CTAs (two per SM), each walking over tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
 *  Everything a tile reads arrives in shared memory asynchronously - the own cells' rows (state, reconstruction
 *  gradient, centre, stencil, area) and the stream-entry metadata (midpoints, normals, lengths, local indices)
 *  by 1-D TMA bulk copies (contiguous per tile), the halo cells' rows by 16-byte cp.async gathers - and the
 *  copies for the NEXT tile are issued as soon as the current tile's phase that reads a buffer is over, so they
 *  land while the current tile is still computing:
 *    group A buffers (state, gradient, centre, midpoints)         read in phase A     -> refilled after phase A
 *    group B buffers (normals, lengths, local indices, halo rows)  read in phase B     -> refilled after phase B
 *    group C buffers (stencil, area)                               read in A and C     -> double-buffered
 *  Per tile, three phases separated by barriers:
 *  A  one thread per own cell: convert to primitive ONCE, extrapolate to each of its <= 4 faces and deposit
 *     the face state in the left or right slot of that face's stream entry.
 *  B  one thread per stream entry (consecutive entries => conflict-free shared-memory rows): both face
 *     states (a side belonging to a halo cell is reconstructed from the staged halo row), boundary ghost,
 *     numerical flux, spectral radii; the flux overwrites the entry's slots. Entries are ordered by kind, so
 *     the halo rows are first needed in a later round and their copies are only waited for there.
 *  C  one thread per own cell: sum the fluxes of its faces in local-face order (deterministic, no
 *     atomics, no scatter), then the residual / time-step or the fused forward-Euler epilogue. */
template <int FLUX, int RECON, int VISC>
__global__ void __launch_bounds__(FACE_BLOCK, FVG_FACE_MINB)
face_kernel(const __grid_constant__ FaceArgs A)
{
	extern __shared__ __align__(1024) unsigned char smraw[];
	const DMesh &M = A.m;
	constexpr bool MIDS = RECON != FR_FIRST;
	constexpr bool LINEAR = RECON == FR_LINEAR;
	const FaceSmem S(M.TC, M.EMAX, M.HMAX, MIDS, LINEAR);
	double2 *const fsL = reinterpret_cast<double2*>(smraw + S.fsL);
	double2 *const fsR = reinterpret_cast<double2*>(smraw + S.fsR);
	const int EP = M.EMAX;                                                  // plane stride of fsL / fsR
	const int BOX = tile_box_rows(M.TC);
	double2 *const sgr = reinterpret_cast<double2*>(smraw + S.sgr);
	double2 *const sn = reinterpret_cast<double2*>(smraw + S.sn);
	double *const slen = reinterpret_cast<double*>(smraw + S.slen);
	unsigned *const sLR = reinterpret_cast<unsigned*>(smraw + S.sLR);
	const unsigned short *const sord = reinterpret_cast<const unsigned short*>(smraw + S.sord);
	double *const hu = reinterpret_cast<double*>(smraw + S.hu);
	double *const hg = reinterpret_cast<double*>(smraw + S.hg);
	double2 *const hrc = reinterpret_cast<double2*>(smraw + S.hrc);
	double *const su = reinterpret_cast<double*>(smraw + S.su);
	double *const sg = reinterpret_cast<double*>(smraw + S.sg);
	double2 *const src = reinterpret_cast<double2*>(smraw + S.src);
	uint64_t *const bar = reinterpret_cast<uint64_t*>(smraw + S.bar);       // [0]: groups A + C, [1]: group B

	const int tid = threadIdx.x;
	const double *const gsrc = RECON == FR_MUSCL ? A.gu : A.lg;     // gradients used by the reconstruction
	// Viscous flux with linear reconstruction and no limiter (the laminar cases of the reference: the gradients of the
	// viscous flux ARE the reconstruction gradients): phase B takes the cell states, gradients, centres and areas of both
	// sides from the staged rows instead of gathering 230 bytes per entry from global memory; the own-cell rows (group
	// A buffers) are then refilled for the next tile after phase B instead of after phase A.
	const bool keepA = VISC != VISC_NONE && RECON == FR_LINEAR && A.gu == A.lg;
	// where the ghost rows are: in the arrays, in the window of a posted exchange (legacy split passes: A.gs_*), or in
	// this evaluation's areas of the fused multi-GPU evaluation (set below, once the evaluation number is known)
	// (kept in shared memory, not in registers: they are read once per tile)
	int4 *const ring = reinterpret_cast<int4*>(smraw + S.ring);
	struct GhostPtrs { unsigned long long dk; const double *u, *g, *v; };
	GhostPtrs *const gp = reinterpret_cast<GhostPtrs*>(smraw + S.gptr);

	// issue helpers (thread 0 only for the bulk copies). 8-byte rows (area) are copied from the even cell below c0.
	auto issue_AC = [&](const TileDesc &D, int buf) {
		const unsigned aoff = (unsigned)(D.c0 & 1);
		const unsigned abytes = (unsigned)((D.nc + aoff + 1) & ~1)*8u;
		// state and gradient rows come through tensor maps in whole boxes (rows past the tile belong to the next tile or
		// are zero fill past the end of the array; either way they are never read)
		const int nbox = (D.nc + BOX - 1)/BOX;
		mbar_expect_tx(bar, (unsigned)(nbox*BOX)*(32u + (LINEAR ? 64u : 0u)) + (unsigned)D.nc*(16u + (LINEAR ? 16u : 0u))
		                    + (MIDS ? (unsigned)D.ne*16u : 0u) + abytes);
		for(int b = 0; b < nbox; b++) {
			tensor_rows_g2s(su + 4*b*BOX, &A.tm_u, D.c0 + b*BOX, bar);
			if(LINEAR) tensor_rows_g2s(sg + 8*b*BOX, &A.tm_g, D.c0 + b*BOX, bar);
		}
		if(LINEAR) bulk_g2s(src, M.rc + D.c0, (unsigned)D.nc*16u, bar);
		if(MIDS) bulk_g2s(sgr, M.fgr + D.e0, (unsigned)D.ne*16u, bar);
		bulk_g2s(smraw + S.scl + buf*S.cbuf, M.cloc + D.c0, (unsigned)D.nc*16u, bar);
		bulk_g2s(smraw + S.sar + buf*S.cbuf, M.area + (D.c0 - (int)aoff), abytes, bar);
	};
	auto issue_B = [&](const TileDesc &D) {
		mbar_expect_tx(bar + 1, (unsigned)D.ne*(16u + 8u + 4u + 2u));
		bulk_g2s(sLR, M.fLR + D.e0, (unsigned)D.ne*4u, bar + 1);
		bulk_g2s(smraw + S.sord, M.ford + D.e0, (unsigned)D.ne*2u, bar + 1);
		bulk_g2s(sn, M.fn + D.e0, (unsigned)D.ne*16u, bar + 1);
		bulk_g2s(slen, M.flen + D.e0, (unsigned)D.ne*8u, bar + 1);
	};
	// halo rows of a tile: thread h gathers halo cell h (device index g), FACE_BLOCK >= HMAX
	auto issue_halo = [&](int nh, int g) {
		if(tid < nh) {
			// a ghost cell's rows may live in a halo window instead of the arrays (in-kernel receive)
			const bool gh = g >= M.ncell;
			const double *const urow = (gh && gp->u) ? gp->u + 4*(size_t)(g - M.ncell) : A.u + 4*(size_t)g;
			cp_async16(hu + 4*tid, urow);
			cp_async16(hu + 4*tid + 2, urow + 2);
			if(MIDS) {
				const double *const grow = (gh && gp->g) ? gp->g + 8*(size_t)(g - M.ncell) : gsrc + 8*(size_t)g;
				#pragma unroll
				for(int q = 0; q < 4; q++) cp_async16(hg + 8*tid + 2*q, grow + 2*q);
				cp_async16(hrc + tid, M.rc + g);
			}
		}
		cp_async_commit();
	};

	// this launch covers positions [tile0, tile1) of the tile sequence A.tdesc (natural order, or interior tiles first)
	const int tend = A.tile1, G = (int)gridDim.x;
	int ti = A.tile0 + (int)blockIdx.x;
	if(tid == 0 && (smem_u32(smraw) & 1023u) != 0) __trap();      // the swizzle formulas assume this alignment
	// descriptor records run two tiles ahead of the computation (shared-memory ring) and the next tile's halo index one
	// tile ahead (a register), so that neither global load is waited for where it is consumed
	fetch_tile_desc(ring, A.tdesc, ti);
	if(ti + G < tend) fetch_tile_desc(ring + 3, A.tdesc, ti + G);
	cp_async_commit();
	if(tid == 0) {
		mbar_init(bar, 1); mbar_init(bar + 1, 1);
		gp->dk = 0; gp->u = A.gs_u.rows; gp->g = A.gs_g.rows; gp->v = A.gs_v.rows;
	}
	pdl_launch_dependents();
	// (programmatic dependent launch) everything above reads only the mesh; the state and gradient rows below are the
	// previous kernels' output
	pdl_wait();
	cp_async_wait_all();
	__syncthreads();
	TileDesc D = unpack_tile_desc(ring);
	int t = D.t;
	if(tid == 0) { issue_AC(D, 0); issue_B(D); }
	if(A.dist.d) {
		const DistDev *const dd = A.dist.d;
		const unsigned long long dk = A.dist.ctl->k;
		if(tid == 0) {
			const int par = (int)(dk & 1ull);
			gp->dk = dk;
			if(A.dist.wait & (1u << X_U)) gp->u = A.dist.ghost[X_U][par];
			if(MIDS && (A.dist.wait & (1u << (RECON == FR_MUSCL ? X_GU : X_LG)))) gp->g = A.dist.ghost[RECON == FR_MUSCL ? X_GU : X_LG][par];
			if(VISC != VISC_NONE && (A.dist.wait & (1u << A.dist.visc_type))) gp->v = A.dist.ghost[A.dist.visc_type][par];
		}
		if(A.dist.first) dist_push_state_prologue(dd, dk, A.u, A.dist.force_push);
		__syncthreads();
	}
	// in-kernel receive: a CTA waits for the neighbours' rows once, right before it gathers the halo of its first
	// tile that sees a ghost cell (interior tiles come first, so for most CTAs the rows have long arrived by then)
	const bool recv_here = A.dist.d ? A.dist.wait != 0 : (A.gs_u.rows != nullptr || A.gs_g.rows != nullptr);
	auto wait_for_ghost_rows = [&]() {
		if(A.dist.d) dist_wait(A.dist.d, A.dist.wait, gp->dk);
		else {
			const GhostSrc &gsw = A.gs_g.rows ? A.gs_g : A.gs_u;
			const unsigned long long wseq = A.gs_g.rows ? (A.gs_g.seq > A.gs_u.seq ? A.gs_g.seq : A.gs_u.seq) : A.gs_u.seq;
			ghost_wait(gsw, wseq);
		}
	};
	bool waited = !recv_here;
	if(!waited && D.ghost) { wait_for_ghost_rows(); waited = true; }
	issue_halo(D.nh, tid < D.nh ? M.thalo[D.h0 + tid] : 0);

	for(int it = 0; ti < tend; it++) {
		const unsigned par = (unsigned)(it & 1);
		const uint4 *const scl = reinterpret_cast<const uint4*>(smraw + S.scl + (int)par*S.cbuf);
		const double *const sar = reinterpret_cast<const double*>(smraw + S.sar + (int)par*S.cbuf);
		const int aoff = D.c0 & 1;
		// the record of the tile after next and this thread's halo index for the next tile are in flight during phase A
		const int tin = ti + G;
		const bool have_next = tin < tend;
		const int4 *const rnext = ring + 3*((it + 1) % 3);
		int gnext = 0;
		if(have_next) { const int nh_n = rnext[1].x, h0_n = rnext[0].w; if(tid < nh_n) gnext = M.thalo[h0_n + tid]; }
		if(tin + G < tend) fetch_tile_desc(ring + 3*((it + 2) % 3), A.tdesc, tin + G);      // joins this tile's halo group

		// ---- phase A: face states of the own cells
		mbar_wait(bar, par);
		double uc0[4] = {1,0,0,1};           // state of cell `tid`, kept for the fused step epilogue
		for(int k = tid; k < D.nc; k += FACE_BLOCK) {
			const uint4 cl = scl[k];
			double uc[4], ga[4] = {0,0,0,0}, gb[4] = {0,0,0,0};
			double2 rc = make_double2(0,0);
			{
				const double2 a = *reinterpret_cast<const double2*>(su + rows32_chunk(k, 0)), b = *reinterpret_cast<const double2*>(su + rows32_chunk(k, 1));
				uc[0] = a.x; uc[1] = a.y; uc[2] = b.x; uc[3] = b.y;
			}
			if(k == tid) { for(int q = 0; q < 4; q++) uc0[q] = uc[q]; }
			if(LINEAR) {
				const double2 g0 = *reinterpret_cast<const double2*>(sg + rows64_chunk(k, 0)), g1 = *reinterpret_cast<const double2*>(sg + rows64_chunk(k, 1));
				const double2 g2 = *reinterpret_cast<const double2*>(sg + rows64_chunk(k, 2)), g3 = *reinterpret_cast<const double2*>(sg + rows64_chunk(k, 3));
				ga[0] = g0.x; ga[1] = g0.y; ga[2] = g1.x; ga[3] = g1.y; gb[0] = g2.x; gb[1] = g2.y; gb[2] = g3.x; gb[3] = g3.y;
				rc = src[k];
			}
			double pc[4];
			if(RECON == FR_FIRST) { for(int q = 0; q < 4; q++) pc[q] = uc[q]; }
			else cons2prim(A.gas, uc, pc);
			const unsigned nb[4] = {cl.x & 0xFFFFu, cl.x >> 16, cl.y & 0xFFFFu, cl.y >> 16};
			const unsigned cf[4] = {cl.z & 0xFFFFu, cl.z >> 16, cl.w & 0xFFFFu, cl.w >> 16};
			// all midpoints are fetched before the extrapolations start (one shared-memory round trip instead of four)
			const bool quad = nb[3] != NB_NONE;           // only the fourth slot can be empty (triangles)
			double2 grj[4];
			#pragma unroll
			for(int j = 0; j < 4; j++) {
				grj[j] = make_double2(0, 0);
				if(RECON == FR_LINEAR && (j < 3 || quad)) grj[j] = sgr[cf[j] & 0x7FFFu];
			}
			#pragma unroll
			for(int j = 0; j < 4; j++) {
				if(j == 3 && !quad) break;
				const int e = (int)(cf[j] & 0x7FFFu);
				double pf[4];
				if(RECON == FR_LINEAR) extrapolate_prim(pc, ga, gb, grj[j].x, grj[j].y, rc.x, rc.y, pf);
				else { for(int q = 0; q < 4; q++) pf[q] = pc[q]; }
				const bool right = (cf[j] & 0x8000u) != 0;
				double2 *const dst = (right ? fsR : fsL) + e;
				dst[0] = make_double2(pf[0], pf[1]);
				dst[EP] = make_double2(pf[2], pf[3]);
				// a face cut by the tile boundary: its other side is a halo cell, reconstructed in phase B, which then
				// finds the face midpoint in that side's (otherwise unused) slot instead of fetching it from global memory
				if(RECON == FR_LINEAR && e >= D.cut0 && e < D.bnd0) (right ? fsL : fsR)[e] = grj[j];
			}
		}
		mbar_wait(bar + 1, par);
		cp_async_wait_all();        // this tile's halo rows: gathered since the previous tile's phase B ended
		__syncthreads();
		// group A buffers are free: the next tile's phase-A inputs (and its stencil/area into the other C buffer)
		if(tid == 0 && have_next && !keepA) { fence_proxy_async(); issue_AC(unpack_tile_desc(rnext), (int)(par ^ 1u)); }

		// ---- phase B: fluxes, one real stream entry per thread and round (the list `sord` skips the padding entries, so
		// the rounds are full: consecutive threads still take nearly consecutive entries)
		for(int ib = 0; ib < D.nreal; ib += FACE_BLOCK) {
			if(ib + tid >= D.nreal) continue;
			const int e = (int)sord[ib + tid];
			const unsigned LR = sLR[e];
			if(LR == LR_PAD) continue;
			const double2 nrm = sn[e];
			const double len = slen[e];
			const unsigned L = LR & 0xFFFFu, Rf = LR >> 16;
			const bool bnd = Rf >= LR_BND;
			const double nx = nrm.x, ny = nrm.y;
			const BCEntry &bc = A.gas.bc[Rf & 15u];
			double sl[4], sr[4];       // face states: conserved (first order) / primitive; MUSCL: cell states
			if(L < (unsigned)D.nc) ldp4(fsL + e, EP, sl);
			else halo_side_state<RECON>(A, hu, hg, hrc, (int)L - D.nc, RECON == FR_LINEAR ? fsL[e] : make_double2(0,0), sl);
			if(!bnd) {
				if(Rf < (unsigned)D.nc) ldp4(fsR + e, EP, sr);
				else halo_side_state<RECON>(A, hu, hg, hrc, (int)Rf - D.nc, RECON == FR_LINEAR ? fsR[e] : make_double2(0,0), sr);
			}
			const bool need_gid = ((VISC != VISC_NONE) && !keepA) || RECON == FR_MUSCL;
			const int gidL = need_gid ? tile_global_h(M, D.h0, D.c0, D.nc, L) : 0;
			const int gidR = need_gid ? (bnd ? gidL : tile_global_h(M, D.h0, D.c0, D.nc, Rf)) : 0;

			Side a, bs;
			double ucl[4], ucr[4];      // conserved cell states for the viscous flux (right = ghost of the cell state)
			if(RECON == FR_FIRST) {
				if(bnd) ghost_state(A.gas, bc, sl, nx, ny, sr);
				a = load_side<true>(A.gas, sl, nx, ny);
				bs = load_side<true>(A.gas, sr, nx, ny);
				if(VISC != VISC_NONE) for(int q = 0; q < 4; q++) { ucl[q] = sl[q]; ucr[q] = sr[q]; }
			}
			else if(RECON == FR_LINEAR) {
				a = side_from_prim<true>(A.gas, sl, nx, ny);
				if(bnd) {
					const double ul[4] = {a.r, a.mx, a.my, a.E};
					double ur[4];
					ghost_state(A.gas, bc, ul, nx, ny, ur);
					bs = load_side<true>(A.gas, ur, nx, ny);
				} else bs = side_from_prim<true>(A.gas, sr, nx, ny);
				if(VISC != VISC_NONE) {
					if(keepA) {
						if(L < (unsigned)D.nc) { const double2 c0_ = *reinterpret_cast<const double2*>(su + rows32_chunk((int)L, 0)), c1_ = *reinterpret_cast<const double2*>(su + rows32_chunk((int)L, 1)); ucl[0] = c0_.x; ucl[1] = c0_.y; ucl[2] = c1_.x; ucl[3] = c1_.y; }
						else lds4(hu + 4*((int)L - D.nc), ucl);
						if(bnd) ghost_state(A.gas, bc, ucl, nx, ny, ucr);
						else if(Rf < (unsigned)D.nc) { const double2 c0_ = *reinterpret_cast<const double2*>(su + rows32_chunk((int)Rf, 0)), c1_ = *reinterpret_cast<const double2*>(su + rows32_chunk((int)Rf, 1)); ucr[0] = c0_.x; ucr[1] = c0_.y; ucr[2] = c1_.x; ucr[3] = c1_.y; }
						else lds4(hu + 4*((int)Rf - D.nc), ucr);
					} else {
						ld4(ghost_aware_row(A.u, gp->u, M.ncell, gidL, 4), ucl);
						if(bnd) ghost_state(A.gas, bc, ucl, nx, ny, ucr);
						else ld4(ghost_aware_row(A.u, gp->u, M.ncell, gidR, 4), ucr);
					}
				}
			}
			else { // MUSCL with Van Albada limiter: sl, sr are the primitive CELL states
				const double2 gr = M.fgr[D.e0 + e];
				const double2 rl = M.rc[gidL];
				double2 rr;
				if(bnd) {
					prim2cons(A.gas, sl, ucl);
					ghost_state(A.gas, bc, ucl, nx, ny, ucr);
					cons2prim(A.gas, ucr, sr);
					rr = make_double2(2.0*gr.x - rl.x, 2.0*gr.y - rl.y);     // ghost centre (aspatial.cpp:98-119)
				} else {
					rr = M.rc[gidR];
					if(VISC != VISC_NONE) { prim2cons(A.gas, sl, ucl); prim2cons(A.gas, sr, ucr); }
				}
				const double dx = rr.x - rl.x, dy = rr.y - rl.y;
				double ga[4], gb[4], pfl[4], pfr[4];
				{ const double *const gl_ = ghost_aware_row(gsrc, gp->g, M.ncell, gidL, 8); ld4(gl_, ga); ld4(gl_ + 4, gb); }
				{
					const double gx[4] = {ga[0], ga[2], gb[0], gb[2]}, gy[4] = {ga[1], ga[3], gb[1], gb[3]};
					for(int q = 0; q < 4; q++) {
						const double dlr = sr[q] - sl[q];
						pfl[q] = sl[q] + muscl_term(2.0*(gx[q]*dx + gy[q]*dy) - dlr, dlr);
					}
				}
				a = side_from_prim<true>(A.gas, pfl, nx, ny);
				if(bnd) {
					const double ul[4] = {a.r, a.mx, a.my, a.E};
					double ur[4];
					ghost_state(A.gas, bc, ul, nx, ny, ur);
					bs = load_side<true>(A.gas, ur, nx, ny);
				} else {
					{ const double *const gr_ = ghost_aware_row(gsrc, gp->g, M.ncell, gidR, 8); ld4(gr_, ga); ld4(gr_ + 4, gb); }
					const double gx[4] = {ga[0], ga[2], gb[0], gb[2]}, gy[4] = {ga[1], ga[3], gb[1], gb[3]};
					for(int q = 0; q < 4; q++) {
						const double dlr = sr[q] - sl[q];
						pfr[q] = sr[q] - muscl_term(2.0*(gx[q]*dx + gy[q]*dy) - dlr, dlr);
					}
					bs = side_from_prim<true>(A.gas, pfr, nx, ny);
				}
			}

			double f[4];
			flux_from_sides<FLUX>(A.gas, a, bs, nx, ny, f);
			for(int q = 0; q < 4; q++) f[q] *= len;
			double sri = (fabs(a.vn) + a.c)*len;
			double srj = (fabs(bs.vn) + bs.c)*len;

			if(VISC != VISC_NONE) {
				const bool ownL = L < (unsigned)D.nc, ownR = !bnd && Rf < (unsigned)D.nc;
				double2 rl, rr;
				double gl[8], grr[8], vf[4];
				if(RECON == FR_LINEAR && keepA) {
					// centres and gradient rows of both cells out of shared memory (own cells: the swizzled TMA rows)
					auto rows = [&](bool own, unsigned loc, double2 &rc_, double g_[8]) {
						if(own) {
							rc_ = src[loc];
							#pragma unroll
							for(int c = 0; c < 4; c++) { const double2 v = *reinterpret_cast<const double2*>(sg + rows64_chunk((int)loc, c)); g_[2*c] = v.x; g_[2*c+1] = v.y; }
						} else {
							const int h = (int)loc - D.nc;
							rc_ = hrc[h];
							lds4(hg + 8*h, g_); lds4(hg + 8*h + 4, g_ + 4);
						}
					};
					rows(ownL, L, rl, gl);
					if(bnd) {
						const double2 gr = sgr[e];
						rr = make_double2(2.0*gr.x - rl.x, 2.0*gr.y - rl.y);
						for(int q = 0; q < 8; q++) grr[q] = gl[q];
					} else rows(ownR, Rf, rr, grr);
				} else {
					rl = M.rc[gidL];
					if(bnd) {
						const double2 gr = M.fgr[D.e0 + e];
						rr = make_double2(2.0*gr.x - rl.x, 2.0*gr.y - rl.y);
					} else rr = M.rc[gidR];
					if(RECON != FR_FIRST) {
						{ const double *const gl_ = ghost_aware_row(A.gu, gp->v, M.ncell, gidL, 8); ld4(gl_, gl); ld4(gl_ + 4, gl+4); }
						if(bnd) for(int q = 0; q < 8; q++) grr[q] = gl[q];
						else { const double *const gr_ = ghost_aware_row(A.gu, gp->v, M.ncell, gidR, 8); ld4(gr_, grr); ld4(gr_ + 4, grr+4); }
					}
				}
				// viscosities of the two face states (needed by the flux and by the spectral radius); a.ir, a.p etc. are at hand
				const double iRe = frcp(A.gas.Reinf);
				const double mui = VISC == VISC_CONST ? iRe : sutherland_fast(A.gas, a.p, a.ir);
				const double muj = VISC == VISC_CONST ? iRe : sutherland_fast(A.gas, bs.p, bs.ir);
				viscous_face_flux_fast<RECON != FR_FIRST>(A.gas, nx, ny, rl.x, rl.y, rr.x, rr.y, ucl, ucr, gl, grr,
					a.vx, a.vy, bs.vx, bs.vy, mui, muj, vf);
				for(int q = 0; q < 4; q++) f[q] += vf[q]*len;
				// max(4/(3 rho), gamma/rho) mu/Pr len^2/area (flow_spatial.cpp:600-617); both terms of the max scale with 1/rho
				const double cmax = fmax(4.0/3.0, A.gas.g)*frcp(A.gas.Pr)*len*len;
				if(RECON == FR_LINEAR && keepA) {
					// (the spectral radius of a side is consumed only by a cell of this tile: a halo side needs none)
					if(ownL) sri += cmax*a.ir*mui*frcp(sar[(int)L + aoff]);
					if(ownR) srj += cmax*bs.ir*muj*frcp(sar[(int)Rf + aoff]);
				} else {
					sri += cmax*a.ir*mui*frcp(M.area[gidL]);
					if(!bnd) srj += cmax*bs.ir*muj*frcp(M.area[gidR]);
				}
			}
			// the entry's slots now carry its flux and the two spectral radii
			fsL[e] = make_double2(f[0], f[1]);
			fsL[e + EP] = make_double2(f[2], f[3]);
			fsR[e] = make_double2(sri, srj);
		}
		__syncthreads();
		// group B buffers are free: the next tile's entry metadata and halo rows
		if(have_next) {
			if(tid == 0) { fence_proxy_async(); if(keepA) issue_AC(unpack_tile_desc(rnext), (int)(par ^ 1u)); issue_B(unpack_tile_desc(rnext)); }
			if(!waited && (rnext[1].w >> 16)) { wait_for_ghost_rows(); waited = true; }
			issue_halo(rnext[1].x, gnext);
		}

		// ---- phase C: per-cell sums in local-face order, then the epilogue
		double part = 0.0;
		for(int k = tid; k < D.nc; k += FACE_BLOCK) {
			const size_t c = (size_t)(D.c0 + k);
			// caller-ordered outputs: the row of res / dtm / unew that belongs to device cell c
			const size_t co = A.dst_idx ? (size_t)A.dst_idx[c] : c;
			const uint4 cl = scl[k];
			const double ar = sar[k + aoff];
			const unsigned nb[4] = {cl.x & 0xFFFFu, cl.x >> 16, cl.y & 0xFFFFu, cl.y >> 16};
			const unsigned cf[4] = {cl.z & 0xFFFFu, cl.z >> 16, cl.w & 0xFFFFu, cl.w >> 16};
			// all rows are fetched before the sums start (no branch between the loads); a missing fourth face reads
			// entry 0 with weight zero. fma(+-1, f, r) is the exact add / subtract, in local-face order.
			double r[4] = {0,0,0,0}, integ = 0.0;
			double f[4][4];
			double2 sr[4];
			double sg_[4];
			#pragma unroll
			for(int j = 0; j < 4; j++) {
				const bool have = !(j == 3 && nb[3] == NB_NONE);
				const int e = have ? (int)(cf[j] & 0x7FFFu) : 0;
				ldp4(fsL + e, EP, f[j]);
				sr[j] = fsR[e];
				sg_[j] = !have ? 0.0 : ((cf[j] & 0x8000u) ? 1.0 : -1.0);
			}
			#pragma unroll
			for(int j = 0; j < 4; j++) {
				#pragma unroll
				for(int v = 0; v < 4; v++) r[v] = fma(sg_[j], f[j][v], r[v]);
				integ += sg_[j] > 0.0 ? sr[j].y : (sg_[j] < 0.0 ? sr[j].x : 0.0);
			}
			if(A.epilogue == EP_RESIDUAL) {
				if(A.accumulate) {
					double o[4];
					ld4c(A.res + 4*co, o);
					for(int v = 0; v < 4; v++) r[v] += o[v];
				}
				st4(A.res + 4*co, r);
				if(A.gettimesteps) A.dtm[co] = ar/integ;
			} else {
				const double dt = ar/integ;
				const double fac = A.cfl*dt/ar;
				double uo[4];
				if(k == tid) { for(int q = 0; q < 4; q++) uo[q] = uc0[q]; } else ld4(A.u + 4*c, uo);
				uo[0] += fac*r[0]; uo[1] += fac*r[1]; uo[2] += fac*r[2]; uo[3] += fac*r[3];
				st4(A.unew + 4*co, uo);
				part += r[3]*r[3]*ar;
			}
		}
		if(A.epilogue == EP_STEP) {
			// fixed-order reduction: warp shuffle tree here, one partial per (tile, warp); the norm kernel sums them in
			// index order (no CTA barrier and no atomics: bitwise reproducible)
			for(int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
			if((tid & 31) == 0) A.partial[(size_t)t*(FACE_BLOCK/32) + (tid >> 5)] = part;
		}
		// phase C reads the flux slots that the next tile's phase A overwrites
		__syncthreads();
		// fused multi-GPU pseudo-time step: the tile's updated state rows go to the neighbours as the state of the next
		// evaluation (the barrier above ordered the stores before the read-back)
		if(A.epilogue == EP_STEP && A.dist.d && (A.dist.push & (1u << X_U))) dist_push_tile(A.dist.d, X_U, gp->dk + 1, t, D.c0, A.unew);
		ti = tin;
		if(have_next) { D = unpack_tile_desc(rnext); t = D.t; }
	}
	if(A.dist.d && A.dist.last) dist_finish_evaluation(A.dist.d, gp->dk, A.epilogue == EP_STEP && (A.dist.push & (1u << X_U)) != 0);
}

template <int FLUX, int RECON, int VISC>
static int launch_one(const FaceArgs &a, cudaStream_t s)
{
	const FaceSmem S(a.m.TC, a.m.EMAX, a.m.HMAX, RECON != FR_FIRST, RECON == FR_LINEAR);
	const size_t smem = (size_t)S.total;
	if(smem > 48*1024) {
		const cudaError_t ea = cudaFuncSetAttribute(face_kernel<FLUX,RECON,VISC>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(ea != cudaSuccess) return cuda_fail(ea, "face_kernel smem attribute", __FILE__, __LINE__);
	}
	if(a.m.HMAX > FACE_BLOCK) { set_error("face kernel: halo capacity exceeds the CTA size"); return FVG_ERR_INVALID; }
	// persistent CTAs: as many as are resident at once (depends on the device and, through the shared-memory size, on
	// the mesh's staging capacities)
	const int ctas = resident_ctas((const void*)face_kernel<FLUX,RECON,VISC>, FACE_BLOCK, smem);
	FaceArgs b = a;
	if(b.tile1 < 0) b.tile1 = b.m.ntile;
	b.tdesc = (b.ordered && b.m.tdesc_ord) ? b.m.tdesc_ord : b.m.tdesc;
	const int nt = b.tile1 - b.tile0;
	if(nt <= 0) return 0;
	{
		cudaLaunchConfig_t cfg{};
		cfg.gridDim = dim3((unsigned)(nt < ctas ? nt : ctas)); cfg.blockDim = dim3(FACE_BLOCK); cfg.dynamicSmemBytes = smem; cfg.stream = s;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
		cfg.attrs = attr; cfg.numAttrs = 1;
		const cudaError_t el = cudaLaunchKernelEx(&cfg, face_kernel<FLUX,RECON,VISC>, b);
		if(el != cudaSuccess) return cuda_fail(el, "face_kernel launch", __FILE__, __LINE__);
	}
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "face_kernel launch", __FILE__, __LINE__);
	return 0;
}

template <int FLUX>
static int launch_flux(int recon, int visc, const FaceArgs &a, cudaStream_t s)
{
#define FVG_CASE(R,V) if(recon == R && visc == V) return launch_one<FLUX,R,V>(a, s);
	FVG_CASE(FR_FIRST, VISC_NONE) FVG_CASE(FR_FIRST, VISC_CONST) FVG_CASE(FR_FIRST, VISC_SUTHERLAND)
	FVG_CASE(FR_LINEAR, VISC_NONE) FVG_CASE(FR_LINEAR, VISC_CONST) FVG_CASE(FR_LINEAR, VISC_SUTHERLAND)
	FVG_CASE(FR_MUSCL, VISC_NONE) FVG_CASE(FR_MUSCL, VISC_CONST) FVG_CASE(FR_MUSCL, VISC_SUTHERLAND)
#undef FVG_CASE
	set_error("face kernel: bad reconstruction/viscosity selector");
	return FVG_ERR_INVALID;
}

} // namespace fvg
