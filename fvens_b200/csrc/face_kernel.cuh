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

/// Device cell index of a tile-local index
__device__ __forceinline__ int tile_global(const DMesh &M, int t, int c0, int nc, unsigned loc) {
	return loc < (unsigned)nc ? c0 + (int)loc : M.thalo[M.thoff[t] + (int)loc - nc];
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
struct FaceSmem {
	int fsL, fsR, sgr, sn, slen, sLR, hu, hg, hrc, su, sg, src, scl, sar, bar, total;   // byte offsets
	__host__ __device__ FaceSmem(int TC, int EMAX, int HMAX, bool mids, bool linear) {
		int o = 0;
		fsL = o; o += EMAX*32;
		fsR = o; o += EMAX*32;
		sgr = o; o += mids ? EMAX*16 : 0;
		sn = o; o += EMAX*16;
		slen = o; o += EMAX*8;
		sLR = o; o += EMAX*4;
		hu = o; o += HMAX*32;
		hg = o; o += mids ? HMAX*64 : 0;
		hrc = o; o += mids ? HMAX*16 : 0;
		su = o; o += TC*32;                    // own cells: state, reconstruction gradient, centre, stencil, area
		sg = o; o += linear ? TC*64 : 0;
		src = o; o += linear ? TC*16 : 0;
		scl = o; o += TC*16;
		sar = o; o += (TC + 2)*8;
		bar = o; o += 16;                      // two mbarriers
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

/** One CTA per tile. Everything the tile reads arrives in shared memory asynchronously - the own cells' rows
 *  (state, reconstruction gradient, centre, stencil, area) and the stream-entry metadata (midpoints, normals,
 *  lengths, local indices) by 1-D TMA bulk copies (contiguous per tile), the halo cells' rows by 16-byte
 *  cp.async gathers - so no thread holds registers for loads in flight, and the second CTA resident on the SM
 *  computes while this one's copies land. Then three phases separated by two barriers:
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
face_kernel(const FaceArgs A)
{
	extern __shared__ __align__(128) unsigned char smraw[];
	const DMesh &M = A.m;
	constexpr bool MIDS = RECON != FR_FIRST;
	constexpr bool LINEAR = RECON == FR_LINEAR;
	const FaceSmem S(M.TC, M.EMAX, M.HMAX, MIDS, LINEAR);
	double *const fsL = reinterpret_cast<double*>(smraw + S.fsL);
	double *const fsR = reinterpret_cast<double*>(smraw + S.fsR);
	double2 *const sgr = reinterpret_cast<double2*>(smraw + S.sgr);
	double2 *const sn = reinterpret_cast<double2*>(smraw + S.sn);
	double *const slen = reinterpret_cast<double*>(smraw + S.slen);
	unsigned *const sLR = reinterpret_cast<unsigned*>(smraw + S.sLR);
	double *const hu = reinterpret_cast<double*>(smraw + S.hu);
	double *const hg = reinterpret_cast<double*>(smraw + S.hg);
	double2 *const hrc = reinterpret_cast<double2*>(smraw + S.hrc);
	double *const su = reinterpret_cast<double*>(smraw + S.su);
	double *const sg = reinterpret_cast<double*>(smraw + S.sg);
	double2 *const src = reinterpret_cast<double2*>(smraw + S.src);
	uint4 *const scl = reinterpret_cast<uint4*>(smraw + S.scl);
	double *const sar = reinterpret_cast<double*>(smraw + S.sar);
	uint64_t *const bar = reinterpret_cast<uint64_t*>(smraw + S.bar);       // [0]: phase A inputs, [1]: the rest
	__shared__ double red_s[FACE_BLOCK/32];

	const int t = blockIdx.x, tid = threadIdx.x;
	const int c0 = M.tcell0[t], nc = M.tcell0[t+1] - c0;
	const int h0 = M.thoff[t], nh = M.thoff[t+1] - h0;
	const int e0 = M.fsoff[t], ne = M.fsoff[t+1] - e0;
	const int ecut = M.tbnd[t].x;                                   // first entry with a halo side
	const double *const gsrc = RECON == FR_MUSCL ? A.gu : A.lg;     // gradients used by the reconstruction
	const int aoff = c0 & 1;                                        // 8-byte rows are copied from the even cell below c0

	if(tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
	__syncthreads();
	if(tid == 0) {
		mbar_expect_tx(bar, (unsigned)nc*(32u + 16u + (LINEAR ? 80u : 0u)) + (MIDS ? (unsigned)ne*16u : 0u));
		bulk_g2s(su, A.u + 4*(size_t)c0, (unsigned)nc*32u, bar);
		bulk_g2s(scl, M.cloc + c0, (unsigned)nc*16u, bar);
		if(LINEAR) { bulk_g2s(sg, gsrc + 8*(size_t)c0, (unsigned)nc*64u, bar); bulk_g2s(src, M.rc + c0, (unsigned)nc*16u, bar); }
		if(MIDS) bulk_g2s(sgr, M.fgr + e0, (unsigned)ne*16u, bar);
		const unsigned abytes = (unsigned)((nc + aoff + 1) & ~1)*8u;
		mbar_expect_tx(bar + 1, (unsigned)ne*(16u + 8u + 4u) + abytes);
		bulk_g2s(sLR, M.fLR + e0, (unsigned)ne*4u, bar + 1);
		bulk_g2s(sn, M.fn + e0, (unsigned)ne*16u, bar + 1);
		bulk_g2s(slen, M.flen + e0, (unsigned)ne*8u, bar + 1);
		bulk_g2s(sar, M.area + (c0 - aoff), abytes, bar + 1);
	}
	for(int h = tid; h < nh; h += FACE_BLOCK) {
		const size_t g = (size_t)M.thalo[h0 + h];
		cp_async16(hu + 4*h, A.u + 4*g);
		cp_async16(hu + 4*h + 2, A.u + 4*g + 2);
		if(MIDS) {
			#pragma unroll
			for(int q = 0; q < 4; q++) cp_async16(hg + 8*h + 2*q, gsrc + 8*g + 2*q);
			cp_async16(hrc + h, M.rc + g);
		}
	}
	cp_async_commit();
	if(tid == 32 && A.prefetch_distance > 0 && t + A.prefetch_distance < M.ntile) {
		// warm L2 with the operands of the tile that runs about one wave of CTAs later
		const int tp = t + A.prefetch_distance;
		const int pc0 = M.tcell0[tp], pnc = M.tcell0[tp+1] - pc0;
		const int pe0 = M.fsoff[tp], pne = M.fsoff[tp+1] - pe0;
		bulk_prefetch_l2(A.u + 4*(size_t)pc0, (unsigned)pnc*32u);
		bulk_prefetch_l2(M.cloc + pc0, (unsigned)pnc*16u);
		if(RECON != FR_FIRST) { bulk_prefetch_l2(gsrc + 8*(size_t)pc0, (unsigned)pnc*64u); bulk_prefetch_l2(M.rc + pc0, (unsigned)pnc*16u); }
		bulk_prefetch_l2(M.fLR + pe0, (unsigned)pne*4u);
		bulk_prefetch_l2(M.fn + pe0, (unsigned)pne*16u);
		bulk_prefetch_l2(M.flen + pe0, (unsigned)pne*8u);
		if(MIDS) bulk_prefetch_l2(M.fgr + pe0, (unsigned)pne*16u);
		bulk_prefetch_l2(M.area + (pc0 & ~1), (unsigned)((pnc + 3) & ~1)*8u);
	}

	// ---- phase A: face states of the own cells
	mbar_wait(bar, 0);
	for(int k = tid; k < nc; k += FACE_BLOCK) {
		const uint4 cl = scl[k];
		double uc[4], ga[4] = {0,0,0,0}, gb[4] = {0,0,0,0};
		double2 rc = make_double2(0,0);
		lds4(su + 4*k, uc);
		if(LINEAR) { lds4(sg + 8*k, ga); lds4(sg + 8*k + 4, gb); rc = src[k]; }
		{
			double pc[4];
			if(RECON == FR_FIRST) { for(int q = 0; q < 4; q++) pc[q] = uc[q]; }
			else cons2prim(A.gas, uc, pc);
			const unsigned nb[4] = {cl.x & 0xFFFFu, cl.x >> 16, cl.y & 0xFFFFu, cl.y >> 16};
			const unsigned cf[4] = {cl.z & 0xFFFFu, cl.z >> 16, cl.w & 0xFFFFu, cl.w >> 16};
			#pragma unroll
			for(int j = 0; j < 4; j++) {
				if(j == 3 && nb[3] == NB_NONE) break;      // only the fourth slot can be empty (triangles)
				const int e = (int)(cf[j] & 0x7FFFu);
				double pf[4];
				if(RECON == FR_LINEAR) {
					const double2 gr = sgr[e];
					extrapolate_prim(pc, ga, gb, gr.x, gr.y, rc.x, rc.y, pf);
				} else { for(int q = 0; q < 4; q++) pf[q] = pc[q]; }
				double *const dst = ((cf[j] & 0x8000u) ? fsR : fsL) + 4*e;
				*reinterpret_cast<double2*>(dst) = make_double2(pf[0], pf[1]);
				*reinterpret_cast<double2*>(dst + 2) = make_double2(pf[2], pf[3]);
			}
		}
	}
	mbar_wait(bar + 1, 0);
	__syncthreads();

	// ---- phase B: fluxes, one stream entry per thread and round. The halo rows are needed from entry `ecut` on:
	// the round that reaches it first waits for the gathers (uniform across the CTA: the test is on the round)
	bool halo_ready = false;
	for(int eb = 0; eb < ne; eb += FACE_BLOCK) {
		if(!halo_ready && eb + FACE_BLOCK > ecut) { cp_async_wait_all(); __syncthreads(); halo_ready = true; }
		const int e = eb + tid;
		if(e >= ne) continue;
		const unsigned LR = sLR[e];
		if(LR == LR_PAD) continue;
		const double2 nrm = sn[e];
		const double len = slen[e];
		const unsigned L = LR & 0xFFFFu, Rf = LR >> 16;
		const bool bnd = Rf >= LR_BND;
		const double nx = nrm.x, ny = nrm.y;
		const BCEntry &bc = A.gas.bc[Rf & 15u];
		double sl[4], sr[4];       // face states: conserved (first order) / primitive; MUSCL: cell states
		if(L < (unsigned)nc) lds4(fsL + 4*e, sl);
		else halo_side_state<RECON>(A, hu, hg, hrc, (int)L - nc, MIDS ? sgr[e] : make_double2(0,0), sl);
		if(!bnd) {
			if(Rf < (unsigned)nc) lds4(fsR + 4*e, sr);
			else halo_side_state<RECON>(A, hu, hg, hrc, (int)Rf - nc, MIDS ? sgr[e] : make_double2(0,0), sr);
		}
		const int gidL = (VISC != VISC_NONE || RECON == FR_MUSCL) ? tile_global(M, t, c0, nc, L) : 0;
		const int gidR = (VISC != VISC_NONE || RECON == FR_MUSCL) ? (bnd ? gidL : tile_global(M, t, c0, nc, Rf)) : 0;

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
				ld4(A.u + 4*(size_t)gidL, ucl);
				if(bnd) ghost_state(A.gas, bc, ucl, nx, ny, ucr);
				else ld4(A.u + 4*(size_t)gidR, ucr);
			}
		}
		else { // MUSCL with Van Albada limiter: sl, sr are the primitive CELL states
			const double2 gr = sgr[e];
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
			ld4(gsrc + 8*(size_t)gidL, ga); ld4(gsrc + 8*(size_t)gidL + 4, gb);
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
				ld4(gsrc + 8*(size_t)gidR, ga); ld4(gsrc + 8*(size_t)gidR + 4, gb);
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
			const double ul[4] = {a.r, a.mx, a.my, a.E}, ur[4] = {bs.r, bs.mx, bs.my, bs.E};
			const double2 rl = M.rc[gidL];
			double2 rr;
			if(bnd) {
				const double2 gr = M.fgr[e0 + e];
				rr = make_double2(2.0*gr.x - rl.x, 2.0*gr.y - rl.y);
			} else rr = M.rc[gidR];
			double gl[8], grr[8], vf[4];
			if(RECON != FR_FIRST) {
				ld4(A.gu + 8*(size_t)gidL, gl); ld4(A.gu + 8*(size_t)gidL + 4, gl+4);
				if(bnd) for(int q = 0; q < 8; q++) grr[q] = gl[q];
				else { ld4(A.gu + 8*(size_t)gidR, grr); ld4(A.gu + 8*(size_t)gidR + 4, grr+4); }
			}
			viscous_face_flux<RECON != FR_FIRST, VISC == VISC_CONST>(A.gas, nx, ny, rl.x, rl.y, rr.x, rr.y,
				ucl, ucr, gl, grr, ul, ur, vf);
			for(int q = 0; q < 4; q++) f[q] += vf[q]*len;
			const double mui = VISC == VISC_CONST ? 1.0/A.gas.Reinf : viscosity_cons(A.gas, ul);
			const double muj = VISC == VISC_CONST ? 1.0/A.gas.Reinf : viscosity_cons(A.gas, ur);
			const double coi = fmax(4.0/(3.0*ul[0]), A.gas.g/ul[0]);
			const double coj = fmax(4.0/(3.0*ur[0]), A.gas.g/ur[0]);
			sri += coi*mui/A.gas.Pr*len*len/M.area[gidL];
			if(!bnd) srj += coj*muj/A.gas.Pr*len*len/M.area[gidR];
		}
		// the entry's slots now carry its flux and the two spectral radii
		*reinterpret_cast<double2*>(fsL + 4*e) = make_double2(f[0], f[1]);
		*reinterpret_cast<double2*>(fsL + 4*e + 2) = make_double2(f[2], f[3]);
		*reinterpret_cast<double2*>(fsR + 4*e) = make_double2(sri, srj);
	}
	__syncthreads();

	// ---- phase C: per-cell sums in local-face order, then the epilogue
	double part = 0.0;
	for(int k = tid; k < nc; k += FACE_BLOCK) {
		const size_t c = (size_t)(c0 + k);
		const uint4 cl = scl[k];
		const double ar = sar[k + aoff];
		const unsigned nb[4] = {cl.x & 0xFFFFu, cl.x >> 16, cl.y & 0xFFFFu, cl.y >> 16};
		const unsigned cf[4] = {cl.z & 0xFFFFu, cl.z >> 16, cl.w & 0xFFFFu, cl.w >> 16};
		double r[4] = {0,0,0,0}, integ = 0.0;
		#pragma unroll
		for(int j = 0; j < 4; j++) {
			if(j == 3 && nb[3] == NB_NONE) break;
			const int e = (int)(cf[j] & 0x7FFFu);
			double f[4];
			lds4(fsL + 4*e, f);
			const double2 sr = *reinterpret_cast<const double2*>(fsR + 4*e);
			if(cf[j] & 0x8000u) { r[0] += f[0]; r[1] += f[1]; r[2] += f[2]; r[3] += f[3]; integ += sr.y; }
			else { r[0] -= f[0]; r[1] -= f[1]; r[2] -= f[2]; r[3] -= f[3]; integ += sr.x; }
		}
		if(A.epilogue == EP_RESIDUAL) {
			if(A.accumulate) {
				double o[4];
				ld4c(A.res + 4*c, o);
				for(int v = 0; v < 4; v++) r[v] += o[v];
			}
			st4(A.res + 4*c, r);
			if(A.gettimesteps) A.dtm[c] = ar/integ;
		} else {
			const double dt = ar/integ;
			const double fac = A.cfl*dt/ar;
			double uo[4];
			lds4(su + 4*k, uo);
			uo[0] += fac*r[0]; uo[1] += fac*r[1]; uo[2] += fac*r[2]; uo[3] += fac*r[3];
			st4(A.unew + 4*c, uo);
			part += r[3]*r[3]*ar;
		}
	}
	if(A.epilogue == EP_STEP) {
		// fixed-order block reduction: warp shuffle tree, then thread 0 sums the warp partials in order
		for(int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
		if((tid & 31) == 0) red_s[tid >> 5] = part;
		__syncthreads();
		if(tid == 0) {
			double s = 0.0;
			for(int w = 0; w < FACE_BLOCK/32; w++) s += red_s[w];
			A.partial[t] = s;
		}
	}
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
	face_kernel<FLUX,RECON,VISC><<<a.m.ntile, FACE_BLOCK, smem, s>>>(a);
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
