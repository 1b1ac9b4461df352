/* The fused face kernel: reconstruct -> (boundary ghost) -> inviscid flux [+ viscous flux] ->
 * spectral radius -> coloured accumulation into the tile's cells -> residual/time-step or fused
 * forward-Euler epilogue. Restates P6-P11 of FlowFV::compute_residual (reference:
 * src/spatial/flow_spatial.cpp:722-812, compute_fluxes :489-563, compute_max_timestep :567-634)
 * and, with the EP_STEP epilogue, the update + norm of SteadyForwardEulerSolver::solve
 * (src/ode/aodesolver.cpp:204-223).
 *
 * One CTA per tile. All of the tile's operands are staged in shared memory first, with copies that
 * hold no registers while in flight: the tile's own cells (state, gradients, centres) and its face
 * stream are contiguous in memory and arrive as 1-D TMA bulk copies counted on an mbarrier; the halo
 * cells (out-of-tile neighbours) are gathered with 16-byte cp.async. The flux phase then runs entirely
 * out of shared memory with 16-bit tile-local indices. Several CTAs are resident per SM, so one
 * tile's staging overlaps its neighbours' arithmetic.
 *
 * Faces cut by a tile boundary appear in both tiles and are evaluated identically in both. The face
 * stream is sorted by colour; no two faces of one colour share a tile cell, so after each colour
 * round a __syncthreads() is all the ordering the shared-memory accumulation needs. No atomics,
 * and the summation order per cell (colour order) is fixed => bitwise reproducible.
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

/// Shared-memory carve-up of the face kernel for the given capacities (host and device agree through this)
struct FaceSmem {
	int su, sg, src, sn, sgr, slen, sres, sLR, bar, total;   // byte offsets
	__host__ __device__ FaceSmem(int TC, int HMAX, int EMAX, bool grads, bool centres) {
		const int CAPC = TC + HMAX;
		int o = 0;
		su = o; o += CAPC*32;
		sg = o; o += grads ? CAPC*64 : 0;
		src = o; o += centres ? CAPC*16 : 0;
		sn = o; o += EMAX*16;
		sgr = o; o += grads ? EMAX*16 : 0;
		slen = o; o += EMAX*8;
		sres = o; o += 5*TC*8;
		sLR = o; o += EMAX*4;
		bar = (o + 7)/8*8; o = bar + 8;
		total = o;
	}
};

template <int FLUX, int RECON, int VISC>
__global__ void __launch_bounds__(FACE_BLOCK, FVG_FACE_MINB)
face_kernel(const FaceArgs A)
{
	extern __shared__ __align__(128) unsigned char smraw[];
	const DMesh &M = A.m;
	constexpr bool GRADS = RECON != FR_FIRST;
	constexpr bool CENTRES = RECON != FR_FIRST || VISC != VISC_NONE;
	const FaceSmem S(M.TC, M.HMAX, M.EMAX, GRADS, CENTRES);
	double *const su = reinterpret_cast<double*>(smraw + S.su);
	double *const sg = reinterpret_cast<double*>(smraw + S.sg);
	double2 *const src = reinterpret_cast<double2*>(smraw + S.src);
	double2 *const sn = reinterpret_cast<double2*>(smraw + S.sn);
	double2 *const sgr = reinterpret_cast<double2*>(smraw + S.sgr);
	double *const slen = reinterpret_cast<double*>(smraw + S.slen);
	double *const res_s = reinterpret_cast<double*>(smraw + S.sres);   // [4][TC] then integ [TC]
	unsigned *const sLR = reinterpret_cast<unsigned*>(smraw + S.sLR);
	uint64_t *const bar = reinterpret_cast<uint64_t*>(smraw + S.bar);
	__shared__ int coloff[MAXCOL+1];
	__shared__ double red_s[FACE_BLOCK/32];

	const int TC = M.TC;
	const int t = blockIdx.x, tid = threadIdx.x;
	const int c0 = M.tcell0[t], nc = M.tcell0[t+1] - c0;
	const int h0 = M.thoff[t], nh = M.thoff[t+1] - h0;
	const int e0 = M.fsoff[t], ne = M.fsoff[t+1] - e0;
	const double *const gsrc = RECON == FR_MUSCL ? A.gu : A.lg;     // gradients used by the reconstruction

	// ---- stage the tile
	if(tid == 0) mbar_init(bar, 1);
	__syncthreads();
	if(tid == 0) {
		unsigned bytes = (unsigned)nc*32u + (unsigned)ne*(16u + 8u + 4u);
		if(GRADS) bytes += (unsigned)nc*64u + (unsigned)ne*16u;
		if(CENTRES) bytes += (unsigned)nc*16u;
		mbar_expect_tx(bar, bytes);
		bulk_g2s(su, A.u + 4*(size_t)c0, (unsigned)nc*32u, bar);
		if(GRADS) bulk_g2s(sg, gsrc + 8*(size_t)c0, (unsigned)nc*64u, bar);
		if(CENTRES) bulk_g2s(src, M.rc + c0, (unsigned)nc*16u, bar);
		bulk_g2s(sLR, M.fLR + e0, (unsigned)ne*4u, bar);
		bulk_g2s(sn, M.fn + e0, (unsigned)ne*16u, bar);
		bulk_g2s(slen, M.flen + e0, (unsigned)ne*8u, bar);
		if(GRADS) bulk_g2s(sgr, M.fgr + e0, (unsigned)ne*16u, bar);
	}
	{
		// halo rows in 16-byte pieces: 2 of the state, 4 of the gradients, 1 of the centre
		constexpr int NP = 2 + (GRADS ? 4 : 0) + (CENTRES ? 1 : 0);
		for(int k = tid; k < nh*NP; k += FACE_BLOCK) {
			const int h = k/NP, piece = k - h*NP;
			const size_t g = (size_t)M.thalo[h0 + h];
			const int row = nc + h;
			if(piece < 2) cp_async16(su + 4*row + 2*piece, A.u + 4*g + 2*piece);
			else if(GRADS && piece < 6) cp_async16(sg + 8*row + 2*(piece-2), gsrc + 8*g + 2*(piece-2));
			else cp_async16(src + row, M.rc + g);
		}
		cp_async_commit();
	}
	for(int k = tid; k < 5*TC; k += FACE_BLOCK) res_s[k] = 0.0;
	if(tid <= MAXCOL) coloff[tid] = M.fcoloff[t*(MAXCOL+1) + tid] - e0;
	cp_async_wait_all();
	mbar_wait(bar, 0);
	__syncthreads();
	if(RECON != FR_FIRST) {
		// conserved -> primitive once per staged cell (the reconstruction works on primitive variables)
		for(int k = tid; k < nc + nh; k += FACE_BLOCK) {
			double uc[4], up[4];
			lds4(su + 4*k, uc);
			cons2prim(A.gas, uc, up);
			*reinterpret_cast<double2*>(su + 4*k) = make_double2(up[0], up[1]);
			*reinterpret_cast<double2*>(su + 4*k + 2) = make_double2(up[2], up[3]);
		}
		__syncthreads();
	}

	// ---- fluxes, one stream entry per thread and round
	for(int base = 0; base < ne; base += FACE_BLOCK) {
		const int e = base + tid;
		const unsigned LR = e < ne ? sLR[e] : LR_PAD;
		const bool valid = LR != LR_PAD;
		const unsigned L = LR & 0xFFFFu, Rf = LR >> 16;
		const bool bnd = Rf >= LR_BND;
		double f[4] = {0,0,0,0};
		double sri = 0, srj = 0;
		if(valid) {
			const double2 nrm = sn[e];
			const double len = slen[e];
			const double nx = nrm.x, ny = nrm.y;
			const BCEntry &bc = A.gas.bc[Rf & 15u];
			double ucl[4], ucr[4];      // conserved cell states (right = ghost of the cell state on a boundary)
			double pl[4];               // primitive left cell state (second order)
			Side a, bs;
			if(RECON == FR_FIRST) lds4(su + 4*L, ucl);
			else {
				lds4(su + 4*L, pl);
				if(VISC != VISC_NONE || RECON == FR_MUSCL) prim2cons(A.gas, pl, ucl);
			}
			if(RECON == FR_FIRST) {
				if(bnd) ghost_state(A.gas, bc, ucl, nx, ny, ucr);
				else lds4(su + 4*Rf, ucr);
				a = load_side<true>(A.gas, ucl, nx, ny);
				bs = load_side<true>(A.gas, ucr, nx, ny);
			}
			else {
				const double2 gr = sgr[e];
				const double2 rl = src[L];
				double pfl[4], pfr[4];
				if(RECON == FR_LINEAR) {
					extrapolate4(pl, sg + 8*L, gr.x - rl.x, gr.y - rl.y, pfl);
					a = side_from_prim<true>(A.gas, pfl, nx, ny);
					if(bnd) {
						const double ul[4] = {a.r, a.mx, a.my, a.E};
						double ur[4];
						ghost_state(A.gas, bc, ul, nx, ny, ur);
						bs = load_side<true>(A.gas, ur, nx, ny);
						if(VISC != VISC_NONE) ghost_state(A.gas, bc, ucl, nx, ny, ucr);
					} else {
						const double2 rr = src[Rf];
						double pr[4];
						lds4(su + 4*Rf, pr);
						if(VISC != VISC_NONE) prim2cons(A.gas, pr, ucr);
						extrapolate4(pr, sg + 8*Rf, gr.x - rr.x, gr.y - rr.y, pfr);
						bs = side_from_prim<true>(A.gas, pfr, nx, ny);
					}
				}
				else { // MUSCL with Van Albada limiter
					double pr[4];
					double2 rr;
					if(bnd) {
						ghost_state(A.gas, bc, ucl, nx, ny, ucr);
						cons2prim(A.gas, ucr, pr);
						rr = make_double2(2.0*gr.x - rl.x, 2.0*gr.y - rl.y);     // ghost centre (aspatial.cpp:98-119)
					} else {
						lds4(su + 4*Rf, pr);
						if(VISC != VISC_NONE) prim2cons(A.gas, pr, ucr);
						rr = src[Rf];
					}
					const double dx = rr.x - rl.x, dy = rr.y - rl.y;
					double ga[4], gb[4];
					lds4(sg + 8*L, ga); lds4(sg + 8*L + 4, gb);
					const double gLx[4] = {ga[0], ga[2], gb[0], gb[2]}, gLy[4] = {ga[1], ga[3], gb[1], gb[3]};
					for(int k = 0; k < 4; k++) {
						const double dlr = pr[k] - pl[k];
						const double dm = 2.0*(gLx[k]*dx + gLy[k]*dy) - dlr;
						pfl[k] = pl[k] + muscl_term(dm, dlr);
					}
					a = side_from_prim<true>(A.gas, pfl, nx, ny);
					if(bnd) {
						const double ul[4] = {a.r, a.mx, a.my, a.E};
						double ur[4];
						ghost_state(A.gas, bc, ul, nx, ny, ur);
						bs = load_side<true>(A.gas, ur, nx, ny);
					} else {
						lds4(sg + 8*Rf, ga); lds4(sg + 8*Rf + 4, gb);
						const double gRx[4] = {ga[0], ga[2], gb[0], gb[2]}, gRy[4] = {ga[1], ga[3], gb[1], gb[3]};
						for(int k = 0; k < 4; k++) {
							const double dlr = pr[k] - pl[k];
							const double dp = 2.0*(gRx[k]*dx + gRy[k]*dy) - dlr;
							pfr[k] = pr[k] - muscl_term(dp, dlr);
						}
						bs = side_from_prim<true>(A.gas, pfr, nx, ny);
					}
				}
			}

			flux_from_sides<FLUX>(A.gas, a, bs, nx, ny, f);
			for(int k = 0; k < 4; k++) f[k] *= len;
			sri = (fabs(a.vn) + a.c)*len;
			srj = (fabs(bs.vn) + bs.c)*len;

			if(VISC != VISC_NONE) {
				const double ul[4] = {a.r, a.mx, a.my, a.E}, ur[4] = {bs.r, bs.mx, bs.my, bs.E};
				const double2 rl = src[L];
				double2 rr;
				if(bnd) {
					const double2 gr = M.fgr[e0 + e];
					rr = make_double2(2.0*gr.x - rl.x, 2.0*gr.y - rl.y);
				} else rr = src[Rf];
				double gl[8], grr[8], vf[4];
				const int gidL = tile_global(M, t, c0, nc, L);
				const int gidR = bnd ? gidL : tile_global(M, t, c0, nc, Rf);
				if(RECON != FR_FIRST) {
					// the unlimited gradients may differ from the staged (limited) ones: read them from global memory
					ld4(A.gu + 8*(size_t)gidL, gl); ld4(A.gu + 8*(size_t)gidL + 4, gl+4);
					if(bnd) for(int k = 0; k < 8; k++) grr[k] = gl[k];
					else { ld4(A.gu + 8*(size_t)gidR, grr); ld4(A.gu + 8*(size_t)gidR + 4, grr+4); }
				}
				viscous_face_flux<RECON != FR_FIRST, VISC == VISC_CONST>(A.gas, nx, ny, rl.x, rl.y, rr.x, rr.y,
					ucl, ucr, gl, grr, ul, ur, vf);
				for(int k = 0; k < 4; k++) f[k] += vf[k]*len;
				const double mui = VISC == VISC_CONST ? 1.0/A.gas.Reinf : viscosity_cons(A.gas, ul);
				const double muj = VISC == VISC_CONST ? 1.0/A.gas.Reinf : viscosity_cons(A.gas, ur);
				const double coi = fmax(4.0/(3.0*ul[0]), A.gas.g/ul[0]);
				const double coj = fmax(4.0/(3.0*ur[0]), A.gas.g/ur[0]);
				sri += coi*mui/A.gas.Pr*len*len/M.area[gidL];
				if(!bnd) srj += coj*muj/A.gas.Pr*len*len/M.area[gidR];
			}
		}

		// colour rounds of this chunk
		int myc = 0, clo = 0, chi = 0;
		{
			const int last = min(base + FACE_BLOCK, ne) - 1;
			#pragma unroll
			for(int c = 1; c < MAXCOL; c++) {
				if(e >= coloff[c]) myc = c;
				if(base >= coloff[c]) clo = c;
				if(last >= coloff[c]) chi = c;
			}
		}
		const bool inL = valid && L < (unsigned)nc;
		const bool inR = valid && Rf < (unsigned)nc;
		for(int c = clo; c <= chi; c++) {
			if(myc == c) {
				if(inL) {
					res_s[L] -= f[0]; res_s[TC+L] -= f[1]; res_s[2*TC+L] -= f[2]; res_s[3*TC+L] -= f[3];
					res_s[4*TC+L] += sri;
				}
				if(inR) {
					res_s[Rf] += f[0]; res_s[TC+Rf] += f[1]; res_s[2*TC+Rf] += f[2]; res_s[3*TC+Rf] += f[3];
					res_s[4*TC+Rf] += srj;
				}
			}
			__syncthreads();
		}
	}

	// ---- epilogue: one thread per tile cell
	if(A.epilogue == EP_RESIDUAL) {
		for(int k = tid; k < nc; k += FACE_BLOCK) {
			const size_t c = (size_t)(c0 + k);
			double r[4] = {res_s[k], res_s[TC+k], res_s[2*TC+k], res_s[3*TC+k]};
			if(A.accumulate) {
				double o[4];
				ld4c(A.res + 4*c, o);
				for(int v = 0; v < 4; v++) r[v] += o[v];
			}
			st4(A.res + 4*c, r);
			if(A.gettimesteps) A.dtm[c] = M.area[c]/res_s[4*TC+k];
		}
	}
	else {
		double part = 0.0;
		for(int k = tid; k < nc; k += FACE_BLOCK) {
			const size_t c = (size_t)(c0 + k);
			const double ar = M.area[c];
			const double dt = ar/res_s[4*TC+k];
			const double fac = A.cfl*dt/ar;
			double uo[4];
			if(RECON == FR_FIRST) lds4(su + 4*k, uo);
			else { double po[4]; lds4(su + 4*k, po); prim2cons(A.gas, po, uo); }
			const double rE = res_s[3*TC+k];
			uo[0] += fac*res_s[k]; uo[1] += fac*res_s[TC+k]; uo[2] += fac*res_s[2*TC+k]; uo[3] += fac*rE;
			st4(A.unew + 4*c, uo);
			part += rE*rE*ar;
		}
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
	const FaceSmem S(a.m.TC, a.m.HMAX, a.m.EMAX, RECON != FR_FIRST, RECON != FR_FIRST || VISC != VISC_NONE);
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
