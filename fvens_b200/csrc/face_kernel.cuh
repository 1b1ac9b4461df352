/* The fused face kernel: reconstruct -> (boundary ghost) -> inviscid flux [+ viscous flux] ->
 * spectral radius -> coloured accumulation into the tile's cells -> residual/time-step or fused
 * forward-Euler epilogue. Restates P6-P11 of FlowFV::compute_residual (reference:
 * src/spatial/flow_spatial.cpp:722-812, compute_fluxes :489-563, compute_max_timestep :567-634)
 * and, with the EP_STEP epilogue, the update + norm of SteadyForwardEulerSolver::solve
 * (src/ode/aodesolver.cpp:204-223).
 *
 * One CTA per tile of TC consecutive cells. The tile's face stream (every face touching one of the
 * tile's cells; faces cut by a tile boundary appear in both tiles and are evaluated identically in
 * both) is sorted by colour; no two faces of one colour share a tile cell, so after each colour
 * round a __syncthreads() is all the ordering the shared-memory accumulation needs. No atomics,
 * and the summation order per cell (colour order) is fixed => bitwise reproducible.
 */
#pragma once
#include "engine.hpp"

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

/// u_face = u_cell + g . (gr - rc) for the four primitive variables (reconstruction_utils.hpp:17-32)
__device__ __forceinline__ void extrapolate4(const double pc[4], const double *g /*8, global*/,
                                             double dx, double dy, double pf[4]) {
	double ga[4], gb[4];
	ld4(g, ga); ld4(g+4, gb);
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

template <int FLUX, int RECON, int VISC>
__global__ void __launch_bounds__(FACE_BLOCK, FVG_FACE_MINB)
face_kernel(const FaceArgs A)
{
	extern __shared__ double sm[];
	const DMesh &M = A.m;
	const int TC = M.TC;
	double *const res_s = sm;             // [4][TC]
	double *const integ_s = sm + 4*TC;    // [TC]
	__shared__ int coloff[MAXCOL+1];
	__shared__ double red_s[FACE_BLOCK/32];

	const int t = blockIdx.x, tid = threadIdx.x;
	const int c0 = t*TC;
	const int nc = min(TC, M.ncell - c0);
	for(int k = tid; k < 5*TC; k += FACE_BLOCK) sm[k] = 0.0;
	if(tid <= MAXCOL) coloff[tid] = M.fcoloff[t*(MAXCOL+1) + tid];
	__syncthreads();
	const int e0 = coloff[0], e1 = coloff[MAXCOL];

	for(int base = e0; base < e1; base += FACE_BLOCK) {
		const int e = base + tid;
		const bool valid = e < e1;
		double f[4] = {0,0,0,0};
		double sri = 0, srj = 0;
		int L = 0, R = -1;
		if(valid) {
			L = M.fL[e]; R = M.fR[e];
			const double2 nrm = M.fn[e];
			const double len = M.flen[e];
			const double nx = nrm.x, ny = nrm.y;
			const bool bnd = R < 0;
			const int b = -2 - R;
			double ucl[4], ucr[4];      // cell states (right = ghost of the cell state on a boundary)
			double ul[4], ur[4];        // face states
			ld4(A.u + 4*(size_t)L, ucl);
			if(RECON == FR_FIRST) {
				for(int k = 0; k < 4; k++) ul[k] = ucl[k];
				if(bnd) ghost_state(A.gas, A.gas.bc[A.bbc[b]], ul, nx, ny, ur);
				else ld4(A.u + 4*(size_t)R, ur);
				if(VISC != VISC_NONE) for(int k = 0; k < 4; k++) ucr[k] = ur[k];
			}
			else {
				const double2 gr = M.fgr[e];
				const double2 rl = M.rc[L];
				double pl[4], pfl[4], pfr[4];
				cons2prim(A.gas, ucl, pl);
				if(RECON == FR_LINEAR) {
					extrapolate4(pl, A.lg + 8*(size_t)L, gr.x - rl.x, gr.y - rl.y, pfl);
					prim2cons(A.gas, pfl, ul);
					if(bnd) {
						ghost_state(A.gas, A.gas.bc[A.bbc[b]], ul, nx, ny, ur);
						if(VISC != VISC_NONE) ghost_state(A.gas, A.gas.bc[A.bbc[b]], ucl, nx, ny, ucr);
					} else {
						const double2 rr = M.rc[R];
						double pr[4];
						ld4(A.u + 4*(size_t)R, ucr);
						cons2prim(A.gas, ucr, pr);
						extrapolate4(pr, A.lg + 8*(size_t)R, gr.x - rr.x, gr.y - rr.y, pfr);
						prim2cons(A.gas, pfr, ur);
					}
				}
				else { // MUSCL with Van Albada limiter
					double pr[4];
					double2 rr;
					if(bnd) {
						ghost_state(A.gas, A.gas.bc[A.bbc[b]], ucl, nx, ny, ucr);
						rr = M.rcbp[b];
					} else {
						ld4(A.u + 4*(size_t)R, ucr);
						rr = M.rc[R];
					}
					cons2prim(A.gas, ucr, pr);
					const double dx = rr.x - rl.x, dy = rr.y - rl.y;
					double ga[4], gb[4];
					ld4(A.gu + 8*(size_t)L, ga); ld4(A.gu + 8*(size_t)L + 4, gb);
					const double gLx[4] = {ga[0], ga[2], gb[0], gb[2]}, gLy[4] = {ga[1], ga[3], gb[1], gb[3]};
					for(int k = 0; k < 4; k++) {
						const double dlr = pr[k] - pl[k];
						const double dm = 2.0*(gLx[k]*dx + gLy[k]*dy) - dlr;
						pfl[k] = pl[k] + muscl_term(dm, dlr);
					}
					prim2cons(A.gas, pfl, ul);
					if(bnd) ghost_state(A.gas, A.gas.bc[A.bbc[b]], ul, nx, ny, ur);
					else {
						ld4(A.gu + 8*(size_t)R, ga); ld4(A.gu + 8*(size_t)R + 4, gb);
						const double gRx[4] = {ga[0], ga[2], gb[0], gb[2]}, gRy[4] = {ga[1], ga[3], gb[1], gb[3]};
						for(int k = 0; k < 4; k++) {
							const double dlr = pr[k] - pl[k];
							const double dp = 2.0*(gRx[k]*dx + gRy[k]*dy) - dlr;
							pfr[k] = pr[k] - muscl_term(dp, dlr);
						}
						prim2cons(A.gas, pfr, ur);
					}
				}
			}

			const Side a = load_side<true>(A.gas, ul, nx, ny);
			const Side bs = load_side<true>(A.gas, ur, nx, ny);
			flux_from_sides<FLUX>(A.gas, a, bs, nx, ny, f);
			for(int k = 0; k < 4; k++) f[k] *= len;
			sri = (fabs(a.vn) + a.c)*len;
			srj = (fabs(bs.vn) + bs.c)*len;

			if(VISC != VISC_NONE) {
				const double2 rl = M.rc[L];
				const double2 rr = bnd ? M.rcbp[b] : M.rc[R];
				double gl[8], grr[8], vf[4];
				if(RECON != FR_FIRST) {
					ld4(A.gu + 8*(size_t)L, gl); ld4(A.gu + 8*(size_t)L + 4, gl+4);
					if(bnd) for(int k = 0; k < 8; k++) grr[k] = gl[k];
					else { ld4(A.gu + 8*(size_t)R, grr); ld4(A.gu + 8*(size_t)R + 4, grr+4); }
				}
				viscous_face_flux<RECON != FR_FIRST, VISC == VISC_CONST>(A.gas, nx, ny, rl.x, rl.y, rr.x, rr.y,
					ucl, ucr, gl, grr, ul, ur, vf);
				for(int k = 0; k < 4; k++) f[k] += vf[k]*len;
				const double mui = VISC == VISC_CONST ? 1.0/A.gas.Reinf : viscosity_cons(A.gas, ul);
				const double muj = VISC == VISC_CONST ? 1.0/A.gas.Reinf : viscosity_cons(A.gas, ur);
				const double coi = fmax(4.0/(3.0*ul[0]), A.gas.g/ul[0]);
				const double coj = fmax(4.0/(3.0*ur[0]), A.gas.g/ur[0]);
				sri += coi*mui/A.gas.Pr*len*len/M.area[L];
				if(!bnd) srj += coj*muj/A.gas.Pr*len*len/M.area[R];
			}
		}

		// colour rounds of this chunk
		int myc = 0, clo = 0, chi = 0;
		{
			const int last = min(base + FACE_BLOCK, e1) - 1;
			#pragma unroll
			for(int c = 1; c < MAXCOL; c++) {
				if(e >= coloff[c]) myc = c;
				if(base >= coloff[c]) clo = c;
				if(last >= coloff[c]) chi = c;
			}
		}
		const int lL = L - c0, lR = R - c0;
		const bool inL = valid && (unsigned)lL < (unsigned)nc;
		const bool inR = valid && R >= 0 && (unsigned)lR < (unsigned)nc;
		for(int c = clo; c <= chi; c++) {
			if(myc == c) {
				if(inL) {
					res_s[lL] -= f[0]; res_s[TC+lL] -= f[1]; res_s[2*TC+lL] -= f[2]; res_s[3*TC+lL] -= f[3];
					integ_s[lL] += sri;
				}
				if(inR) {
					res_s[lR] += f[0]; res_s[TC+lR] += f[1]; res_s[2*TC+lR] += f[2]; res_s[3*TC+lR] += f[3];
					integ_s[lR] += srj;
				}
			}
			__syncthreads();
		}
	}

	// epilogue: one thread per tile cell
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
			if(A.gettimesteps) A.dtm[c] = M.area[c]/integ_s[k];
		}
	}
	else {
		double part = 0.0;
		for(int k = tid; k < nc; k += FACE_BLOCK) {
			const size_t c = (size_t)(c0 + k);
			const double ar = M.area[c];
			const double dt = ar/integ_s[k];
			const double fac = A.cfl*dt/ar;
			double uo[4];
			ld4(A.u + 4*c, uo);
			const double rE = res_s[3*TC+k];
			uo[0] += fac*res_s[k]; uo[1] += fac*res_s[TC+k]; uo[2] += fac*res_s[2*TC+k]; uo[3] += fac*rE;
			st4(A.unew + 4*c, uo);
			part += rE*rE*ar;
		}
		// fixed-order block reduction: warp shuffle tree, then warp 0 sums the warp partials in order
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
	const size_t smem = (size_t)5*a.m.TC*sizeof(double);
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
