/** Explicit multi-stage (strong-stability-preserving Runge-Kutta) time stepping on top of fvg_residual.
 *
 * The reference's TVDRKSolver::solve (ode/aodesolver.cpp:672-785) carries the Shu-Osher coefficient table
 * (initialize_TVDRK_Coeffs, :45-67) but (1) evaluates the residual at the step's initial state in every stage and
 * (2) subtracts dt/area * residual although compute_residual leaves -r(u) there (the forward-Euler loop adds it,
 * :208). SURVEY 8f-4 asks for the scheme as it is meant:
 *     u^(0) = u^n;  u^(i+1) = a_i u^n + b_i u^(i) + c_i dt/area * (-r(u^(i)));  u^(n+1) = u^(order)
 * with the reference's table (a, b, c), its global time step dt = cfl * min_cells dtm(u^n) taken in the first stage,
 * its loop condition (time <= finaltime - 1e-12, last step not clipped) and its divergence check on dt.
 *
 * Everything stays on the device; the host reads one double (dt) per step, as it must for the loop condition.
 */
#include "engine.hpp"
#include <cmath>
#include <vector>

namespace fvg {

__device__ __forceinline__ double nan_min(double a, double b)
{
	// fmin drops NaNs; the solver has to see them
	return (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000ll) : fmin(a, b);
}

__global__ void min_partial_kernel(const double *__restrict__ x, long long n, double *__restrict__ partial)
{
	__shared__ double s[256];
	double acc = __longlong_as_double(0x7ff0000000000000ll);   // +inf
	for(long long i = (long long)blockIdx.x*256 + threadIdx.x; i < n; i += (long long)gridDim.x*256) acc = nan_min(acc, x[i]);
	s[threadIdx.x] = acc;
	__syncthreads();
	for(int o = 128; o > 0; o >>= 1) {
		if(threadIdx.x < o) s[threadIdx.x] = nan_min(s[threadIdx.x], s[threadIdx.x + o]);
		__syncthreads();
	}
	if(threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

/// out = a u0 + b us + c * cfl * dtmin / area * res   (rows of `nvars`; us and out may alias)
__global__ void rk_stage_kernel(const double *__restrict__ u0, const double *us, const double *__restrict__ res,
                                const double *__restrict__ area, const double *__restrict__ dtmin,
                                double a, double b, double ccfl, int ncell, int nvars, double *out)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= (long long)ncell*nvars) return;
	const double fac = ccfl*dtmin[0]/area[i/nvars];
	out[i] = a*u0[i] + b*us[i] + fac*res[i];
}

static int launch_min(const double *x, long long n, double *partial, int nblk, double *out, cudaStream_t s)
{
	min_partial_kernel<<<nblk, 256, 0, s>>>(x, n, partial);
	cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "min_partial_kernel launch", __FILE__, __LINE__);
	min_partial_kernel<<<1, 256, 0, s>>>(partial, nblk, out);
	e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "min_partial_kernel launch", __FILE__, __LINE__);
	return 0;
}

static int launch_rk_stage(const double *u0, const double *us, const double *res, const double *area, const double *dtmin,
                           double a, double b, double ccfl, int ncell, int nvars, double *out, cudaStream_t s)
{
	const long long n = (long long)ncell*nvars;
	if(n == 0) return 0;
	rk_stage_kernel<<<(unsigned)((n + 255)/256), 256, 0, s>>>(u0, us, res, area, dtmin, a, b, ccfl, ncell, nvars, out);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "rk_stage_kernel launch", __FILE__, __LINE__);
	return 0;
}

} // namespace fvg

using namespace fvg;

extern "C" int fvg_tvdrk_coefficients(int order, double *h_coeffs)
{
	// initialize_TVDRK_Coeffs, ode/aodesolver.cpp:45-67 (rows: stage; columns: weight of u^n, of the stage state, of the update)
	static const double c1[3] = {1.0, 0.0, 1.0};
	static const double c2[6] = {1.0, 0.0, 1.0,   0.5, 0.5, 0.5};
	static const double c3[9] = {1.0, 0.0, 1.0,   0.75, 0.25, 0.25,   0.3333333333333333, 0.6666666666666667, 0.6666666666666667};
	if(!h_coeffs || order < 1 || order > 3) { set_error("fvg_tvdrk_coefficients: temporal order " + std::to_string(order) + " not available"); return FVG_ERR_INVALID; }
	const double *c = order == 1 ? c1 : order == 2 ? c2 : c3;
	for(int k = 0; k < 3*order; k++) h_coeffs[k] = c[k];
	return 0;
}

extern "C" int fvg_tvdrk_solve(fvg_flow *f, double *d_u, int order, double cfl, double finaltime, int maxsteps,
                               int *h_steps, double *h_time)
{
	if(!f || !d_u || !h_steps || !h_time || !(cfl > 0.0)) { set_error("fvg_tvdrk_solve: bad argument"); return FVG_ERR_INVALID; }
	double coef[9];
	int rc = fvg_tvdrk_coefficients(order, coef);
	if(rc != 0) return rc;
	if(f->mesh->nranks > 1) { set_error("fvg_tvdrk_solve: single-process driver"); return FVG_ERR_UNSUPPORTED; }
	*h_steps = 0; *h_time = 0.0;
	FVG_CUDA(cudaSetDevice(f->mesh->device));
	cudaStream_t s = nullptr;
	const DMesh &D = f->mesh->d;
	const int n = D.ncell;
	const int nblk = 1024;
	// scratch arrays are owned by the flow (allocated on first use, freed with it), as for the other drivers:
	// stage state [4n], residual [4n], local steps [n], caller-ordered areas [n], reduction blocks, the global step
	if(!f->d_rk && (rc = flow_dev_alloc(f, &f->d_rk, 10*(size_t)n + (size_t)nblk + 1)) != 0) return rc;
	double *const us = f->d_rk, *const res = us + 4*(size_t)n, *const dtm = res + 4*(size_t)n, *const area_own = dtm + n,
	       *const part = area_own + n, *const dtmin = part + nblk;
	// cell areas in the caller's cell order
	const double *area = D.area;
	if(!f->mesh->identity_perm) {
		if((rc = launch_permute_rows(D.area, area_own, D.new2old, n, 1, false, false, s)) != 0) return rc;
		f->launches++;
		area = area_own;
	}
	FVG_CUDA(cudaMemcpyAsync(us, d_u, 4*(size_t)n*sizeof(double), cudaMemcpyDeviceToDevice, s));

	int step = 0, status = FVG_OK;
	double time = 0.0;
	while(time <= finaltime - 1e-12 && (maxsteps <= 0 || step < maxsteps)) {
		for(int istage = 0; istage < order; istage++) {
			if((rc = fvg_residual(f, us, res, 0, istage == 0, dtm, s)) != 0) return rc;
			if(istage == 0) {
				if((rc = launch_min(dtm, n, part, nblk, dtmin, s)) != 0) return rc;
				f->launches += 2;
			}
			if((rc = launch_rk_stage(d_u, us, res, area, dtmin, coef[3*istage], coef[3*istage+1], coef[3*istage+2]*cfl, n, 4, us, s)) != 0) return rc;
			f->launches++;
		}
		// the host needs the step for the time and the stopping rule, as in the reference (aodesolver.cpp:719-726): one read
		// per step through the flow's pinned scalar
		FVG_CUDA(cudaMemcpyAsync(f->h_norm, dtmin, sizeof(double), cudaMemcpyDeviceToHost, s));
		FVG_CUDA(cudaStreamSynchronize(s));
		const double h_dtmin = *f->h_norm;
		if(!std::isfinite(h_dtmin)) { status = FVG_ERR_NUMERICAL; break; }      // the state of the last good step is kept
		FVG_CUDA(cudaMemcpyAsync(d_u, us, 4*(size_t)n*sizeof(double), cudaMemcpyDeviceToDevice, s));
		step++;
		time += h_dtmin*cfl;
	}
	FVG_CUDA(cudaStreamSynchronize(s));
	*h_steps = step; *h_time = time;
	if(status == FVG_ERR_NUMERICAL) set_error("TVDRK solver diverged - dtmin is Nan or inf!");
	return status;
}
