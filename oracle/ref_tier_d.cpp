/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Tier-D reference build: tier C (oracle/ref_tier_c.cpp: the reference's FlowFV::compute_residual and everything
 * under it) plus the reference's own explicit pseudo-time driver, SteadyForwardEulerSolver::solve
 * (ode/aodesolver.cpp), and its output translation unit (spatial/aoutput.cpp: convergence-history format, entropy
 * norm, point data), compiled UNMODIFIED, in place from /root/reference/src against ref_shim_b/. The implicit solvers
 * in the same file compile against inert PETSc KSP/Mat stand-ins and are never called. This file adds
 * createSystemVector (linalg/alinalg.cpp:9-15 on the serial Vec stand-in) and a C interface; no reference code.
 * Built serially into oracle/_ref/libfvens_ref_d.so.
 */
#include "ref_tier_c.cpp"
#include "spatial/aoutput.cpp"
#include "ode/aodesolver.cpp"
#include <sstream>

namespace fvens {
StatusCode createSystemVector(const UMesh<freal,NDIM> *const m, const int nvars, Vec *const v)
{
	*v = new _p_Vec;
	(*v)->nlocal = m->gnelem()*nvars; (*v)->nghost = 0;
	(*v)->a.assign((size_t)(*v)->nlocal, 0.0);
	return 0;
}
}

extern "C" {

/** SteadyForwardEulerSolver<NVARS>(space, u, {lognres = true, ..., cflinit = cflfin = cfl, tol, maxiter}).solve(u) of
 * the reference on the flow of ref_flow_create. u [nelem][4] in/out; hist_rel, hist_abs [maxiter] (the solver stores
 * them as float); returns 0 converged, 1 Tolerance_error (max iterations), 2 Numerical_error, else the status code. */
int ref_flow_forward_euler(void *hv, double cfl, double tol, int maxiter, double *u, int *steps, double *hist_rel, double *hist_abs)
{
	RefFlow *h = static_cast<RefFlow*>(hv);
	_p_Vec uv;
	uv.a.assign(u, u + h->uv.a.size()); uv.nlocal = h->uv.nlocal; uv.nghost = 0;
	const SteadySolverConfig conf { true, "ref-tier-d", false, cfl, cfl, 0, 0, tol, maxiter, 0, 0 };
	SteadyForwardEulerSolver<NVARS> solver(h->prob.get(), &uv, conf);
	int code = 0;
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());        // the solver logs every 50th step to stdout
	try { code = solver.solve(&uv); }
	catch(Tolerance_error&) { code = 1; }
	catch(Numerical_error&) { code = 2; }
	std::cout.rdbuf(old);
	const TimingData td = solver.getTimingData();
	*steps = td.num_timesteps;
	for(size_t i = 0; i < td.convhis.size() && (int)i < maxiter; i++) { hist_rel[i] = td.convhis[i].rmsres; hist_abs[i] = td.convhis[i].absrmsres; }
	std::copy(uv.a.begin(), uv.a.end(), u);
	return code;
}

/** TVDRKSolver<NVARS>(space, u, order, logfile, cfl).solve(finaltime) of the reference (ode/aodesolver.cpp:647-785), as
 * it is: used to document what that loop computes (every stage evaluated at the step's initial state, update
 * subtracted) next to the scheme the product implements. u [nelem][4] in/out. Returns 0, 2 on Numerical_error. */
int ref_flow_tvdrk(void *hv, int order, double cfl, double finaltime, const char *logfile, double *u)
{
	RefFlow *h = static_cast<RefFlow*>(hv);
	_p_Vec uv;
	uv.a.assign(u, u + h->uv.a.size()); uv.nlocal = h->uv.nlocal; uv.nghost = 0;
	int code = 0;
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	try {
		TVDRKSolver<NVARS> solver(h->prob.get(), &uv, order, logfile, cfl);
		code = solver.solve(finaltime);
	}
	catch(Numerical_error&) { code = 2; }
	std::cout.rdbuf(old);
	std::copy(uv.a.begin(), uv.a.end(), u);
	return code;
}

/// The reference's convergence-history writer (spatial/aoutput.cpp:617-636) into a string buffer
int ref_convergence_history_text(int nsteps, const int *step, const float *rel, const float *abs_, const float *wtime,
                                 const float *cfl, char *out, int outlen)
{
	std::stringstream ss;
	writeConvergenceHistoryHeader(ss);
	for(int i = 0; i < nsteps; i++) {
		const SteadyStepMonitor s { step[i], rel[i], abs_[i], wtime[i], 0.0f, 0, cfl[i] };
		writeStepToConvergenceHistory(s, ss);
	}
	const std::string t = ss.str();
	if((int)t.size() + 1 > outlen) return -1;
	std::copy(t.begin(), t.end(), out); out[t.size()] = 0;
	return (int)t.size();
}

}
