/* ode_b200.hpp - the second, optional file of the reference-side binding (next to flow_spatial_b200.hpp; it would live
 * in FVENS as src/ode/ode_b200.hpp): the explicit pseudo-time driver with the state resident on the GPU.
 * SteadyForwardEulerSolver_B200 is a SteadySolver<NVARS> of the reference (ode/aodesolver.hpp:70-100) with the public
 * behaviour of its SteadyForwardEulerSolver::solve (ode/aodesolver.cpp:136-282): same stopping rule, same TimingData and
 * convergence history, same exceptions - but one upload of u, the whole loop in fvg_forward_euler_solve (fused residual +
 * local time step + update + norm per step), one download. Written against FVENS's headers; compiled against them by
 * oracle/ref_binding.cpp.
 */
#ifndef FVENS_ODE_B200_H
#define FVENS_ODE_B200_H

#include <chrono>
#include <cmath>
#include <iostream>
#include "ode/aodesolver.hpp"
#include "utilities/aerrorhandling.hpp"
#include "flow_spatial_b200.hpp"

namespace fvens {

class SteadyForwardEulerSolver_B200 : public SteadySolver<NVARS>
{
public:
	SteadyForwardEulerSolver_B200(const Spatial<freal,NVARS> *const euler, const Vec, const SteadySolverConfig& conf)
		: SteadySolver<NVARS>(euler, conf), engine(dynamic_cast<const B200Engine*>(euler))
	{
		if(!engine) throw std::runtime_error("SteadyForwardEulerSolver_B200 needs a FlowFV_B200 spatial discretization");
	}

	StatusCode solve(Vec u)
	{
		const UMesh<freal,NDIM> *const m = space->mesh();
		tdata.nelem = m->gnelem();
		if(config.maxiter <= 0) {
			std::cout << " SteadyForwardEulerSolver: solve(): No iterations to be done.\n";
			return 0;
		}
		const auto t0 = std::chrono::steady_clock::now();
		const unsigned long long bytes = (unsigned long long)m->gnelem()*NVARS*sizeof(double);
		std::vector<double> hist((size_t)config.maxiter, 0.0);
		int steps = 0, code;
		{
			MutableGhostedVecHandler<PetscScalar> uh(u);
			void *d_u = nullptr;
			if(fvg_malloc(&d_u, bytes)) throw std::runtime_error(fvg_last_error());
			code = fvg_memcpy(d_u, uh.getArray(), bytes, 0);
			// the reference applies cflinit on every step (aodesolver.cpp:208); norm read back every step, as its Allreduce
			if(code == 0) code = fvg_forward_euler_solve(engine->engine_flow(), static_cast<double*>(d_u), config.cflinit, config.tol,
			                                             config.maxiter, 1, &steps, hist.data());
			if(code == FVG_OK || code == FVG_ERR_TOLERANCE || code == FVG_ERR_NUMERICAL) {
				const int rc = fvg_memcpy(uh.getArray(), d_u, bytes, 1);
				if(rc) code = rc;
			}
			fvg_free(d_u);
		}
		const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		tdata.ode_walltime += wall; tdata.num_timesteps = steps;
		const double initres = steps > 0 ? hist[0] : 1.0;
		for(int s = 0; s < steps; s++) {
			const SteadyStepMonitor mon { s+1, (float)(hist[s]/initres), (float)hist[s], (float)(wall*(s+1)/steps), 0.0f, 0,
			                              (float)config.cflinit };
			tdata.convhis.push_back(mon);
		}
		if(code == FVG_ERR_NUMERICAL) throw Numerical_error("Steady forward Euler diverged - residual is Nan or inf!");
		if(code == FVG_ERR_TOLERANCE) {
			tdata.converged = false;
			throw Tolerance_error("Steady forward Euler did not converge to specified tolerance!");
		}
		if(code != FVG_OK) throw std::runtime_error(fvg_last_error());
		tdata.converged = true;
		return 0;
	}

private:
	using SteadySolver<NVARS>::space;
	using SteadySolver<NVARS>::config;
	using SteadySolver<NVARS>::tdata;
	const B200Engine *const engine;
};

}
#endif
