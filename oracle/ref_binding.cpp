/* ORACLE - TEST INFRASTRUCTURE ONLY.
 * Compile-and-run check of the reference-side binding (fvens_b200/host/reference_binding/flow_spatial_b200.hpp): the
 * binding is built here against the reference's own, unmodified headers and sources (everything of tier E) and linked
 * with libfvens_b200.so, which proves that the class in INTEGRATION.md is real code for the real FVENS. On a GPU box the
 * reference's own SteadyForwardEulerSolver then drives the CUDA residual through it (tests/test_post_r1_c_reference_binding.py).
 * Built into oracle/_ref/libfvens_ref_binding.so (needs /root/reference and the product library).
 */
#include "ref_tier_e.cpp"
#include "../fvens_b200/host/reference_binding/flow_spatial_b200.hpp"
#include "../fvens_b200/host/reference_binding/ode_b200.hpp"

extern "C" {

/// As ref_e_flow_create, but the Spatial object is the binding: FlowFV_B200 on the mesh of the case.
/// Returns 0, or 1 with the message in ref_binding_error().
static std::string binding_error;
const char* ref_binding_error() { return binding_error.c_str(); }

int ref_e_flow_create_b200(void *hv, const double *phys, const char *flux, const char *gradient, const char *recon, double limiter_param,
                           int order2, int viscous, int const_visc, int nbc, const int *bc_tag_type, const double *bc_vals)
{
	RefCase *h = static_cast<RefCase*>(hv);
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	int rc = 0;
	try {
		std::vector<FlowBCConfig> bcs;
		for(int i = 0; i < nbc; i++) {
			FlowBCConfig c;
			c.bc_tag = bc_tag_type[2*i]; c.bc_type = static_cast<BCType>(bc_tag_type[2*i+1]);
			c.bc_vals = {bc_vals[2*i], bc_vals[2*i+1]};
			bcs.push_back(c);
		}
		const FlowPhysicsConfig pconf { phys[0], phys[1], phys[2], phys[3], phys[4], phys[5], viscous != 0, const_visc != 0, bcs };
		const FlowNumericsConfig nconf { flux, flux, gradient, recon, limiter_param, order2 != 0 };
		h->prob.reset(create_flowSpatialDiscretization_b200(h->m.get(), pconf, nconf));
		const size_t ne = h->m->gnelem();
		h->uv.a.assign(ne*NVARS, 0.0); h->uv.nlocal = (PetscInt)(ne*NVARS); h->uv.nghost = 0;
		h->rv = h->uv;
		h->dv.a.assign(ne, 0.0); h->dv.nlocal = (PetscInt)ne; h->dv.nghost = 0;
	} catch(std::exception& e) { binding_error = e.what(); rc = 1; }
	std::cout.rdbuf(old);
	return rc;
}

/// As ref_e_flow_forward_euler, with the device-resident driver of the binding (ode_b200.hpp) in place of the
/// reference's SteadyForwardEulerSolver; the Spatial object must be the binding's (ref_e_flow_create_b200).
/// Returns 0 converged, 1 Tolerance_error, 2 Numerical_error, 3 other failure (message in ref_binding_error()).
int ref_e_flow_forward_euler_b200(void *hv, double cfl, double tol, int maxiter, double *u, int *steps, double *hist_rel, double *hist_abs)
{
	RefCase *h = static_cast<RefCase*>(hv);
	_p_Vec uv;
	uv.a.assign(u, u + h->uv.a.size()); uv.nlocal = h->uv.nlocal; uv.nghost = 0;
	const SteadySolverConfig conf { true, "ref-binding", false, cfl, cfl, 0, 0, tol, maxiter, 0, 0 };
	int code = 0;
	*steps = 0;
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	try {
		SteadyForwardEulerSolver_B200 solver(h->prob.get(), &uv, conf);
		try { code = solver.solve(&uv); }
		catch(Tolerance_error&) { code = 1; }
		catch(Numerical_error&) { code = 2; }
		const TimingData td = solver.getTimingData();
		*steps = td.num_timesteps;
		for(size_t i = 0; i < td.convhis.size() && (int)i < maxiter; i++) { hist_rel[i] = td.convhis[i].rmsres; hist_abs[i] = td.convhis[i].absrmsres; }
	} catch(std::exception& e) { binding_error = e.what(); code = 3; }
	std::cout.rdbuf(old);
	std::copy(uv.a.begin(), uv.a.end(), u);
	return code;
}

}
