/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Tier-A reference build: the reference's own ens_gasdynamics sources (numerical fluxes, boundary
 * conditions, gas physics, viscous flux) compiled UNMODIFIED, in place from /root/reference, against
 * the Eigen forward-declaration stub in ref_shim/. This file only adds a C interface; it contains
 * no reference code. Built by oracle/Makefile into oracle/_ref/libfvens_ref_a.so (git-ignored).
 */
#include "physics/aphysics.cpp"
#include "physics/viscousphysics.cpp"
#include "spatial/anumericalflux.cpp"
#include "spatial/abc.cpp"
#include <memory>
#include <string>

using namespace fvens;

extern "C" {

/// flux ids as oracle/orc_physics.hpp FluxId
void ref_flux(int flux_id, const double *phys, int n, const double *ul, const double *ur,
              const double *nrm, double *out)
{
	const IdealGasPhysics<freal> p(phys[0], phys[1], phys[2], phys[3], phys[4]);
	std::unique_ptr<InviscidFlux<freal>> f;
	switch(flux_id) {
	case 0: f.reset(new LocalLaxFriedrichsFlux<freal>(&p)); break;
	case 1: f.reset(new VanLeerFlux<freal>(&p)); break;
	case 2: f.reset(new AUSMFlux<freal>(&p)); break;
	case 3: f.reset(new AUSMPlusFlux<freal>(&p)); break;
	case 4: f.reset(new RoeFlux<freal>(&p)); break;
	case 5: f.reset(new HLLFlux<freal>(&p)); break;
	default: f.reset(new HLLCFlux<freal>(&p)); break;
	}
	for(int i = 0; i < n; i++)
		f->get_flux(ul+4*i, ur+4*i, nrm+2*i, out+4*i);
}

void ref_ghost_state(int bc_type, const double *bc_vals, const double *phys, double aoa, int n,
                     const double *ins, const double *nrm, double *out)
{
	const IdealGasPhysics<freal> p(phys[0], phys[1], phys[2], phys[3], phys[4]);
	const std::array<freal,NVARS> uinf = p.compute_freestream_state(aoa);
	FlowBCConfig conf;
	conf.bc_tag = 0; conf.bc_type = static_cast<BCType>(bc_type);
	conf.bc_vals = {bc_vals[0], bc_vals[1]};
	std::vector<FlowBCConfig> confs{conf};
	std::map<int,const FlowBC<freal>*> bcs = create_const_flowBCs<freal>(confs, p, uinf);
	for(int i = 0; i < n; i++)
		bcs.at(0)->computeGhostState(ins+4*i, nrm+2*i, out+4*i);
	delete bcs.at(0);
}

/// Same steps as FlowFV::compute_viscous_flux (spatial/flow_spatial.cpp:349-395) except the face
/// gradient, which lives in Spatial (not part of ens_gasdynamics): the caller passes the face
/// gradient grad[2][4] directly. Checks computeViscousFlux + getPrimitive2StatesAndGradients.
void ref_viscous_flux_from_facegrad(const double *phys, int const_visc, int n, const double *nrm,
                                    const double *grad, const double *ul, const double *ur, double *out)
{
	const IdealGasPhysics<freal> p(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++) {
		freal g[NDIM][NVARS];
		for(int a = 0; a < NDIM; a++) for(int b = 0; b < NVARS; b++) g[a][b] = grad[8*i+a*NVARS+b];
		if(const_visc)
			computeViscousFlux<freal,NDIM,NVARS,true>(p, nrm+2*i, g, ul+4*i, ur+4*i, out+4*i);
		else
			computeViscousFlux<freal,NDIM,NVARS,false>(p, nrm+2*i, g, ul+4*i, ur+4*i, out+4*i);
	}
}

/// gradl/gradr: [dim][var] row-major, converted in place; uctl/uctr out
void ref_prim2_states_grads(const double *phys, int order2, int n, const double *ucl, const double *ucr,
                            double *gradl, double *gradr, double *uctl, double *uctr)
{
	const IdealGasPhysics<freal> p(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++) {
		if(order2)
			getPrimitive2StatesAndGradients<freal,NDIM,true>(p, ucl+4*i, ucr+4*i, gradl+8*i, gradr+8*i,
			                                                 uctl+4*i, uctr+4*i, gradl+8*i, gradr+8*i);
		else
			getPrimitive2StatesAndGradients<freal,NDIM,false>(p, ucl+4*i, ucr+4*i, gradl+8*i, gradr+8*i,
			                                                  uctl+4*i, uctr+4*i, gradl+8*i, gradr+8*i);
	}
}

void ref_cons2prim(const double *phys, int n, const double *uc, double *up) {
	const IdealGasPhysics<freal> p(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++) p.getPrimitiveFromConserved(uc+4*i, up+4*i);
}
void ref_prim2cons(const double *phys, int n, const double *up, double *uc) {
	const IdealGasPhysics<freal> p(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++) p.getConservedFromPrimitive(up+4*i, uc+4*i);
}
void ref_freestream(const double *phys, double aoa, double *uinf) {
	const IdealGasPhysics<freal> p(phys[0], phys[1], phys[2], phys[3], phys[4]);
	const std::array<freal,NVARS> u = p.compute_freestream_state(aoa);
	for(int i = 0; i < NVARS; i++) uinf[i] = u[i];
}
/// {p, c, T, mu_sutherland, entropy} from conserved
void ref_scalars(const double *phys, int n, const double *uc, double *out) {
	const IdealGasPhysics<freal> p(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++) {
		out[5*i+0] = p.getPressureFromConserved(uc+4*i);
		out[5*i+1] = p.getSoundSpeedFromConserved(uc+4*i);
		out[5*i+2] = p.getTemperatureFromConserved(uc+4*i);
		out[5*i+3] = p.getViscosityCoeffFromConserved(uc+4*i);
		out[5*i+4] = p.getEntropyFromConserved(uc+4*i);
	}
}

}
