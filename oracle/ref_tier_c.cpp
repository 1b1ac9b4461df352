/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Tier-C reference build: the reference's own FlowFV::compute_residual with everything it calls - gas physics,
 * numerical fluxes, boundary conditions, viscous flux (the tier-A sources), gradient schemes and reconstructions (the
 * tier-B sources), Spatial geometry and the modified-average face gradient (spatial/aspatial.cpp), the trace vector
 * (linalg/tracevector.cpp), the Vec handlers (linalg/petscutils.cpp) and spatial/flow_spatial.cpp itself - compiled
 * UNMODIFIED, in place from /root/reference/src, against the stand-ins of ref_shim_b/ (Eigen-lite, serial PETSc Vec,
 * single-process MPI, boost::bimap for the BC name table, a UMesh with the reference's accessors over the harness's
 * arrays, the factory declarations). This file adds the three factory definitions (same keys -> same reference
 * classes as utilities/afactory.cpp:30-217, whose own file needs the ODE and Boost control-file layers) and a C
 * interface; it contains no reference code. Built twice: serially (no -fopenmp) into oracle/_ref/libfvens_ref_c.so for
 * the parity tests (deterministic), and with the reference's own OpenMP pragmas enabled into
 * oracle/_ref/libfvens_ref_c_omp.so, which is what `bench.py --impl reference` times on the host cores.
 */
#include "ref_sources_spatial.hpp"

using namespace fvens;

extern "C" {

/** FlowFV<freal,order2,constVisc>::compute_residual(u, r, gettimesteps, dtm) of the reference on a serial mesh.
 * sizes = {npoin, nelem, nbface, naface, maxnnode}; btags [nbface]; phys = {gamma, Minf, Tinf, Reinf, Pr, aoa};
 * keys: flux / gradient / reconstruction names as in the control files (upper case); bcs: nbc x {tag, type} and
 * nbc x 2 values; u [nelem][4] conserved -> res [nelem][4] (the reference ADDS -r(u) into a zeroed vector), dtm [nelem].
 * Returns the reference's status code. */
/// Persistent variant for timing: the reference's FlowFV object, its mesh view and its Vecs are built once.
struct RefFlow {
	UMesh<freal,NDIM> m;
	std::vector<double> coords, facemetric, area;
	std::vector<int> nnode, inpoel, esuel, elemface, intfac, btags;
	std::unique_ptr<const Spatial<freal,NVARS>> prob;
	_p_Vec uv, rv, dv;
};

void* ref_flow_create(const int *sizes, const double *coords, const int *nnode, const int *inpoel, const int *esuel,
                      const int *elemface, const int *intfac, const double *facemetric, const double *area, const int *btags,
                      const double *phys, const char *flux, const char *gradient, const char *recon, double limiter_param,
                      int order2, int viscous, int const_visc, int nbc, const int *bc_tag_type, const double *bc_vals)
{
	RefFlow *h = new RefFlow;
	const size_t np = sizes[0], ne = sizes[1], nb = sizes[2], nf = sizes[3], mw = sizes[4];
	h->coords.assign(coords, coords + 2*np); h->nnode.assign(nnode, nnode + ne); h->inpoel.assign(inpoel, inpoel + ne*mw);
	h->esuel.assign(esuel, esuel + ne*mw); h->elemface.assign(elemface, elemface + ne*mw); h->intfac.assign(intfac, intfac + 4*nf);
	h->facemetric.assign(facemetric, facemetric + 3*nf); h->area.assign(area, area + ne); h->btags.assign(btags, btags + nb);
	UMesh<freal,NDIM> &m = h->m;
	m.npoin = (fint)np; m.nelem = (fint)ne; m.nbface = (fint)nb; m.naface = (fint)nf; m.maxnnode = (int)mw;
	m.coords = h->coords.data(); m.nnode = h->nnode.data(); m.inpoel = h->inpoel.data(); m.esuel = h->esuel.data();
	m.elemface = h->elemface.data(); m.intfac = h->intfac.data(); m.facemetric = h->facemetric.data(); m.area = h->area.data();
	m.btags = h->btags.data();
	std::vector<FlowBCConfig> bcs;
	for(int i = 0; i < nbc; i++) {
		FlowBCConfig c;
		c.bc_tag = bc_tag_type[2*i]; c.bc_type = static_cast<BCType>(bc_tag_type[2*i+1]);
		c.bc_vals = {bc_vals[2*i], bc_vals[2*i+1]};
		bcs.push_back(c);
	}
	const FlowPhysicsConfig pconf { phys[0], phys[1], phys[2], phys[3], phys[4], phys[5], viscous != 0, const_visc != 0, bcs };
	const FlowNumericsConfig nconf { flux, flux, gradient, recon, limiter_param, order2 != 0 };
	if(order2) { if(const_visc) h->prob.reset(new FlowFV<freal,true,true>(&m, pconf, nconf)); else h->prob.reset(new FlowFV<freal,true,false>(&m, pconf, nconf)); }
	else { if(const_visc) h->prob.reset(new FlowFV<freal,false,true>(&m, pconf, nconf)); else h->prob.reset(new FlowFV<freal,false,false>(&m, pconf, nconf)); }
	h->uv.a.assign(ne*NVARS, 0.0); h->uv.nlocal = (PetscInt)(ne*NVARS); h->uv.nghost = 0;
	h->rv = h->uv;
	h->dv.a.assign(ne, 0.0); h->dv.nlocal = (PetscInt)ne; h->dv.nghost = 0;
	return h;
}

/// One evaluation as the reference's explicit solver makes it (ode/aodesolver.cpp:177-190): zero the residual Vec,
/// compute_residual(u, r, gettimesteps, dtm). res / dtm may be NULL (timing only).
int ref_flow_residual(void *hv, const double *u, int gettimesteps, double *res, double *dtm)
{
	RefFlow *h = static_cast<RefFlow*>(hv);
	std::copy(u, u + h->uv.a.size(), h->uv.a.begin());
	std::fill(h->rv.a.begin(), h->rv.a.end(), 0.0);
	const int ierr = h->prob->compute_residual(&h->uv, &h->rv, gettimesteps != 0, &h->dv);
	if(res) std::copy(h->rv.a.begin(), h->rv.a.end(), res);
	if(dtm && gettimesteps) std::copy(h->dv.a.begin(), h->dv.a.end(), dtm);
	return ierr;
}

void ref_flow_destroy(void *hv) { delete static_cast<RefFlow*>(hv); }

int ref_num_threads()
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

int ref_residual(const int *sizes, const double *coords, const int *nnode, const int *inpoel, const int *esuel,
                 const int *elemface, const int *intfac, const double *facemetric, const double *area, const int *btags,
                 const double *phys, const char *flux, const char *gradient, const char *recon, double limiter_param,
                 int order2, int viscous, int const_visc, int nbc, const int *bc_tag_type, const double *bc_vals,
                 const double *u, int gettimesteps, double *res, double *dtm)
{
	UMesh<freal,NDIM> m;
	m.npoin = sizes[0]; m.nelem = sizes[1]; m.nbface = sizes[2]; m.naface = sizes[3]; m.maxnnode = sizes[4];
	m.coords = coords; m.nnode = nnode; m.inpoel = inpoel; m.esuel = esuel; m.elemface = elemface; m.intfac = intfac;
	m.facemetric = facemetric; m.area = area; m.btags = btags;
	std::vector<FlowBCConfig> bcs;
	for(int i = 0; i < nbc; i++) {
		FlowBCConfig c;
		c.bc_tag = bc_tag_type[2*i]; c.bc_type = static_cast<BCType>(bc_tag_type[2*i+1]);
		c.bc_vals = {bc_vals[2*i], bc_vals[2*i+1]};
		bcs.push_back(c);
	}
	const FlowPhysicsConfig pconf { phys[0], phys[1], phys[2], phys[3], phys[4], phys[5], viscous != 0, const_visc != 0, bcs };
	const FlowNumericsConfig nconf { flux, flux, gradient, recon, limiter_param, order2 != 0 };
	std::unique_ptr<const Spatial<freal,NVARS>> prob;
	if(order2) { if(const_visc) prob.reset(new FlowFV<freal,true,true>(&m, pconf, nconf)); else prob.reset(new FlowFV<freal,true,false>(&m, pconf, nconf)); }
	else { if(const_visc) prob.reset(new FlowFV<freal,false,true>(&m, pconf, nconf)); else prob.reset(new FlowFV<freal,false,false>(&m, pconf, nconf)); }
	_p_Vec uv, rv, dv;
	uv.a.assign(u, u + (size_t)m.nelem*NVARS); uv.nlocal = m.nelem*NVARS; uv.nghost = 0;
	rv.a.assign((size_t)m.nelem*NVARS, 0.0); rv.nlocal = m.nelem*NVARS; rv.nghost = 0;
	dv.a.assign((size_t)m.nelem, 0.0); dv.nlocal = m.nelem; dv.nghost = 0;
	const int ierr = prob->compute_residual(&uv, &rv, gettimesteps != 0, &dv);
	for(size_t i = 0; i < rv.a.size(); i++) res[i] = rv.a[i];
	if(gettimesteps) for(size_t i = 0; i < dv.a.size(); i++) dtm[i] = dv.a[i];
	return ierr;
}

}
