/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_physics.hpp header).
 * C interface (for ctypes) over the CPU restatement. Built by oracle/Makefile into
 * oracle/liborc.so. Nothing in fvens_b200/ may load this library.
 */
#include "orc_spatial.hpp"
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
	omp_set_num_threads(n);
#else
	(void)n;
#endif
}

// ---- pointwise ---------------------------------------------------------------------------------

/// phys = {gamma, Minf, Tinf, Reinf, Pr}
void orc_flux(int flux_id, const double *phys, int n, const double *ul, const double *ur,
              const double *nrm, double *out)
{
	const Gas g(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++)
		inviscid_flux(flux_id, g, ul+4*i, ur+4*i, nrm+2*i, out+4*i);
}

void orc_ghost_state(int bc_type, const double *bc_vals, const double *phys, double aoa, int n,
                     const double *ins, const double *nrm, double *out)
{
	const Gas g(phys[0], phys[1], phys[2], phys[3], phys[4]);
	double uinf[4]; g.freestreamState(aoa, uinf);
	BC bc; bc.tag = 0; bc.type = bc_type; bc.vals[0] = bc_vals[0]; bc.vals[1] = bc_vals[1];
	for(int i = 0; i < n; i++)
		ghost_state(bc, g, uinf, ins+4*i, nrm+2*i, out+4*i);
}

/// Viscous flux through FlowFV::compute_viscous_flux glue. grads in GradBlock layout (8 per side).
void orc_viscous_flux(const double *phys, int order2, int const_visc, int n, const double *nrm,
                      const double *rcl, const double *rcr, const double *ucl, const double *ucr,
                      const double *gl, const double *gr, const double *ul, const double *ur,
                      double *out)
{
	const Gas g(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++)
		cell_viscous_flux(g, order2, const_visc, nrm+2*i, rcl+2*i, rcr+2*i, ucl+4*i, ucr+4*i,
		                  gl+8*i, gr+8*i, ul+4*i, ur+4*i, out+4*i);
}

void orc_cons2prim(const double *phys, int n, const double *uc, double *up) {
	const Gas g(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++) g.primitiveFromConserved(uc+4*i, up+4*i);
}
void orc_prim2cons(const double *phys, int n, const double *up, double *uc) {
	const Gas g(phys[0], phys[1], phys[2], phys[3], phys[4]);
	for(int i = 0; i < n; i++) g.conservedFromPrimitive(up+4*i, uc+4*i);
}
void orc_freestream(const double *phys, double aoa, double *uinf) {
	const Gas g(phys[0], phys[1], phys[2], phys[3], phys[4]);
	g.freestreamState(aoa, uinf);
}

// ---- mesh --------------------------------------------------------------------------------------

void *orc_mesh_read(const char *path)
{
	try {
		Mesh *m = new Mesh(read_mesh(path));
		finalize_mesh(*m);
		return m;
	} catch(std::exception &e) {
		std::fprintf(stderr, "orc_mesh_read: %s\n", e.what());
		return nullptr;
	}
}

/// inpoel [nelem][4] (-1 padded), bface [nbface][3] = n0, n1, tag
void *orc_mesh_from_arrays(int npoin, const double *coords, int nelem, const int *nnode,
                           const int *inpoel, int nbface, const int *bface)
{
	try {
		Mesh *m = new Mesh;
		m->npoin = npoin; m->nelem = nelem; m->nbface = nbface;
		m->coords.assign(coords, coords+2*(size_t)npoin);
		m->nnode.assign(nnode, nnode+nelem);
		m->inpoel.assign(inpoel, inpoel+4*(size_t)nelem);
		m->bface.assign(bface, bface+3*(size_t)nbface);
		finalize_mesh(*m);
		return m;
	} catch(std::exception &e) {
		std::fprintf(stderr, "orc_mesh_from_arrays: %s\n", e.what());
		return nullptr;
	}
}

void orc_mesh_free(void *mp) { delete static_cast<Mesh*>(mp); }

/// out = npoin, nelem, nbface, naface, ninface
void orc_mesh_sizes(const void *mp, int *out)
{
	const Mesh *m = static_cast<const Mesh*>(mp);
	out[0] = m->npoin; out[1] = m->nelem; out[2] = m->nbface; out[3] = m->naface; out[4] = m->ninface;
}

#define COPY(vec, dst) std::memcpy(dst, (vec).data(), sizeof((vec)[0])*(vec).size())
void orc_mesh_get(const void *mp, double *coords, int *inpoel, int *nnode, int *bface, int *esuel,
                  int *elemface, int *intfac, int *btags, double *facemetric, double *area)
{
	const Mesh *m = static_cast<const Mesh*>(mp);
	if(coords) COPY(m->coords, coords);
	if(inpoel) COPY(m->inpoel, inpoel);
	if(nnode) COPY(m->nnode, nnode);
	if(bface) COPY(m->bface, bface);
	if(esuel) COPY(m->esuel, esuel);
	if(elemface) COPY(m->elemface, elemface);
	if(intfac) COPY(m->intfac, intfac);
	if(btags) COPY(m->btags, btags);
	if(facemetric) COPY(m->facemetric, facemetric);
	if(area) COPY(m->area, area);
}

// ---- flow --------------------------------------------------------------------------------------

/// phys = {gamma, Minf, Tinf, Reinf, Pr, aoa}; iopts = {viscous, const_visc, flux, gradient, recon,
/// order2, bnd_policy}; bcs: nbc x {tag, type}, bcvals: nbc x 2
void *orc_flow_create(const void *mp, const double *phys, const int *iopts, double limiter_param,
                      int nbc, const int *bcs, const double *bcvals)
{
	PhysConf pc;
	pc.gamma = phys[0]; pc.Minf = phys[1]; pc.Tinf = phys[2]; pc.Reinf = phys[3]; pc.Pr = phys[4];
	pc.aoa = phys[5]; pc.viscous = iopts[0]; pc.const_visc = iopts[1];
	Numerics nc;
	nc.flux = iopts[2]; nc.gradient = iopts[3]; nc.recon = iopts[4]; nc.order2 = iopts[5];
	nc.bnd_policy = iopts[6]; nc.limiter_param = limiter_param;
	std::vector<BC> bl(nbc);
	for(int i = 0; i < nbc; i++) {
		bl[i].tag = bcs[2*i]; bl[i].type = bcs[2*i+1];
		bl[i].vals[0] = bcvals[2*i]; bl[i].vals[1] = bcvals[2*i+1];
	}
	return new Flow(static_cast<const Mesh*>(mp), pc, nc, bl);
}
void orc_flow_free(void *fp) { delete static_cast<Flow*>(fp); }

void orc_flow_geometry(const void *fp, double *rc, double *gr, double *rcbp, double *V, double *clength)
{
	const Flow *f = static_cast<const Flow*>(fp);
	if(rc) COPY(f->rc, rc);
	if(gr) COPY(f->gr, gr);
	if(rcbp) COPY(f->rcbp, rcbp);
	if(V && !f->V.empty()) COPY(f->V, V);
	if(clength && !f->clength.empty()) COPY(f->clength, clength);
}

int orc_flow_residual(const void *fp, const double *u, double *res, int gettimesteps, double *dtm,
                      double *grad_out, double *face_out)
{
	try {
		static_cast<const Flow*>(fp)->compute_residual(u, res, gettimesteps, dtm, grad_out, face_out);
	} catch(std::exception &e) {
		std::fprintf(stderr, "orc_flow_residual: %s\n", e.what());
		return 1;
	}
	return 0;
}

/// Plug-in level entry points (GradientScheme::compute_gradients, SolutionReconstruction::
/// compute_face_values) on primitive cell states u and primitive boundary ghost states ug
void orc_flow_gradients(const void *fp, const double *u, const double *ug, double *grad) {
	static_cast<const Flow*>(fp)->gradients(u, ug, grad);
}
void orc_flow_face_values(const void *fp, const double *u, const double *ug, const double *grad,
                          double *ufl, double *ufr) {
	static_cast<const Flow*>(fp)->face_values(u, ug, grad, ufl, ufr);
}
void orc_flow_boundary_states(const void *fp, const double *ins, double *gs) {
	static_cast<const Flow*>(fp)->boundary_states(ins, gs);
}
void orc_flow_get_gradients(const void *fp, const double *u, double *grads) {
	static_cast<const Flow*>(fp)->get_gradients(u, grads);
}
void orc_flow_surface_data(const void *fp, const double *u, const double *grads, int marker, double *out3) {
	static_cast<const Flow*>(fp)->surface_data(u, grads, marker, out3);
}
double orc_flow_entropy(const void *fp, const double *u) {
	return static_cast<const Flow*>(fp)->entropy_error(u);
}
int orc_forward_euler(const void *fp, double *u, double cfl, double tol, int maxiter, int *steps,
                      double *hist) {
	return forward_euler(*static_cast<const Flow*>(fp), u, cfl, tol, maxiter, steps, hist);
}

} // extern "C"
