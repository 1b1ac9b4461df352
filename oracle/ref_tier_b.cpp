/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Tier-B reference build: the reference's own gradient and reconstruction translation units
 *   spatial/agradientschemes.cpp        (ZeroGradients, GreenGaussGradients, WeightedLeastSquaresGradients)
 *   spatial/areconstruction.cpp         (LinearUnlimitedReconstruction)
 *   spatial/limitedlinearreconstruction.cpp (WENOReconstruction, BarthJespersenLimiter, VenkatakrishnanLimiter)
 *   spatial/musclreconstruction.cpp     (MUSCLVanAlbada)
 * compiled UNMODIFIED, in place from /root/reference/src, against the stand-ins in ref_shim_b/ (an Eigen-lite that
 * restates the few Eigen features these files use, a single-process mpi.h, and a UMesh with the reference's accessor
 * names over arrays the harness passes in). This file only adds a C interface; it contains no reference code.
 * Built by oracle/Makefile into oracle/_ref/libfvens_ref_b.so (git-ignored). The parity tests use it to check the
 * oracle's restatement of these loops (oracle/orc_spatial.hpp) against the reference's own object code, which pins
 * the limiters no reference test covers. Built single-threaded: the Green-Gauss boundary loop races
 * under OpenMP.
 */
#include "spatial/agradientschemes.cpp"
#include "spatial/areconstruction.cpp"
#include "spatial/limitedlinearreconstruction.cpp"
#include "spatial/musclreconstruction.cpp"
#include "utilities/aarray2d.cpp"
#include <memory>

using namespace fvens;

namespace {
UMesh<freal,NDIM> make_mesh(const int *sizes, const double *coords, const int *nnode, const int *inpoel, const int *esuel,
                            const int *elemface, const int *intfac, const double *facemetric, const double *area)
{
	UMesh<freal,NDIM> m;
	m.npoin = sizes[0]; m.nelem = sizes[1]; m.nbface = sizes[2]; m.naface = sizes[3]; m.maxnnode = sizes[4];
	m.coords = coords; m.nnode = nnode; m.inpoel = inpoel; m.esuel = esuel; m.elemface = elemface; m.intfac = intfac;
	m.facemetric = facemetric; m.area = area;
	return m;
}
}

extern "C" {

/// gradient: 0 zero, 1 Green-Gauss, 2 weighted least squares (ids of oracle/orc_spatial.hpp).
/// rc [nelem][2] cell centres, rcbp [nbface][2] ghost centres, u [nelem][4], ug [nbface][4] -> grad [nelem][8]
void ref_gradients(int gradient, const int *sizes, const double *coords, const int *nnode, const int *inpoel,
                   const int *esuel, const int *elemface, const int *intfac, const double *facemetric, const double *area,
                   const double *rc, const double *rcbp, const double *u, const double *ug, double *grad)
{
	const UMesh<freal,NDIM> m = make_mesh(sizes, coords, nnode, inpoel, esuel, elemface, intfac, facemetric, area);
	std::unique_ptr<GradientScheme<freal,NVARS>> g;
	if(gradient == 1) g.reset(new GreenGaussGradients<freal,NVARS>(&m, rc, rcbp));
	else if(gradient == 2) g.reset(new WeightedLeastSquaresGradients<freal,NVARS>(&m, rc, rcbp));
	else g.reset(new ZeroGradients<freal,NVARS>(&m, rc, rcbp));
	g->compute_gradients(amat::Array2dView<freal>(u, m.gnelem(), NVARS), amat::Array2dView<freal>(ug, m.gnbface(), NVARS), grad);
}

/// recon: 0 linear, 1 WENO, 2 Van Albada MUSCL, 3 Barth-Jespersen, 4 Venkatakrishnan (ids of oracle/orc_spatial.hpp).
/// gr [naface][2] face midpoints; ufl, ufr [naface][4] (in/out: entries the class does not write are left alone)
void ref_face_values(int recon, double param, const int *sizes, const double *coords, const int *nnode, const int *inpoel,
                     const int *esuel, const int *elemface, const int *intfac, const double *facemetric, const double *area,
                     const double *rc, const double *rcbp, const double *gr, const double *u, const double *ug,
                     const double *grad, double *ufl, double *ufr)
{
	const UMesh<freal,NDIM> m = make_mesh(sizes, coords, nnode, inpoel, esuel, elemface, intfac, facemetric, area);
	amat::Array2d<freal> gauss(m.gnaface(), NDIM);
	for(fint f = 0; f < m.gnaface(); f++) for(int d = 0; d < NDIM; d++) gauss(f,d) = gr[(size_t)f*NDIM+d];
	std::unique_ptr<SolutionReconstruction<freal,NVARS>> r;
	switch(recon) {
	case 1: r.reset(new WENOReconstruction<freal,NVARS>(&m, rc, rcbp, gauss, param)); break;
	case 2: r.reset(new MUSCLVanAlbada<freal,NVARS>(&m, rc, rcbp, gauss)); break;
	case 3: r.reset(new BarthJespersenLimiter<freal,NVARS>(&m, rc, rcbp, gauss)); break;
	case 4: r.reset(new VenkatakrishnanLimiter<freal,NVARS>(&m, rc, rcbp, gauss, param)); break;
	default: r.reset(new LinearUnlimitedReconstruction<freal,NVARS>(&m, rc, rcbp, gauss)); break;
	}
	// The limiters index the cell-state matrix with esuel, which is nelem + face for a boundary neighbour: one row
	// past the end per boundary face in the reference. The matrix handed over here has those rows, holding
	// the boundary ghost states, so the reference's code reads defined data - the "ghost" boundary policy.
	MVector<freal> um(m.gnelem() + m.gnbface(), NVARS);
	for(fint i = 0; i < m.gnelem(); i++) for(int k = 0; k < NVARS; k++) um(i,k) = u[(size_t)i*NVARS+k];
	for(fint f = 0; f < m.gnbface(); f++) for(int k = 0; k < NVARS; k++) um(m.gnelem()+f,k) = ug[(size_t)f*NVARS+k];
	r->compute_face_values(um, amat::Array2dView<freal>(ug, m.gnbface(), NVARS), grad,
	                       amat::Array2dMutableView<freal>(ufl, m.gnaface(), NVARS), amat::Array2dMutableView<freal>(ufr, m.gnaface(), NVARS));
}

}
