/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * The reference translation units of the residual path, included in place and unmodified from /root/reference/src,
 * plus the harness definitions they need at link time: createGhostedSystemVector (linalg/alinalg.cpp:17-29 on the
 * serial Vec stand-in) and the three string-keyed factories (same keys -> same reference classes as
 * utilities/afactory.cpp:30-217, whose own file needs the ODE and Boost control-file layers). Shared by
 * ref_tier_c.cpp / ref_tier_d.cpp (mesh stand-in) and ref_tier_e.cpp (the reference's own UMesh). No reference code here.
 */
#ifndef FVENS_B200_REF_SOURCES_SPATIAL
#define FVENS_B200_REF_SOURCES_SPATIAL
#include "physics/aphysics.cpp"
#include "physics/viscousphysics.cpp"
#include "spatial/anumericalflux.cpp"
#include "spatial/abc.cpp"
#include "spatial/abctypemap.cpp"
#include "spatial/agradientschemes.cpp"
#include "spatial/areconstruction.cpp"
#include "spatial/limitedlinearreconstruction.cpp"
#include "spatial/musclreconstruction.cpp"
#include "utilities/aarray2d.cpp"
#include "utilities/aerrorhandling.cpp"
#include "utilities/mpiutils.cpp"
#include "linalg/petscutils.cpp"
#include "linalg/tracevector.cpp"
#include "spatial/aspatial.cpp"
#include "spatial/flow_spatial.cpp"
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace fvens {

/// linalg/alinalg.cpp:17-29 with the serial Vec stand-in: (nelem + connectivity ghosts) blocks of nvars
StatusCode createGhostedSystemVector(const UMesh<freal,NDIM> *const m, const int nvars, Vec *const v)
{
	*v = new _p_Vec;
	(*v)->nlocal = m->gnelem()*nvars; (*v)->nghost = m->gnConnFace()*nvars;
	(*v)->a.assign((size_t)((*v)->nlocal + (*v)->nghost), 0.0);
	return 0;
}

template <typename scalar>
const InviscidFlux<scalar>* create_const_inviscidflux(const std::string& type, const IdealGasPhysics<scalar> *const p)
{
	if(type == "VANLEER") return new VanLeerFlux<scalar>(p);
	if(type == "ROE") return new RoeFlux<scalar>(p);
	if(type == "HLL") return new HLLFlux<scalar>(p);
	if(type == "HLLC") return new HLLCFlux<scalar>(p);
	if(type == "LLF") return new LocalLaxFriedrichsFlux<scalar>(p);
	if(type == "AUSM") return new AUSMFlux<scalar>(p);
	if(type == "AUSMPLUS") return new AUSMPlusFlux<scalar>(p);
	std::cout << " InviscidFluxFactory: Invalid flux!\n";
	return nullptr;
}

template <typename scalar, int nvars>
const GradientScheme<scalar,nvars>* create_const_gradientscheme(const std::string& type, const UMesh<scalar,NDIM> *const m,
                                                                const scalar *const rc, const scalar *const rcbp)
{
	if(type == "LEASTSQUARES") return new WeightedLeastSquaresGradients<scalar,nvars>(m, rc, rcbp);
	if(type == "GREENGAUSS") return new GreenGaussGradients<scalar,nvars>(m, rc, rcbp);
	return new ZeroGradients<scalar,nvars>(m, rc, rcbp);
}

template <typename scalar, int nvars>
const SolutionReconstruction<scalar,nvars>* create_const_reconstruction(const std::string& type, const UMesh<scalar,NDIM> *const m,
                                                                        const scalar *const rc, const scalar *const rcbp,
                                                                        const amat::Array2d<scalar>& gr, const freal param)
{
	if(type == "NONE") return new LinearUnlimitedReconstruction<scalar,nvars>(m, rc, rcbp, gr);
	if(type == "WENO") return new WENOReconstruction<scalar,nvars>(m, rc, rcbp, gr, param);
	if(type == "VANALBADA") return new MUSCLVanAlbada<scalar,nvars>(m, rc, rcbp, gr);
	if(type == "BARTHJESPERSEN") return new BarthJespersenLimiter<scalar,nvars>(m, rc, rcbp, gr);
	if(type == "VENKATAKRISHNAN") return new VenkatakrishnanLimiter<scalar,nvars>(m, rc, rcbp, gr, param);
	std::cout << " !ReconstructionFactory: Invalid reconstruction!!\n";
	return nullptr;
}

template const InviscidFlux<freal>* create_const_inviscidflux<freal>(const std::string&, const IdealGasPhysics<freal> *const);
template const GradientScheme<freal,NVARS>* create_const_gradientscheme<freal,NVARS>(const std::string&, const UMesh<freal,NDIM> *const, const freal *const, const freal *const);
template const SolutionReconstruction<freal,NVARS>* create_const_reconstruction<freal,NVARS>(const std::string&, const UMesh<freal,NDIM> *const, const freal *const, const freal *const, const amat::Array2d<freal>&, const freal);

}

#endif
