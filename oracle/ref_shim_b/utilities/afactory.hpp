/* The factory declarations flow_spatial.cpp needs, with the reference's signatures (utilities/afactory.hpp:38-112).
 * The reference's header also drags in the ODE and control-file layers (Boost); the definitions the harness links
 * are in oracle/ref_tier_c.cpp and map the same string keys to the same reference classes. TEST INFRASTRUCTURE ONLY. */
#ifndef AFACTORY_H
#define AFACTORY_H
#include <string>
#include "utilities/aarray2d.hpp"
#include "spatial/anumericalflux.hpp"
#include "spatial/agradientschemes.hpp"
#include "spatial/areconstruction.hpp"

namespace fvens {
template <typename scalar>
const InviscidFlux<scalar>* create_const_inviscidflux(const std::string& type, const IdealGasPhysics<scalar> *const p);
template <typename scalar, int nvars>
const GradientScheme<scalar,nvars>* create_const_gradientscheme(const std::string& type, const UMesh<scalar,NDIM> *const m,
                                                                const scalar *const rc, const scalar *const rcbp);
template <typename scalar, int nvars>
const SolutionReconstruction<scalar,nvars>* create_const_reconstruction(const std::string& type, const UMesh<scalar,NDIM> *const m,
                                                                        const scalar *const rc, const scalar *const rcbp,
                                                                        const amat::Array2d<scalar>& gr, const freal param);
}
#endif
