/* flow_spatial_b200.hpp - THE REFERENCE-SIDE BINDING: the one file a FVENS maintainer adds (as src/spatial/
 * flow_spatial_b200.hpp) to route FlowFV::compute_residual through libfvens_b200. It is written against FVENS's own
 * headers - not against this repository's host layer - and is compiled against them here: oracle/ref_binding.cpp builds
 * it with the unmodified reference sources.
 *
 * FlowFV_B200<order2,constVisc> derives from the reference's FlowFV<freal,order2,constVisc> and overrides only
 * compute_residual (spatial/flow_spatial.hpp:199-200): gradients for output, surface data and the Jacobian blocks of the
 * implicit path stay the reference's CPU code. The mesh reaches the library through UMesh's own accessors, the numerics
 * through the string keys of the reference's factories (utilities/afactory.cpp:38-81, 111-127, 178-211).
 */
#ifndef FVENS_FLOW_SPATIAL_B200_H
#define FVENS_FLOW_SPATIAL_B200_H

#include <stdexcept>
#include <string>
#include <vector>
#include "spatial/flow_spatial.hpp"
#include "linalg/petscutils.hpp"
#include "fvens_b200.h"

namespace fvens {

/// How the device-resident driver (ode_b200.hpp) finds the engine behind a Spatial pointer
struct B200Engine { virtual fvg_flow* engine_flow() const = 0; virtual ~B200Engine() {} };

template <bool secondOrderRequested, bool constVisc>
class FlowFV_B200 : public FlowFV<freal,secondOrderRequested,constVisc>, public B200Engine
{
public:
	FlowFV_B200(const UMesh<freal,NDIM> *const mesh, const FlowPhysicsConfig& pc, const FlowNumericsConfig& nc,
	            const int reorder = FVG_REORDER_HILBERT, const int tile_cells = 256)
		: FlowFV<freal,secondOrderRequested,constVisc>(mesh, pc, nc), dmesh(nullptr), flow(nullptr)
	{
		// 1. the mesh arrays, copied once through UMesh's accessors (they return by value) into contiguous storage
		const fint npoin = mesh->gnpoin(), nelem = mesh->gnelem(), nbface = mesh->gnbface(), naface = mesh->gnaface();
		const int mw = (int)mesh->gmaxnfael();
		std::vector<double> coords((size_t)npoin*NDIM), facemetric((size_t)naface*3), area(nelem);
		std::vector<int> inpoel((size_t)nelem*mw, -1), esuel((size_t)nelem*mw, -1), elemface((size_t)nelem*mw, -1), nnode(nelem),
			intfac((size_t)naface*4), btags((size_t)nbface*std::max(mesh->gnbtag(), 1));
		for(fint p = 0; p < npoin; p++)
			for(int d = 0; d < NDIM; d++) coords[(size_t)p*NDIM+d] = mesh->gcoords(p,d);
		for(fint e = 0; e < nelem; e++) {
			nnode[e] = mesh->gnnode(e);
			area[e] = mesh->garea(e);
			for(int j = 0; j < mesh->gnnode(e); j++) inpoel[(size_t)e*mw+j] = (int)mesh->ginpoel(e,j);
			for(int j = 0; j < mesh->gnfael(e); j++) {
				esuel[(size_t)e*mw+j] = (int)mesh->gesuel(e,j);
				elemface[(size_t)e*mw+j] = (int)mesh->gelemface(e,j);
			}
		}
		for(fint f = 0; f < naface; f++) {
			for(int j = 0; j < 4; j++) intfac[(size_t)f*4+j] = (int)mesh->gintfac(f,j);
			for(int j = 0; j < 3; j++) facemetric[(size_t)f*3+j] = mesh->gfacemetric(f,j);
		}
		for(fint f = 0; f < nbface; f++)
			for(int j = 0; j < mesh->gnbtag(); j++) btags[(size_t)f*mesh->gnbtag()+j] = mesh->gbtags(f,j);

		fvg_host_mesh v;
		v.npoin = (int)npoin; v.nelem = (int)nelem; v.nbface = (int)nbface; v.naface = (int)naface;
		v.nconnface = (int)mesh->gnConnFace(); v.ninface = v.naface - v.nbface - v.nconnface;
		v.maxnnode = mw; v.nbtag = mesh->gnbtag();
		v.coords = coords.data(); v.inpoel = inpoel.data(); v.nnode = nnode.data(); v.esuel = esuel.data();
		v.elemface = elemface.data(); v.intfac = intfac.data(); v.btags = btags.data();
		v.facemetric = facemetric.data(); v.area = area.data();
		v.bpartner = nullptr;      // the reference cannot run periodic boundaries: none to forward
		const fvg_mesh_opts mo = { reorder, tile_cells, -1 };
		if(fvg_mesh_create(&v, &mo, &dmesh)) throw std::runtime_error(fvg_last_error());

		// 2. physics, numerics and boundary conditions
		const fvg_physics p = { pc.gamma, pc.Minf, pc.Tinf, pc.Reinf, pc.Pr, pc.aoa, pc.viscous_sim ? 1 : 0, pc.const_visc ? 1 : 0 };
		const fvg_numerics n = { key(nc.conv_numflux, {"LLF","VANLEER","AUSM","AUSMPLUS","ROE","HLL","HLLC"}, -1),
		                         key(nc.gradientscheme, {"","GREENGAUSS","LEASTSQUARES"}, FVG_GRAD_ZERO),
		                         key(nc.reconstruction, {"NONE","WENO","VANALBADA","BARTHJESPERSEN","VENKATAKRISHNAN"}, -1),
		                         nc.limiter_param, nc.order2 ? 1 : 0, FVG_BND_GHOST };
		std::vector<fvg_bc> bcs;
		for(const FlowBCConfig& b : pc.bcconf) {
			const fvg_bc c = { b.bc_tag, (int)b.bc_type, { b.bc_vals.size() > 0 ? b.bc_vals[0] : 0.0,
			                                               b.bc_vals.size() > 1 ? b.bc_vals[1] : 0.0 } };
			bcs.push_back(c);
		}
		if(fvg_flow_create(dmesh, &p, &n, bcs.data(), (int)bcs.size(), &flow)) {
			const std::string msg = fvg_last_error();
			fvg_mesh_destroy(dmesh);
			throw std::runtime_error(msg);
		}
	}

	~FlowFV_B200() { fvg_flow_destroy(flow); fvg_mesh_destroy(dmesh); }

	/// The reference's contract (spatial/aspatial.hpp:44-63): ADDS -r(u) into `residual`; `dtm` only if asked
	StatusCode compute_residual(const Vec u, Vec residual, const bool gettimesteps, Vec dtm) const
	{
		ConstGhostedVecHandler<PetscScalar> uh(u);
		MutableVecHandler<PetscScalar> rh(residual);
		if(gettimesteps) {
			MutableVecHandler<PetscScalar> dth(dtm);
			return fvg_residual_host(flow, uh.getArray(), rh.getArray(), 1, 1, dth.getArray());
		}
		return fvg_residual_host(flow, uh.getArray(), rh.getArray(), 1, 0, nullptr);     // 0 = success, as CHKERRQ expects
	}

	fvg_flow* engine_flow() const { return flow; }

private:
	fvg_mesh *dmesh;
	fvg_flow *flow;

	static int key(const std::string& name, const std::vector<std::string>& names, const int otherwise) {
		for(size_t i = 0; i < names.size(); i++) if(!names[i].empty() && names[i] == name) return (int)i;
		if(otherwise < 0) throw std::runtime_error("FlowFV_B200: unknown key " + name);
		return otherwise;
	}
};

/// What create_mutable_flowSpatialDiscretization (utilities/afactory.cpp:252-275) returns when the run asks for the GPU
inline const Spatial<freal,NVARS>* create_flowSpatialDiscretization_b200(const UMesh<freal,NDIM> *const m,
                                                                          const FlowPhysicsConfig& pconf,
                                                                          const FlowNumericsConfig& nconf)
{
	if(nconf.order2) {
		if(pconf.const_visc) return new FlowFV_B200<true,true>(m, pconf, nconf);
		return new FlowFV_B200<true,false>(m, pconf, nconf);
	}
	if(pconf.const_visc) return new FlowFV_B200<false,true>(m, pconf, nconf);
	return new FlowFV_B200<false,false>(m, pconf, nconf);
}

}
#endif
