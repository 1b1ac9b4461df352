/* C++ host surface of fvens_b200: the reference's class names and call signatures for the residual path,
 * forwarding to the C ABI (include/fvens_b200.h). Header-only above the ABI; link libfvens_b200.so.
 *
 * Mirrors (reference paths relative to src/):
 *   Vec + createSystemVector/createGhostedSystemVector   linalg/alinalg.hpp:21-52 (PETSc Vec replaced by a shim
 *                                                        that owns host or device memory)
 *   amat::Array2dView / Array2dMutableView               utilities/aarray2d.hpp
 *   GradBlock_t, StatusCode, freal/fint, NDIM/NVARS      aconstants.hpp:32-93
 *   Numerical_error, Tolerance_error, ...                utilities/aerrorhandling.hpp:16-60
 *   FlowBCConfig, FlowBC family, create_const_flowBCs    spatial/abc.hpp:34-345
 *   InviscidFlux family, create_const_inviscidflux       spatial/anumericalflux.hpp:19-300, utilities/afactory.cpp:30-87
 *   GradientScheme family, create_const_gradientscheme   spatial/agradientschemes.hpp:18-130, afactory.cpp:100-133
 *   SolutionReconstruction family, create_const_reconstruction  spatial/areconstruction.hpp:19-80,
 *                                                        limitedlinearreconstruction.hpp, musclreconstruction.hpp, afactory.cpp:160-217
 *   Spatial, FlowFV_base, FlowFV, configs, create_const_flowSpatialDiscretization
 *                                                        spatial/aspatial.hpp:36-170, flow_spatial.hpp:33-262, afactory.cpp:252-275
 *   SteadySolverConfig, TimingData, SteadySolver, SteadyForwardEulerSolver   ode/aodesolver.hpp:18-137
 *
 * Every method that computes launches CUDA kernels through the ABI; there is no CPU implementation behind
 * any of them. Jacobian-related members of the reference are not part of this path and throw.
 */
#ifndef FVENS_B200_HOST_HPP
#define FVENS_B200_HOST_HPP

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "mesh.hpp"
#include "../../include/fvens_b200.h"

namespace fvens {

using StatusCode = int;

// ---------------------------------------------------------------------------------------- errors

class Numerical_error : public std::logic_error {
public:
	Numerical_error(const std::string& msg) : std::logic_error(msg) {}
};
class Tolerance_error : public Numerical_error {
public:
	Tolerance_error(const std::string& msg) : Numerical_error(msg) {}
};
class UnsupportedOptionError : public std::runtime_error {
public:
	UnsupportedOptionError(const std::string& msg) : std::runtime_error(msg) {}
};
/// Thrown when the engine reports an error that has no counterpart in the reference (CUDA failures etc.)
class Engine_error : public std::runtime_error {
public:
	Engine_error(const std::string& msg) : std::runtime_error(msg) {}
};

/// Reference: fvens_throw (utilities/aerrorhandling.hpp:64-70)
inline void fvens_throw(const int ierr, const std::string& msg) {
	if(ierr != 0) throw std::runtime_error(msg);
}
/// Turns an ABI status into the reference's exception taxonomy
inline void fvg_throw(const int code, const char *where) {
	if(code == FVG_OK) return;
	const std::string msg = std::string(where) + ": " + fvg_last_error();
	if(code == FVG_ERR_TOLERANCE) throw Tolerance_error(msg);
	if(code == FVG_ERR_NUMERICAL) throw Numerical_error(msg);
	if(code == FVG_ERR_UNSUPPORTED) throw UnsupportedOptionError(msg);
	throw Engine_error(msg);
}

// ---------------------------------------------------------------------------------------- containers

template <typename scalar, int ndim, int nvars>
struct GradBlock_t {
	scalar v[ndim*nvars];    ///< column-major ndim x nvars: v[idim + ndim*ivar] (aconstants.hpp:80-90)
	scalar& operator()(const int idim, const int ivar) { return v[idim + ndim*ivar]; }
	const scalar& operator()(const int idim, const int ivar) const { return v[idim + ndim*ivar]; }
};

namespace amat {
/// Non-owning row-major views (utilities/aarray2d.hpp)
template <typename T> class Array2dView {
public:
	Array2dView(const T *const data, const fint nr, const int nc) : p(data), nr_(nr), nc_(nc) {}
	const T& operator()(const fint i, const int j) const { return p[(size_t)i*nc_+j]; }
	const T* data() const { return p; }
	fint rows() const { return nr_; }
	int cols() const { return nc_; }
private:
	const T *p; fint nr_; int nc_;
};
template <typename T> class Array2dMutableView {
public:
	Array2dMutableView(T *const data, const fint nr, const int nc) : p(data), nr_(nr), nc_(nc) {}
	T& operator()(const fint i, const int j) { return p[(size_t)i*nc_+j]; }
	const T& operator()(const fint i, const int j) const { return p[(size_t)i*nc_+j]; }
	T* data() { return p; }
	const T* data() const { return p; }
	fint rows() const { return nr_; }
	int cols() const { return nc_; }
private:
	T *p; fint nr_; int nc_;
};
/// Owning row-major array
template <typename T> class Array2d {
public:
	Array2d() : nr_(0), nc_(0) {}
	Array2d(const fint nr, const int nc) : d((size_t)nr*nc), nr_(nr), nc_(nc) {}
	void resize(const fint nr, const int nc) { d.assign((size_t)nr*nc, T()); nr_ = nr; nc_ = nc; }
	T& operator()(const fint i, const int j) { return d[(size_t)i*nc_+j]; }
	const T& operator()(const fint i, const int j) const { return d[(size_t)i*nc_+j]; }
	T* data() { return d.data(); }
	const T* data() const { return d.data(); }
	fint rows() const { return nr_; }
	int cols() const { return nc_; }
private:
	std::vector<T> d; fint nr_; int nc_;
};
}

/// The reference's MVector is a dynamic row-major Eigen matrix; here the owning 2D array
template <typename scalar> using MVector = amat::Array2d<scalar>;

// ---------------------------------------------------------------------------------------- Vec shim

enum VecPlace { VEC_HOST = 0, VEC_DEVICE = 1 };

/// Stand-in for a PETSc ghosted block Vec (linalg/alinalg.cpp:9-29): nlocal blocks followed by nghost
/// ghost blocks of `bs` doubles, in host memory (drop-in mode: every compute_residual copies) or in device
/// memory (resident mode: the state never leaves the GPU).
struct _p_Vec {
	fint nlocal = 0, nghost = 0;
	int bs = 1;
	VecPlace place = VEC_HOST;
	std::vector<double> host;
	double *dev = nullptr;
	size_t size() const { return (size_t)(nlocal + nghost)*bs; }
};
typedef _p_Vec* Vec;

inline StatusCode VecCreateBlocked(const fint nlocal, const fint nghost, const int bs, const VecPlace place, Vec *const v) {
	std::unique_ptr<_p_Vec> p(new _p_Vec);
	p->nlocal = nlocal; p->nghost = nghost; p->bs = bs; p->place = place;
	if(place == VEC_HOST) p->host.assign(p->size(), 0.0);
	else {
		void *d = nullptr;
		int rc = fvg_malloc(&d, p->size()*sizeof(double)); if(rc) return rc;
		p->dev = static_cast<double*>(d);
		rc = fvg_memset(d, 0, p->size()*sizeof(double)); if(rc) { fvg_free(d); return rc; }
	}
	*v = p.release();
	return 0;
}
inline StatusCode VecDestroy(Vec *const v) {
	if(v && *v) { if((*v)->dev) fvg_free((*v)->dev); delete *v; *v = nullptr; }
	return 0;
}
inline StatusCode VecDuplicate(const Vec a, Vec *const b) {
	return VecCreateBlocked(a->nlocal, a->nghost, a->bs, a->place, b);
}
/// Host access; only valid for host Vecs (device Vecs are read with VecCopyToHost)
inline StatusCode VecGetArray(Vec v, double **arr) {
	if(v->place != VEC_HOST) return FVG_ERR_INVALID;
	*arr = v->host.data(); return 0;
}
inline StatusCode VecGetArrayRead(const Vec v, const double **arr) {
	if(v->place != VEC_HOST) return FVG_ERR_INVALID;
	*arr = v->host.data(); return 0;
}
inline StatusCode VecRestoreArray(Vec, double **) { return 0; }
inline StatusCode VecRestoreArrayRead(const Vec, const double **) { return 0; }
inline StatusCode VecGetLocalSize(const Vec v, fint *const n) { *n = v->nlocal*v->bs; return 0; }
inline StatusCode VecSet(Vec v, const double val) {
	if(v->place == VEC_HOST) { std::fill(v->host.begin(), v->host.end(), val); return 0; }
	if(val == 0.0) return fvg_memset(v->dev, 0, v->size()*sizeof(double));
	std::vector<double> tmp(v->size(), val);
	return fvg_memcpy(v->dev, tmp.data(), tmp.size()*sizeof(double), 0);
}
inline StatusCode VecCopyFromHost(Vec v, const double *const src) {
	if(v->place == VEC_HOST) { std::memcpy(v->host.data(), src, v->size()*sizeof(double)); return 0; }
	return fvg_memcpy(v->dev, src, v->size()*sizeof(double), 0);
}
inline StatusCode VecCopyToHost(const Vec v, double *const dst) {
	if(v->place == VEC_HOST) { std::memcpy(dst, v->host.data(), v->size()*sizeof(double)); return 0; }
	return fvg_memcpy(dst, v->dev, v->size()*sizeof(double), 1);
}

/// Reference: createGhostedSystemVector / createSystemVector (linalg/alinalg.cpp:9-52)
inline StatusCode createGhostedSystemVector(const UMesh<freal,NDIM> *const m, const int nvars, Vec *const v,
                                            const VecPlace place = VEC_HOST) {
	return VecCreateBlocked(m->gnelem(), m->gnConnFace(), nvars, place, v);
}
inline StatusCode createSystemVector(const UMesh<freal,NDIM> *const m, const int nvars, Vec *const v,
                                     const VecPlace place = VEC_HOST) {
	return VecCreateBlocked(m->gnelem(), 0, nvars, place, v);
}

// ---------------------------------------------------------------------------------------- physics + configs

enum BCType { SLIP_WALL_BC, FARFIELD_BC, INFLOW_OUTFLOW_BC, SUBSONIC_INFLOW_BC, EXTRAPOLATION_BC, PERIODIC_BC,
              ISOTHERMAL_WALL_BC, ADIABATIC_WALL_BC };

struct FlowBCConfig {
	int bc_tag;
	BCType bc_type;
	std::vector<freal> bc_vals;
	std::vector<int> bc_opts;
};

struct FlowPhysicsConfig {
	freal gamma, Minf, Tinf, Reinf, Pr, aoa;
	bool viscous_sim, const_visc;
	std::vector<FlowBCConfig> bcconf;
};

struct FlowNumericsConfig {
	std::string conv_numflux, conv_numflux_jac, gradientscheme, reconstruction;
	freal limiter_param;
	bool order2;
};

/// Carries the gas constants (physics/aphysics.hpp:60-75). The arithmetic lives on the device; the host
/// object is the parameter block the plug-in classes hand to the engine.
template <typename scalar>
class IdealGasPhysics {
public:
	IdealGasPhysics(const freal _g, const freal M_inf, const freal T_inf, const freal Re_inf, const freal _Pr)
		: g(_g), Minf(M_inf), Tinf(T_inf), Reinf(Re_inf), Pr(_Pr) {}
	const freal g, Minf, Tinf, Reinf, Pr;
	fvg_physics abi(const freal aoa = 0.0, const bool viscous = false, const bool constvisc = false) const {
		fvg_physics p; p.gamma = g; p.Minf = Minf; p.Tinf = Tinf; p.Reinf = Reinf; p.Pr = Pr; p.aoa = aoa;
		p.viscous_sim = viscous; p.const_visc = constvisc; return p;
	}
	/// Reference: compute_freestream_state (physics/aphysics.cpp:44-58)
	std::array<scalar,NVARS> compute_freestream_state(const freal aoa) const {
		std::array<scalar,NVARS> u; const fvg_physics p = abi(aoa);
		fvg_throw(fvg_freestream(&p, u.data()), "compute_freestream_state"); return u;
	}
};

// ---------------------------------------------------------------------------------------- boundary conditions

template <typename scalar, typename j_real = freal>
class FlowBC {
public:
	FlowBC(const BCType bt, const int bc_tag, const IdealGasPhysics<scalar>& gasphysics,
	       const freal v0 = 0.0, const freal v1 = 0.0)
		: btype(bt), btag(bc_tag), phy(gasphysics), vals{v0, v1} {}
	virtual ~FlowBC() {}
	int bctag() const { return btag; }
	BCType type() const { return btype; }
	/// Ghost state from the interior state and the unit normal (spatial/abc.hpp:72-73)
	virtual void computeGhostState(const scalar *const uin, const scalar *const n, scalar *const ughost) const {
		computeGhostStates(1, uin, n, ughost);
	}
	/// Batched form of the same call (one kernel for n states)
	void computeGhostStates(const int n, const scalar *const uin, const scalar *const nrm, scalar *const ughost) const {
		const fvg_bc bc = abi(); const fvg_physics p = phy.abi(aoa_);
		fvg_throw(fvg_bc_pointwise(&bc, &p, n, uin, nrm, ughost), "FlowBC::computeGhostState");
	}
	virtual void computeGhostStateAndJacobian(const scalar *const, const scalar *const, scalar *const, scalar *const) const {
		throw UnsupportedOptionError("FlowBC::computeGhostStateAndJacobian is not on the explicit residual path");
	}
	fvg_bc abi() const { fvg_bc b; b.tag = btag; b.type = (int)btype; b.vals[0] = vals[0]; b.vals[1] = vals[1]; return b; }
	void set_aoa(const freal a) { aoa_ = a; }
protected:
	const BCType btype;
	const int btag;
	const IdealGasPhysics<scalar>& phy;
	const freal vals[2];
	freal aoa_ = 0.0;      ///< the free-stream direction enters far-field / in-out-flow through uinf
};

#define FVENS_BC_CLASS(Name, TYPE) \
	template <typename scalar, typename j_real = freal> class Name : public FlowBC<scalar,j_real> { public: \
		Name(const int bc_tag, const IdealGasPhysics<scalar>& gp) : FlowBC<scalar,j_real>(TYPE, bc_tag, gp) {} };
FVENS_BC_CLASS(Slipwall, SLIP_WALL_BC)
FVENS_BC_CLASS(Extrapolation, EXTRAPOLATION_BC)
#undef FVENS_BC_CLASS

/// Far field and in/out flow take the free-stream state; the engine derives it from the physics (Minf, aoa),
/// so only the direction is read back from `u_far` (spatial/abc.hpp:112-200).
template <typename scalar, typename j_real = freal>
class Farfield : public FlowBC<scalar,j_real> {
public:
	Farfield(const int bc_tag, const IdealGasPhysics<scalar>& gp, const std::array<scalar,NVARS>& u_far)
		: FlowBC<scalar,j_real>(FARFIELD_BC, bc_tag, gp) { this->set_aoa(std::atan2(u_far[2], u_far[1])); }
};
template <typename scalar, typename j_real = freal>
class InOutFlow : public FlowBC<scalar,j_real> {
public:
	InOutFlow(const int bc_tag, const IdealGasPhysics<scalar>& gp, const std::array<scalar,NVARS>& u_far)
		: FlowBC<scalar,j_real>(INFLOW_OUTFLOW_BC, bc_tag, gp) { this->set_aoa(std::atan2(u_far[2], u_far[1])); }
};
template <typename scalar, typename j_real = freal>
class InFlow : public FlowBC<scalar,j_real> {
public:
	InFlow(const int bc_tag, const IdealGasPhysics<scalar>& gp, const freal totalpressure, const freal totaltemperature)
		: FlowBC<scalar,j_real>(SUBSONIC_INFLOW_BC, bc_tag, gp, totalpressure, totaltemperature) {}
};
template <typename scalar, typename j_real = freal>
class Adiabaticwall2D : public FlowBC<scalar,j_real> {
public:
	Adiabaticwall2D(const int bc_tag, const IdealGasPhysics<scalar>& gp, const freal wall_tangential_velocity)
		: FlowBC<scalar,j_real>(ADIABATIC_WALL_BC, bc_tag, gp, wall_tangential_velocity) {}
};
template <typename scalar, typename j_real = freal>
class Isothermalwall2D : public FlowBC<scalar,j_real> {
public:
	Isothermalwall2D(const int bc_tag, const IdealGasPhysics<scalar>& gp, const freal wall_tangential_velocity,
	                 const freal wall_temperature)
		: FlowBC<scalar,j_real>(ISOTHERMAL_WALL_BC, bc_tag, gp, wall_tangential_velocity, wall_temperature) {}
};

/// Periodic marker. NOT in the reference (its factory throws for PERIODIC_BC, abc.cpp:493-494, and its face loop would
/// count a periodic edge twice): here the faces of the marker are paired by UMesh::compute_periodic_map
/// and the device mesh makes interior faces of them, so no ghost state is ever asked of this object.
template <typename scalar, typename j_real = freal>
class PeriodicBC : public FlowBC<scalar,j_real> {
public:
	PeriodicBC(const int bc_tag, const IdealGasPhysics<scalar>& gp) : FlowBC<scalar,j_real>(PERIODIC_BC, bc_tag, gp) {}
	void computeGhostState(const scalar *const, const scalar *const, scalar *const) const {
		throw UnsupportedOptionError("PeriodicBC has no ghost state: its faces are interior faces of the device mesh");
	}
};

/// Reference: create_const_flowBCs (spatial/abc.cpp:461-500); unknown types throw std::runtime_error
template <typename scalar>
std::map<int,const FlowBC<scalar>*> create_const_flowBCs(const std::vector<FlowBCConfig>& conf,
                                                         const IdealGasPhysics<scalar>& physics,
                                                         const std::array<freal,NVARS>& uinf)
{
	std::map<int,const FlowBC<scalar>*> bcmap;
	for(auto it = conf.begin(); it != conf.end(); it++) {
		const FlowBC<scalar>* bc;
		switch(it->bc_type) {
		case SLIP_WALL_BC: bc = new Slipwall<scalar>(it->bc_tag, physics); break;
		case FARFIELD_BC: bc = new Farfield<scalar>(it->bc_tag, physics, uinf); break;
		case INFLOW_OUTFLOW_BC: bc = new InOutFlow<scalar>(it->bc_tag, physics, uinf); break;
		case SUBSONIC_INFLOW_BC: bc = new InFlow<scalar>(it->bc_tag, physics, it->bc_vals.at(0), it->bc_vals.at(1)); break;
		case EXTRAPOLATION_BC: bc = new Extrapolation<scalar>(it->bc_tag, physics); break;
		case ADIABATIC_WALL_BC: bc = new Adiabaticwall2D<scalar>(it->bc_tag, physics, it->bc_vals.at(0)); break;
		case ISOTHERMAL_WALL_BC: bc = new Isothermalwall2D<scalar>(it->bc_tag, physics, it->bc_vals.at(0), it->bc_vals.at(1)); break;
		case PERIODIC_BC: bc = new PeriodicBC<scalar>(it->bc_tag, physics); break;      // (the reference throws here)
		default: throw std::runtime_error("BC type not implemented yet!");
		}
		bcmap[it->bc_tag] = bc;
	}
	return bcmap;
}

// ---------------------------------------------------------------------------------------- inviscid fluxes

template <typename scalar, typename j_real = freal>
class InviscidFlux {
public:
	InviscidFlux(const IdealGasPhysics<scalar> *const analyticalflux, const int flux_id) : physics(analyticalflux), id(flux_id) {}
	virtual ~InviscidFlux() {}
	/// Numerical flux normal to the face (spatial/anumericalflux.hpp:32-34)
	virtual void get_flux(const scalar *const uleft, const scalar *const uright, const scalar *const n,
	                      scalar *const flux) const {
		get_fluxes(1, uleft, uright, n, flux);
	}
	/// Batched form of the same call (one kernel for nf faces)
	void get_fluxes(const int nf, const scalar *const ul, const scalar *const ur, const scalar *const nrm,
	                scalar *const flux) const {
		const fvg_physics p = physics->abi();
		fvg_throw(fvg_flux_pointwise(id, &p, nf, ul, ur, nrm, flux), "InviscidFlux::get_flux");
	}
	virtual void get_jacobian(const freal *const, const freal *const, const freal *const, freal *const, freal *const) const {
		throw UnsupportedOptionError("InviscidFlux::get_jacobian is not on the explicit residual path");
	}
	int abi_id() const { return id; }
protected:
	const IdealGasPhysics<scalar> *const physics;
	const int id;
};

#define FVENS_FLUX_CLASS(Name, ID) \
	template <typename scalar, typename j_real = freal> class Name : public InviscidFlux<scalar,j_real> { public: \
		Name(const IdealGasPhysics<scalar> *const p) : InviscidFlux<scalar,j_real>(p, ID) {} };
FVENS_FLUX_CLASS(LocalLaxFriedrichsFlux, FVG_FLUX_LLF)
FVENS_FLUX_CLASS(VanLeerFlux, FVG_FLUX_VANLEER)
FVENS_FLUX_CLASS(AUSMFlux, FVG_FLUX_AUSM)
FVENS_FLUX_CLASS(AUSMPlusFlux, FVG_FLUX_AUSMPLUS)
FVENS_FLUX_CLASS(RoeFlux, FVG_FLUX_ROE)
FVENS_FLUX_CLASS(HLLFlux, FVG_FLUX_HLL)
FVENS_FLUX_CLASS(HLLCFlux, FVG_FLUX_HLLC)
#undef FVENS_FLUX_CLASS

/// Reference: create_const_inviscidflux (utilities/afactory.cpp:30-87); unknown key prints and returns nullptr
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
	std::cout << " ! FluxFactory: Invalid inviscid flux!!\n";
	return nullptr;
}

// ---------------------------------------------------------------------------------------- engine handle

/// View of a host UMesh in the layout of the C ABI (non-owning)
inline fvg_host_mesh host_mesh_view(const UMesh<freal,NDIM> *const m) {
	fvg_host_mesh v;
	v.npoin = m->gnpoin(); v.nelem = m->gnelem(); v.nbface = m->gnbface(); v.naface = m->gnaface();
	v.ninface = m->gninface(); v.nconnface = m->gnConnFace(); v.maxnnode = m->gmaxnnode(); v.nbtag = m->gnbtag();
	v.coords = m->coordsData(); v.inpoel = m->inpoelData(); v.nnode = m->nnodeData(); v.esuel = m->esuelData();
	v.elemface = m->elemfaceData(); v.intfac = m->intfacData(); v.btags = m->btagsData();
	v.facemetric = m->facemetricData(); v.area = m->areaData();
	v.bpartner = m->periodicmapData();       // periodic pairs (UMesh::compute_periodic_map), or null
	return v;
}

/// Device mesh + flow context shared by the plug-in classes and FlowFV. Owns its ABI handles.
/// With a cell -> rank map it is one rank's subdomain of a mesh distributed over the GPUs of a box (the job of the
/// reference's ReplicatedGlobalMeshPartitioner::restrictMeshToPartitions, mesh/meshpartitioning.cpp:24-159).
class EngineContext {
public:
	EngineContext(const UMesh<freal,NDIM> *const m, const fvg_physics& phys, const fvg_numerics& num,
	              const std::vector<fvg_bc>& bcs, const int reorder = FVG_REORDER_HILBERT, const int tile_cells = 256,
	              const int *const cell_rank = nullptr, const int rank = 0, const int nranks = 1, const int device = -1)
	{
		const fvg_host_mesh v = host_mesh_view(m);
		fvg_mesh_opts o; o.reorder = reorder; o.tile_cells = tile_cells; o.device = device;
		fvg_throw(cell_rank ? fvg_mesh_create_part(&v, &o, cell_rank, rank, nranks, &mesh) : fvg_mesh_create(&v, &o, &mesh), "fvg_mesh_create");
		const int rc = fvg_flow_create(mesh, &phys, &num, bcs.data(), (int)bcs.size(), &flow);
		if(rc) { fvg_mesh_destroy(mesh); mesh = nullptr; fvg_throw(rc, "fvg_flow_create"); }
	}
	~EngineContext() { if(flow) fvg_flow_destroy(flow); if(mesh) fvg_mesh_destroy(mesh); }
	EngineContext(const EngineContext&) = delete;
	EngineContext& operator=(const EngineContext&) = delete;
	fvg_mesh *mesh = nullptr;
	fvg_flow *flow = nullptr;
};

/// Scratch device copy of a host array (RAII)
class DeviceScratch {
public:
	DeviceScratch(const size_t count, const double *const host_src = nullptr) : n(count) {
		void *d = nullptr; fvg_throw(fvg_malloc(&d, n*sizeof(double)), "fvg_malloc"); p = static_cast<double*>(d);
		if(host_src) fvg_throw(fvg_memcpy(p, host_src, n*sizeof(double), 0), "fvg_memcpy");
	}
	~DeviceScratch() { fvg_free(p); }
	void download(double *const dst) const { fvg_throw(fvg_memcpy(dst, p, n*sizeof(double), 1), "fvg_memcpy"); }
	double *p = nullptr; size_t n;
};

inline int gradient_id(const std::string& type) {
	if(type == "LEASTSQUARES") return FVG_GRAD_LEASTSQUARES;
	if(type == "GREENGAUSS") return FVG_GRAD_GREENGAUSS;
	return FVG_GRAD_ZERO;                      // reference: any other key selects ZeroGradients (afactory.cpp:123-127)
}
inline int reconstruction_id(const std::string& type) {
	if(type == "NONE") return FVG_RECON_NONE;
	if(type == "WENO") return FVG_RECON_WENO;
	if(type == "VANALBADA") return FVG_RECON_VANALBADA;
	if(type == "BARTHJESPERSEN") return FVG_RECON_BARTHJESPERSEN;
	if(type == "VENKATAKRISHNAN") return FVG_RECON_VENKATAKRISHNAN;
	return -1;
}
inline int flux_id(const std::string& type) {
	static const char *keys[] = {"LLF", "VANLEER", "AUSM", "AUSMPLUS", "ROE", "HLL", "HLLC"};
	for(int i = 0; i < 7; i++) if(type == keys[i]) return i;
	return -1;
}

/// A context for the stand-alone plug-in objects: they need a mesh on the device but no particular physics.
/// Every boundary marker present in the mesh gets an extrapolation BC (the plug-in calls take ghost states as input).
inline std::shared_ptr<EngineContext> make_plugin_context(const UMesh<freal,NDIM> *const m, const int grad, const int recon,
                                                          const freal param)
{
	fvg_physics p; p.gamma = 1.4; p.Minf = 1.0; p.Tinf = 288.15; p.Reinf = 1.0e6; p.Pr = 0.72; p.aoa = 0.0; p.viscous_sim = 0; p.const_visc = 0;
	fvg_numerics n; n.flux = FVG_FLUX_LLF; n.gradient = grad; n.reconstruction = recon; n.limiter_param = param; n.order2 = 1; n.bnd_policy = FVG_BND_GHOST;
	std::vector<int> tags;
	for(fint f = 0; f < m->gnbface(); f++) tags.push_back(m->gbtags(f, 0));
	std::sort(tags.begin(), tags.end()); tags.erase(std::unique(tags.begin(), tags.end()), tags.end());
	std::vector<fvg_bc> bcs;
	for(int t : tags) { fvg_bc b; b.tag = t; b.type = EXTRAPOLATION_BC; b.vals[0] = b.vals[1] = 0.0; bcs.push_back(b); }
	return std::make_shared<EngineContext>(m, p, n, bcs);
}

// ---------------------------------------------------------------------------------------- gradient schemes

template <typename scalar, int nvars>
class GradientScheme {
	static_assert(nvars == NVARS, "the engine differentiates the four flow variables");
public:
	GradientScheme(const UMesh<scalar,2> *const mesh, const scalar *const _rc, const scalar *const _rcbp, const int grad_id)
		: m(mesh), rc(_rc), rcbp(_rcbp), ctx(make_plugin_context(mesh, grad_id, FVG_RECON_NONE, 0.0)) {}
	virtual ~GradientScheme() {}
	/// Gradients of the cell states `unk` with boundary ghost states `unkg` (spatial/agradientschemes.hpp:44-48);
	/// grads: nelem GradBlock_t's, host memory
	virtual void compute_gradients(const amat::Array2dView<scalar> unk, const amat::Array2dView<scalar> unkg,
	                               scalar *const grads) const {
		const size_t ne = m->gnelem(), nb = m->gnbface();
		DeviceScratch du(ne*nvars, unk.data()), dg(std::max<size_t>(nb, 1)*nvars, nb ? unkg.data() : nullptr), dgr(ne*NDIM*nvars);
		fvg_throw(fvg_gradients(ctx->flow, du.p, dg.p, dgr.p, nullptr), "GradientScheme::compute_gradients");
		dgr.download(grads);
	}
protected:
	const UMesh<scalar,2> *const m;
	const scalar *const rc;
	const scalar *const rcbp;
	std::shared_ptr<EngineContext> ctx;
};
template <typename scalar, int nvars> class ZeroGradients : public GradientScheme<scalar,nvars> { public:
	ZeroGradients(const UMesh<scalar,2> *const mesh, const scalar *const _rc, const scalar *const _rcbp)
		: GradientScheme<scalar,nvars>(mesh, _rc, _rcbp, FVG_GRAD_ZERO) {} };
template <typename scalar, int nvars> class GreenGaussGradients : public GradientScheme<scalar,nvars> { public:
	GreenGaussGradients(const UMesh<scalar,2> *const mesh, const scalar *const _rc, const scalar *const _rcbp)
		: GradientScheme<scalar,nvars>(mesh, _rc, _rcbp, FVG_GRAD_GREENGAUSS) {} };
template <typename scalar, int nvars> class WeightedLeastSquaresGradients : public GradientScheme<scalar,nvars> { public:
	WeightedLeastSquaresGradients(const UMesh<scalar,2> *const mesh, const scalar *const _rc, const scalar *const _rcbp)
		: GradientScheme<scalar,nvars>(mesh, _rc, _rcbp, FVG_GRAD_LEASTSQUARES) {} };

/// Reference: create_const_gradientscheme (utilities/afactory.cpp:100-133)
template <typename scalar, int nvars>
const GradientScheme<scalar,nvars>* create_const_gradientscheme(const std::string& type, const UMesh<scalar,NDIM> *const m,
                                                                const scalar *const rc, const scalar *const rcbp)
{
	if(type == "LEASTSQUARES") return new WeightedLeastSquaresGradients<scalar,nvars>(m, rc, rcbp);
	if(type == "GREENGAUSS") return new GreenGaussGradients<scalar,nvars>(m, rc, rcbp);
	return new ZeroGradients<scalar,nvars>(m, rc, rcbp);
}

// ---------------------------------------------------------------------------------------- reconstruction

template <typename scalar, int nvars>
class SolutionReconstruction {
	static_assert(nvars == NVARS, "the engine reconstructs the four flow variables");
public:
	SolutionReconstruction(const UMesh<scalar,2> *const mesh, const scalar *const c_centres, const scalar *const c_centres_ghost,
	                       const amat::Array2d<scalar>& gauss_r, const int recon_id, const freal param = 0.0)
		: m(mesh), ri(c_centres), ribp(c_centres_ghost), gr(gauss_r), ctx(make_plugin_context(mesh, FVG_GRAD_ZERO, recon_id, param)) {}
	virtual ~SolutionReconstruction() {}
	/// Left/right face values from cell states, ghost states and cell gradients (spatial/areconstruction.hpp:37-41).
	/// uface_left/right: naface x nvars host arrays; right values are written for interior faces only.
	virtual void compute_face_values(const MVector<scalar>& unknowns, const amat::Array2dView<scalar> unknow_ghost,
	                                 const scalar *const grads, amat::Array2dMutableView<scalar> uface_left,
	                                 amat::Array2dMutableView<scalar> uface_right) const {
		const size_t ne = m->gnelem(), nb = m->gnbface(), nf = m->gnaface();
		DeviceScratch du(ne*nvars, unknowns.data()), dg(std::max<size_t>(nb, 1)*nvars, nb ? unknow_ghost.data() : nullptr),
			dgr(ne*NDIM*nvars, grads), dl(nf*nvars, uface_left.data()), dr(nf*nvars, uface_right.data());
		fvg_throw(fvg_face_values(ctx->flow, du.p, dg.p, dgr.p, dl.p, dr.p, nullptr), "SolutionReconstruction::compute_face_values");
		dl.download(uface_left.data()); dr.download(uface_right.data());
	}
protected:
	const UMesh<scalar,2> *const m;
	const scalar *const ri;
	const scalar *const ribp;
	const amat::Array2d<scalar>& gr;
	std::shared_ptr<EngineContext> ctx;
};
#define FVENS_RECON_CLASS(Name, ID) \
	template <typename scalar, int nvars> class Name : public SolutionReconstruction<scalar,nvars> { public: \
		Name(const UMesh<scalar,2> *const mesh, const scalar *const c, const scalar *const cg, const amat::Array2d<scalar>& g) \
			: SolutionReconstruction<scalar,nvars>(mesh, c, cg, g, ID) {} };
FVENS_RECON_CLASS(LinearUnlimitedReconstruction, FVG_RECON_NONE)
FVENS_RECON_CLASS(MUSCLVanAlbada, FVG_RECON_VANALBADA)
FVENS_RECON_CLASS(BarthJespersenLimiter, FVG_RECON_BARTHJESPERSEN)
#undef FVENS_RECON_CLASS
template <typename scalar, int nvars> class WENOReconstruction : public SolutionReconstruction<scalar,nvars> { public:
	WENOReconstruction(const UMesh<scalar,2> *const mesh, const scalar *const c, const scalar *const cg,
	                   const amat::Array2d<scalar>& g, const freal central_weight)
		: SolutionReconstruction<scalar,nvars>(mesh, c, cg, g, FVG_RECON_WENO, central_weight) {} };
template <typename scalar, int nvars> class VenkatakrishnanLimiter : public SolutionReconstruction<scalar,nvars> { public:
	VenkatakrishnanLimiter(const UMesh<scalar,2> *const mesh, const scalar *const c, const scalar *const cg,
	                       const amat::Array2d<scalar>& g, const freal k_param)
		: SolutionReconstruction<scalar,nvars>(mesh, c, cg, g, FVG_RECON_VENKATAKRISHNAN, k_param) {} };

/// Reference: create_const_reconstruction (utilities/afactory.cpp:160-217); unknown key prints and returns nullptr
template <typename scalar, int nvars>
const SolutionReconstruction<scalar,nvars>* create_const_reconstruction(const std::string& type, const UMesh<scalar,NDIM> *const m,
	const scalar *const rc, const scalar *const rcbp, const amat::Array2d<scalar>& gr, const freal param)
{
	if(type == "NONE") return new LinearUnlimitedReconstruction<scalar,nvars>(m, rc, rcbp, gr);
	if(type == "WENO") return new WENOReconstruction<scalar,nvars>(m, rc, rcbp, gr, param);
	if(type == "VANALBADA") return new MUSCLVanAlbada<scalar,nvars>(m, rc, rcbp, gr);
	if(type == "BARTHJESPERSEN") return new BarthJespersenLimiter<scalar,nvars>(m, rc, rcbp, gr);
	if(type == "VENKATAKRISHNAN") return new VenkatakrishnanLimiter<scalar,nvars>(m, rc, rcbp, gr, param);
	std::cout << " !ReconstructionFactory: Invalid reconstruction!!\n";
	return nullptr;
}

// ---------------------------------------------------------------------------------------- spatial discretisation

/// Reference: Spatial<scalar,nvars> (spatial/aspatial.hpp:36-170). The constructor computes the same host
/// geometry (cell centres, face midpoints, ghost-cell centres; spatial/aspatial.cpp:37-119).
template <typename scalar, int nvars>
class Spatial {
public:
	Spatial(const UMesh<scalar,NDIM> *const mesh) : m(mesh) {
		rch.assign((size_t)m->gnelem()*NDIM, 0.0);
		m->compute_cell_centres(rch.data());
		gr.resize(m->gnaface(), NDIM);
		for(fint f = 0; f < m->gnaface(); f++)
			for(int d = 0; d < NDIM; d++)
				gr(f,d) = (m->gcoords(m->gintfac(f,2),d) + m->gcoords(m->gintfac(f,3),d))/2.0;
		rcbp.resize(m->gnbface(), NDIM);
		for(fint f = 0; f < m->gnbface(); f++)
			for(int d = 0; d < NDIM; d++)
				rcbp(f,d) = 2.0*gr(f,d) - rch[(size_t)m->gintfac(f,0)*NDIM+d];
	}
	virtual ~Spatial() {}
	/// Adds -r(u) into `residual` (caller zeroes) and, if asked, writes the local time steps (aspatial.hpp:62-63)
	virtual StatusCode compute_residual(const Vec u, Vec residual, const bool gettimesteps, Vec dtm) const = 0;
	virtual void getGradients(const Vec u, GradBlock_t<freal,NDIM,nvars> *const grads) const = 0;
	StatusCode assemble_jacobian(const Vec, void*) const {
		throw UnsupportedOptionError("Spatial::assemble_jacobian is not on the explicit residual path");
	}
	const UMesh<scalar,2>* mesh() const { return m; }
protected:
	const UMesh<scalar,NDIM> *const m;
	std::vector<scalar> rch;             ///< cell centres [nelem][NDIM]
	amat::Array2d<scalar> rcbp;          ///< ghost-cell centres of the physical boundary faces
	amat::Array2d<scalar> gr;            ///< face midpoints
};

template <typename scalar>
class FlowFV_base : public Spatial<scalar,NVARS> {
public:
	/// tile_cells / reorder are engine knobs with no reference counterpart; the defaults are the tuned ones
	FlowFV_base(const UMesh<scalar,NDIM> *const mesh, const FlowPhysicsConfig& pconf, const FlowNumericsConfig& nconf,
	            const int bnd_policy = FVG_BND_GHOST, const int reorder = FVG_REORDER_HILBERT, const int tile_cells = 256,
	            const int *const cell_rank = nullptr, const int rank = 0, const int nranks = 1, const int device = -1)
		: Spatial<scalar,NVARS>(mesh), pconfig(pconf), nconfig(nconf),
		  physics(pconf.gamma, pconf.Minf, pconf.Tinf, pconf.Reinf, pconf.Pr),
		  uinf(physics.compute_freestream_state(pconf.aoa)),
		  inviflux(create_const_inviscidflux<scalar>(nconf.conv_numflux, &physics)),
		  bcs(create_const_flowBCs<scalar>(pconf.bcconf, physics, uinf))
	{
		if(!inviflux) throw UnsupportedOptionError("unknown inviscid flux " + nconf.conv_numflux);
		const int rid = reconstruction_id(nconf.reconstruction);
		if(rid < 0) throw UnsupportedOptionError("unknown reconstruction " + nconf.reconstruction);
		fvg_numerics n; n.flux = inviflux->abi_id(); n.gradient = gradient_id(nconf.gradientscheme); n.reconstruction = rid;
		n.limiter_param = nconf.limiter_param; n.order2 = nconf.order2; n.bnd_policy = bnd_policy;
		std::vector<fvg_bc> b;
		for(auto it = bcs.begin(); it != bcs.end(); ++it) b.push_back(it->second->abi());
		ctx.reset(new EngineContext(mesh, physics.abi(pconf.aoa, pconf.viscous_sim, pconf.const_visc), n, b, reorder, tile_cells,
		                            cell_rank, rank, nranks, device));
	}
	virtual ~FlowFV_base() {
		delete inviflux;
		for(auto it = bcs.begin(); it != bcs.end(); ++it) delete it->second;
	}

	virtual StatusCode compute_residual(const Vec u, Vec residual, const bool gettimesteps, Vec timesteps) const = 0;

	/// Gradients of the CONSERVED variables (spatial/flow_spatial.cpp:96-112)
	void getGradients(const Vec u, GradBlock_t<scalar,NDIM,NVARS> *const grads) const {
		const size_t ne = this->m->gnelem();
		DeviceScratch dg(ne*NDIM*NVARS);
		if(u->place == VEC_DEVICE) fvg_throw(fvg_get_gradients(ctx->flow, u->dev, dg.p, nullptr), "getGradients");
		else { DeviceScratch du(ne*NVARS, u->host.data()); fvg_throw(fvg_get_gradients(ctx->flow, du.p, dg.p, nullptr), "getGradients"); }
		dg.download(reinterpret_cast<double*>(grads));
	}

	/// Cl, Cdp, Cdf over the faces with marker iwbcm (spatial/flow_spatial.cpp:131-310); the per-face
	/// Cp/Cf table of the reference's `output` argument is not produced here
	std::tuple<scalar,scalar,scalar> computeSurfaceData(const amat::Array2dView<scalar> u,
	                                                    const GradBlock_t<scalar,NDIM,NVARS> *const grad, const int iwbcm,
	                                                    MVector<scalar>& /*output*/) const {
		const size_t ne = this->m->gnelem();
		DeviceScratch du(ne*NVARS, u.data()), dg(ne*NDIM*NVARS, reinterpret_cast<const double*>(grad));
		double out[3];
		fvg_throw(fvg_surface_data(ctx->flow, du.p, dg.p, iwbcm, out), "computeSurfaceData");
		return std::make_tuple(out[0], out[1], out[2]);
	}

	/// Reference: compute_entropy_cell (spatial/aoutput.cpp:28-63)
	scalar compute_entropy_cell(const Vec u) const {
		double e = 0;
		if(u->place == VEC_DEVICE) fvg_throw(fvg_entropy_error(ctx->flow, u->dev, &e), "compute_entropy_cell");
		else { DeviceScratch du((size_t)this->m->gnelem()*NVARS, u->host.data()); fvg_throw(fvg_entropy_error(ctx->flow, du.p, &e), "compute_entropy_cell"); }
		return e;
	}

	const std::array<freal,NVARS>& freestream() const { return uinf; }
	fvg_flow* engine_flow() const { return ctx->flow; }
	fvg_mesh* engine_mesh() const { return ctx->mesh; }

protected:
	const FlowPhysicsConfig pconfig;
	const FlowNumericsConfig nconfig;
	const IdealGasPhysics<scalar> physics;
	const std::array<freal,NVARS> uinf;
	const InviscidFlux<scalar> *const inviflux;
	const std::map<int,const FlowBC<scalar>*> bcs;
	std::unique_ptr<EngineContext> ctx;

	/// Reference: compute_boundary_states (spatial/flow_spatial.cpp:74-93); host arrays [nbface][NVARS]
	void compute_boundary_states(const scalar *const instates, scalar *const ghoststates) const {
		const size_t nb = this->m->gnbface();
		if(nb == 0) return;
		DeviceScratch di(nb*NVARS, instates), dgs(nb*NVARS);
		fvg_throw(fvg_boundary_states(ctx->flow, di.p, dgs.p, nullptr), "compute_boundary_states");
		dgs.download(ghoststates);
	}
};

template <typename scalar, bool secondOrderRequested, bool constVisc>
class FlowFV : public FlowFV_base<scalar> {
public:
	FlowFV(const UMesh<scalar,NDIM> *const mesh, const FlowPhysicsConfig& pconf, const FlowNumericsConfig& nconf,
	       const int bnd_policy = FVG_BND_GHOST, const int reorder = FVG_REORDER_HILBERT, const int tile_cells = 256,
	       const int *const cell_rank = nullptr, const int rank = 0, const int nranks = 1, const int device = -1)
		: FlowFV_base<scalar>(mesh, fix_p(pconf), fix_n(nconf), bnd_policy, reorder, tile_cells, cell_rank, rank, nranks, device) {}

	/// spatial/flow_spatial.cpp:637-816. Host Vecs: upload u (and the residual it adds into), kernels,
	/// download residual and time steps. Device Vecs: kernels only, asynchronous on the default stream.
	StatusCode compute_residual(const Vec u, Vec residual, const bool gettimesteps, Vec timesteps) const {
		if(!u || !residual || (gettimesteps && !timesteps)) return FVG_ERR_INVALID;
		if(u->place != residual->place || (gettimesteps && timesteps->place != u->place)) return FVG_ERR_INVALID;
		fvg_flow *const f = this->ctx->flow;
		if(u->place == VEC_HOST)
			return fvg_residual_host(f, u->host.data(), residual->host.data(), 1, gettimesteps,
			                         gettimesteps ? timesteps->host.data() : nullptr);
		return fvg_residual(f, u->dev, residual->dev, 1, gettimesteps, gettimesteps ? timesteps->dev : nullptr, nullptr);
	}

	void compute_local_jacobian_interior(const fint, const freal *const, const freal *const, void*, void*) const {
		throw UnsupportedOptionError("FlowFV::compute_local_jacobian_interior is not on the explicit residual path");
	}
	void compute_local_jacobian_boundary(const fint, const freal *const, void*) const {
		throw UnsupportedOptionError("FlowFV::compute_local_jacobian_boundary is not on the explicit residual path");
	}
private:
	static FlowPhysicsConfig fix_p(FlowPhysicsConfig p) { p.const_visc = constVisc; return p; }
	static FlowNumericsConfig fix_n(FlowNumericsConfig n) { n.order2 = secondOrderRequested; return n; }
};

/// Reference: create_const_flowSpatialDiscretization (utilities/afactory.cpp:252-275)
template <typename scalar>
const FlowFV_base<scalar>* create_const_flowSpatialDiscretization(const UMesh<scalar,NDIM> *const m,
                                                                  const FlowPhysicsConfig& pconf, const FlowNumericsConfig& nconf)
{
	if(nconf.order2) {
		if(pconf.const_visc) return new FlowFV<scalar,true,true>(m, pconf, nconf);
		return new FlowFV<scalar,true,false>(m, pconf, nconf);
	}
	if(pconf.const_visc) return new FlowFV<scalar,false,true>(m, pconf, nconf);
	return new FlowFV<scalar,false,false>(m, pconf, nconf);
}

// ---------------------------------------------------------------------------------------- multi-GPU

/// All-gather used once at set-up: every rank contributes `bytes` bytes at `mine`, `all` receives nranks*bytes in rank
/// order. The transport is the caller's (MPI_Allgather in an MPI program - exactly what the reference's drivers have at
/// hand - a socket, shared files ...); nothing else of the data path goes through it.
typedef std::function<void(const void *mine, void *all, size_t bytes)> AllGather;

/// FlowFV on one rank's subdomain of a mesh distributed over the GPUs of one NVLink/NVSwitch box: the reference's
/// FlowFV over a ReplicatedGlobalMeshPartitioner subdomain (mesh/meshpartitioning.cpp:24-159) with its L2TraceVector
/// exchange (linalg/tracevector.cpp:214-325) and ghost updates (spatial/flow_spatial.cpp:711-788), here the fused
/// evaluation fvg_dist_*. Every rank passes the same GLOBAL mesh and cell -> rank map (fvg_partition_sfc / _rcb, a Scotch
/// map read by fvg_partition_read_scotch_map, or the reference's trivial partition). Device Vecs in device order:
/// [ncell() own rows (+ ghost rows, ignored)]; global_cell(i) says which cell of the global mesh row i is.
template <typename scalar, bool secondOrderRequested, bool constVisc>
class DistributedFlowFV : public FlowFV<scalar,secondOrderRequested,constVisc> {
public:
	DistributedFlowFV(const UMesh<scalar,NDIM> *const globalmesh, const FlowPhysicsConfig& pconf, const FlowNumericsConfig& nconf,
	                  const std::vector<int>& cell_rank, const int rank, const int nranks, const int device, const AllGather& allgather,
	                  const int bnd_policy = FVG_BND_GHOST, const int tile_cells = 256)
		: FlowFV<scalar,secondOrderRequested,constVisc>(globalmesh, pconf, nconf, bnd_policy, FVG_REORDER_HILBERT, tile_cells,
		                                                cell_rank.data(), rank, nranks, device), nr(nranks)
	{
		if((fint)cell_rank.size() != globalmesh->gnelem()) throw std::runtime_error("DistributedFlowFV: cell_rank needs one entry per cell of the global mesh");
		fvg_throw(fvg_dist_create(this->ctx->flow, &dist), "fvg_dist_create");
		fvg_mesh_info info;
		fvg_throw(fvg_mesh_get_info(this->ctx->mesh, &info), "fvg_mesh_get_info");
		nown = info.ncell; nghost = info.nghost;
		// set-up exchange: 64-byte window handle + this rank's receive counts
		std::vector<int> sc((size_t)nranks), rc((size_t)nranks);
		fvg_throw(fvg_mesh_halo_lists(this->ctx->mesh, sc.data(), rc.data(), nullptr), "fvg_mesh_halo_lists");
		const size_t rec = 64 + sizeof(int)*(size_t)nranks;
		std::vector<unsigned char> mine(rec), all(rec*(size_t)nranks);
		fvg_throw(fvg_dist_ipc_handle(dist, mine.data()), "fvg_dist_ipc_handle");
		std::memcpy(mine.data() + 64, rc.data(), sizeof(int)*(size_t)nranks);
		allgather(mine.data(), all.data(), rec);
		std::vector<unsigned char> handles(64*(size_t)nranks);
		std::vector<int> counts((size_t)nranks*(size_t)nranks);
		for(int r = 0; r < nranks; r++) {
			std::memcpy(handles.data() + 64*(size_t)r, all.data() + rec*(size_t)r, 64);
			std::memcpy(counts.data() + (size_t)r*(size_t)nranks, all.data() + rec*(size_t)r + 64, sizeof(int)*(size_t)nranks);
		}
		fvg_throw(fvg_dist_connect(dist, handles.data(), counts.data()), "fvg_dist_connect");
		perm.resize((size_t)nown + (size_t)nghost);
		fvg_throw(fvg_mesh_permutation(this->ctx->mesh, perm.data()), "fvg_mesh_permutation");
	}
	~DistributedFlowFV() { if(dist) fvg_dist_destroy(dist); }

	fint ncell() const { return nown; }
	fint nghostcell() const { return nghost; }
	fint global_cell(const fint i) const { return perm[(size_t)i]; }
	/// This rank's rows of a global cell array (host), `width` doubles per cell
	std::vector<scalar> restrict_to_rank(const scalar *const global, const int width) const {
		std::vector<scalar> loc((size_t)nown*(size_t)width);
		for(fint i = 0; i < nown; i++) for(int k = 0; k < width; k++) loc[(size_t)i*width+k] = global[(size_t)perm[(size_t)i]*width+k];
		return loc;
	}

	/// spatial/flow_spatial.cpp:637-816 on the subdomain; device Vecs ([ncell()] own rows at least), adds into `residual`
	StatusCode compute_residual(const Vec u, Vec residual, const bool gettimesteps, Vec timesteps) const {
		if(!u || !residual || (gettimesteps && !timesteps)) return FVG_ERR_INVALID;
		if(u->place != VEC_DEVICE || residual->place != VEC_DEVICE || (gettimesteps && timesteps->place != VEC_DEVICE)) return FVG_ERR_INVALID;
		return fvg_dist_residual(dist, u->dev, residual->dev, 1, gettimesteps, gettimesteps ? timesteps->dev : nullptr, nullptr);
	}
	/// FVG_ERR_COMM (as Engine_error) once a neighbour rank has stopped delivering rows
	void check_neighbours() const { unsigned long long t = 0; fvg_throw(fvg_dist_status(dist, &t), "DistributedFlowFV"); }
	fvg_dist* engine_dist() const { return dist; }
	int nranks() const { return nr; }
private:
	fvg_dist *dist = nullptr;
	fint nown = 0, nghost = 0;
	int nr = 1;
	std::vector<int> perm;
};

// ---------------------------------------------------------------------------------------- pseudo-time

struct SteadySolverConfig {
	bool lognres; std::string logfile; bool write_final_lin_sys;
	freal cflinit, cflfin; int rampstart, rampend;
	freal tol; int maxiter; int linmaxiterstart, linmaxiterend;
};
struct SteadyStepMonitor { int step; float rmsres, absrmsres, odewalltime, linwalltime; int linits; float cfl; };
struct TimingData {
	fint nelem = 0; int num_threads = 1;
	double lin_walltime = 0, lin_cputime = 0, ode_walltime = 0, ode_cputime = 0;
	int total_lin_iters = 0, avg_lin_iters = 0, num_timesteps = 0;
	bool converged = false;
	double precsetup_walltime = 0, precapply_walltime = 0, prec_cputime = 0;
	std::vector<SteadyStepMonitor> convhis;
};

template <int nvars>
class SteadySolver {
public:
	SteadySolver(const Spatial<freal,nvars> *const spatial, const SteadySolverConfig& conf) : space(spatial), config(conf) {}
	virtual ~SteadySolver() {}
	TimingData getTimingData() const { return tdata; }
	virtual StatusCode solve(Vec u) = 0;
protected:
	const Spatial<freal,nvars> *const space;
	const SteadySolverConfig& config;
	TimingData tdata;
};

/// Reference: SteadyForwardEulerSolver (ode/aodesolver.cpp:136-282). With the engine's FlowFV the whole loop
/// runs on the device (fused residual + local dt + update + norm per step; one host read of the norm per step,
/// as the reference's MPI_Allreduce). With any other Spatial the generic loop below drives compute_residual.
template <int nvars>
class SteadyForwardEulerSolver : public SteadySolver<nvars> {
public:
	SteadyForwardEulerSolver(const Spatial<freal,nvars> *const euler, const Vec /*x*/, const SteadySolverConfig& conf)
		: SteadySolver<nvars>(euler, conf) {}

	StatusCode solve(Vec u) {
		const SteadySolverConfig& config = this->config;
		TimingData& tdata = this->tdata;
		const UMesh<freal,NDIM> *const m = this->space->mesh();
		tdata.nelem = m->gnelem();
		if(config.maxiter <= 0) { std::cout << " SteadyForwardEulerSolver: solve(): No iterations to be done.\n"; return 0; }
		const auto t0 = std::chrono::steady_clock::now();
		int steps = 0;
		std::vector<double> hist((size_t)config.maxiter, 0.0);
		int code;
		const FlowFV_base<freal> *const eng = dynamic_cast<const FlowFV_base<freal>*>(this->space);
		fvg_dist *const dist = distributed_engine(this->space);
		if(dist) {
			// partitioned mesh: the loop of aodesolver.cpp:177-251 with its ghost update (:212) and MPI_Allreduce (:227)
			// folded into the kernels and the peer windows (fvg_dist_forward_euler_solve); device Vec of own rows
			if(u->place != VEC_DEVICE) throw UnsupportedOptionError("SteadyForwardEulerSolver on a distributed FlowFV takes a device Vec");
			code = fvg_dist_forward_euler_solve(dist, u->dev, config.cflinit, config.tol, config.maxiter, 1, &steps, hist.data());
		}
		else if(eng) {
			// note H4 of SURVEY.md: the reference applies cflinit on every step
			if(u->place == VEC_DEVICE)
				code = fvg_forward_euler_solve(eng->engine_flow(), u->dev, config.cflinit, config.tol, config.maxiter, 1, &steps, hist.data());
			else {
				DeviceScratch du(u->size(), u->host.data());
				code = fvg_forward_euler_solve(eng->engine_flow(), du.p, config.cflinit, config.tol, config.maxiter, 1, &steps, hist.data());
				if(code == FVG_OK || code == FVG_ERR_TOLERANCE || code == FVG_ERR_NUMERICAL) du.download(u->host.data());
			}
		}
		else code = generic_loop(u, steps, hist);
		const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		tdata.ode_walltime += wall; tdata.num_timesteps = steps;
		const double initres = steps > 0 ? hist[0] : 1.0;
		for(int s = 0; s < steps; s++) {
			SteadyStepMonitor mon; mon.step = s+1; mon.absrmsres = (float)hist[s]; mon.rmsres = (float)(hist[s]/initres);
			mon.odewalltime = (float)(wall*(s+1)/steps); mon.linwalltime = 0; mon.linits = 0; mon.cfl = (float)config.cflinit;
			tdata.convhis.push_back(mon);
			if((s+1) % 50 == 0 || s == 0) std::cout << "  SteadyForwardEulerSolver: solve(): Step " << s+1 << ", rel residual " << hist[s]/initres << std::endl;
		}
		if(code == FVG_ERR_NUMERICAL) throw Numerical_error("Steady forward Euler diverged - residual is " + std::to_string(hist[std::max(steps-1,0)]));
		if(code == FVG_ERR_TOLERANCE) { tdata.converged = false; throw Tolerance_error("Steady forward Euler did not converge to specified tolerance!"); }
		fvg_throw(code, "SteadyForwardEulerSolver::solve");
		tdata.converged = true;
		return 0;
	}
private:
	/// the exchange engine behind a DistributedFlowFV of any template flavour (null otherwise)
	static fvg_dist* distributed_engine(const Spatial<freal,nvars> *const sp) {
		if(auto p = dynamic_cast<const DistributedFlowFV<freal,true,true>*>(sp)) return p->engine_dist();
		if(auto p = dynamic_cast<const DistributedFlowFV<freal,true,false>*>(sp)) return p->engine_dist();
		if(auto p = dynamic_cast<const DistributedFlowFV<freal,false,true>*>(sp)) return p->engine_dist();
		if(auto p = dynamic_cast<const DistributedFlowFV<freal,false,false>*>(sp)) return p->engine_dist();
		return nullptr;
	}
	/// aodesolver.cpp:177-251 for a Spatial that is not the engine's: orchestration only, host Vecs
	int generic_loop(Vec u, int& step, std::vector<double>& hist) {
		const SteadySolverConfig& config = this->config;
		const UMesh<freal,NDIM> *const m = this->space->mesh();
		if(u->place != VEC_HOST) return FVG_ERR_UNSUPPORTED;
		Vec r = nullptr, dt = nullptr;
		VecDuplicate(u, &r); VecCreateBlocked(m->gnelem(), 0, 1, VEC_HOST, &dt);
		double resi = 1.0, initres = 1.0; int code = FVG_OK;
		step = 0;
		while(resi/initres > config.tol && step < config.maxiter) {
			VecSet(r, 0.0);
			const int ierr = this->space->compute_residual(u, r, true, dt);
			if(ierr) { code = ierr; break; }
			double sum = 0;
			for(fint i = 0; i < m->gnelem(); i++) {
				for(int k = 0; k < nvars; k++) u->host[(size_t)i*nvars+k] += config.cflinit*dt->host[i]/m->garea(i)*r->host[(size_t)i*nvars+k];
				sum += r->host[(size_t)i*nvars+nvars-1]*r->host[(size_t)i*nvars+nvars-1]*m->garea(i);
			}
			resi = std::sqrt(sum);
			if(step == 0) initres = resi;
			hist[step++] = resi;
			if(!std::isfinite(resi)) { code = FVG_ERR_NUMERICAL; break; }
		}
		// aodesolver.cpp:253-256: reaching maxiter is an error whatever the last residual (as fvg_forward_euler_solve)
		if(code == FVG_OK && step == config.maxiter) code = FVG_ERR_TOLERANCE;
		VecDestroy(&r); VecDestroy(&dt);
		return code;
	}
};

/// Reference: UnsteadySolver (ode/aodesolver.hpp:195-226)
template <int nvars>
class UnsteadySolver {
public:
	UnsteadySolver(const Spatial<freal,nvars> *const spatial, Vec soln, const int temporal_order, const std::string log_file)
		: space(spatial), uvec(soln), order(temporal_order), cputime(0.0), walltime(0.0), logfile(log_file) {}
	virtual ~UnsteadySolver() {}
	std::tuple<double,double> getRunTimes() const { return std::make_tuple(walltime, cputime); }
	virtual StatusCode solve(const freal time) = 0;
protected:
	const Spatial<freal,nvars> *space;
	Vec uvec;
	const int order;
	double cputime, walltime;
	const std::string logfile;
};

/// Reference: TVDRKSolver (ode/aodesolver.hpp:229-258, aodesolver.cpp:647-785), total-variation-diminishing
/// Runge-Kutta up to order 3 with the reference's coefficient table and global time step cfl * min(dtm). The stages
/// are evaluated at the stage state and the update has the sign of the forward-Euler loop - the reference's loop does
/// neither (see fvg_tvdrk_solve in include/fvens_b200.h). With the engine's FlowFV the loop runs on the device; with
/// any other Spatial the generic loop below drives compute_residual on host Vecs.
template <int nvars>
class TVDRKSolver : public UnsteadySolver<nvars> {
public:
	TVDRKSolver(const Spatial<freal,nvars> *const spatial, Vec soln, const int temporal_order, const std::string log_file,
	            const double cfl_num)
		: UnsteadySolver<nvars>(spatial, soln, temporal_order, log_file), cfl(cfl_num), tvdcoeffs(3*std::max(temporal_order,1), 0.0),
		  nsteps(0), phytime(0.0)
	{
		if(temporal_order < 1 || temporal_order > 3) std::cout << "! Temporal order " << temporal_order << " not available!\n";
		else fvg_tvdrk_coefficients(temporal_order, tvdcoeffs.data());
		std::cout << " TVDRKSolver: Initialized TVD RK solver of order " << temporal_order << ", CFL = " << cfl << std::endl;
	}

	StatusCode solve(const freal finaltime) {
		if(this->order < 1 || this->order > 3) throw UnsupportedOptionError("TVDRKSolver: temporal order not available");
		Vec u = this->uvec;
		const auto t0 = std::chrono::steady_clock::now();
		const double c0 = (double)clock()/(double)CLOCKS_PER_SEC;
		int code;
		const FlowFV_base<freal> *const eng = dynamic_cast<const FlowFV_base<freal>*>(this->space);
		if(eng) {
			if(u->place == VEC_DEVICE) code = fvg_tvdrk_solve(eng->engine_flow(), u->dev, this->order, cfl, finaltime, 0, &nsteps, &phytime);
			else {
				DeviceScratch du(u->size(), u->host.data());
				code = fvg_tvdrk_solve(eng->engine_flow(), du.p, this->order, cfl, finaltime, 0, &nsteps, &phytime);
				if(code == FVG_OK || code == FVG_ERR_NUMERICAL) du.download(u->host.data());
			}
		}
		else code = generic_loop(u, finaltime);
		this->walltime += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		this->cputime += (double)clock()/(double)CLOCKS_PER_SEC - c0;
		if(code == FVG_ERR_NUMERICAL) throw Numerical_error("TVDRK solver diverged - dtmin is Nan or inf!");
		fvg_throw(code, "TVDRKSolver::solve");
		std::cout << " TVDRKSolver: solve(): Done, steps = " << nsteps << ", phy time = " << phytime << "\n\n";
		std::cout << " TVDRKSolver: solve(): Time taken by ODE solver:\n";
		std::cout << "                                   CPU time = " << this->cputime << ", wall time = " << this->walltime << std::endl << std::endl;
		if(!this->logfile.empty()) {
			// the reference appends threads, wall and CPU time to the log file (aodesolver.cpp:769-771); 0 threads = no OpenMP
			std::ofstream outf(this->logfile, std::ofstream::app);
			outf << "\t" << 0 << "\t" << this->walltime << "\t" << this->cputime << "\n";
		}
		return 0;
	}

	int numSteps() const { return nsteps; }
	double physicalTime() const { return phytime; }
protected:
	const double cfl;
	std::vector<double> tvdcoeffs;       ///< [order][3]
private:
	int nsteps; double phytime;

	int generic_loop(Vec u, const freal finaltime) {
		const UMesh<freal,NDIM> *const m = this->space->mesh();
		if(u->place != VEC_HOST) return FVG_ERR_UNSUPPORTED;
		const fint n = m->gnelem();
		Vec us = nullptr, r = nullptr, dt = nullptr;
		VecDuplicate(u, &us); VecDuplicate(u, &r); VecCreateBlocked(n, 0, 1, VEC_HOST, &dt);
		us->host = u->host;
		int code = FVG_OK;
		nsteps = 0; phytime = 0.0;
		while(phytime <= finaltime - 1e-12 && code == FVG_OK) {
			double dtmin = 0.0;
			for(int istage = 0; istage < this->order && code == FVG_OK; istage++) {
				VecSet(r, 0.0);
				const int ierr = this->space->compute_residual(us, r, true, dt);
				if(ierr) { code = ierr; break; }
				if(istage == 0) {
					dtmin = n > 0 ? dt->host[0] : 0.0;
					for(fint i = 0; i < n; i++) { if(!(dt->host[i] == dt->host[i])) dtmin = dt->host[i]; else if(dt->host[i] < dtmin) dtmin = dt->host[i]; }
				}
				if(!std::isfinite(dtmin)) { code = FVG_ERR_NUMERICAL; break; }
				const double a = tvdcoeffs[3*istage], b = tvdcoeffs[3*istage+1], c = tvdcoeffs[3*istage+2];
				for(fint i = 0; i < n; i++)
					for(int k = 0; k < nvars; k++) {
						const size_t j = (size_t)i*nvars + k;
						us->host[j] = a*u->host[j] + b*us->host[j] + c*dtmin*cfl/m->garea(i)*r->host[j];
					}
			}
			if(code != FVG_OK) break;
			u->host = us->host;
			nsteps++;
			phytime += dtmin*cfl;
		}
		VecDestroy(&us); VecDestroy(&r); VecDestroy(&dt);
		return code;
	}
};

/// Reference: MatrixFreeSpatialJacobian (linalg/alinalg.hpp:60-110, alinalg.cpp:122-230): the action of the
/// pseudo-time-shifted residual Jacobian on a vector by a finite difference of two residuals. set_state stores
/// (non-owning) the state, the residual compute_residual left for it and the diagonal shift; apply(x, y) costs one
/// residual evaluation on the device. Device Vecs are used in place, host Vecs are staged.
template <int nvars>
class MatrixFreeSpatialJacobian {
public:
	explicit MatrixFreeSpatialJacobian(const Spatial<freal,nvars> *const s, const freal difference_step = 1e-7)
		: spatial(s), eps(difference_step), u(nullptr), res(nullptr), mdt(nullptr) {}

	int set_state(const Vec u_state, const Vec r_state, const Vec dtms) { u = u_state; res = r_state; mdt = dtms; return 0; }

	StatusCode apply(const Vec x, Vec y) const {
		const FlowFV_base<freal> *const eng = dynamic_cast<const FlowFV_base<freal>*>(spatial);
		if(!eng || !u || !res || !mdt || !x || !y) return FVG_ERR_INVALID;
		const size_t n = (size_t)spatial->mesh()->gnelem();
		if(u->place == VEC_DEVICE && x->place == VEC_DEVICE && y->place == VEC_DEVICE && res->place == VEC_DEVICE && mdt->place == VEC_DEVICE)
			return fvg_jacobian_vector_product(eng->engine_flow(), u->dev, res->dev, mdt->dev, x->dev, eps, y->dev, nullptr);
		if(u->place != VEC_HOST || x->place != VEC_HOST || y->place != VEC_HOST || res->place != VEC_HOST || mdt->place != VEC_HOST)
			return FVG_ERR_INVALID;
		DeviceScratch du(n*nvars, u->host.data()), dr(n*nvars, res->host.data()), dm(n, mdt->host.data()), dx(n*nvars, x->host.data()), dy(n*nvars);
		const int rc = fvg_jacobian_vector_product(eng->engine_flow(), du.p, dr.p, dm.p, dx.p, eps, dy.p, nullptr);
		if(rc == 0) dy.download(y->host.data());
		return rc;
	}
private:
	const Spatial<freal,nvars> *const spatial;
	const freal eps;
	Vec u, res, mdt;
};

/// Reference: initializeSystemVector (utilities/casesolvers.cpp:52-69): u := free stream everywhere
inline StatusCode initializeSystemVector(const FlowPhysicsConfig& pconf, const UMesh<freal,NDIM>& m, Vec u)
{
	const IdealGasPhysics<freal> phy(pconf.gamma, pconf.Minf, pconf.Tinf, pconf.Reinf, pconf.Pr);
	const std::array<freal,NVARS> uinf = phy.compute_freestream_state(pconf.aoa);
	std::vector<double> h(u->size());
	for(size_t i = 0; i < h.size()/NVARS; i++) for(int k = 0; k < NVARS; k++) h[i*NVARS+k] = uinf[k];
	(void)m;
	return VecCopyFromHost(u, h.data());
}

}
#endif
