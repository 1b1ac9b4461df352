/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Tier-E reference build: the WHOLE chain in the reference's own object code - its mesh class and readers
 * (mesh/mesh.cpp, mesh/meshreaders.cpp: Gmsh 2 and SU2 readers, boundary-face orientation, topology, areas, face
 * metrics), its FlowFV::compute_residual with everything under it (the tier-C sources), its explicit pseudo-time
 * solver and output unit (ode/aodesolver.cpp, spatial/aoutput.cpp) - compiled UNMODIFIED, in place from
 * /root/reference/src against ref_shim_b/ (Eigen-lite, serial PETSc Vec, single-process MPI, boost::split, boost::bimap).
 * Also its mesh partitioner (mesh/meshpartitioning.cpp: restrictMeshToPartitions, the subdomain + connectivity-face
 * construction every rank of an MPI run performs), driven rank by rank through a settable rank/size in the MPI stand-in.
 * No mesh stand-in here: from the mesh file to the residual every instruction is the reference's. This file adds the
 * mesh construction sequence of the reference's constructMesh / preprocessMesh for one rank without reordering
 * (mesh/ameshutils.cpp:97-153, 39-96), createSystemVector, and a C interface; it contains no reference code.
 * Built serially into oracle/_ref/libfvens_ref_e.so.
 */
#include "mesh/mesh.cpp"
#include "mesh/meshreaders.cpp"
#include "mesh/meshpartitioning.cpp"
#include "ref_sources_spatial.hpp"
#include "spatial/aoutput.cpp"
#include "ode/aodesolver.cpp"
#include <sstream>

namespace fvens {
StatusCode createSystemVector(const UMesh<freal,NDIM> *const m, const int nvars, Vec *const v)
{
	*v = new _p_Vec;
	(*v)->nlocal = m->gnelem()*nvars; (*v)->nghost = 0;
	(*v)->a.assign((size_t)(*v)->nlocal, 0.0);
	return 0;
}
}

using namespace fvens;

namespace {

struct RefCase {
	std::unique_ptr<UMesh<freal,NDIM>> m;
	std::unique_ptr<const Spatial<freal,NVARS>> prob;
	_p_Vec uv, rv, dv;
};

/// constructMesh + preprocessMesh for a serial run without reordering: orientation, topology, areas, face data
void finish_mesh(UMesh<freal,NDIM>& m)
{
	m.correctBoundaryFaceOrientation();
	m.compute_topological();
	m.compute_areas();
	m.compute_face_data();
}

void make_flow(RefCase *h, const double *phys, const char *flux, const char *gradient, const char *recon, double limiter_param,
               int order2, int viscous, int const_visc, int nbc, const int *bc_tag_type, const double *bc_vals)
{
	std::vector<FlowBCConfig> bcs;
	for(int i = 0; i < nbc; i++) {
		FlowBCConfig c;
		c.bc_tag = bc_tag_type[2*i]; c.bc_type = static_cast<BCType>(bc_tag_type[2*i+1]);
		c.bc_vals = {bc_vals[2*i], bc_vals[2*i+1]};
		bcs.push_back(c);
	}
	const FlowPhysicsConfig pconf { phys[0], phys[1], phys[2], phys[3], phys[4], phys[5], viscous != 0, const_visc != 0, bcs };
	const FlowNumericsConfig nconf { flux, flux, gradient, recon, limiter_param, order2 != 0 };
	const UMesh<freal,NDIM> *const m = h->m.get();
	if(order2) { if(const_visc) h->prob.reset(new FlowFV<freal,true,true>(m, pconf, nconf)); else h->prob.reset(new FlowFV<freal,true,false>(m, pconf, nconf)); }
	else { if(const_visc) h->prob.reset(new FlowFV<freal,false,true>(m, pconf, nconf)); else h->prob.reset(new FlowFV<freal,false,false>(m, pconf, nconf)); }
	const size_t ne = m->gnelem();
	h->uv.a.assign(ne*NVARS, 0.0); h->uv.nlocal = (PetscInt)(ne*NVARS); h->uv.nghost = 0;
	h->rv = h->uv;
	h->dv.a.assign(ne, 0.0); h->dv.nlocal = (PetscInt)ne; h->dv.nghost = 0;
}

}

extern "C" {

/// readMesh (Gmsh 2 .msh or .su2 by extension) + the reference's preprocessing; NULL on failure
void* ref_e_mesh_read(const char *path)
{
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	RefCase *h = nullptr;
	try {
		h = new RefCase;
		h->m.reset(new UMesh<freal,NDIM>(readMesh(path)));
		finish_mesh(*h->m);
	} catch(std::exception&) { delete h; h = nullptr; }
	std::cout.rdbuf(old);
	return h;
}

/// The same from arrays: coords [npoin][2], nnode [nelem], inpoel [nelem][maxnnode] (-1 padded), bface [nbface][3] (two nodes, tag)
void* ref_e_mesh_from_arrays(int npoin, const double *coords, int nelem, int maxnnode, const int *nnode, const int *inpoel,
                             int nbface, const int *bface)
{
	MeshData md;
	md.npoin = npoin; md.nelem = nelem; md.nbface = nbface; md.maxnnode = maxnnode; md.maxnfael = maxnnode; md.nnofa = 2;
	md.nbtag = 1; md.ndtag = 0;
	md.nnode.assign(nnode, nnode + nelem); md.nfael.assign(nnode, nnode + nelem);
	md.coords.resize(npoin, NDIM);
	for(int i = 0; i < npoin; i++) for(int d = 0; d < NDIM; d++) md.coords(i,d) = coords[(size_t)i*NDIM+d];
	md.inpoel.resize(nelem, maxnnode);
	for(int i = 0; i < nelem; i++) for(int j = 0; j < maxnnode; j++) md.inpoel(i,j) = inpoel[(size_t)i*maxnnode+j];
	if(nbface > 0) md.bface.resize(nbface, 3);
	for(int i = 0; i < nbface; i++) for(int j = 0; j < 3; j++) md.bface(i,j) = bface[(size_t)i*3+j];
	md.vol_regions.resize(nelem, 0);
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	RefCase *h = nullptr;
	try {
		h = new RefCase;
		h->m.reset(new UMesh<freal,NDIM>(md));
		finish_mesh(*h->m);
	} catch(std::exception&) { delete h; h = nullptr; }
	std::cout.rdbuf(old);
	return h;
}

namespace {
/// restrictMeshToPartitions with a given cell -> rank map (the reference's partitioners only differ in how they fill it)
struct GivenPartitioner : public ReplicatedGlobalMeshPartitioner {
	GivenPartitioner(const UMesh<freal,NDIM>& gm, const int *dist) : ReplicatedGlobalMeshPartitioner(gm) { elemdist.assign(dist, dist + gm.gnelem()); }
	void compute_partition() {}
};
struct TrivialProbe : public TrivialReplicatedGlobalMeshPartitioner {
	using TrivialReplicatedGlobalMeshPartitioner::TrivialReplicatedGlobalMeshPartitioner;
	const std::vector<int>& dist() const { return elemdist; }
};
}

/// TrivialReplicatedGlobalMeshPartitioner::compute_partition for `nranks` ranks: cell -> rank into dist [nelem]
void ref_e_trivial_partition(void *hv, int nranks, int *dist)
{
	RefCase *h = static_cast<RefCase*>(hv);
	mpi_lite_size = nranks; mpi_lite_rank = 0;
	TrivialProbe p(*h->m);
	p.compute_partition();
	std::copy(p.dist().begin(), p.dist().end(), dist);
	mpi_lite_size = 1;
}

/// The subdomain mesh of `rank` as the reference builds it: restrictMeshToPartitions on the global mesh of hv (which
/// must have its topology), then the preprocessing. Returns a new handle (mesh only) or NULL.
void* ref_e_restrict_to_rank(void *hv, const int *dist, int nranks, int rank)
{
	RefCase *g = static_cast<RefCase*>(hv);
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	RefCase *h = nullptr;
	mpi_lite_size = nranks; mpi_lite_rank = rank;
	try {
		GivenPartitioner p(*g->m, dist);
		h = new RefCase;
		h->m.reset(new UMesh<freal,NDIM>(p.restrictMeshToPartitions()));
		h->m->compute_topological();
		h->m->compute_areas();
		h->m->compute_face_data();
	} catch(std::exception&) { delete h; h = nullptr; }
	mpi_lite_size = 1; mpi_lite_rank = 0;
	std::cout.rdbuf(old);
	return h;
}

/// getCellAdjLists of the reference (mesh/meshpartitioning.cpp:376-430): the CSR graph it would hand to Scotch.
/// Returns the number of adjacency entries; ptrs [nelem+1], store [that number] (either may be NULL).
int ref_e_cell_adjacency(void *hv, int *ptrs, int *store)
{
	RefCase *h = static_cast<RefCase*>(hv);
	const ListOfArrays<fint> loa = getCellAdjLists(*h->m);
	if(ptrs) std::copy(loa.ptrs.begin(), loa.ptrs.end(), ptrs);
	if(store) std::copy(loa.store.begin(), loa.store.end(), store);
	return (int)loa.store.size();
}

/// The body of the reference's utilities/convertformat.cpp main (the file itself is a program, so its three statements
/// are restated): readMesh -> UMesh(md) -> writeGmsh2 | writeMeshToVtu, no preprocessing in between
int ref_e_convertformat(const char *inmesh, const char *outmesh, const char *outformat)
{
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	int rc = 0;
	try {
		const MeshData md = readMesh(inmesh);
		const UMesh<freal,NDIM> m(md);
		if(std::string(outformat) == "msh") m.writeGmsh2(outmesh);
		else if(std::string(outformat) == "vtu") writeMeshToVtu(outmesh, m);
		else rc = -1;
	} catch(std::exception&) { rc = 1; }
	std::cout.rdbuf(old);
	return rc;
}

/// writeMeshToVtu of the reference (spatial/aoutput.cpp:557-615)
int ref_e_write_mesh_vtu(void *hv, const char *path)
{
	RefCase *h = static_cast<RefCase*>(hv);
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	writeMeshToVtu(path, *h->m);
	std::cout.rdbuf(old);
	return 0;
}

/// UMesh::writeGmsh2 of the reference (mesh/mesh.cpp:205-286)
int ref_e_write_gmsh2(void *hv, const char *path)
{
	RefCase *h = static_cast<RefCase*>(hv);
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	h->m->writeGmsh2(path);
	std::cout.rdbuf(old);
	return 0;
}

/// number of connectivity faces; glob [nelem] global cell ids; conn [nconn][5] = gconnface(i, 0..4)
int ref_e_connectivity(void *hv, int *glob, int *conn)
{
	const UMesh<freal,NDIM>& m = *static_cast<RefCase*>(hv)->m;
	if(glob) for(fint i = 0; i < m.gnelem(); i++) glob[i] = m.gglobalElemIndex(i);
	if(conn) for(fint i = 0; i < m.gnConnFace(); i++) for(int j = 0; j < 5; j++) conn[(size_t)i*5+j] = m.gconnface(i,j);
	return m.gnConnFace();
}

/// UMesh::reorder_cells (new cell i = old cell perm[i], mesh/mesh.cpp:85-99) followed by the preprocessing again
void ref_e_mesh_reorder(void *hv, const int *perm)
{
	RefCase *h = static_cast<RefCase*>(hv);
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	h->m->reorder_cells(perm);
	h->m->compute_topological();
	h->m->compute_areas();
	h->m->compute_face_data();
	std::cout.rdbuf(old);
}

void ref_e_destroy(void *hv) { delete static_cast<RefCase*>(hv); }

/// sizes = {npoin, nelem, nbface, naface, maxnnode}
void ref_e_mesh_sizes(void *hv, int *sizes)
{
	const UMesh<freal,NDIM>& m = *static_cast<RefCase*>(hv)->m;
	sizes[0] = m.gnpoin(); sizes[1] = m.gnelem(); sizes[2] = m.gnbface(); sizes[3] = m.gnaface(); sizes[4] = m.gmaxnfael();
}

/// The derived arrays of the reference's UMesh; padded entries of 4-wide cell arrays are -1
void ref_e_mesh_get(void *hv, double *coords, int *inpoel, int *nnode, int *bface, int *esuel, int *elemface, int *intfac,
                    int *btags, double *facemetric, double *area)
{
	const UMesh<freal,NDIM>& m = *static_cast<RefCase*>(hv)->m;
	for(fint i = 0; i < m.gnpoin(); i++) for(int d = 0; d < NDIM; d++) coords[(size_t)i*NDIM+d] = m.gcoords(i,d);
	for(fint i = 0; i < m.gnelem(); i++) {
		nnode[i] = m.gnnode(i); area[i] = m.garea(i);
		for(int j = 0; j < 4; j++) {
			const bool have = j < m.gnnode(i);
			inpoel[(size_t)i*4+j] = have ? m.ginpoel(i,j) : -1;
			esuel[(size_t)i*4+j] = have ? m.gesuel(i,j) : -1;
			elemface[(size_t)i*4+j] = have ? m.gelemface(i,j) : -1;
		}
	}
	for(fint f = 0; f < m.gnbface(); f++) {
		bface[(size_t)f*3] = m.gbface(f,0); bface[(size_t)f*3+1] = m.gbface(f,1); bface[(size_t)f*3+2] = m.gbface(f,2);
		btags[f] = m.gbtags(f,0);
	}
	for(fint f = 0; f < m.gnaface(); f++) {
		for(int j = 0; j < 4; j++) intfac[(size_t)f*4+j] = m.gintfac(f,j);
		for(int j = 0; j < 3; j++) facemetric[(size_t)f*3+j] = m.gfacemetric(f,j);
	}
}

int ref_e_flow_create(void *hv, const double *phys, const char *flux, const char *gradient, const char *recon, double limiter_param,
                      int order2, int viscous, int const_visc, int nbc, const int *bc_tag_type, const double *bc_vals)
{
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	int rc = 0;
	try { make_flow(static_cast<RefCase*>(hv), phys, flux, gradient, recon, limiter_param, order2, viscous, const_visc, nbc, bc_tag_type, bc_vals); }
	catch(std::exception&) { rc = 1; }
	std::cout.rdbuf(old);
	return rc;
}

int ref_e_flow_residual(void *hv, const double *u, int gettimesteps, double *res, double *dtm)
{
	RefCase *h = static_cast<RefCase*>(hv);
	std::copy(u, u + h->uv.a.size(), h->uv.a.begin());
	std::fill(h->rv.a.begin(), h->rv.a.end(), 0.0);
	const int ierr = h->prob->compute_residual(&h->uv, &h->rv, gettimesteps != 0, &h->dv);
	if(res) std::copy(h->rv.a.begin(), h->rv.a.end(), res);
	if(dtm && gettimesteps) std::copy(h->dv.a.begin(), h->dv.a.end(), dtm);
	return ierr;
}

/// as ref_flow_forward_euler of tier D
int ref_e_flow_forward_euler(void *hv, double cfl, double tol, int maxiter, double *u, int *steps, double *hist_rel, double *hist_abs)
{
	RefCase *h = static_cast<RefCase*>(hv);
	_p_Vec uv;
	uv.a.assign(u, u + h->uv.a.size()); uv.nlocal = h->uv.nlocal; uv.nghost = 0;
	const SteadySolverConfig conf { true, "ref-tier-e", false, cfl, cfl, 0, 0, tol, maxiter, 0, 0 };
	SteadyForwardEulerSolver<NVARS> solver(h->prob.get(), &uv, conf);
	int code = 0;
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	try { code = solver.solve(&uv); }
	catch(Tolerance_error&) { code = 1; }
	catch(Numerical_error&) { code = 2; }
	std::cout.rdbuf(old);
	const TimingData td = solver.getTimingData();
	*steps = td.num_timesteps;
	for(size_t i = 0; i < td.convhis.size() && (int)i < maxiter; i++) { hist_rel[i] = td.convhis[i].rmsres; hist_abs[i] = td.convhis[i].absrmsres; }
	std::copy(uv.a.begin(), uv.a.end(), u);
	return code;
}

/// The reference's output artefacts for the state u: <basename>-surf_w<m>.out / -surf_o<m>.out (FlowOutput::exportSurfaceData),
/// the VTU file (postprocess_point + writeScalarsVectorToVtu_PointData) and <volprefix>-vol.out (exportVolumeData)
int ref_e_write_outputs(void *hv, const double *u, double aoa, const double *phys, int nwall, const int *walls, int nother,
                        const int *others, const char *basename, const char *vtufile, const char *volprefix)
{
	RefCase *h = static_cast<RefCase*>(hv);
	const FlowFV_base<freal> *const fv = dynamic_cast<const FlowFV_base<freal>*>(h->prob.get());
	if(!fv) return 1;
	std::copy(u, u + h->uv.a.size(), h->uv.a.begin());
	const IdealGasPhysics<freal> phy(phys[0], phys[1], phys[2], phys[3], phys[4]);
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	int rc = 0;
	try {
		FlowOutput out(fv, &phy, aoa);
		out.exportSurfaceData(&h->uv, std::vector<int>(walls, walls + nwall), std::vector<int>(others, others + nother), basename);
		amat::Array2d<freal> scalars, velocities;
		out.postprocess_point(&h->uv, scalars, velocities);
		std::string scalarnames[] = {"density", "mach-number", "pressure", "temperature"};
		writeScalarsVectorToVtu_PointData(vtufile, *h->m, scalars, scalarnames, velocities, "velocity");
		const fint ne = h->m->gnelem();
		MVector<freal> umat(ne, NVARS);
		for(fint i = 0; i < ne; i++) for(int j = 0; j < NVARS; j++) umat(i,j) = u[(size_t)i*NVARS+j];
		out.exportVolumeData(umat, volprefix);
	} catch(std::exception&) { rc = 2; }
	std::cout.rdbuf(old);
	return rc;
}

/// Cl, Cdp, Cdf of the reference's computeSurfaceData on gradients from its getGradients; entropy norm from FlowOutput
int ref_e_surface_and_entropy(void *hv, const double *u, int marker, double aoa, const double *phys, double *out4)
{
	RefCase *h = static_cast<RefCase*>(hv);
	const FlowFV_base<freal> *const fv = dynamic_cast<const FlowFV_base<freal>*>(h->prob.get());
	if(!fv) return 1;
	std::copy(u, u + h->uv.a.size(), h->uv.a.begin());
	const fint ne = h->m->gnelem();
	std::vector<GradBlock_t<freal,NDIM,NVARS>> grad(ne);
	std::stringstream sink;
	std::streambuf *const old = std::cout.rdbuf(sink.rdbuf());
	fv->getGradients(&h->uv, &grad[0]);
	const amat::Array2dView<freal> ua(h->uv.a.data(), ne, NVARS);
	fint nface = 0;
	for(fint f = h->m->gPhyBFaceStart(); f < h->m->gPhyBFaceEnd(); f++) if(h->m->gbtags(f,0) == marker) nface++;
	MVector<freal> output(nface, NDIM+2);
	const std::tuple<freal,freal,freal> t = fv->computeSurfaceData(ua, &grad[0], marker, output);
	const IdealGasPhysics<freal> phy(phys[0], phys[1], phys[2], phys[3], phys[4]);
	FlowOutput fo(fv, &phy, aoa);
	out4[3] = fo.compute_entropy_cell(&h->uv);
	std::cout.rdbuf(old);
	out4[0] = std::get<0>(t); out4[1] = std::get<1>(t); out4[2] = std::get<2>(t);
	return 0;
}

}
