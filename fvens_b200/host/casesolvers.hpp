/* Case drivers and output with the reference's names (reference: src/utilities/casesolvers.{hpp,cpp},
 * src/spatial/aoutput.{hpp,cpp}): constructMeshFlow, initializeSystemVector, FlowOutput (entropy norm, surface
 * files, volume file, point data), the VTU writer, the residual-history log, FlowCase / SteadyFlowCase with
 * run / run_output / execute_starter / execute_main. Explicit pseudo-time only: an IMPLICIT control file is
 * rejected with UnsupportedOptionError (needs PETSc matrices and KSP; out of scope).
 * The solver loops run on the GPU through SteadyForwardEulerSolver; everything in this file is host-side
 * orchestration and post-processing on downloaded arrays (O(N) once per run, as in the reference).
 */
#ifndef FVENS_B200_CASESOLVERS_HPP
#define FVENS_B200_CASESOLVERS_HPP

#include "controlparser.hpp"
#include <iomanip>

namespace fvens {

/// Reference: constructMeshFlow (mesh/ameshutils.cpp:39-153). Cell reordering (-mesh_reorder) is the engine's job.
inline UMesh<freal,NDIM> constructMeshFlow(const FlowParserOptions& opts, const std::string& mesh_suffix) {
	UMesh<freal,NDIM> m = constructMesh(opts.meshfile + mesh_suffix);
	// utilities/casesolvers.cpp:31-35: a `periodic` boundary condition carries (marker, axis) in its `options`; the
	// pairs become interior faces of the device mesh (the reference pairs them too but has no FlowBC to run them with)
	for(auto it = opts.bcconf.begin(); it != opts.bcconf.end(); it++)
		if(it->bc_type == PERIODIC_BC) {
			if(it->bc_opts.size() < 2) throw std::runtime_error("periodic boundary condition needs `options marker axis`");
			m.compute_periodic_map(it->bc_opts[0], it->bc_opts[1]);
		}
	return m;
}

/// Reference: initializeSystemVector (utilities/casesolvers.cpp:52-69)
inline StatusCode initializeSystemVector(const FlowParserOptions& opts, const UMesh<freal,NDIM>& m, Vec *const u,
                                         const VecPlace place = VEC_DEVICE) {
	const StatusCode ierr = createGhostedSystemVector(&m, NVARS, u, place);
	if(ierr) return ierr;
	return initializeSystemVector(extract_spatial_physics_config(opts), m, *u);
}

// ---------------------------------------------------------------------------------------- convergence log

/// Reference: writeConvergenceHistoryHeader / writeStepToConvergenceHistory (spatial/aoutput.cpp:617-636)
inline void writeConvergenceHistoryHeader(std::ostream& outf) {
	using std::setw;
	outf << '#' << setw(6) << "NStep" << setw(14) << "Log rel resi" << setw(14) << "Log abs resi"
	     << setw(12) << "Tot.Wtime" << setw(12) << "Lin.Wtime" << setw(12) << "Lin.iters" << setw(10) << "CFL" << '\n';
	outf << "#----------------------------------------------------------------------------------\n";
}
inline void writeStepToConvergenceHistory(const SteadyStepMonitor s, std::ostream& outf) {
	using std::setw;
	outf << std::setprecision(6);
	outf << setw(7) << s.step << setw(14) << std::log10(s.rmsres) << setw(14) << std::log10(s.absrmsres);
	outf << std::setprecision(4);
	outf << setw(12) << s.odewalltime << setw(12) << s.linwalltime << setw(12) << s.linits << setw(10) << s.cfl << '\n';
	outf << std::flush;
}

// ---------------------------------------------------------------------------------------- gas relations (host, output only)

namespace hostgas {
inline freal pressure(const freal g, const freal *u) { return (g - 1.0)*(u[3] - 0.5*(u[1]*u[1] + u[2]*u[2])/u[0]); }
inline freal soundspeed(const freal g, const freal *u) { return std::sqrt(g*pressure(g, u)/u[0]); }
inline freal temperature(const freal g, const freal Minf, const freal *u) { return g*Minf*Minf*pressure(g, u)/u[0]; }
/// physics/aphysics_defs.hpp: Sutherland's law in the reference's non-dimensionalisation, or 1/Re
inline freal viscosity(const FlowPhysicsConfig& p, const freal *u) {
	if(!p.viscous_sim) return 0.0;
	if(p.const_visc) return 1.0/p.Reinf;
	const freal T = temperature(p.gamma, p.Minf, u), C = 110.5/p.Tinf;
	return (1.0 + C)/(T + C)*std::pow(T, 1.5)/p.Reinf;
}
}

// ---------------------------------------------------------------------------------------- VTU

/// Reference: writeScalarsVectorToVtu_PointData (spatial/aoutput.cpp:427-610), ASCII unstructured grid with point data
inline void writeScalarsVectorToVtu_PointData(const std::string& fname, const UMesh<freal,NDIM>& m,
                                              const amat::Array2d<freal>& x, const std::string scaname[],
                                              const amat::Array2d<freal>& y, const std::string& vecname)
{
	std::ofstream out(fname);
	if(!out) throw std::runtime_error("cannot open " + fname + " for writing");
	out << std::setprecision(10);
	const int nscalars = (int)x.cols();
	out << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
	out << "<UnstructuredGrid>\n";
	out << "\t<Piece NumberOfPoints=\"" << m.gnpoin() << "\" NumberOfCells=\"" << m.gnelem() << "\">\n";
	if(x.rows() > 0 || y.rows() > 0) {
		out << "\t\t<PointData ";
		if(x.rows() > 0) out << "Scalars=\"" << scaname[0] << "\" ";
		if(y.rows() > 0) out << "Vectors=\"" << vecname << "\"";
		out << ">\n";
		for(int in = 0; in < nscalars; in++) {
			out << "\t\t\t<DataArray type=\"Float64\" Name=\"" << scaname[in] << "\" Format=\"ascii\">\n";
			for(fint i = 0; i < m.gnpoin(); i++) out << "\t\t\t\t" << x(i,in) << '\n';
			out << "\t\t\t</DataArray>\n";
		}
		if(y.rows() > 0) {
			out << "\t\t\t<DataArray type=\"Float64\" Name=\"" << vecname << "\" NumberOfComponents=\"3\" Format=\"ascii\">\n";
			for(fint i = 0; i < m.gnpoin(); i++) {
				out << "\t\t\t\t";
				for(int idim = 0; idim < NDIM; idim++) out << y(i,idim) << " ";
				out << "0.0 " << '\n';
			}
			out << "\t\t\t</DataArray>\n";
		}
		out << "\t\t</PointData>\n";
	}
	out << "\t\t<Points>\n";
	out << "\t\t<DataArray type=\"Float64\" NumberOfComponents=\"3\" Format=\"ascii\">\n";
	for(fint i = 0; i < m.gnpoin(); i++) {
		out << "\t\t\t";
		for(int idim = 0; idim < NDIM; idim++) out << m.gcoords(i,idim) << " ";
		out << "0.0 " << '\n';
	}
	out << "\t\t</DataArray>\n\t\t</Points>\n";
	out << "\t\t<Cells>\n";
	out << "\t\t\t<DataArray type=\"UInt32\" Name=\"connectivity\" Format=\"ascii\">\n";
	for(fint i = 0; i < m.gnelem(); i++) {
		out << "\t\t\t\t";
		for(int j = 0; j < m.gnnode(i); j++) out << m.ginpoel(i,j) << " ";
		out << '\n';
	}
	out << "\t\t\t</DataArray>\n";
	out << "\t\t\t<DataArray type=\"UInt32\" Name=\"offsets\" Format=\"ascii\">\n";
	fint totalcells = 0;
	for(fint i = 0; i < m.gnelem(); i++) { totalcells += m.gnnode(i); out << "\t\t\t\t" << totalcells << '\n'; }
	out << "\t\t\t</DataArray>\n";
	out << "\t\t\t<DataArray type=\"Int32\" Name=\"types\" Format=\"ascii\">\n";
	for(fint i = 0; i < m.gnelem(); i++) out << "\t\t\t\t" << (m.gnnode(i) == 3 ? 5 : 9) << '\n';      // VTK_TRIANGLE, VTK_QUAD
	out << "\t\t\t</DataArray>\n";
	out << "\t\t</Cells>\n";
	out << "\t</Piece>\n</UnstructuredGrid>\n</VTKFile>";
	out.close();
	std::cout << "Vtu file written.\n";
}

/// Reference: writeMeshToVtu (spatial/aoutput.cpp:557-615): the grid alone. The reference writes the cell offsets as
/// nnode(i)*(i+1), which is the running node count only when all cells have the same type; the running count is
/// written here (same file for all-triangle and all-quadrangle meshes, a readable one for hybrid meshes).
inline void writeMeshToVtu(const std::string& fname, const UMesh<freal,NDIM>& m)
{
	std::cout << "Writing vtu output...\n";
	std::ofstream out(fname);
	if(!out) throw std::runtime_error("cannot open " + fname + " for writing");
	out << std::setprecision(10);
	out << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n<UnstructuredGrid>\n";
	out << "\t<Piece NumberOfPoints=\"" << m.gnpoin() << "\" NumberOfCells=\"" << m.gnelem() << "\">\n";
	out << "\t\t<Points>\n\t\t<DataArray type=\"Float64\" NumberOfComponents=\"3\" Format=\"ascii\">\n";
	for(fint i = 0; i < m.gnpoin(); i++) out << "\t\t\t" << m.gcoords(i,0) << " " << m.gcoords(i,1) << " " << 0.0 << '\n';
	out << "\t\t</DataArray>\n\t\t</Points>\n\t\t<Cells>\n";
	out << "\t\t\t<DataArray type=\"UInt32\" Name=\"connectivity\" Format=\"ascii\">\n";
	for(fint i = 0; i < m.gnelem(); i++) {
		out << "\t\t\t\t";
		for(int j = 0; j < m.gnnode(i); j++) out << m.ginpoel(i,j) << " ";
		out << '\n';
	}
	out << "\t\t\t</DataArray>\n\t\t\t<DataArray type=\"UInt32\" Name=\"offsets\" Format=\"ascii\">\n";
	fint nodes = 0;
	for(fint i = 0; i < m.gnelem(); i++) { nodes += m.gnnode(i); out << "\t\t\t\t" << nodes << '\n'; }
	out << "\t\t\t</DataArray>\n\t\t\t<DataArray type=\"Int32\" Name=\"types\" Format=\"ascii\">\n";
	for(fint i = 0; i < m.gnelem(); i++) out << "\t\t\t\t" << (m.gnnode(i) == 3 ? 5 : 9) << '\n';
	out << "\t\t\t</DataArray>\n\t\t</Cells>\n\t</Piece>\n</UnstructuredGrid>\n</VTKFile>";
	out.close();
	std::cout << "Vtu file written.\n";
}

// ---------------------------------------------------------------------------------------- output on host arrays

/// Area-weighted cell->point averaging, then density, Mach number, pressure, temperature and velocity
/// (FlowOutput::postprocess_point, aoutput.cpp:97-148); u [nelem][NVARS] conserved
inline void postprocess_point(const UMesh<freal,NDIM>& m, const FlowPhysicsConfig& pconf, const freal *const u,
                              amat::Array2d<freal>& scalars, amat::Array2d<freal>& velocities)
{
	scalars.resize(m.gnpoin(), 4); velocities.resize(m.gnpoin(), NDIM);
	std::vector<freal> up((size_t)m.gnpoin()*NVARS, 0.0), areasum(m.gnpoin(), 0.0);
	// The reference adds the cell area to the point's area sum inside its loop over the variables
	// (aoutput.cpp:115-123), i.e. NVARS times per node, so its point values of the CONSERVED variables - and with them
	// the density and pressure it writes - are 1/NVARS of the area-weighted averages (velocity, Mach number and
	// temperature are ratios and come out right). Here the area is added once per node: true averages.
	for(fint ielem = 0; ielem < m.gnelem(); ielem++)
		for(int inode = 0; inode < m.gnnode(ielem); inode++) {
			const fint p = m.ginpoel(ielem,inode);
			for(int ivar = 0; ivar < NVARS; ivar++) up[(size_t)p*NVARS+ivar] += u[(size_t)ielem*NVARS+ivar]*m.garea(ielem);
			areasum[p] += m.garea(ielem);
		}
	for(fint ip = 0; ip < m.gnpoin(); ip++) {
		freal *const q = &up[(size_t)ip*NVARS];
		for(int ivar = 0; ivar < NVARS; ivar++) q[ivar] /= areasum[ip];
	}
	for(fint ip = 0; ip < m.gnpoin(); ip++) {
		const freal *const q = &up[(size_t)ip*NVARS];
		scalars(ip,0) = q[0];
		for(int idim = 0; idim < NDIM; idim++) velocities(ip,idim) = q[idim+1]/q[0];
		const freal vmag2 = velocities(ip,0)*velocities(ip,0) + velocities(ip,1)*velocities(ip,1);
		scalars(ip,2) = hostgas::pressure(pconf.gamma, q);
		scalars(ip,1) = std::sqrt(vmag2)/hostgas::soundspeed(pconf.gamma, q);
		scalars(ip,3) = hostgas::temperature(pconf.gamma, pconf.Minf, q);
	}
}

/// x y rho u v p T M per cell (FlowOutput::exportVolumeData, aoutput.cpp:150-176)
inline void exportVolumeData(const UMesh<freal,NDIM>& m, const FlowPhysicsConfig& pconf, const freal *const u, const std::string& volfile)
{
	std::ofstream fout(volfile + "-vol.out");
	if(!fout) throw std::runtime_error("cannot open " + volfile + "-vol.out");
	fout << "#   x    y    rho     u      v      p      T      M \n";
	for(fint iel = 0; iel < m.gnelem(); iel++) {
		const freal *const q = &u[(size_t)iel*NVARS];
		const freal T = hostgas::temperature(pconf.gamma, pconf.Minf, q), c = hostgas::soundspeed(pconf.gamma, q), p = hostgas::pressure(pconf.gamma, q);
		const freal vmag = std::sqrt(q[1]/q[0]*q[1]/q[0] + q[2]/q[0]*q[2]/q[0]);
		freal rc[NDIM] = {0.0, 0.0};
		for(int ino = 0; ino < m.gnnode(iel); ino++) for(int j = 0; j < NDIM; j++) rc[j] += m.gcoords(m.ginpoel(iel,ino),j);
		for(int j = 0; j < NDIM; j++) rc[j] /= m.gnnode(iel);
		fout << rc[0] << " " << rc[1] << " " << q[0] << " " << q[1]/q[0] << " " << q[2]/q[0] << " " << p << " " << T << " " << vmag/c << '\n';
	}
}

/// The per-face table of computeSurfaceData (flow_spatial.cpp:131-310) for the faces with marker iwbcm: x, y, Cp, Cf per
/// face, from the cell state and the conserved-variable gradients; also the integrated Cl, Cdp, Cdf the reference derives
/// from the same rows (the product takes those from the device, fvg_surface_data; here they serve the host-only checks).
inline std::tuple<freal,freal,freal> surfaceFaceTable(const UMesh<freal,NDIM>& m, const FlowPhysicsConfig& pconf, const freal *const u,
                                                      const GradBlock_t<freal,NDIM,NVARS> *const grad, const int iwbcm,
                                                      std::vector<std::array<freal,4>>& rows)
{
	rows.clear();
	const freal pinf = 1.0/(pconf.gamma*pconf.Minf*pconf.Minf);
	const freal wind[NDIM] = {std::cos(pconf.aoa), std::sin(pconf.aoa)}, flownormal[NDIM] = {-wind[1], wind[0]};
	freal totalarea = 0, Cl = 0, Cdp = 0, Cdf = 0;
	for(fint iface = m.gPhyBFaceStart(); iface < m.gPhyBFaceEnd(); iface++) {
		if(m.gbtags(iface,0) != iwbcm) continue;
		const fint lelem = m.gintfac(iface,0);
		const std::array<freal,NDIM> n = m.gnormal(iface);
		const freal area = m.gfacemetric(iface,NDIM);
		const freal tangf[NDIM] = {n[1], -n[0]};
		freal fcen[NDIM];
		for(int j = 0; j < NDIM; j++) {
			fcen[j] = 0;
			for(int inofa = 0; inofa < m.gnnofa(iface); inofa++) fcen[j] += m.gcoords(m.gintfac(iface,2+inofa),j);
			fcen[j] /= m.gnnofa(iface);
		}
		const freal *const q = &u[(size_t)lelem*NVARS];
		const freal cp = (hostgas::pressure(pconf.gamma, q) - pinf)*2.0;
		const freal muhat = hostgas::viscosity(pconf, q);
		freal gradu[NDIM][NDIM];
		for(int i = 0; i < NDIM; i++) for(int j = 0; j < NDIM; j++)
			gradu[i][j] = (grad[lelem](j,i+1)*q[0] - q[i+1]*grad[lelem](j,0))/(q[0]*q[0]);
		freal force[NDIM];
		for(int i = 0; i < NDIM; i++) { force[i] = 0; for(int j = 0; j < NDIM; j++) force[i] += (gradu[i][j] + gradu[j][i])*n[j]; }
		const freal tauw = muhat*(force[0]*tangf[0] + force[1]*tangf[1]);
		rows.push_back({fcen[0], fcen[1], cp, 2*tauw});
		totalarea += area;
		Cl += cp*(n[0]*flownormal[0] + n[1]*flownormal[1])*area;
		Cdp += cp*(n[0]*wind[0] + n[1]*wind[1])*area;
		Cdf += 2*tauw*(tangf[0]*wind[0] + tangf[1]*wind[1])*area;
	}
	return std::make_tuple(Cl/totalarea, Cdp/totalarea, Cdf/totalarea);
}

/// <basename>-surf_w<marker>.out in the reference's format (FlowOutput::exportSurfaceData, aoutput.cpp:209-241)
inline void writeWallSurfaceFile(const std::string& fname, const std::vector<std::array<freal,4>>& rows, const freal Cl, const freal Cdp, const freal Cdf)
{
	std::ofstream fout(fname);
	if(!fout) throw std::runtime_error("cannot open " + fname);
	fout << "#  x \t y \t Cp  \t Cf \n";
	for(const auto& r : rows) { for(int j = 0; j < 4; j++) fout << "  " << r[j]; fout << '\n'; }
	fout << "# Cl      Cdp      Cdf\n";
	fout << "# " << Cl << "  " << Cdp << "  " << Cdf << '\n';
}

/// <basename>-surf_o<marker>.out: face centre and cell velocity per face of an "other" boundary (aoutput.cpp:243-290)
inline void writeOtherSurfaceFile(const std::string& fname, const UMesh<freal,NDIM>& m, const freal *const u, const int marker)
{
	std::ofstream fout(fname);
	if(!fout) throw std::runtime_error("cannot open " + fname);
	fout << "#   x         y          u           v\n";
	for(fint iface = m.gPhyBFaceStart(); iface < m.gPhyBFaceEnd(); iface++) {
		if(m.gbtags(iface,0) != marker) continue;
		const fint lelem = m.gintfac(iface,0);
		const freal *const q = &u[(size_t)lelem*NVARS];
		freal coord[NDIM];
		for(int j = 0; j < NDIM; j++) {
			coord[j] = 0;
			for(int inofa = 0; inofa < m.gnnofa(iface); inofa++) coord[j] += m.gcoords(m.gintfac(iface,2+inofa),j);
			coord[j] /= m.gnnofa(iface);
		}
		fout << "  " << coord[0] << "  " << coord[1] << "  " << q[1]/q[0] << "  " << q[2]/q[0] << '\n';
	}
}

// ---------------------------------------------------------------------------------------- FlowOutput

/// Reference: FlowOutput (spatial/aoutput.hpp:60-110, aoutput.cpp:20-290)
class FlowOutput {
public:
	FlowOutput(const FlowFV_base<freal> *const fv, const FlowPhysicsConfig& pc, const freal aoa)
		: space(fv), m(fv->mesh()), pconf(pc), av(aoa) {}

	/// || (s - s_inf)/s_inf ||_{L2, area} over the cells (aoutput.cpp:28-63), computed on the device
	freal compute_entropy_cell(const Vec u) const {
		const freal e = space->compute_entropy_cell(u);
		std::cout << "FlowOutput: log mesh size and log entropy:   " << std::log10(1.0/std::sqrt((freal)m->gnelem())) << "  "
		          << std::setprecision(10) << std::log10(e) << std::endl;
		return e;
	}

	StatusCode postprocess_point(const std::vector<freal>& u, amat::Array2d<freal>& scalars, amat::Array2d<freal>& velocities) const {
		fvens::postprocess_point(*m, pconf, u.data(), scalars, velocities);
		return 0;
	}

	void exportVolumeData(const std::vector<freal>& u, const std::string& volfile) const { fvens::exportVolumeData(*m, pconf, u.data(), volfile); }

	/// Per wall marker: x, y, Cp, Cf of every face with that marker and the integrated Cl, Cdp, Cdf; per other marker:
	/// x, y and the velocity (aoutput.cpp:181-290). The integrals are the device's (fvg_surface_data); the per-face table
	/// is computed on the host from the same cell state and conserved-variable gradients (surfaceFaceTable).
	void exportSurfaceData(const Vec u, const std::vector<int>& wbcm, const std::vector<int>& obcm, const std::string& basename) const {
		const fint ne = m->gnelem();
		std::vector<freal> uh((size_t)ne*NVARS);
		VecCopyToHost(u, uh.data());
		std::vector<GradBlock_t<freal,NDIM,NVARS>> grad(ne);
		space->getGradients(u, &grad[0]);
		const amat::Array2dView<freal> ua(uh.data(), ne, NVARS);
		for(size_t im = 0; im < wbcm.size(); im++) {
			MVector<freal> dummy;
			freal Cl, Cdp, Cdf;
			std::tie(Cl, Cdp, Cdf) = space->computeSurfaceData(ua, &grad[0], wbcm[im], dummy);
			std::vector<std::array<freal,4>> rows;
			surfaceFaceTable(*m, pconf, uh.data(), &grad[0], wbcm[im], rows);
			writeWallSurfaceFile(basename + "-surf_w" + std::to_string(wbcm[im]) + ".out", rows, Cl, Cdp, Cdf);
			std::cout << "FlowOutput: CL = " << Cl << "   CDp = " << Cdp << "    CDf = " << Cdf << std::endl;
		}
		for(size_t im = 0; im < obcm.size(); im++)
			writeOtherSurfaceFile(basename + "-surf_o" + std::to_string(obcm[im]) + ".out", *m, uh.data(), obcm[im]);
	}

private:
	const FlowFV_base<freal> *const space;
	const UMesh<freal,NDIM> *const m;
	const FlowPhysicsConfig pconf;
	const freal av;
};

// ---------------------------------------------------------------------------------------- cases

/// Reference: FlowSolutionFunctionals (utilities/casesolvers.hpp:24-33)
struct FlowSolutionFunctionals { freal meshSizeParameter, entropy, cl, cdp, cdf; };

/// Reference: createFlowSpatial (utilities/casesolvers.cpp:40-50)
inline const FlowFV_base<freal>* createFlowSpatial(const FlowParserOptions& opts, const UMesh<freal,NDIM>& m) {
	std::cout << "Setting up main spatial scheme.\n";
	return create_const_flowSpatialDiscretization<freal>(&m, extract_spatial_physics_config(opts), extract_spatial_numerics_config(opts));
}

/// Reference: FlowCase / SteadyFlowCase (utilities/casesolvers.cpp:71-420), explicit pseudo-time only
class SteadyFlowCase {
public:
	explicit SteadyFlowCase(const FlowParserOptions& options) : opts(options) {
		if(opts.pseudotimetype == "IMPLICIT")
			throw UnsupportedOptionError("pseudotime_stepping_type implicit (needs PETSc matrices and Krylov solvers; this build runs the explicit path)");
		if(opts.sim_type != "STEADY") throw UnsupportedOptionError("simulation_type " + opts.sim_type);
	}

	/// First-order starting solve when the control file has an `initialization` block (casesolvers.cpp:224-314)
	int execute_starter(const Spatial<freal,NVARS> *const prob, Vec u) const {
		if(opts.usestarter == 0) return 0;
		const UMesh<freal,NDIM> *const m = prob->mesh();
		std::cout << "\nSetting up spatial scheme for the initial guess.\n";
		std::unique_ptr<const FlowFV_base<freal>> startprob(create_const_flowSpatialDiscretization<freal>(
			m, extract_spatial_physics_config(opts), firstorder_spatial_numerics_config(opts)));
		const SteadySolverConfig starttconf { opts.lognres, opts.logfile + "-init", false, opts.firstinitcfl, opts.firstendcfl,
			opts.firstrampstart, opts.firstrampend, opts.firsttolerance, opts.firstmaxiter, 0, 0 };
		SteadyForwardEulerSolver<NVARS> starttime(startprob.get(), u, starttconf);
		std::cout << "Set up explicit forward Euler temporal scheme for startup solve.\n***\n";
		// a starting solve that does not reach its tolerance is not an error
		try { return starttime.solve(u); }
		catch(Tolerance_error& e) { std::cout << e.what() << std::endl; }
		return 0;
	}

	/// Main solve (casesolvers.cpp:316-384)
	TimingData execute_main(const Spatial<freal,NVARS> *const prob, Vec u) const {
		const SteadySolverConfig maintconf { opts.lognres, opts.logfile, opts.write_final_lin_sys, opts.initcfl, opts.endcfl,
			opts.rampstart, opts.rampend, opts.tolerance, opts.maxiter, 0, 0 };
		SteadyForwardEulerSolver<NVARS> time(prob, u, maintconf);
		std::cout << "\nSet up explicit forward Euler temporal scheme for main solve.\n";
		try { time.solve(u); }
		catch(Tolerance_error&) {
			// the reference's solver returns its timing data with converged = false in this case and execute() throws
			TimingData td = time.getTimingData(); td.converged = false; return td;
		}
		catch(Numerical_error& e) {
			std::cout << "FVENS: Main solve failed: " << e.what() << std::endl;
			TimingData td = time.getTimingData(); td.converged = false; return td;
		}
		std::cout << "***\n";
		return time.getTimingData();
	}

	/// Starter + main solve; writes <log_file_prefix>-residual_history.log when asked (casesolvers.cpp:386-420).
	/// One deliberate difference: the reference loses the history when the main solve stops at max_timesteps (the
	/// solver's Tolerance_error passes through execute() before the file is written); here the file is written first.
	int execute(const Spatial<freal,NVARS> *const prob, const bool outhist, Vec u) const {
		fvens_throw(execute_starter(prob, u), "Startup solve failed!");
		const TimingData td = execute_main(prob, u);
		if(outhist) {
			std::ofstream convout(opts.logfile + "-residual_history.log");
			writeConvergenceHistoryHeader(convout);
			for(size_t istp = 0; istp < td.convhis.size(); istp++) writeStepToConvergenceHistory(td.convhis[istp], convout);
		}
		if(!td.converged) throw Tolerance_error("Main flow solve did not converge!");
		return 0;
	}

	int run(const UMesh<freal,NDIM>& m, Vec u) const {
		std::unique_ptr<const FlowFV_base<freal>> prob(createFlowSpatial(opts, m));
		return execute(prob.get(), false, u);
	}

	/// Solve, then entropy norm, surface / VTU / volume files and the lift and drag of the first output wall
	/// (casesolvers.cpp:87-168)
	FlowSolutionFunctionals run_output(const bool surface_file_needed, const bool vtu_output_needed,
	                                   const UMesh<freal,NDIM>& m, Vec u) const {
		std::unique_ptr<const FlowFV_base<freal>> prob(createFlowSpatial(opts, m));
		const freal h = 1.0/std::pow((freal)m.gnelem(), 1.0/NDIM);
		try { execute(prob.get(), opts.lognres, u); }
		catch(Tolerance_error& e) { std::cout << e.what() << std::endl; }

		const FlowPhysicsConfig pconf = extract_spatial_physics_config(opts);
		FlowOutput out(prob.get(), pconf, opts.alpha);
		const freal entropy = out.compute_entropy_cell(u);
		std::vector<freal> uh((size_t)m.gnelem()*NVARS);
		VecCopyToHost(u, uh.data());
		if(surface_file_needed && (opts.num_out_walls > 0 || opts.num_out_others > 0)) {
			try { out.exportSurfaceData(u, opts.lwalls, opts.lothers, opts.surfnameprefix); }
			catch(std::exception& e) { std::cout << e.what() << std::endl; }
		}
		if(vtu_output_needed) {
			amat::Array2d<freal> scalars, velocities;
			out.postprocess_point(uh, scalars, velocities);
			const std::string scalarnames[] = {"density", "mach-number", "pressure", "temperature"};
			writeScalarsVectorToVtu_PointData(opts.vtu_output_file, m, scalars, scalarnames, velocities, "velocity");
		}
		if(opts.vol_output_reqd == "YES") out.exportVolumeData(uh, opts.volnameprefix);

		freal cl = 0, cdp = 0, cdf = 0;
		if(!opts.lwalls.empty()) {
			std::vector<GradBlock_t<freal,NDIM,NVARS>> grad(m.gnelem());
			prob->getGradients(u, &grad[0]);
			const amat::Array2dView<freal> ua(uh.data(), m.gnelem(), NVARS);
			MVector<freal> output;
			std::tie(cl, cdp, cdf) = prob->computeSurfaceData(ua, &grad[0], opts.lwalls[0], output);
		}
		return FlowSolutionFunctionals{h, entropy, cl, cdp, cdf};
	}

protected:
	const FlowParserOptions opts;
};

/// Reference: UnsteadyFlowCase (utilities/casesolvers.hpp:212-220, casesolvers.cpp:422-445): only TVDRK exists
class UnsteadyFlowCase {
public:
	explicit UnsteadyFlowCase(const FlowParserOptions& options) : opts(options) {}

	int execute(const Spatial<freal,NVARS> *const prob, Vec u) const {
		if(opts.time_integrator != "TVDRK") throw UnsupportedOptionError("Nothing but TVDRK is implemented yet!");
		TVDRKSolver<NVARS> time(prob, u, opts.time_order, opts.logfile, opts.phy_cfl);
		return time.solve(opts.final_time);
	}
protected:
	const FlowParserOptions opts;
};

}
#endif
