/* CPU-only: parses a control file with the host front end (fvens_b200/host/controlparser.hpp) and prints the
 * FlowParserOptions as JSON for tests/test_controlfile.py. usage: test_controlparser <ctrl> [key value]... */
#include "../../fvens_b200/host/casesolvers.hpp"
#include <cstdio>
using namespace fvens;

static std::string q(const std::string& s) { return "\"" + s + "\""; }

int main(int argc, char *argv[])
{
	if(argc < 2) return 2;
	if(std::string(argv[1]) == "--history") {
		// the residual-history writer of casesolvers.hpp on monitors given as: step rel abs wtime cfl ...
		writeConvergenceHistoryHeader(std::cout);
		for(int i = 2; i + 4 < argc; i += 5) {
			SteadyStepMonitor s; s.step = std::atoi(argv[i]); s.rmsres = (float)std::atof(argv[i+1]); s.absrmsres = (float)std::atof(argv[i+2]);
			s.odewalltime = (float)std::atof(argv[i+3]); s.linwalltime = 0; s.linits = 0; s.cfl = (float)std::atof(argv[i+4]);
			writeStepToConvergenceHistory(s, std::cout);
		}
		return 0;
	}
	if(std::string(argv[1]) == "--outputs") {
		// the host-side writers of casesolvers.hpp on given arrays (no GPU): --outputs mesh ufile gradfile gamma Minf Tinf Reinf Pr
		// aoa viscous constvisc wallmarker othermarker prefix
		if(argc < 16) return 2;
		const UMesh<freal,NDIM> m = constructMesh(argv[2]);
		const size_t ne = m.gnelem();
		std::vector<double> u(ne*NVARS);
		std::vector<GradBlock_t<freal,NDIM,NVARS>> grad(ne);
		{ std::ifstream f(argv[3], std::ios::binary); f.read(reinterpret_cast<char*>(u.data()), u.size()*sizeof(double)); if(!f) return 3; }
		{ std::ifstream f(argv[4], std::ios::binary); f.read(reinterpret_cast<char*>(grad.data()), ne*sizeof(grad[0])); if(!f) return 3; }
		FlowPhysicsConfig pc { std::atof(argv[5]), std::atof(argv[6]), std::atof(argv[7]), std::atof(argv[8]), std::atof(argv[9]),
		                       std::atof(argv[10]), std::atoi(argv[11]) != 0, std::atoi(argv[12]) != 0, {} };
		const int wall = std::atoi(argv[13]), other = std::atoi(argv[14]);
		const std::string prefix = argv[15];
		std::vector<std::array<freal,4>> rows;
		freal Cl, Cdp, Cdf;
		std::tie(Cl, Cdp, Cdf) = surfaceFaceTable(m, pc, u.data(), grad.data(), wall, rows);
		writeWallSurfaceFile(prefix + "-surf_w" + std::to_string(wall) + ".out", rows, Cl, Cdp, Cdf);
		writeOtherSurfaceFile(prefix + "-surf_o" + std::to_string(other) + ".out", m, u.data(), other);
		amat::Array2d<freal> scalars, velocities;
		postprocess_point(m, pc, u.data(), scalars, velocities);
		const std::string names[] = {"density", "mach-number", "pressure", "temperature"};
		writeScalarsVectorToVtu_PointData(prefix + ".vtu", m, scalars, names, velocities, "velocity");
		exportVolumeData(m, pc, u.data(), prefix);
		return 0;
	}
	std::map<std::string,std::string> cmd;
	for(int i = 2; i + 1 < argc; i += 2) cmd[argv[i]] = argv[i+1];
	try {
		const FlowParserOptions o = parse_flow_controlfile(argv[1], cmd);
		const FlowNumericsConfig n = extract_spatial_numerics_config(o), n1 = firstorder_spatial_numerics_config(o);
		const FlowPhysicsConfig p = extract_spatial_physics_config(o);
		std::printf("{\"meshfile\": %s, \"vtu\": %s, \"logfile\": %s, \"lognres\": %d, \"flowtype\": %s, \"gamma\": %.17g, \"alpha\": %.17g, "
		            "\"Minf\": %.17g, \"viscsim\": %d, \"Tinf\": %.17g, \"Reinf\": %s, \"Pr\": %s, \"useconstvisc\": %d, ",
		            q(o.meshfile).c_str(), q(o.vtu_output_file).c_str(), q(o.logfile).c_str(), (int)o.lognres, q(o.flowtype).c_str(), o.gamma, o.alpha,
		            o.Minf, (int)o.viscsim, o.Tinf, std::isfinite(o.Reinf) ? std::to_string(o.Reinf).c_str() : "null",
		            std::isfinite(o.Pr) ? std::to_string(o.Pr).c_str() : "null", (int)o.useconstvisc);
		std::printf("\"invflux\": %s, \"invfluxjac\": %s, \"gradient\": %s, \"limiter\": %s, \"limiter_param\": %.17g, \"order2\": %d, "
		            "\"pseudotimetype\": %s, \"initcfl\": %.17g, \"endcfl\": %.17g, \"tolerance\": %.17g, \"maxiter\": %d, \"usestarter\": %d, "
		            "\"firstinitcfl\": %.17g, \"firstendcfl\": %.17g, \"firsttolerance\": %.17g, \"firstmaxiter\": %d, \"sim_type\": %s, \"surfnameprefix\": %s, "
		            "\"vol_output_reqd\": %s, ",
		            q(o.invflux).c_str(), q(o.invfluxjac).c_str(), q(o.gradientmethod).c_str(), q(o.limiter).c_str(), o.limiter_param, (int)o.order2,
		            q(o.pseudotimetype).c_str(), o.initcfl, o.endcfl, o.tolerance, o.maxiter, (int)o.usestarter, o.firstinitcfl, o.firstendcfl, o.firsttolerance,
		            o.firstmaxiter, q(o.sim_type).c_str(), q(o.surfnameprefix).c_str(), q(o.vol_output_reqd).c_str());
		std::printf("\"lwalls\": [");
		for(size_t i = 0; i < o.lwalls.size(); i++) std::printf("%s%d", i ? ", " : "", o.lwalls[i]);
		std::printf("], \"bcs\": [");
		for(size_t i = 0; i < o.bcconf.size(); i++) {
			std::printf("%s{\"tag\": %d, \"type\": %d, \"vals\": [", i ? ", " : "", o.bcconf[i].bc_tag, (int)o.bcconf[i].bc_type);
			for(size_t k = 0; k < o.bcconf[i].bc_vals.size(); k++) std::printf("%s%.17g", k ? ", " : "", o.bcconf[i].bc_vals[k]);
			std::printf("]}");
		}
		std::printf("], \"first_order\": {\"gradient\": %s, \"limiter\": %s, \"order2\": %d, \"flux\": %s}, \"main\": {\"gradient\": %s, \"order2\": %d}, "
		            "\"phys_nbc\": %d}\n", q(n1.gradientscheme).c_str(), q(n1.reconstruction).c_str(), (int)n1.order2, q(n1.conv_numflux).c_str(),
		            q(n.gradientscheme).c_str(), (int)n.order2, (int)p.bcconf.size());
	}
	catch(std::exception& e) { std::printf("{\"error\": %s}\n", q(e.what()).c_str()); return 1; }
	return 0;
}
