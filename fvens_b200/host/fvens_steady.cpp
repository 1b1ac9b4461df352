/* fvens_steady <control file> [--mesh_file <file>] [--source_dir <dir>] [--log_file_prefix <prefix>]
 *              [--max_timesteps <n>] [--host_vec]
 * The reference's steady-state driver (src/fvens_steady.cpp:20-70) on the B200 engine: reads the control file,
 * builds the mesh, starts from the free stream, runs the explicit pseudo-time solve (first-order starter when the
 * control file asks for one), writes the residual history, surface, volume and VTU files and prints Cl / Cd.
 * --max_timesteps caps both solves (for smoke runs); --host_vec keeps the state in host memory (drop-in mode,
 * every step through fvg_residual_host) instead of on the device.
 */
#include "casesolvers.hpp"

using namespace fvens;

int main(int argc, char *argv[])
{
	if(argc < 2) { std::cerr << "! Please give a control file name.\n"; return 2; }
	std::map<std::string,std::string> cmdvars;
	bool host_vec = false;
	int cap = -1;
	for(int i = 2; i < argc; i++) {
		const std::string a = argv[i];
		if(a == "--host_vec") { host_vec = true; continue; }
		if(a.size() > 2 && a.compare(0, 2, "--") == 0 && i + 1 < argc) {
			if(a == "--max_timesteps") cap = std::atoi(argv[++i]);
			else { cmdvars[a.substr(2)] = argv[i+1]; i++; }
			continue;
		}
		std::cerr << "! Unknown argument " << a << "\n";
		return 2;
	}
	try {
		FlowParserOptions opts = parse_flow_controlfile(argv[1], cmdvars);
		if(cap >= 0) { opts.maxiter = std::min(opts.maxiter, cap); opts.firstmaxiter = std::min(opts.firstmaxiter, cap); }
		SteadyFlowCase case1(opts);              // rejects what this build cannot run (implicit, unsteady) before any work
		const UMesh<freal,NDIM> m = constructMeshFlow(opts, "");
		std::cout << "Mesh: " << m.gnelem() << " cells, " << m.gnaface() << " faces, " << m.gnbface() << " boundary faces\n";
		Vec u = nullptr;
		fvens_throw(initializeSystemVector(opts, m, &u, host_vec ? VEC_HOST : VEC_DEVICE), "could not create the state vector");
		const FlowSolutionFunctionals fnls = case1.run_output(true, true, m, u);
		std::cout << std::setprecision(12) << "Functionals: h " << fnls.meshSizeParameter << " entropy " << fnls.entropy
		          << " CL " << fnls.cl << " CDp " << fnls.cdp << " CDf " << fnls.cdf << "\n";
		VecDestroy(&u);
	}
	catch(std::exception& e) {
		std::cerr << "fvens_steady: " << e.what() << std::endl;
		return 1;
	}
	std::cout << "\n\n--------------- End --------------------- \n\n";
	return 0;
}
