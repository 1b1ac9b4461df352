/* flow_conv <control file> --number_of_meshes <n> --mesh_file <prefix> [--source_dir <dir>] [--log_file_prefix <p>]
 * The reference's grid-convergence acceptance test (tests/flow_conv.cpp) on the host class surface: for meshes
 * <prefix>0.msh ... <prefix>(n-1).msh run the steady case of the control file from the free stream, take the entropy
 * error and the mesh size parameter, and pass iff the observed order between the two finest meshes lies in [1.65, 2.1].
 * The reference registers it for the explicit solver as Flow_Explicit_Euler_Cylinder_GreenGauss_Roe_Tri_
 * EntropyConvergence (tests/inv-2dcyl/CMakeLists.txt, 3 meshes of testcases/2dcylinder/grids). Needs a B200.
 * Run by tests/test_post_r1_b_flow_conv.py.
 */
#include "../../fvens_b200/host/casesolvers.hpp"

using namespace fvens;

int main(int argc, char *argv[])
{
	if(argc < 2) { std::cerr << "! Please give a control file name.\n"; return 2; }
	std::map<std::string,std::string> cmdvars;
	int nmesh = 0;
	for(int i = 2; i + 1 < argc; i += 2) {
		const std::string a = argv[i];
		if(a.compare(0, 2, "--") != 0) { std::cerr << "! Unknown argument " << a << "\n"; return 2; }
		if(a == "--number_of_meshes") nmesh = std::atoi(argv[i+1]);
		else cmdvars[a.substr(2)] = argv[i+1];
	}
	if(nmesh < 2) { std::cerr << "! --number_of_meshes must be at least 2\n"; return 2; }
	try {
		const FlowParserOptions opts = parse_flow_controlfile(argv[1], cmdvars);
		SteadyFlowCase case1(opts);
		std::vector<double> lh(nmesh), lerrors(nmesh), slopes(nmesh-1);
		for(int imesh = 0; imesh < nmesh; imesh++) {
			const UMesh<freal,NDIM> m = constructMeshFlow(opts, std::to_string(imesh) + ".msh");
			Vec u = nullptr;
			fvens_throw(initializeSystemVector(opts, m, &u, VEC_DEVICE), "could not create the state vector");
			FlowSolutionFunctionals fnls {0, 0, 0, 0, 0};
			try { fnls = case1.run_output(false, false, m, u); }
			catch(Numerical_error& e) { std::cout << e.what() << std::endl; }
			std::cout << std::setprecision(12) << "Log of Mesh size and error are " << std::log10(fnls.meshSizeParameter) << "  "
			          << std::log10(fnls.entropy) << std::endl;
			lh[imesh] = std::log10(fnls.meshSizeParameter);
			lerrors[imesh] = std::log10(fnls.entropy);
			if(imesh > 0) slopes[imesh-1] = (lerrors[imesh] - lerrors[imesh-1])/(lh[imesh] - lh[imesh-1]);
			VecDestroy(&u);
		}
		std::cout << ">> Spatial orders = \n";
		for(int i = 0; i < nmesh-1; i++) std::cout << "   " << slopes[i] << std::endl;
		// the same window for LEASTSQUARES and GREENGAUSS ("the lower limit is chosen from experience")
		const bool passed = slopes[nmesh-2] <= 2.1 && slopes[nmesh-2] >= 1.65;
		std::cout << "\n--------------- End --------------------- \n";
		return passed ? 0 : 1;
	}
	catch(std::exception& e) { std::cerr << "flow_conv: " << e.what() << std::endl; return 3; }
}
