/* convertformat <input mesh> <output file> <msh|vtu>: mesh format conversion, the reference's utility of the same name
 * (src/utilities/convertformat.cpp). Reads Gmsh 2.2 ASCII or SU2, writes Gmsh 2.2 ASCII (UMesh::writeGmsh2) or the grid
 * as a VTU file (writeMeshToVtu). Host only. */
#include "fvens_b200.hpp"
#include "controlparser.hpp"
#include "casesolvers.hpp"

using namespace fvens;

int main(int argc, char *argv[])
{
	if(argc < 4) {
		std::cout << "Need: 1. Input mesh file, 2. Output mesh file 3. Output format.\n" << std::endl;
		return 2;                         // (the reference goes on and reads argv past the end)
	}
	const std::string inmesh = argv[1], outmesh = argv[2], outformat = argv[3];
	try {
		const UMesh<freal,NDIM> m(readMesh(inmesh));
		if(outformat == "msh") m.writeGmsh2(outmesh);
		else if(outformat == "vtu") writeMeshToVtu(outmesh, m);
		else { std::cout << "Invalid format. Exiting." << std::endl; return -1; }
	} catch(std::exception& e) { std::cerr << "convertformat: " << e.what() << std::endl; return 1; }
	std::cout << std::endl;
	return 0;
}
