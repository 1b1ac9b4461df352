/* Stand-in for the reference's mesh/mesh.hpp in the Tier-B build (oracle/ref_tier_b.cpp): a UMesh with the
 * reference's accessor names (src/mesh/mesh.hpp:60-260) over arrays handed in by the test harness (which builds
 * them with the oracle's restated topology code, itself pinned by the reference's mesh tests). Serial: no
 * connectivity faces. The reference's own UMesh needs Boost, MPI and PETSc. TEST INFRASTRUCTURE ONLY. */
#ifndef FVENS_B200_MESH_SHIM
#define FVENS_B200_MESH_SHIM
#include <array>
#include "aconstants.hpp"
#include "utilities/aarray2d.hpp"

namespace fvens {

typedef int EIndex;
typedef int FIndex;

template <typename scalar, int ndim>
class UMesh {
public:
	fint nelem = 0, nbface = 0, naface = 0, npoin = 0;
	int maxnnode = 4;
	const scalar *coords = nullptr, *area = nullptr, *facemetric = nullptr;
	const fint *inpoel = nullptr, *esuel = nullptr, *elemface = nullptr, *intfac = nullptr;
	const int *nnode = nullptr, *btags = nullptr;

	fint gnelem() const { return nelem; }
	fint gnpoin() const { return npoin; }
	fint gnbface() const { return nbface; }
	fint gnaface() const { return naface; }
	fint gnConnFace() const { return 0; }
	fint gPhyBFaceStart() const { return 0; }
	fint gPhyBFaceEnd() const { return nbface; }
	fint gSubDomFaceStart() const { return nbface; }
	fint gSubDomFaceEnd() const { return naface; }
	fint gConnBFaceStart() const { return naface; }
	fint gConnBFaceEnd() const { return naface; }
	fint gDomFaceStart() const { return nbface; }
	fint gDomFaceEnd() const { return naface; }
	fint gFaceStart() const { return 0; }
	fint gFaceEnd() const { return naface; }
	int gnnode(const fint i) const { return nnode[i]; }
	int gnfael(const fint i) const { return nnode[i]; }
	int gnnofa(const fint) const { return 2; }
	scalar gcoords(const fint p, const int d) const { return coords[(size_t)p*ndim+d]; }
	fint ginpoel(const fint i, const int j) const { return inpoel[(size_t)i*maxnnode+j]; }
	fint gesuel(const fint i, const int j) const { return esuel[(size_t)i*maxnnode+j]; }
	fint gelemface(const fint i, const int j) const { return elemface[(size_t)i*maxnnode+j]; }
	fint gintfac(const fint f, const int j) const { return intfac[(size_t)f*4+j]; }
	scalar gfacemetric(const fint f, const int j) const { return facemetric[(size_t)f*3+j]; }
	scalar garea(const fint i) const { return area[i]; }
	int gbtags(const fint face, const int) const { return btags[face]; }
	/// mean of the nodes (reference: mesh/mesh.cpp:317-328)
	void compute_cell_centres(scalar *const centres) const {
		for(fint i = 0; i < nelem; i++)
			for(int idim = 0; idim < ndim; idim++) {
				centres[i*ndim+idim] = 0;
				for(int jnode = 0; jnode < nnode[i]; jnode++) centres[i*ndim+idim] += gcoords(ginpoel(i,jnode),idim);
				centres[i*ndim+idim] /= (scalar)nnode[i];
			}
	}
	/// serial: no connectivity faces, these are never reached
	fint gconnface(const fint, const int) const { return -1; }
	fint gglobalElemIndex(const fint i) const { return i; }
	fint gnelemglobal() const { return nelem; }
	std::array<scalar,ndim> gnormal(const fint f) const { std::array<scalar,ndim> n; for(int d = 0; d < ndim; d++) n[d] = facemetric[(size_t)f*3+d]; return n; }
};

}
#endif
