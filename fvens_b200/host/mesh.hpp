/* Host-side hybrid triangle/quad mesh with the class surface of the reference's UMesh<scalar,2>
 * (reference: src/mesh/mesh.hpp:25-500, src/mesh/meshreaders.hpp:29-66). Same accessor names, same
 * numbering conventions (faces ordered physical-boundary -> interior -> connectivity; left cell is
 * the smaller index; normals point left -> right), because everything downstream - including the
 * device mesh in csrc/device_mesh.cu - is defined in terms of them. The algorithms that build the
 * derived arrays are new (edge buckets keyed by the smaller node instead of points-surrounding-
 * points searches) and linear in the mesh size, so a 10M-cell mesh preprocesses in seconds.
 */
#ifndef FVENS_B200_HOST_MESH_HPP
#define FVENS_B200_HOST_MESH_HPP

#include <vector>
#include <string>
#include <array>
#include <cassert>

namespace fvens {

using freal = double;
using fint = int;
using EIndex = int;
constexpr int NDIM = 2;
constexpr int NVARS = 4;

/// Raw mesh as read from a file (reference: src/mesh/meshreaders.hpp:29-56). Row-major arrays.
struct MeshData
{
	fint npoin = 0, nelem = 0, nbface = 0;
	std::vector<int> nnode, nfael;
	int maxnnode = 0, maxnfael = 0, nnofa = 2, nbtag = 0, ndtag = 0;
	std::vector<freal> coords;      ///< [npoin][2]
	std::vector<fint> inpoel;       ///< [nelem][maxnnode], -1 padded
	std::vector<fint> bface;        ///< [nbface][nnofa+nbtag]
	std::vector<int> vol_regions;   ///< [nelem][ndtag]
};

/// Reads Gmsh-2.2 ASCII (.msh) or SU2 (.su2) (reference: src/mesh/meshreaders.cpp:35-64)
MeshData readMesh(const std::string mfile);

/// One connectivity (inter-subdomain) face; columns as the reference's connface (mesh.hpp:60-70)
struct ConnFace { fint elem; EIndex eface; int nbrrank; fint nbrglobalelem; fint globalface; };

template <typename scalar, int ndim>
class UMesh
{
	static_assert(ndim == 2, "only 2D meshes");
public:
	UMesh();
	UMesh(const MeshData& md);
	~UMesh();

	scalar gcoords(const fint pointno, const int dim) const { return coords[(size_t)pointno*ndim+dim]; }
	fint ginpoel(const fint elemnum, const int localnodenum) const { return inpoel[(size_t)elemnum*maxnnode+localnodenum]; }
	fint gbface(const fint facenum, const int locindex) const { return bface[(size_t)facenum*(nnofa+nbtag)+locindex]; }
	fint gconnface(const fint icface, const int infoindex) const {
		const ConnFace& c = connface[icface];
		switch(infoindex) { case 0: return c.elem; case 1: return c.eface; case 2: return c.nbrrank;
		                    case 3: return c.nbrglobalelem; default: return c.globalface; }
	}
	fint gesup(const fint i) const { return esup[i]; }
	fint gesup_p(const fint i) const { return esup_p[i]; }
	fint gesuel(const fint ielem, const int jface) const { return esuel[(size_t)ielem*maxnfael+jface]; }
	fint gelemface(const fint ielem, const EIndex ifael) const { return elemface[(size_t)ielem*maxnfael+ifael]; }
	fint gglobalElemIndex(const fint iel) const { return globalElemIndex.empty() ? iel : globalElemIndex[iel]; }
	fint gintfac(const fint face, const int i) const { return intfac[(size_t)face*4+i]; }

	fint gPhyBFaceStart() const { return 0; }
	fint gPhyBFaceEnd() const { return nbface; }
	fint gSubDomFaceStart() const { return nbface; }
	fint gSubDomFaceEnd() const { return nbface+ninface; }
	fint gConnBFaceStart() const { return nbface+ninface; }
	fint gConnBFaceEnd() const { return naface; }
	fint gDomFaceStart() const { return nbface; }
	fint gDomFaceEnd() const { return naface; }
	fint gFaceStart() const { return 0; }
	fint gFaceEnd() const { return naface; }

	int gbtags(const fint face, const int i) const { return btags[(size_t)(face-gPhyBFaceStart())*nbtag+i]; }
	scalar garea(const fint ielem) const { return area[ielem]; }
	scalar gfacemetric(const fint iface, const int index) const { return facemetric[(size_t)iface*3+index]; }
	std::array<scalar,ndim> gnormal(const fint iface) const {
		return {facemetric[(size_t)iface*3], facemetric[(size_t)iface*3+1]};
	}

	fint gnelemglobal() const { return nelemglobal; }
	fint gnpoinglobal() const { return npoinglobal; }
	fint gnpoin() const { return npoin; }
	fint gnelem() const { return nelem; }
	fint gnbface() const { return nbface; }
	int gnnode(const int ielem) const { return nnode[ielem]; }
	fint gnaface() const { return naface; }
	fint gninface() const { return ninface; }
	fint gnConnFace() const { return nconnface; }
	int gnfael(const int ielem) const { return nfael[ielem]; }
	int gnnofa(const int) const { return nnofa; }
	int gnbtag() const { return nbtag; }
	int gndtag() const { return ndtag; }
	fint gmaxnfael() const { return maxnfael; }
	std::vector<fint> getConnectivityGlobalIndices() const;

	/// Writes the mesh as Gmsh 2.2 ASCII (reference: mesh.cpp:205-286; the job of utilities/convertformat.cpp)
	void writeGmsh2(const std::string mfile) const;
	void correctBoundaryFaceOrientation();
	void scoords(const fint pointno, const int dim, const scalar value) {
		assert(pointno < npoin); assert(dim < ndim);
		coords[(size_t)pointno*ndim+dim] = value;
	}
	/// New cell i is old cell permvec[i] (reference: mesh.cpp:85-99). Topology must be recomputed.
	void reorder_cells(const int *const permvec);
	void compute_areas();
	void compute_cell_centres(scalar *const centres) const;
	void compute_topological();
	void compute_face_data();
	/// Pairs up faces of the periodic boundary with marker bcm along `axis` (mesh.cpp:369-424)
	void compute_periodic_map(const int bcm, const int axis);
	fint gperiodicmap(const fint iface) const { return periodicmap.empty() ? -1 : periodicmap[iface]; }
	EIndex getNodeEIndex(const fint ielem, const EIndex iface, const int inode) const {
		return (iface + inode) % nnode[ielem];
	}
	EIndex getFaceEIndex(const bool phyboundary, const fint iface, const fint elem) const;

	/// Installs inter-subdomain faces (used by the partitioner); call before compute_topological
	void setConnectivity(const std::vector<ConnFace>& cf, const std::vector<fint>& globalElemIdx,
	                     fint nelemglob, fint npoinglob);

	// raw array access for the C ABI (non-owning views of the storage above)
	const scalar* coordsData() const { return coords.data(); }
	const fint* inpoelData() const { return inpoel.data(); }
	const int* nnodeData() const { return nnode.data(); }
	const fint* esuelData() const { return esuel.data(); }
	const fint* elemfaceData() const { return elemface.data(); }
	const fint* intfacData() const { return intfac.data(); }
	const int* btagsData() const { return btags.data(); }
	const scalar* facemetricData() const { return facemetric.data(); }
	const scalar* areaData() const { return area.data(); }
	const fint* bfaceData() const { return bface.data(); }
	/// partner boundary face of every boundary face (-1: none); null before any compute_periodic_map
	const fint* periodicmapData() const { return periodicmap.empty() ? nullptr : periodicmap.data(); }
	int gmaxnnode() const { return maxnnode; }

private:
	fint npoinglobal = 0, nelemglobal = 0;
	fint npoin = 0, nelem = 0, nbface = 0, naface = 0, ninface = 0, nconnface = 0;
	std::vector<int> nnode, nfael;
	int maxnnode = 0, maxnfael = 0, nnofa = 2, nbtag = 0, ndtag = 0;
	std::vector<scalar> coords;
	std::vector<fint> inpoel, bface;
	std::vector<int> vol_regions;
	std::vector<ConnFace> connface;
	std::vector<fint> globalElemIndex;
	std::vector<fint> esup_p, esup;
	std::vector<fint> esuel, elemface, intfac;
	std::vector<int> btags;
	std::vector<scalar> area, facemetric;
	std::vector<fint> periodicmap;

	struct HalfEdge { fint hi; fint cellface; };   ///< cellface = 4*cell + local face
	/// Buckets of half edges keyed by the smaller node index
	void build_edge_buckets(std::vector<fint>& start, std::vector<HalfEdge>& edges) const;
	/// Host cell and local face of every physical boundary face
	void find_bface_hosts(std::vector<fint>& host, std::vector<EIndex>& lface) const;
	void compute_elementsSurroundingPoints();
};

/// Reads, orients, builds topology and metrics: the serial call order of the reference's
/// constructMesh + preprocessMesh (src/mesh/ameshutils.cpp:39-153) without reordering.
UMesh<freal,NDIM> constructMesh(const std::string mesh_path);

}
#endif
