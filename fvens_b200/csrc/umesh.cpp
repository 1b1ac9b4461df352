/* Host mesh: readers, topology and metrics for fvens::UMesh<double,2> (class surface in
 * ../host/mesh.hpp). Behavioural contract = the reference's src/mesh/mesh.cpp and
 * src/mesh/meshreaders.cpp; the construction algorithms are linear-time edge-bucket matches.
 */
#include "../host/mesh.hpp"
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace fvens {

// ------------------------------------------------------------------------------------- readers

namespace {

std::runtime_error mesh_error(const std::string& msg) { return std::runtime_error("UMesh: " + msg); }

/// Gmsh 2.2 ASCII: $MeshFormat / $Nodes / $Elements. Linear edges (1), triangles (2), quads (3).
/// Edge records must precede cell records, as the reference assumes (meshreaders.cpp:237-258).
MeshData readGmsh2(const std::string& mfile)
{
	std::ifstream in(mfile);
	if(!in) throw mesh_error("could not open " + mfile);
	MeshData m;
	std::string line;
	for(int i = 0; i < 4; i++)
		if(!std::getline(in, line)) throw mesh_error("truncated Gmsh header in " + mfile);
	if(!(in >> m.npoin) || m.npoin <= 0) throw mesh_error("bad node count in " + mfile);
	m.coords.resize((size_t)m.npoin*NDIM);
	for(fint i = 0; i < m.npoin; i++) {
		long id; double z;
		in >> id >> m.coords[(size_t)i*2] >> m.coords[(size_t)i*2+1] >> z;
	}
	std::string tok;
	in >> tok >> tok;
	long nrec = 0;
	if(!(in >> nrec) || nrec <= 0) throw mesh_error("bad element count in " + mfile);

	struct Rec { int type, ntags; int tags[6]; fint nodes[4]; };
	std::vector<Rec> recs((size_t)nrec);
	for(long i = 0; i < nrec; i++) {
		Rec& r = recs[i];
		long id;
		in >> id >> r.type >> r.ntags;
		for(int j = 0; j < r.ntags; j++) { int t; in >> t; if(j < 6) r.tags[j] = t; }
		int nn;
		if(r.type == 1) { nn = 2; m.nbface++; m.nbtag = std::max(m.nbtag, r.ntags); }
		else if(r.type == 2) { nn = 3; m.nelem++; m.ndtag = std::max(m.ndtag, r.ntags); }
		else if(r.type == 3) { nn = 4; m.nelem++; m.ndtag = std::max(m.ndtag, r.ntags); }
		else throw mesh_error("only linear edge/triangle/quad Gmsh elements are supported");
		for(int j = 0; j < nn; j++) in >> r.nodes[j];
		if(!in) throw mesh_error("truncated element list in " + mfile);
	}
	if(m.nbtag > 6 || m.ndtag > 6) throw mesh_error("too many tags per element");
	for(fint i = 0; i < m.nbface; i++)
		if(recs[i].type != 1) throw mesh_error("boundary edges must come first in the element list");

	m.nnofa = 2;
	m.maxnnode = 3; m.maxnfael = 3;
	for(long i = m.nbface; i < nrec; i++) {
		if(recs[i].type == 1) throw mesh_error("boundary edges must come first in the element list");
		if(recs[i].type == 3) { m.maxnnode = 4; m.maxnfael = 4; }
	}
	const int bw = m.nnofa + m.nbtag;
	m.bface.assign((size_t)m.nbface*bw, 0);
	for(fint i = 0; i < m.nbface; i++) {
		m.bface[(size_t)i*bw] = recs[i].nodes[0]-1;
		m.bface[(size_t)i*bw+1] = recs[i].nodes[1]-1;
		for(int j = 0; j < m.nbtag; j++) m.bface[(size_t)i*bw+2+j] = j < recs[i].ntags ? recs[i].tags[j] : 0;
	}
	m.inpoel.assign((size_t)m.nelem*m.maxnnode, -1);
	m.vol_regions.assign((size_t)m.nelem*m.ndtag, 0);
	m.nnode.resize(m.nelem); m.nfael.resize(m.nelem);
	for(fint i = 0; i < m.nelem; i++) {
		const Rec& r = recs[(size_t)i+m.nbface];
		const int nn = r.type == 2 ? 3 : 4;
		m.nnode[i] = nn; m.nfael[i] = nn;
		for(int j = 0; j < nn; j++) m.inpoel[(size_t)i*m.maxnnode+j] = r.nodes[j]-1;
		for(int j = 0; j < m.ndtag; j++) m.vol_regions[(size_t)i*m.ndtag+j] = j < r.ntags ? r.tags[j] : 0;
	}
	return m;
}

/// Value after '=' on the next "KEY= value" line
long su2_value(std::ifstream& in, const char *what)
{
	std::string s;
	while(std::getline(in, s)) {
		const size_t eq = s.find('=');
		if(eq == std::string::npos) continue;
		return std::stol(s.substr(eq+1));
	}
	throw mesh_error(std::string("SU2: missing ") + what);
}

/// SU2 native format with integer marker tags (meshreaders.cpp:267-395)
MeshData readSU2(const std::string& mfile)
{
	std::ifstream in(mfile);
	if(!in) throw mesh_error("could not open " + mfile);
	MeshData m;
	if(su2_value(in, "NDIME") != 2) throw mesh_error("SU2: only 2D meshes");
	m.nelem = (fint)su2_value(in, "NELEM");
	std::vector<fint> tmp((size_t)m.nelem*4, -1);
	m.nnode.resize(m.nelem); m.nfael.resize(m.nelem);
	m.maxnnode = 3;
	for(fint i = 0; i < m.nelem; i++) {
		int vtk; long idx;
		in >> vtk;
		if(vtk == 5) m.nnode[i] = 3;
		else if(vtk == 9) { m.nnode[i] = 4; m.maxnnode = 4; }
		else throw mesh_error("SU2: unsupported element type");
		m.nfael[i] = m.nnode[i];
		for(int j = 0; j < m.nnode[i]; j++) in >> tmp[(size_t)i*4+j];
		in >> idx;
	}
	m.maxnfael = m.maxnnode;
	m.inpoel.assign((size_t)m.nelem*m.maxnnode, -1);
	for(fint i = 0; i < m.nelem; i++)
		for(int j = 0; j < m.nnode[i]; j++) m.inpoel[(size_t)i*m.maxnnode+j] = tmp[(size_t)i*4+j];
	std::string rest; std::getline(in, rest);
	m.npoin = (fint)su2_value(in, "NPOIN");
	m.coords.resize((size_t)m.npoin*2);
	for(fint i = 0; i < m.npoin; i++) {
		long idx;
		in >> m.coords[(size_t)i*2] >> m.coords[(size_t)i*2+1] >> idx;
	}
	std::getline(in, rest);
	const long nmark = su2_value(in, "NMARK");
	m.nnofa = 2; m.nbtag = 1; m.ndtag = 0;
	for(long im = 0; im < nmark; im++) {
		const int tag = (int)su2_value(in, "MARKER_TAG");
		const long nf = su2_value(in, "MARKER_ELEMS");
		for(long k = 0; k < nf; k++) {
			int vtk; fint a, b;
			in >> vtk >> a >> b;
			m.bface.push_back(a); m.bface.push_back(b); m.bface.push_back(tag);
		}
		m.nbface += (fint)nf;
		std::getline(in, rest);
	}
	if(!in && !in.eof()) throw mesh_error("SU2: read error in " + mfile);
	return m;
}

} // anonymous

MeshData readMesh(const std::string mfile)
{
	const size_t dot = mfile.find_last_of('.');
	const std::string ext = dot == std::string::npos ? "" : mfile.substr(dot+1);
	if(ext == "su2") return readSU2(mfile);
	return readGmsh2(mfile);
}

// ------------------------------------------------------------------------------------- UMesh

template <typename scalar, int ndim> UMesh<scalar,ndim>::UMesh() { }
template <typename scalar, int ndim> UMesh<scalar,ndim>::~UMesh() { }

template <typename scalar, int ndim>
UMesh<scalar,ndim>::UMesh(const MeshData& md)
	: npoinglobal(md.npoin), nelemglobal(md.nelem), npoin(md.npoin), nelem(md.nelem), nbface(md.nbface),
	  nnode(md.nnode), nfael(md.nfael), maxnnode(md.maxnnode), maxnfael(md.maxnfael), nnofa(md.nnofa),
	  nbtag(md.nbtag), ndtag(md.ndtag), coords(md.coords.begin(), md.coords.end()),
	  inpoel(md.inpoel), bface(md.bface), vol_regions(md.vol_regions)
{
	if(nnofa != 2) throw mesh_error("only linear faces are supported");
	if(maxnnode > 4 || maxnfael > 4) throw mesh_error("only triangles and quadrangles are supported");
}

template <typename scalar, int ndim>
void UMesh<scalar,ndim>::setConnectivity(const std::vector<ConnFace>& cf, const std::vector<fint>& gei,
                                         fint nelemglob, fint npoinglob)
{
	connface = cf; nconnface = (fint)cf.size(); globalElemIndex = gei;
	nelemglobal = nelemglob; npoinglobal = npoinglob;
}

template <typename scalar, int ndim>
std::vector<fint> UMesh<scalar,ndim>::getConnectivityGlobalIndices() const
{
	std::vector<fint> g(nconnface);
	for(fint i = 0; i < nconnface; i++) g[i] = connface[i].nbrglobalelem;
	return g;
}

/// Counting sort of all (cell, local face) half edges into buckets keyed by min(node a, node b).
template <typename scalar, int ndim>
void UMesh<scalar,ndim>::build_edge_buckets(std::vector<fint>& start, std::vector<HalfEdge>& edges) const
{
	start.assign((size_t)npoin+1, 0);
	for(fint ie = 0; ie < nelem; ie++) {
		const int nn = nnode[ie];
		for(int j = 0; j < nn; j++) {
			const fint a = ginpoel(ie,j), b = ginpoel(ie,(j+1)%nn);
			start[std::min(a,b)+1]++;
		}
	}
	for(fint i = 0; i < npoin; i++) start[i+1] += start[i];
	edges.resize(start[npoin]);
	std::vector<fint> pos(start.begin(), start.end()-1);
	for(fint ie = 0; ie < nelem; ie++) {
		const int nn = nnode[ie];
		for(int j = 0; j < nn; j++) {
			const fint a = ginpoel(ie,j), b = ginpoel(ie,(j+1)%nn);
			edges[pos[std::min(a,b)]++] = HalfEdge{std::max(a,b), 4*ie+j};
		}
	}
}

template <typename scalar, int ndim>
void UMesh<scalar,ndim>::find_bface_hosts(std::vector<fint>& host, std::vector<EIndex>& lface) const
{
	std::vector<fint> start; std::vector<HalfEdge> edges;
	build_edge_buckets(start, edges);
	host.assign(nbface, -1); lface.assign(nbface, -1);
	for(fint f = 0; f < nbface; f++) {
		const fint a = gbface(f,0), b = gbface(f,1);
		const fint lo = std::min(a,b), hi = std::max(a,b);
		int nfound = 0;
		for(fint k = start[lo]; k < start[lo+1]; k++)
			if(edges[k].hi == hi) { host[f] = edges[k].cellface/4; lface[f] = edges[k].cellface%4; nfound++; }
		if(nfound != 1)
			throw mesh_error("boundary face " + std::to_string(f) + " does not have exactly one host cell");
	}
}

template <typename scalar, int ndim>
void UMesh<scalar,ndim>::correctBoundaryFaceOrientation()
{
	std::vector<fint> host; std::vector<EIndex> lface;
	find_bface_hosts(host, lface);
	const int bw = nnofa+nbtag;
	for(fint f = 0; f < nbface; f++) {
		const fint h = host[f];
		const fint n0 = ginpoel(h, getNodeEIndex(h,lface[f],0)), n1 = ginpoel(h, getNodeEIndex(h,lface[f],1));
		if(n0 != gbface(f,0) || n1 != gbface(f,1))
			std::swap(bface[(size_t)f*bw], bface[(size_t)f*bw+1]);
	}
}

template <typename scalar, int ndim>
void UMesh<scalar,ndim>::reorder_cells(const int *const permvec)
{
	const std::vector<fint> oldinpoel = inpoel;
	const std::vector<int> oldnnode = nnode, oldnfael = nfael, oldvr = vol_regions;
	for(fint i = 0; i < nelem; i++) {
		const fint o = permvec[i];
		for(int j = 0; j < maxnnode; j++) inpoel[(size_t)i*maxnnode+j] = oldinpoel[(size_t)o*maxnnode+j];
		nnode[i] = oldnnode[o]; nfael[i] = oldnfael[o];
		for(int j = 0; j < ndtag; j++) vol_regions[(size_t)i*ndtag+j] = oldvr[(size_t)o*ndtag+j];
	}
}

template <typename scalar, int ndim>
void UMesh<scalar,ndim>::compute_elementsSurroundingPoints()
{
	esup_p.assign((size_t)npoin+1, 0);
	for(fint ie = 0; ie < nelem; ie++)
		for(int j = 0; j < nnode[ie]; j++) esup_p[ginpoel(ie,j)+1]++;
	for(fint i = 0; i < npoin; i++) esup_p[i+1] += esup_p[i];
	esup.resize(esup_p[npoin]);
	std::vector<fint> pos(esup_p.begin(), esup_p.end()-1);
	for(fint ie = 0; ie < nelem; ie++)
		for(int j = 0; j < nnode[ie]; j++) esup[pos[ginpoel(ie,j)]++] = ie;
}

/// esup, esuel, intfac, elemface, btags. Face numbering and orientation as mesh.cpp:660-762.
template <typename scalar, int ndim>
void UMesh<scalar,ndim>::compute_topological()
{
	if(maxnfael == 0) { maxnfael = 3; for(int n : nfael) maxnfael = std::max(maxnfael, n); }
	compute_elementsSurroundingPoints();

	std::vector<fint> start; std::vector<HalfEdge> edges;
	build_edge_buckets(start, edges);
	esuel.assign((size_t)nelem*maxnfael, -1);
	for(fint p = 0; p < npoin; p++)
		for(fint k = start[p]; k < start[p+1]; k++)
			for(fint l = k+1; l < start[p+1]; l++)
				if(edges[k].hi == edges[l].hi) {
					const fint ck = edges[k].cellface, cl = edges[l].cellface;
					esuel[(size_t)(ck/4)*maxnfael + ck%4] = cl/4;
					esuel[(size_t)(cl/4)*maxnfael + cl%4] = ck/4;
				}

	ninface = 0;
	for(fint ie = 0; ie < nelem; ie++)
		for(int j = 0; j < nfael[ie]; j++) {
			const fint je = gesuel(ie,j);
			if(je > ie && je < nelem) ninface++;
		}
	naface = ninface + nbface + nconnface;
	intfac.assign((size_t)naface*4, -1);
	elemface.assign((size_t)nelem*maxnfael, -1);
	btags.assign((size_t)nbface*nbtag, 0);

	std::vector<fint> host; std::vector<EIndex> lface;
	host.assign(nbface, -1); lface.assign(nbface, -1);
	for(fint f = 0; f < nbface; f++) {
		const fint a = gbface(f,0), b = gbface(f,1);
		const fint lo = std::min(a,b), hi = std::max(a,b);
		int nfound = 0;
		for(fint k = start[lo]; k < start[lo+1]; k++)
			if(edges[k].hi == hi) { host[f] = edges[k].cellface/4; lface[f] = edges[k].cellface%4; nfound++; }
		if(nfound != 1)
			throw mesh_error("boundary face " + std::to_string(f) + " does not have exactly one host cell");
		intfac[(size_t)f*4] = host[f];
		intfac[(size_t)f*4+1] = nelem + nconnface + f;
		intfac[(size_t)f*4+2] = a;
		intfac[(size_t)f*4+3] = b;
		for(int j = 0; j < nbtag; j++) btags[(size_t)f*nbtag+j] = (int)gbface(f,nnofa+j);
		esuel[(size_t)host[f]*maxnfael+lface[f]] = nelem + nconnface + f;
		elemface[(size_t)host[f]*maxnfael+lface[f]] = f;
	}

	// interior faces in ascending (cell, local face) order; the matching local face of the right
	// cell is found from its half edge in the same bucket
	fint fi = nbface;
	for(fint ie = 0; ie < nelem; ie++) {
		const int nn = nnode[ie];
		for(int j = 0; j < nn; j++) {
			const fint je = gesuel(ie,j);
			if(!(je > ie && je < nelem)) continue;
			const fint a = ginpoel(ie,j), b = ginpoel(ie,(j+1)%nn);
			intfac[(size_t)fi*4] = ie; intfac[(size_t)fi*4+1] = je;
			intfac[(size_t)fi*4+2] = a; intfac[(size_t)fi*4+3] = b;
			elemface[(size_t)ie*maxnfael+j] = fi;
			const fint lo = std::min(a,b), hi = std::max(a,b);
			for(fint k = start[lo]; k < start[lo+1]; k++)
				if(edges[k].hi == hi && edges[k].cellface/4 == je)
					elemface[(size_t)je*maxnfael + edges[k].cellface%4] = fi;
			fi++;
		}
	}

	for(fint ic = 0; ic < nconnface; ic++) {
		const fint f = nbface + ninface + ic;
		const fint e = connface[ic].elem; const EIndex lf = connface[ic].eface;
		intfac[(size_t)f*4] = e;
		intfac[(size_t)f*4+1] = nelem + ic;
		intfac[(size_t)f*4+2] = ginpoel(e, getNodeEIndex(e,lf,0));
		intfac[(size_t)f*4+3] = ginpoel(e, getNodeEIndex(e,lf,1));
		esuel[(size_t)e*maxnfael+lf] = nelem + ic;
		elemface[(size_t)e*maxnfael+lf] = f;
	}

	for(fint ie = 0; ie < nelem; ie++)
		for(int j = 0; j < nfael[ie]; j++)
			if(gesuel(ie,j) < 0)
				throw mesh_error("cell " + std::to_string(ie) + " has an unmatched face that is not a boundary face");
}

template <typename scalar, int ndim>
EIndex UMesh<scalar,ndim>::getFaceEIndex(const bool, const fint iface, const fint elem) const
{
	for(int j = 0; j < nfael[elem]; j++)
		if(gelemface(elem,j) == iface) return j;
	return -1;
}

/// Shoelace formula over the first triangle (+ second triangle of a quad), mesh.cpp:289-313
template <typename scalar, int ndim>
void UMesh<scalar,ndim>::compute_areas()
{
	area.resize(nelem);
	for(fint i = 0; i < nelem; i++) {
		const scalar x0 = gcoords(ginpoel(i,0),0), y0 = gcoords(ginpoel(i,0),1);
		const scalar x1 = gcoords(ginpoel(i,1),0), y1 = gcoords(ginpoel(i,1),1);
		const scalar x2 = gcoords(ginpoel(i,2),0), y2 = gcoords(ginpoel(i,2),1);
		scalar a = 0.5*(x0*(y1 - y2) - y0*(x1 - x2) + x1*y2 - x2*y1);
		if(nnode[i] == 4) {
			const scalar x3 = gcoords(ginpoel(i,3),0), y3 = gcoords(ginpoel(i,3),1);
			a += 0.5*(x0*(y2 - y3) - y0*(x2 - x3) + x2*y3 - x3*y2);
		}
		area[i] = a;
	}
}

template <typename scalar, int ndim>
void UMesh<scalar,ndim>::compute_cell_centres(scalar *const centres) const
{
	for(fint i = 0; i < nelem; i++)
		for(int d = 0; d < ndim; d++) {
			scalar c = 0;
			for(int j = 0; j < nnode[i]; j++) c += gcoords(ginpoel(i,j),d);
			centres[(size_t)i*ndim+d] = c/(scalar)nnode[i];
		}
}

/// Unit normal (rotated edge vector, pointing from the left to the right cell) and length
template <typename scalar, int ndim>
void UMesh<scalar,ndim>::compute_face_data()
{
	facemetric.resize((size_t)naface*3);
	for(fint f = 0; f < naface; f++) {
		const fint p0 = gintfac(f,2), p1 = gintfac(f,3);
		scalar nx = gcoords(p1,1) - gcoords(p0,1);
		scalar ny = -1.0*(gcoords(p1,0) - gcoords(p0,0));
		const scalar len = std::sqrt(nx*nx + ny*ny);
		facemetric[(size_t)f*3] = nx/len;
		facemetric[(size_t)f*3+1] = ny/len;
		facemetric[(size_t)f*3+2] = len;
	}
}

/// Faces with marker bcm whose midpoints agree in the coordinate other than `axis` are partners.
template <typename scalar, int ndim>
void UMesh<scalar,ndim>::compute_periodic_map(const int bcm, const int axis)
{
	if(bcm < 0 || axis < 0 || axis > 1) return;
	// one call per periodic direction (each with its own marker): earlier pairings are kept
	if((fint)periodicmap.size() != nbface) periodicmap.assign(nbface, -1);
	const int ax = 1-axis;
	struct Key { scalar c; fint f; };
	std::vector<Key> keys;
	for(fint f = 0; f < nbface; f++)
		if(gbtags(f,0) == bcm)
			keys.push_back(Key{ (scalar)0.5*(gcoords(gintfac(f,2),ax) + gcoords(gintfac(f,3),ax)), f });
	std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b){ return a.c < b.c || (a.c == b.c && a.f < b.f); });
	const scalar tol = 1e-8;
	for(size_t k = 0; k+1 < keys.size(); k++)
		if(std::fabs(keys[k].c - keys[k+1].c) <= tol && periodicmap[keys[k].f] < 0) {
			periodicmap[keys[k].f] = keys[k+1].f;
			periodicmap[keys[k+1].f] = keys[k].f;
			k++;
		}
}

/// Gmsh 2.2 ASCII: nodes with a zero z coordinate; boundary faces first (line = type 1), then cells (triangle = 2,
/// quadrangle = 3). Gmsh wants at least two tags per element: a missing second tag is 1 + the last physical tag
/// (1 when there is no tag at all), so that different physical groups keep different elementary ids.
template <typename scalar, int ndim>
void UMesh<scalar,ndim>::writeGmsh2(const std::string mfile) const
{
	std::ofstream out(mfile);
	if(!out) throw std::runtime_error("UMesh: writeGmsh2(): cannot open " + mfile + " for writing");
	out.precision(20);
	out << "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n" << npoin << '\n';
	for(fint p = 0; p < npoin; p++)
		out << p+1 << ' ' << coords[(size_t)p*ndim] << ' ' << coords[(size_t)p*ndim+1] << ' ' << 0.0 << '\n';
	out << "$EndNodes\n$Elements\n" << nelem + nbface << '\n';

	const auto tags = [&out](const int *const t, const int nt) {
		out << ' ' << std::max(nt, 2);
		for(int k = 0; k < nt; k++) out << ' ' << t[k];
		for(int k = nt; k < 2; k++) out << ' ' << (nt == 0 ? 1 : 1 + t[nt-1]);
	};
	const int bw = nnofa + nbtag;
	for(fint f = 0; f < nbface; f++) {
		out << f+1 << ' ' << (nnofa == 3 ? 8 : 1);
		tags(&bface[(size_t)f*bw] + nnofa, nbtag);
		for(int k = 0; k < nnofa; k++) out << ' ' << bface[(size_t)f*bw+k] + 1;
		out << '\n';
	}
	for(fint e = 0; e < nelem; e++) {
		out << nbface+e+1 << ' ' << (nnode[e] == 3 ? 2 : 3);
		tags(ndtag ? &vol_regions[(size_t)e*ndtag] : nullptr, ndtag);
		for(int k = 0; k < nnode[e]; k++) out << ' ' << inpoel[(size_t)e*maxnnode+k] + 1;
		out << '\n';
	}
	out << "$EndElements\n";
}

template class UMesh<freal,NDIM>;

UMesh<freal,NDIM> constructMesh(const std::string mesh_path)
{
	UMesh<freal,NDIM> m(readMesh(mesh_path));
	m.correctBoundaryFaceOrientation();
	m.compute_topological();
	m.compute_areas();
	m.compute_face_data();
	return m;
}

}
