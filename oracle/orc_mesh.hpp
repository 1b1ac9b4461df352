/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_physics.hpp header).
 *
 * CPU restatement of the hybrid tri/quad mesh of FVENS: readers, boundary-face orientation,
 * topology (esup -> esuel -> intfac/elemface), areas, face metrics. Follows mesh/mesh.cpp and
 * mesh/meshreaders.cpp of the reference; the face numbering produced here is the reference's
 * (boundary faces first, then interior faces in ascending (cell, local face) order) because the
 * single-threaded accumulation order of the residual depends on it.
 */
#ifndef ORC_MESH_HPP
#define ORC_MESH_HPP

#include <vector>
#include <string>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <algorithm>
#include <cmath>
#include <iostream>

namespace orc {

struct Mesh {
	int npoin = 0, nelem = 0, nbface = 0, naface = 0, ninface = 0;
	int nbtag = 1;
	std::vector<double> coords;     ///< [npoin][2]
	std::vector<int> inpoel;        ///< [nelem][4], -1 padded
	std::vector<int> nnode;         ///< [nelem] 3 or 4 (== nfael for linear cells)
	std::vector<int> bface;         ///< [nbface][3]: node0, node1, first tag

	// derived
	std::vector<int> esup_p, esup;
	std::vector<int> esuel;         ///< [nelem][4]
	std::vector<int> elemface;      ///< [nelem][4]
	std::vector<int> intfac;        ///< [naface][4]: L, R, n0, n1
	std::vector<int> btags;         ///< [nbface]
	std::vector<double> facemetric; ///< [naface][3]: nx, ny, len
	std::vector<double> area;       ///< [nelem]

	double x(int ip, int d) const { return coords[2*ip+d]; }
};

// ---- readers ------------------------------------------------------------------------------

/// mesh/meshreaders.cpp:66-265. Only linear edges (1), triangles (2), quads (3).
static inline Mesh read_gmsh2(const std::string &fname)
{
	std::ifstream in(fname);
	if(!in) throw std::runtime_error("oracle: cannot open " + fname);
	Mesh m;
	std::string line;
	for(int i = 0; i < 4; i++) std::getline(in, line);   // $MeshFormat, version, $End, $Nodes
	in >> m.npoin;
	m.coords.resize(2*(size_t)m.npoin);
	for(int i = 0; i < m.npoin; i++) {
		int id; double z;
		in >> id >> m.coords[2*i] >> m.coords[2*i+1] >> z;
	}
	std::string tok;
	in >> tok; in >> tok;       // $EndNodes $Elements
	int nelm; in >> nelm;
	struct Rec { int type; int tags[8]; int ntags; int nodes[4]; };
	std::vector<Rec> recs(nelm);
	m.nbface = 0; m.nelem = 0; m.nbtag = 0;
	for(int i = 0; i < nelm; i++) {
		int id; Rec &r = recs[i];
		in >> id >> r.type >> r.ntags;
		for(int j = 0; j < r.ntags; j++) { int t; in >> t; if(j < 8) r.tags[j] = t; }
		int nn;
		switch(r.type) {
		case 1: nn = 2; m.nbface++; if(r.ntags > m.nbtag) m.nbtag = r.ntags; break;
		case 2: nn = 3; m.nelem++; break;
		case 3: nn = 4; m.nelem++; break;
		default: throw std::runtime_error("oracle: unsupported gmsh element type");
		}
		for(int j = 0; j < nn; j++) in >> r.nodes[j];
	}
	// the first nbface records are taken to be the boundary faces (meshreaders.cpp:237-258)
	m.bface.resize(3*(size_t)m.nbface);
	for(int i = 0; i < m.nbface; i++) {
		m.bface[3*i+0] = recs[i].nodes[0]-1;
		m.bface[3*i+1] = recs[i].nodes[1]-1;
		m.bface[3*i+2] = recs[i].tags[0];
	}
	m.inpoel.assign(4*(size_t)m.nelem, -1);
	m.nnode.resize(m.nelem);
	for(int i = 0; i < m.nelem; i++) {
		const Rec &r = recs[i+m.nbface];
		const int nn = (r.type == 2) ? 3 : 4;
		m.nnode[i] = nn;
		for(int j = 0; j < nn; j++) m.inpoel[4*i+j] = r.nodes[j]-1;
	}
	return m;
}

/// mesh/meshreaders.cpp:267-395
static inline Mesh read_su2(const std::string &fname)
{
	std::ifstream fin(fname);
	if(!fin) throw std::runtime_error("oracle: cannot open " + fname);
	Mesh m;
	std::string dum;
	std::getline(fin, dum, '='); std::getline(fin, dum);     // NDIME
	std::getline(fin, dum, '='); std::getline(fin, dum);
	m.nelem = std::stoi(dum);
	m.inpoel.assign(4*(size_t)m.nelem, -1);
	m.nnode.resize(m.nelem);
	for(int iel = 0; iel < m.nelem; iel++) {
		int id, ddum; fin >> id;
		if(id == 5) m.nnode[iel] = 3;
		else if(id == 9) m.nnode[iel] = 4;
		else throw std::runtime_error("oracle: unknown SU2 element");
		for(int i = 0; i < m.nnode[iel]; i++) fin >> m.inpoel[4*iel+i];
		fin >> ddum;
	}
	std::getline(fin, dum);
	std::getline(fin, dum, '='); std::getline(fin, dum);
	m.npoin = std::stoi(dum);
	m.coords.resize(2*(size_t)m.npoin);
	for(int ip = 0; ip < m.npoin; ip++) {
		int ddum;
		fin >> m.coords[2*ip] >> m.coords[2*ip+1] >> ddum;
	}
	std::getline(fin, dum);
	std::getline(fin, dum, '='); std::getline(fin, dum);
	const int nbmarkers = std::stoi(dum);
	m.nbface = 0;
	for(int ib = 0; ib < nbmarkers; ib++) {
		std::getline(fin, dum, '='); std::getline(fin, dum);
		const int tag = std::stoi(dum);
		std::getline(fin, dum, '='); std::getline(fin, dum);
		const int nf = std::stoi(dum);
		for(int i = 0; i < nf; i++) {
			int ddum, a, b; fin >> ddum >> a >> b;
			m.bface.push_back(a); m.bface.push_back(b); m.bface.push_back(tag);
		}
		m.nbface += nf;
		if(ib < nbmarkers-1) std::getline(fin, dum);
	}
	return m;
}

/// mesh/meshreaders.cpp:35-64
static inline Mesh read_mesh(const std::string &fname)
{
	const size_t dot = fname.find_last_of('.');
	if(dot != std::string::npos && fname.substr(dot+1) == "su2") return read_su2(fname);
	return read_gmsh2(fname);
}

// ---- topology -----------------------------------------------------------------------------

/// mesh/mesh.cpp:427-468
static inline void compute_esup(Mesh &m)
{
	m.esup_p.assign(m.npoin+1, 0);
	for(int i = 0; i < m.nelem; i++)
		for(int j = 0; j < m.nnode[i]; j++)
			m.esup_p[m.inpoel[4*i+j]+1] += 1;
	for(int i = 1; i < m.npoin+1; i++) m.esup_p[i] += m.esup_p[i-1];
	m.esup.assign(m.esup_p[m.npoin], 0);
	for(int i = 0; i < m.nelem; i++)
		for(int j = 0; j < m.nnode[i]; j++) {
			const int ip = m.inpoel[4*i+j];
			m.esup[m.esup_p[ip]] = i;
			m.esup_p[ip] += 1;
		}
	for(int i = m.npoin; i >= 1; i--) m.esup_p[i] = m.esup_p[i-1];
	m.esup_p[0] = 0;
}

/// mesh/mesh.cpp:470-545. Local face i of a cell has local nodes (i, i+1 mod nnode).
static inline void compute_esuel(Mesh &m)
{
	m.esuel.assign(4*(size_t)m.nelem, -1);
	for(int ielem = 0; ielem < m.nelem; ielem++) {
		const int nn = m.nnode[ielem];
		for(int ifael = 0; ifael < nn; ifael++) {
			const int a = m.inpoel[4*ielem + ifael];
			const int b = m.inpoel[4*ielem + (ifael+1)%nn];
			for(int istor = m.esup_p[a]; istor < m.esup_p[a+1]; istor++) {
				const int jelem = m.esup[istor];
				if(jelem == ielem) continue;
				const int nj = m.nnode[jelem];
				for(int jfael = 0; jfael < nj; jfael++) {
					const int c = m.inpoel[4*jelem + jfael];
					const int d = m.inpoel[4*jelem + (jfael+1)%nj];
					if((c == a || c == b) && (d == a || d == b)) {
						m.esuel[4*ielem+ifael] = jelem;
						m.esuel[4*jelem+jfael] = ielem;
					}
				}
			}
		}
	}
}

/// mesh/mesh.cpp:602-658 and 547-600: host cell and local face index of each boundary face.
static inline void bface_hosts(const Mesh &m, std::vector<int> &host, std::vector<int> &lface)
{
	host.resize(m.nbface); lface.resize(m.nbface);
	for(int iface = 0; iface < m.nbface; iface++) {
		const int a = m.bface[3*iface], b = m.bface[3*iface+1];
		int found = -1, count = 0;
		for(int ia = m.esup_p[a]; ia < m.esup_p[a+1]; ia++)
			for(int ib = m.esup_p[b]; ib < m.esup_p[b+1]; ib++)
				if(m.esup[ia] == m.esup[ib]) { found = m.esup[ia]; count++; }
		if(count != 1)
			throw std::logic_error("oracle: boundary face does not have exactly one host cell");
		host[iface] = found;
		const int nn = m.nnode[found];
		int lf = -1;
		for(int k = 0; k < nn; k++) {
			const int c = m.inpoel[4*found+k], d = m.inpoel[4*found+(k+1)%nn];
			if((c == a || c == b) && (d == a || d == b)) { lf = k; break; }
		}
		if(lf < 0) throw std::logic_error("oracle: boundary face not found in its host cell");
		lface[iface] = lf;
	}
}

/// mesh/mesh.cpp:55-82
static inline void correct_bface_orientation(Mesh &m)
{
	compute_esup(m);
	std::vector<int> host, lface;
	bface_hosts(m, host, lface);
	for(int iface = 0; iface < m.nbface; iface++) {
		const int h = host[iface], nn = m.nnode[h], k = lface[iface];
		if(m.inpoel[4*h+k] != m.bface[3*iface] || m.inpoel[4*h+(k+1)%nn] != m.bface[3*iface+1])
			std::swap(m.bface[3*iface], m.bface[3*iface+1]);
	}
}

/// mesh/mesh.cpp:660-762 (nconnface = 0)
static inline void compute_face_connectivity(Mesh &m)
{
	m.ninface = 0;
	for(int ie = 0; ie < m.nelem; ie++)
		for(int in = 0; in < m.nnode[ie]; in++) {
			const int je = m.esuel[4*ie+in];
			if(je > ie && je < m.nelem) m.ninface++;
		}
	m.naface = m.ninface + m.nbface;
	m.intfac.assign(4*(size_t)m.naface, -1);
	m.elemface.assign(4*(size_t)m.nelem, -1);
	m.btags.resize(m.nbface);

	std::vector<int> host, lface;
	bface_hosts(m, host, lface);
	for(int iface = 0; iface < m.nbface; iface++) {
		m.intfac[4*iface+0] = host[iface];
		m.intfac[4*iface+1] = m.nelem + iface;
		m.intfac[4*iface+2] = m.bface[3*iface];
		m.intfac[4*iface+3] = m.bface[3*iface+1];
		m.btags[iface] = m.bface[3*iface+2];
		m.esuel[4*host[iface]+lface[iface]] = m.nelem + iface;
		m.elemface[4*host[iface]+lface[iface]] = iface;
	}

	int faceindex = m.nbface;
	for(int ie = 0; ie < m.nelem; ie++) {
		const int nn = m.nnode[ie];
		for(int in = 0; in < nn; in++) {
			const int je = m.esuel[4*ie+in];
			if(je > ie && je < m.nelem) {
				const int in1 = (in+1)%nn;
				m.intfac[4*faceindex+0] = ie;
				m.intfac[4*faceindex+1] = je;
				m.intfac[4*faceindex+2] = m.inpoel[4*ie+in];
				m.intfac[4*faceindex+3] = m.inpoel[4*ie+in1];
				m.elemface[4*ie+in] = faceindex;
				for(int jnode = 0; jnode < m.nnode[je]; jnode++)
					if(m.inpoel[4*ie+in1] == m.inpoel[4*je+jnode])
						m.elemface[4*je+jnode] = faceindex;
				faceindex++;
			}
		}
	}
}

/// mesh/mesh.cpp:289-313
static inline void compute_areas(Mesh &m)
{
	m.area.resize(m.nelem);
	auto X = [&](int i, int k, int d) { return m.x(m.inpoel[4*i+k], d); };
	for(int i = 0; i < m.nelem; i++) {
		double a = 0.5*(X(i,0,0)*(X(i,1,1) - X(i,2,1)) - X(i,0,1)*(X(i,1,0) - X(i,2,0))
		                + X(i,1,0)*X(i,2,1) - X(i,2,0)*X(i,1,1));
		if(m.nnode[i] == 4)
			a += 0.5*(X(i,0,0)*(X(i,2,1) - X(i,3,1)) - X(i,0,1)*(X(i,2,0) - X(i,3,0))
			          + X(i,2,0)*X(i,3,1) - X(i,3,0)*X(i,2,1));
		m.area[i] = a;
	}
}

/// mesh/mesh.cpp:347-365
static inline void compute_face_data(Mesh &m)
{
	m.facemetric.resize(3*(size_t)m.naface);
	for(int i = 0; i < m.naface; i++) {
		const int p0 = m.intfac[4*i+2], p1 = m.intfac[4*i+3];
		double nx = m.x(p1,1) - m.x(p0,1);
		double ny = -1.0*(m.x(p1,0) - m.x(p0,0));
		const double len = std::sqrt(nx*nx + ny*ny);
		nx /= len;
		ny /= len;
		m.facemetric[3*i] = nx; m.facemetric[3*i+1] = ny; m.facemetric[3*i+2] = len;
	}
}

/// The call order of mesh/ameshutils.cpp:102-153 + preprocessMesh :39-100 without reordering
static inline void finalize_mesh(Mesh &m)
{
	correct_bface_orientation(m);
	compute_esup(m);
	compute_esuel(m);
	compute_face_connectivity(m);
	compute_areas(m);
	compute_face_data(m);
}

} // namespace orc
#endif
