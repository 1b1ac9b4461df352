/* Device mesh construction: cell renumbering (Hilbert curve / reverse Cuthill-McKee), tiling, the
 * per-tile face streams with their edge colouring, geometry precomputation, upload.
 *
 * Geometry restated from the reference (src/ relative): cell centres mesh/mesh.cpp:317-328, face
 * midpoints and ghost centres spatial/aspatial.cpp:37-119, least-squares matrices
 * spatial/agradientschemes.cpp:219-317, Venkatakrishnan length spatial/limitedlinearreconstruction.cpp:189-205.
 * Orientation (left/right cell, normal) of every face is carried over from the reference numbering
 * and is NOT re-derived from the device numbering.
 */
#include "engine.hpp"
#include <algorithm>
#include <array>
#include <numeric>
#include <cmath>
#include <cstring>
#include <queue>
#include <climits>
#include <cstdlib>

namespace fvg {

static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
const std::string &last_error() { return g_last_error; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
	g_last_error = std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + what + " at " + file
	               + ":" + std::to_string(line);
	return FVG_ERR_CUDA;
}

/// Hilbert curve index of (x,y) on a 2^order x 2^order grid
static inline uint64_t hilbert_d(uint32_t x, uint32_t y, int order)
{
	const uint32_t n = 1u << order;
	uint64_t d = 0;
	for(uint32_t s = n >> 1; s > 0; s >>= 1) {
		const uint32_t rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
		d += (uint64_t)s*s*((3*rx) ^ ry);
		if(ry == 0) {
			if(rx == 1) { x = n-1-x; y = n-1-y; }
			std::swap(x, y);
		}
	}
	return d;
}

void hilbert_order(int n, const double *rc, std::vector<int> &new2old)
{
	double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
	for(int i = 0; i < n; i++)
		for(int d = 0; d < 2; d++) { lo[d] = std::min(lo[d], rc[2*i+d]); hi[d] = std::max(hi[d], rc[2*i+d]); }
	const int order = 20;
	const double span = std::max(std::max(hi[0]-lo[0], hi[1]-lo[1]), 1e-300);
	const double scale = ((double)((1u << order) - 1))/span;
	std::vector<std::pair<uint64_t,int>> keys((size_t)n);
	for(int i = 0; i < n; i++) {
		const uint32_t x = (uint32_t)((rc[2*i]-lo[0])*scale), y = (uint32_t)((rc[2*i+1]-lo[1])*scale);
		keys[i] = std::make_pair(hilbert_d(x, y, order), i);
	}
	std::sort(keys.begin(), keys.end());
	new2old.resize(n);
	for(int i = 0; i < n; i++) new2old[i] = keys[i].second;
}

/// Reverse Cuthill-McKee on the cell adjacency graph (esuel), components started from a
/// minimum-degree cell. perm[new] = old.
void rcm_order(int n, int mw, const int *esuel, const int *nnode, std::vector<int> &new2old)
{
	std::vector<int> deg(n, 0);
	for(int i = 0; i < n; i++)
		for(int j = 0; j < nnode[i]; j++) { const int e = esuel[(size_t)i*mw+j]; if(e >= 0 && e < n) deg[i]++; }
	std::vector<int> order_by_deg(n);
	std::iota(order_by_deg.begin(), order_by_deg.end(), 0);
	std::stable_sort(order_by_deg.begin(), order_by_deg.end(), [&](int a, int b){ return deg[a] < deg[b]; });
	std::vector<char> seen(n, 0);
	std::vector<int> cm; cm.reserve(n);
	for(int s : order_by_deg) {
		if(seen[s]) continue;
		seen[s] = 1;
		size_t head = cm.size();
		cm.push_back(s);
		while(head < cm.size()) {
			const int c = cm[head++];
			int nb[4], k = 0;
			for(int j = 0; j < nnode[c]; j++) {
				const int e = esuel[(size_t)c*mw+j];
				if(e >= 0 && e < n && !seen[e]) { seen[e] = 1; nb[k++] = e; }
			}
			for(int p = 1; p < k; p++)     // insertion sort by (degree, index); k <= 4
				for(int q = p; q > 0 && (deg[nb[q]] < deg[nb[q-1]] || (deg[nb[q]] == deg[nb[q-1]] && nb[q] < nb[q-1])); q--)
					std::swap(nb[q], nb[q-1]);
			for(int q = 0; q < k; q++) cm.push_back(nb[q]);
		}
	}
	new2old.assign(cm.rbegin(), cm.rend());
}

template <typename T>
static int upload(fvg_mesh *m, const std::vector<T> &h, const T **dptr)
{
	if(m->device < 0) { *dptr = nullptr; return 0; }   // host-only inspection build
	void *p = nullptr;
	const size_t bytes = std::max<size_t>(h.size(), 1)*sizeof(T);
	FVG_CUDA(cudaMalloc(&p, bytes));
	m->allocs.push_back(p);
	if(!h.empty()) FVG_CUDA(cudaMemcpy(p, h.data(), h.size()*sizeof(T), cudaMemcpyHostToDevice));
	*dptr = static_cast<const T*>(p);
	return 0;
}

/** Builds the device mesh of one subdomain (or of the whole mesh when cell_rank is null).
 * Device cell numbering: the rank's own cells first, in the global locality order restricted to the
 * rank; then its ghost cells (cells of other ranks across a cut face - one layer, as the reference's
 * connectivity ghosts, src/mesh/mesh.hpp:60-70), grouped by owner rank and sorted by the owner's own
 * device index, so that a ghost block is exactly what the owner's pack kernel emits for this rank.
 * Cut faces keep the GLOBAL left/right roles and normal: both ranks evaluate the identical
 * expression, which keeps the scheme conservative to the last bit and the 1/2/4/8-GPU results equal. */
static int build(const fvg_host_mesh *hm, const fvg_mesh_opts *opts, const int *cell_rank, int rank, int nranks,
                 fvg_mesh *m)
{
	const int n = hm->nelem, nb = hm->nbface, nf = hm->naface, mw = hm->maxnnode;
	if(n <= 0 || nf <= 0 || !hm->coords || !hm->inpoel || !hm->nnode || !hm->esuel || !hm->elemface ||
	   !hm->intfac || !hm->facemetric || !hm->area || (nb > 0 && !hm->btags)) {
		set_error("fvg_mesh_create: incomplete host mesh");
		return FVG_ERR_INVALID;
	}
	if(hm->nconnface != 0) {
		set_error("fvg_mesh_create: pass the GLOBAL mesh plus a cell->rank map; pre-partitioned host meshes are not accepted");
		return FVG_ERR_UNSUPPORTED;
	}
	if(mw < 3 || mw > 4) { set_error("fvg_mesh_create: maxnnode must be 3 or 4"); return FVG_ERR_INVALID; }
	if(nranks < 1 || rank < 0 || rank >= nranks) { set_error("fvg_mesh_create: bad rank"); return FVG_ERR_INVALID; }
	const int TC = opts && opts->tile_cells > 0 ? opts->tile_cells : 256;
	if(TC % 32 != 0 || TC > 1024) { set_error("fvg_mesh_create: tile_cells must be a multiple of 32, <= 1024"); return FVG_ERR_INVALID; }
	// capacities of a tile's shared-memory staging areas: halo cells and stream entries
	const int HMAX = std::max(32, (3*TC/8 + 31)/32*32);
	const int EMAX = (2*TC + TC/8 + 31)/32*32;
	const int LCAP = (2*TC + TC/4 + 31)/32*32;    // cap on a tile's stream segment including bank padding
	// The face kernel takes a tile's entries FACE_BLOCK at a time: a tile just over a multiple of that size pays a whole
	// round for a handful of entries (all-quad meshes: 2 entries per cell + half the perimeter = 544 for a 16 x 16 tile,
	// which used to fall back to 204 cells). Tiles are cut at the largest multiple within the capacity instead
	// (FVG_TILE_ENTRY_CAP=0: the capacity itself, the cut of the earlier rounds).
	int ECAP = EMAX >= FACE_BLOCK ? EMAX/FACE_BLOCK*FACE_BLOCK : EMAX;
	if(const char *ev = getenv("FVG_TILE_ENTRY_CAP")) { const int v = atoi(ev); ECAP = v > 0 ? std::min(v, EMAX) : EMAX; }
	const int reorder = opts ? opts->reorder : FVG_REORDER_NONE;
	auto rank_of = [&](int o) { return cell_rank ? cell_rank[o] : 0; };
	if(cell_rank) for(int o = 0; o < n; o++) if(cell_rank[o] < 0 || cell_rank[o] >= nranks) { set_error("fvg_mesh_create: cell_rank entry out of range"); return FVG_ERR_INVALID; }

	// ---- geometry in reference numbering
	std::vector<double> rc(2*(size_t)n);
	for(int i = 0; i < n; i++)
		for(int d = 0; d < 2; d++) {
			double c = 0;
			for(int j = 0; j < hm->nnode[i]; j++) c += hm->coords[2*(size_t)hm->inpoel[(size_t)i*mw+j]+d];
			rc[2*(size_t)i+d] = c/(double)hm->nnode[i];
		}

	// ---- boundary markers -> slots of the flow's BC table (global, so that every rank agrees)
	m->h_btag.resize(nb);
	for(int b = 0; b < nb; b++) m->h_btag[b] = hm->btags[(size_t)b*hm->nbtag];
	m->h_markers.assign(m->h_btag.begin(), m->h_btag.end());
	std::sort(m->h_markers.begin(), m->h_markers.end());
	m->h_markers.erase(std::unique(m->h_markers.begin(), m->h_markers.end()), m->h_markers.end());
	if((int)m->h_markers.size() > MAX_BC) { set_error("fvg_mesh_create: more than 16 distinct boundary markers"); return FVG_ERR_UNSUPPORTED; }
	std::vector<int> bslot(nb);
	for(int b = 0; b < nb; b++)
		bslot[b] = (int)(std::lower_bound(m->h_markers.begin(), m->h_markers.end(), m->h_btag[b]) - m->h_markers.begin());
	// a marker is periodic when every one of its faces has a partner (then no boundary condition is ever evaluated for it)
	m->h_marker_periodic.assign(m->h_markers.size(), hm->bpartner ? 1 : 0);
	{
		std::vector<int> any(m->h_markers.size(), 0);
		for(int b = 0; b < nb; b++) {
			const bool paired = hm->bpartner && hm->bpartner[b] >= 0;
			if(!paired) m->h_marker_periodic[bslot[b]] = 0; else any[bslot[b]] = 1;
		}
		for(size_t k = 0; k < any.size(); k++) {
			if(any[k] && !m->h_marker_periodic[k]) { set_error("fvg_mesh_create: marker " + std::to_string(m->h_markers[k]) + " has periodic and non-periodic faces"); return FVG_ERR_INVALID; }
			if(!any[k]) m->h_marker_periodic[k] = 0;
		}
	}

	// ---- global locality order, then the rank's own cells and ghosts
	std::vector<int> gorder;
	if(reorder == FVG_REORDER_HILBERT) hilbert_order(n, rc.data(), gorder);
	else if(reorder == FVG_REORDER_RCM) rcm_order(n, mw, hm->esuel, hm->nnode, gorder);
	else if(reorder == FVG_REORDER_NONE) { gorder.resize(n); std::iota(gorder.begin(), gorder.end(), 0); }
	else { set_error("fvg_mesh_create: unknown reorder option"); return FVG_ERR_INVALID; }
	m->reorder = reorder;

	std::vector<int> rankidx((size_t)n);          // position of every cell inside its owner's own-cell sequence
	{
		std::vector<int> cnt(nranks, 0);
		for(int k = 0; k < n; k++) { const int o = gorder[k]; rankidx[o] = cnt[rank_of(o)]++; }
	}
	std::vector<int> &d2g = m->h_new2old, &g2d = m->h_old2new;
	g2d.assign(n, -1);
	for(int k = 0; k < n; k++) { const int o = gorder[k]; if(rank_of(o) == rank) { g2d[o] = (int)d2g.size(); d2g.push_back(o); } }
	const int nown = (int)d2g.size();
	if(nown == 0) { set_error("fvg_mesh_create: this rank owns no cells"); return FVG_ERR_INVALID; }
	m->send_counts.assign(nranks, 0); m->recv_counts.assign(nranks, 0);
	// Periodic boundaries (hm->bpartner, UMesh::compute_periodic_map): a boundary face f with partner fp becomes an
	// interior face whose right cell is a GHOST copy of the partner's cell b, displaced by the period (centre rc_b +
	// mid_f - mid_fp). Such ghosts are rows like the ghosts of a partition - filled by the same exchange, possibly from
	// this very rank - so everything downstream (tiles, halos, streams, kernels) sees an ordinary interior face.
	// Each of the two partner faces is evaluated once, by the tile of its own left cell.
	const int *const bpart = hm->bpartner;
	auto face_mid = [&](int f, int d) { return 0.5*(hm->coords[2*(size_t)hm->intfac[4*(size_t)f+2]+d] + hm->coords[2*(size_t)hm->intfac[4*(size_t)f+3]+d]); };
	if(bpart)
		for(int b = 0; b < nb; b++) {
			const int q = bpart[b];
			if(q < 0) continue;
			if(q >= nb || q == b || bpart[q] != b || m->h_btag[q] != m->h_btag[b]) { set_error("fvg_mesh_create: inconsistent periodic pairing of the boundary faces"); return FVG_ERR_INVALID; }
		}
	auto partner_of = [&](int f) { return (bpart && f >= 0 && f < nb) ? bpart[f] : -1; };
	std::vector<int> pg_dev((size_t)std::max(nb, 1), -1);      // boundary face -> device index of its periodic ghost cell
	std::vector<double2> gshift;                                // displacement of every ghost cell's centre (zero for partition ghosts)
	m->h_periodic_faces = 0;
	{
		struct GhostRec { long long key; int pface; int cell; };      // pface: the partner's boundary face (owned by `cell`), -1 for a partition ghost
		struct SendRec { int peer, idx, pface; };
		std::vector<GhostRec> ghosts;
		std::vector<SendRec> sends;
		for(int i = 0; i < nown; i++) {
			const int o = d2g[i];
			for(int j = 0; j < hm->nnode[o]; j++) {
				const int e = hm->esuel[(size_t)o*mw+j];
				if(e >= n) {
					const int f = e - n, fp = partner_of(f);
					if(fp < 0) continue;
					const int b = hm->intfac[4*(size_t)fp];
					ghosts.push_back({((long long)rank_of(b) << 32) + rankidx[b], fp, b});
					sends.push_back({rank_of(b), i, f});      // the cell behind fp sees this cell through its face fp, whose partner is f
					m->h_periodic_faces++;
					continue;
				}
				if(e < 0 || rank_of(e) == rank) continue;
				ghosts.push_back({((long long)rank_of(e) << 32) + rankidx[e], -1, e});
				sends.push_back({rank_of(e), i, -1});
			}
		}
		auto gless = [](const GhostRec &a, const GhostRec &b) { return a.key < b.key || (a.key == b.key && a.pface < b.pface); };
		std::sort(ghosts.begin(), ghosts.end(), gless);
		ghosts.erase(std::unique(ghosts.begin(), ghosts.end(), [](const GhostRec &a, const GhostRec &b) { return a.key == b.key && a.pface == b.pface; }), ghosts.end());
		for(const GhostRec &gst : ghosts) {
			const int dev = (int)d2g.size();
			d2g.push_back(gst.cell);
			m->recv_counts[rank_of(gst.cell)]++;
			if(gst.pface < 0) { g2d[gst.cell] = dev; gshift.push_back(make_double2(0.0, 0.0)); }
			else {
				const int f = bpart[gst.pface];      // this rank's face
				pg_dev[f] = dev;
				gshift.push_back(make_double2(face_mid(f, 0) - face_mid(gst.pface, 0), face_mid(f, 1) - face_mid(gst.pface, 1)));
			}
		}
		auto sless = [](const SendRec &a, const SendRec &b) { return a.peer != b.peer ? a.peer < b.peer : (a.idx != b.idx ? a.idx < b.idx : a.pface < b.pface); };
		std::sort(sends.begin(), sends.end(), sless);
		sends.erase(std::unique(sends.begin(), sends.end(), [](const SendRec &a, const SendRec &b) { return a.peer == b.peer && a.idx == b.idx && a.pface == b.pface; }), sends.end());
		m->h_send_idx.clear();
		for(const SendRec &sd : sends) { m->h_send_idx.push_back(sd.idx); m->send_counts[sd.peer]++; }
	}
	const int ntot = (int)d2g.size();
	m->nghost = ntot - nown;
	m->nranks = nranks; m->rank = rank;
	// subdomain meshes expose device-ordered arrays (own rows, then ghost rows): nothing to permute at the API
	m->identity_perm = true;
	if(nranks == 1) for(int i = 0; i < nown; i++) if(d2g[i] != i) { m->identity_perm = false; break; }

	// faces that touch an own cell, ascending global id
	auto is_own = [&](int dev) { return dev >= 0 && dev < nown; };
	auto faceL = [&](int f) { return g2d[hm->intfac[4*(size_t)f]]; };
	auto faceR = [&](int f) { return f < nb ? pg_dev[f] : g2d[hm->intfac[4*(size_t)f+1]]; };
	std::vector<int> rfaces;
	for(int f = 0; f < nf; f++) {
		const int L = faceL(f), R = faceR(f);
		if(is_own(L) || is_own(R)) rfaces.push_back(f);
	}
	const int nrf = (int)rfaces.size();
	// neighbour (device numbering) of own device cell i across local face j: >= 0 cell (own or ghost), -1 boundary
	auto nbr_of = [&](int i, int j) {
		const int e = hm->esuel[(size_t)d2g[i]*mw+j];
		return e < n ? g2d[e] : pg_dev[e - n];
	};
	// centre of a device cell (periodic ghosts: the partner's centre displaced by the period)
	auto centre = [&](int dev, int d) {
		const double c = rc[2*(size_t)d2g[dev]+d];
		if(dev < nown) return c;
		const double2 sh = gshift[(size_t)(dev - nown)];
		return c + (d == 0 ? sh.x : sh.y);
	};

	// ---- tiles: greedy ranges of consecutive own cells within the capacities
	std::vector<int> tcell0(1, 0), thoff(1, 0), thalo;
	std::vector<int> stamp((size_t)ntot, -1);
	long long distsum = 0;
	{
		std::vector<int> halo;
		int s = 0;
		while(s < nown) {
			int nc = std::min(TC, nown - s);
			const int t = (int)tcell0.size() - 1;
			if(ECAP < EMAX) {
				// entries of the first k cells of the run, k = 1..nc, in one sweep: a cell adds its faces but those it
				// shares with an earlier cell of the run (counted when that cell came in). The run ends at the entry cap.
				int E = 0, k = 0;
				for(; k < nc; k++) {
					const int i = s + k, nn = hm->nnode[d2g[i]];
					int add = 0;
					for(int j = 0; j < nn; j++) { const int q = nbr_of(i, j); if(!(q >= s && q < i)) add++; }
					if(E + add > ECAP && k > 0) break;
					E += add;
				}
				nc = k;
			}
			for(int attempt = 0; ; attempt++) {
				halo.clear();
				int E = 0, nbnd = 0;
				const int tag = t*64 + (attempt & 63);      // unique stamp per (tile, attempt)
				for(int i = s; i < s+nc; i++) {
					const int nn = hm->nnode[d2g[i]];
					for(int j = 0; j < nn; j++) {
						const int q = nbr_of(i, j);
						if(q < 0) { E++; nbnd++; continue; }
						if(q >= s && q < s+nc) { if(i < q) E++; continue; }
						E++;
						if(stamp[q] != tag) { stamp[q] = tag; halo.push_back(q); }
					}
				}
				// boundary ghost cells are staged like halo cells, so they share the row capacity
				if(((int)halo.size() + nbnd <= HMAX && E <= EMAX) || nc == 1) break;
				nc = std::max(1, std::min(nc - 1, ECAP < EMAX ? nc - nc/16 : (int)(nc*0.8)));
			}
			if((int)halo.size() > HMAX) { set_error("fvg_mesh_create: a single cell exceeds the halo capacity"); return FVG_ERR_INVALID; }
			std::sort(halo.begin(), halo.end());
			thalo.insert(thalo.end(), halo.begin(), halo.end());
			thoff.push_back((int)thalo.size());
			s += nc;
			tcell0.push_back(s);
		}
	}
	const int ntile = (int)tcell0.size() - 1;
	// Subdomain of a partitioned mesh: renumber the own cells so that the tiles which see a ghost cell come LATE. The
	// kernels then meet them last in their natural tile order - no tile list, no indirection on any CTA's critical path -
	// and the rows the neighbours push have had the whole interior to arrive (the same tiles are the ones that push).
	// Tiles stay what they are (the same cells, in the same relative order); halo lists are re-sorted. Ghost rows and
	// send lists are ordered by the owner's position in the GLOBAL locality order (rankidx), which no rank's tiling changes.
	if(nranks > 1 && ntot > nown) {
		std::vector<char> is_bnd((size_t)ntile, 0);
		for(int t = 0; t < ntile; t++) is_bnd[t] = thoff[t+1] > thoff[t] && thalo[thoff[t+1]-1] >= nown;
		// ... last but for about two waves of interior tiles, behind which the latency of their pushes (a system-scope
		// fence over NVLink per tile) hides instead of sitting in the kernel's tail
		std::vector<int> torder, inner;
		for(int t = 0; t < ntile; t++) if(!is_bnd[t]) inner.push_back(t);
		const size_t after = std::min<size_t>(inner.size()/3, 1024);
		torder.assign(inner.begin(), inner.end() - (long)after);
		for(int t = 0; t < ntile; t++) if(is_bnd[t]) torder.push_back(t);
		torder.insert(torder.end(), inner.end() - (long)after, inner.end());
		std::vector<int> newpos((size_t)nown), d2g_new(d2g), tcell0_new(1, 0), thoff_new(1, 0), thalo_new;
		for(int t : torder) {
			for(int i = tcell0[t]; i < tcell0[t+1]; i++) { newpos[i] = tcell0_new.back() + (i - tcell0[t]); d2g_new[(size_t)newpos[i]] = d2g[i]; }
			tcell0_new.push_back(tcell0_new.back() + tcell0[t+1] - tcell0[t]);
		}
		for(int t : torder) {
			const size_t b0 = thalo_new.size();
			for(int h = thoff[t]; h < thoff[t+1]; h++) thalo_new.push_back(thalo[h] < nown ? newpos[thalo[h]] : thalo[h]);
			std::sort(thalo_new.begin() + (long)b0, thalo_new.end());
			thoff_new.push_back((int)thalo_new.size());
		}
		d2g.swap(d2g_new); tcell0.swap(tcell0_new); thoff.swap(thoff_new); thalo.swap(thalo_new);
		for(int i = 0; i < nown; i++) g2d[d2g[i]] = i;
		for(int &c : m->h_send_idx) c = newpos[c];
	}
	std::vector<int> tile_order;
	int ntile_interior = ntile;
	if(ntot > nown) {
		std::vector<int> bnd;
		for(int t = 0; t < ntile; t++) {
			bool ghost = false;
			for(int h = thoff[t]; h < thoff[t+1] && !ghost; h++) ghost = thalo[h] >= nown;
			(ghost ? bnd : tile_order).push_back(t);
		}
		ntile_interior = (int)tile_order.size();
		tile_order.insert(tile_order.end(), bnd.begin(), bnd.end());
	}
	std::vector<int> tile_of((size_t)ntot, -1);
	for(int t = 0; t < ntile; t++) for(int i = tcell0[t]; i < tcell0[t+1]; i++) tile_of[i] = t;

	// ---- face streams. Raw per-tile entry lists in reference face order first (f, or -1-f for the second copy
	// of a face cut by a tile boundary), then per tile: kinds, edge colours, shared-memory bank residues, positions.
	std::vector<int> roff((size_t)ntile+1, 0);
	int ninterior = 0;
	for(int f : rfaces) {
		const int L = faceL(f), R = faceR(f);
		const int tL = L >= 0 ? tile_of[L] : -1, tR = R >= 0 ? tile_of[R] : -1;
		if(tL >= 0) roff[tL+1]++;
		if(tR >= 0 && tR != tL) roff[tR+1]++;
		if(R >= 0) { distsum += std::abs(L-R); ninterior++; }
	}
	for(int t = 0; t < ntile; t++) roff[t+1] += roff[t];
	const int ncopies = roff[ntile];
	m->ncut_dup = ncopies - nrf;
	m->mean_nbr_dist = ninterior ? (double)distsum/(double)ninterior : 0.0;
	std::vector<int> rent((size_t)ncopies);
	{
		std::vector<int> pos(roff.begin(), roff.end()-1);
		for(int f : rfaces) {
			const int L = faceL(f), R = faceR(f);
			const int tL = L >= 0 ? tile_of[L] : -1, tR = R >= 0 ? tile_of[R] : -1;
			if(tL >= 0) {
				rent[pos[tL]++] = f;
				if(tR >= 0 && tR != tL) rent[pos[tR]++] = -1-f;
			} else rent[pos[tR]++] = f;                  // left cell is a ghost: the right cell's tile holds the only copy
		}
	}
	// local face slot of global face f in own device cell i
	auto slot_of = [&](int i, int f) {
		const int o = d2g[i];
		for(int j = 0; j < hm->nnode[o]; j++) if(hm->elemface[(size_t)o*mw+j] == f) return j;
		return 0;
	};

	// Greedy edge colouring per tile (kept as metadata: fvg_mesh_stream exposes it and a test checks that no two
	// entries of one colour touch the same tile cell). The kernels gather per cell and do not need the colours;
	// the stream is ordered by KIND: faces with both cells in the tile, then faces cut by the tile boundary (one
	// side is a halo cell), then physical-boundary faces, then padding.
	//
	// Inside the first two kinds the POSITION of an entry is chosen for the shared-memory banks. The kernels keep
	// per-entry data in arrays of 16-byte elements (midpoints, the two halves of each face state / flux, spectral
	// radii), so entry e lives in bank group e mod 8. Those arrays are written and read per CELL (phases A and C of
	// the face kernel, the limiter of the cell kernel): a 128-bit shared-memory access is served per quarter warp,
	// i.e. the 8 consecutive cells 8q .. 8q+7 access the entries of their local face j together. The entries of such
	// a (q, j) group are therefore given distinct residues mod 8 - an edge colouring with 8 "bank colours" of the
	// graph whose vertices are the groups - which makes these scatters and gathers conflict-free; entries are then
	// placed at 8*i + residue, the unused positions of a residue class are padding entries.
	const bool bank_colour = !(getenv("FVG_BANK_COLOUR") && atoi(getenv("FVG_BANK_COLOUR")) == 0);
	std::vector<int> fsoff((size_t)ntile+1, 0);
	std::vector<int4> tbnd((size_t)ntile);      // tile-local index of the first cut entry, of the first boundary entry, number of boundary entries, padding
	std::vector<int> epos((size_t)ncopies), ecol((size_t)ncopies, MAXCOL-1);   // tile-local position / edge colour of every raw entry
	m->max_colours = 0;
	long long ngroups = 0, nconfl = 0;
	int emax_seen = 0;
	{
		const int NG = (TC/8 + 1)*4;
		std::vector<unsigned char> used(TC);
		std::vector<unsigned char> gcnt((size_t)NG*8);
		std::vector<int> at((size_t)NG*8);           // entry holding residue c at group g (-1: free)
		std::vector<int> kind, res, eg, path;
		for(int t = 0; t < ntile; t++) {
			const int c0 = tcell0[t], r0 = roff[t], nr = roff[t+1] - r0;
			std::fill(used.begin(), used.end(), 0);
			std::fill(at.begin(), at.end(), -1);
			kind.assign(nr, 0); res.assign(nr, -1); eg.assign(2*(size_t)nr, -1);
			int cnt[3] = {0, 0, 0};
			int rcnt[2][8] = {{0,0,0,0,0,0,0,0},{0,0,0,0,0,0,0,0}};
			int seq[2] = {0, 0};
			auto setres = [&](int k, int c) {
				res[k] = c; rcnt[kind[k]][c]++;
				for(int s2 = 0; s2 < 2; s2++) if(eg[2*k+s2] >= 0) at[(size_t)eg[2*k+s2]*8+c] = k;
			};
			for(int pass = 0; pass < 2; pass++)          // entries with two groups in the tile first: they are the constrained ones
			for(int k = 0; k < nr; k++) {
				const int sf = rent[r0+k];
				const int f = sf >= 0 ? sf : -1-sf;
				const int L = faceL(f), R = faceR(f);
				const bool inL = tile_of[L] == t, inR = R >= 0 && tile_of[R] == t;
				const int kd = R < 0 ? 2 : ((inL && inR) ? 0 : 1);
				if(pass == 0) {
					kind[k] = kd; cnt[kd]++;
					unsigned mask = 0;
					if(inL) mask |= used[L-c0];
					if(inR) mask |= used[R-c0];
					int c = 0;
					while(mask & (1u << c)) c++;
					if(c >= MAXCOL) { set_error("fvg_mesh_create: edge colouring needs more than 8 colours"); return FVG_ERR_INVALID; }
					if(inL) used[L-c0] |= (unsigned char)(1u << c);
					if(inR) used[R-c0] |= (unsigned char)(1u << c);
					ecol[r0+k] = c;
					m->max_colours = std::max(m->max_colours, c+1);
				}
				if(kd == 2 || kd != pass) continue;
				if(!bank_colour) { res[k] = seq[kd]++ & 7; rcnt[kd][res[k]]++; continue; }
				const int gL = inL ? ((L-c0) >> 3)*4 + slot_of(L, f) : -1;
				const int gR = inR ? ((R-c0) >> 3)*4 + slot_of(R, f) : -1;
				eg[2*k] = gL; eg[2*k+1] = gR;
				unsigned freeL = 0xFF, freeR = 0xFF;
				for(int c = 0; c < 8; c++) {
					if(gL >= 0 && at[(size_t)gL*8+c] >= 0) freeL &= ~(1u << c);
					if(gR >= 0 && at[(size_t)gR*8+c] >= 0) freeR &= ~(1u << c);
				}
				int best = -1;
				for(int c = 0; c < 8; c++)
					if((freeL & freeR & (1u << c)) && (best < 0 || rcnt[kd][c] < rcnt[kd][best])) best = c;
				if(best < 0 && gL >= 0 && gR >= 0) {
					// no common free residue: a is free at the left group, b at the right one; flip a <-> b along the alternating
					// path that starts at the right group with a, unless it ends in the left group (then a would be taken there)
					// (dir 0), or the mirror image: the path that starts at the left group with b, unless it ends in the right one
					for(int dir = 0; dir < 2 && best < 0; dir++)
					for(int a = 0; a < 8 && best < 0; a++) {
						if(!(freeL & (1u << a))) continue;
						for(int b = 0; b < 8 && best < 0; b++) {
							if(!(freeR & (1u << b))) continue;
							path.clear();
							const int from = dir == 0 ? gR : gL, avoid = dir == 0 ? gL : gR;
							int cur = from, col = dir == 0 ? a : b, prev = -1;
							bool ok = true;
							for(int step = 0; step < 200; step++) {
								const int e = at[(size_t)cur*8+col];
								if(e < 0) break;
								if(e == prev || res[e] != col || std::find(path.begin(), path.end(), e) != path.end()) { ok = false; break; }
								path.push_back(e);
								const int other = eg[2*e] == cur ? eg[2*e+1] : eg[2*e];
								prev = e;
								if(other < 0) break;               // an entry with one group ends the path
								cur = other; col = col == a ? b : a;
								if(cur == avoid || cur == from) { ok = false; break; }
								if(step == 199) ok = false;
							}
							if(!ok) continue;
							for(int e : path) {
								const int oc = res[e];
								rcnt[kind[e]][oc]--;
								for(int s2 = 0; s2 < 2; s2++) if(eg[2*e+s2] >= 0 && at[(size_t)eg[2*e+s2]*8+oc] == e) at[(size_t)eg[2*e+s2]*8+oc] = -1;
							}
							for(int e : path) setres(e, res[e] == a ? b : a);
							best = dir == 0 ? a : b;
						}
					}
				}
				if(best < 0) {        // give up: least-loaded residue that is free in at least one of the groups
					for(int c = 0; c < 8; c++)
						if(((freeL | freeR) & (1u << c)) && (best < 0 || rcnt[kd][c] < rcnt[kd][best])) best = c;
					if(best < 0) best = 0;
				}
				setres(k, best);
			}
			// even out the residue classes of each kind (their longest class sets the padding): move entries from the
			// fullest class to emptier ones where the target residue is still free in the entry's groups
			for(int q = 0; q < 2 && bank_colour; q++)
				for(int iter = 0; iter < 64; iter++) {
					int cmax = 0;
					for(int c = 1; c < 8; c++) if(rcnt[q][c] > rcnt[q][cmax]) cmax = c;
					bool moved = false;
					for(int k = 0; k < nr && !moved; k++) {
						if(kind[k] != q || res[k] != cmax) continue;
						int tgt = -1;
						for(int c = 0; c < 8; c++) {
							if(rcnt[q][c] + 1 >= rcnt[q][cmax] || (tgt >= 0 && rcnt[q][c] >= rcnt[q][tgt])) continue;
							bool fr = true;
							for(int s2 = 0; s2 < 2; s2++) if(eg[2*k+s2] >= 0 && at[(size_t)eg[2*k+s2]*8+c] >= 0) fr = false;
							if(fr) tgt = c;
						}
						if(tgt < 0) continue;
						rcnt[q][cmax]--;
						for(int s2 = 0; s2 < 2; s2++) if(eg[2*k+s2] >= 0 && at[(size_t)eg[2*k+s2]*8+cmax] == k) at[(size_t)eg[2*k+s2]*8+cmax] = -1;
						setres(k, tgt);
						moved = true;
					}
					if(!moved) break;
				}
			// positions: kind 0 at 8*i + residue from 0, kind 1 likewise after it, boundary entries contiguous after both
			int seg[2], nextb = 0;
			auto place = [&]() {
				for(int q = 0; q < 2; q++) { int mx = 0; for(int c = 0; c < 8; c++) mx = std::max(mx, rcnt[q][c]); seg[q] = 8*mx; }
				int next[2][8];
				for(int q = 0; q < 2; q++) for(int c = 0; c < 8; c++) next[q][c] = (q == 0 ? 0 : seg[0]) + c;
				nextb = seg[0] + seg[1];
				for(int k = 0; k < nr; k++) {
					if(kind[k] == 2) epos[r0+k] = nextb++;
					else { epos[r0+k] = next[kind[k]][res[k]]; next[kind[k]][res[k]] += 8; }
				}
				return (nextb + 7)/8*8;
			};
			int len = place();
			if(len > LCAP) {
				// the padding would push the kernels' per-entry staging past the shared-memory budget of two CTAs per SM:
				// this tile keeps its entries densely packed instead (bank conflicts in this tile only)
				int sq[2] = {0, 0};
				for(int q = 0; q < 2; q++) for(int c = 0; c < 8; c++) rcnt[q][c] = 0;
				for(int k = 0; k < nr; k++) if(kind[k] < 2) { res[k] = sq[kind[k]]++ & 7; rcnt[kind[k]][res[k]]++; }
				len = place();
			}
			std::fill(gcnt.begin(), gcnt.end(), 0);
			for(int k = 0; k < nr; k++)          // an entry whose two cells share a group (same address: a broadcast) counts once
				for(int s2 = 0; s2 < 2; s2++) if(eg[2*k+s2] >= 0 && !(s2 == 1 && eg[2*k] == eg[2*k+1])) gcnt[(size_t)eg[2*k+s2]*8+res[k]]++;
			for(size_t g = 0; g < gcnt.size()/8; g++) {
				int members = 0, distinct = 0;
				for(int c = 0; c < 8; c++) { members += gcnt[g*8+c]; distinct += gcnt[g*8+c] > 0; }
				if(members > 0) { ngroups++; if(distinct < members) nconfl++; }
			}
			// w: padding entries of the segment; bit 16: the tile's halo contains a ghost cell of another rank (the halo
			// list is ascending, ghosts are numbered after the own cells)
			const bool ghost_tile = thoff[t+1] > thoff[t] && thalo[thoff[t+1]-1] >= nown;
			tbnd[t] = make_int4(seg[0], seg[0] + seg[1], cnt[2], (len - nr) | (ghost_tile ? 0x10000 : 0));
			fsoff[t+1] = fsoff[t] + len;
			emax_seen = std::max(emax_seen, len);
		}
	}
	m->bank_groups = ngroups; m->bank_conflict_groups = nconfl;
	const int ns = fsoff[ntile];
	const int EMAX_LAYOUT = std::max(32, (emax_seen + 31)/32*32);      // capacity of the kernels' per-entry staging arrays
	const int PAD = INT_MIN;
	std::vector<int> sface((size_t)ns, PAD);    // global face of each entry, -1-f for the second copy, PAD for padding
	std::vector<int> scolour((size_t)ns, MAXCOL-1);
	// ford: the positions of a tile's real (non-padding) entries in ascending order; the flux phase walks this list so
	// that its rounds are full whatever the padding
	std::vector<unsigned short> ford((size_t)ns, 0);
	for(int t = 0; t < ntile; t++) {
		for(int k = roff[t]; k < roff[t+1]; k++) { sface[(size_t)fsoff[t] + epos[k]] = rent[k]; scolour[(size_t)fsoff[t] + epos[k]] = ecol[k]; }
		int i = 0;
		for(int e = fsoff[t]; e < fsoff[t+1]; e++) if(sface[e] != PAD) ford[(size_t)fsoff[t] + i++] = (unsigned short)(e - fsoff[t]);
	}

	// ---- per-entry arrays with tile-local cell indices
	std::vector<unsigned> &fLR = m->h_fLR;
	fLR.assign((size_t)ns, LR_PAD);
	std::vector<double2> fn((size_t)ns, make_double2(1.0, 0.0)), fgr((size_t)ns, make_double2(0.0, 0.0));
	std::vector<double> flen((size_t)ns, 0.0);
	std::vector<double2> fgw((size_t)ns, make_double2(0.5, 0.5)), fgln((size_t)ns, make_double2(0.0, 0.0));
	std::vector<int> own_entry((size_t)nf, -1), dup_entry((size_t)nf, -1);
	m->h_fref.assign(ns, PAD); m->h_fcolour = scolour; m->h_ftile.resize(ns);
	auto local_of = [&](int t, int g) -> unsigned {
		if(tile_of[g] == t) return (unsigned)(g - tcell0[t]);
		const int *hb = &thalo[thoff[t]], *he = &thalo[thoff[t+1]];
		const int *it = std::lower_bound(hb, he, g);
		return (unsigned)((tcell0[t+1]-tcell0[t]) + (int)(it - hb));
	};
	for(int t = 0; t < ntile; t++)
		for(int e = fsoff[t]; e < fsoff[t+1]; e++) {
			m->h_ftile[e] = t;
			if(sface[e] == PAD) continue;
			const int f = sface[e] >= 0 ? sface[e] : -1-sface[e];
			const int L = faceL(f), R = faceR(f);
			const unsigned lL = local_of(t, L);
			const unsigned lR = R >= 0 ? local_of(t, R) : (LR_BND | (unsigned)bslot[f]);
			fLR[e] = lL | (lR << 16);
			fn[e] = make_double2(hm->facemetric[3*(size_t)f], hm->facemetric[3*(size_t)f+1]);
			flen[e] = hm->facemetric[3*(size_t)f+2];
			const int p0 = hm->intfac[4*(size_t)f+2], p1 = hm->intfac[4*(size_t)f+3];
			double g[2];
			for(int d = 0; d < 2; d++) {
				double sum = 0;
				sum += hm->coords[2*(size_t)p0+d];
				sum += hm->coords[2*(size_t)p1+d];
				g[d] = sum/2;
			}
			fgr[e] = make_double2(g[0], g[1]);
			{
				// Green-Gauss: u_face = (u_L/d_L + u_R/d_R)/(1/d_L + 1/d_R), d = distance of the midpoint to the cell centre
				// (boundary: to the mirrored ghost centre, aspatial.cpp:98-119)
				const double lx = centre(L, 0), ly = centre(L, 1);
				const double rx = R >= 0 ? centre(R, 0) : 2.0*g[0] - lx, ry = R >= 0 ? centre(R, 1) : 2.0*g[1] - ly;
				const double di = 1.0/std::sqrt((g[0]-lx)*(g[0]-lx) + (g[1]-ly)*(g[1]-ly));
				const double dj = 1.0/std::sqrt((g[0]-rx)*(g[0]-rx) + (g[1]-ry)*(g[1]-ry));
				fgw[e] = make_double2(di/(di + dj), dj/(di + dj));
				fgln[e] = make_double2(flen[e]*fn[e].x, flen[e]*fn[e].y);
			}
			if(sface[e] >= 0) own_entry[f] = e; else dup_entry[f] = e;
			m->h_fref[e] = sface[e];
		}

	// ---- per-cell arrays in device order (centres also for the ghosts)
	std::vector<uint4> cloc((size_t)nown);
	std::vector<double2> drc((size_t)ntot);
	std::vector<double> area((size_t)ntot + 1, 1.0), clength((size_t)nown + 1, 1.0);     // one pad entry: tiles copy 16-byte granules
	for(int i = nown; i < ntot; i++) area[i] = hm->area[d2g[i]];                          // (ghost areas: the viscous spectral radius of a cut face's far side)
	for(int i = 0; i < ntot; i++) drc[i] = make_double2(centre(i, 0), centre(i, 1));
	for(int i = 0; i < nown; i++) {
		const int o = d2g[i], t = tile_of[i];
		unsigned a[4] = {NB_NONE, NB_NONE, NB_NONE, NB_NONE}, c[4] = {0,0,0,0};
		for(int j = 0; j < hm->nnode[o]; j++) {
			const int e = hm->esuel[(size_t)o*mw+j];
			const int f = hm->elemface[(size_t)o*mw+j];
			if(e < 0 || f < 0 || f >= nf) { set_error("fvg_mesh_create: inconsistent esuel/elemface"); return FVG_ERR_INVALID; }
			if(e >= n && (e-n != f || f >= nb)) { set_error("fvg_mesh_create: boundary ghost index does not match its face"); return FVG_ERR_INVALID; }
			a[j] = e < n ? local_of(t, g2d[e]) : (pg_dev[e - n] >= 0 ? local_of(t, pg_dev[e - n]) : NB_BND);
			const bool isL = hm->intfac[4*(size_t)f] == o;
			if(!isL && (f < nb || hm->intfac[4*(size_t)f+1] != o)) { set_error("fvg_mesh_create: elemface/intfac mismatch"); return FVG_ERR_INVALID; }
			// the copy of the face that lives in this cell's tile
			int entry = own_entry[f];
			if(!isL && dup_entry[f] >= 0) entry = dup_entry[f];
			if(entry < 0 || m->h_ftile[entry] != t) { set_error("fvg_mesh_create: internal error, face copy not in the cell's tile"); return FVG_ERR_INVALID; }
			c[j] = (unsigned)(entry - fsoff[t]) | (isL ? 0u : 0x8000u);
		}
		cloc[i] = make_uint4(a[0] | (a[1] << 16), a[2] | (a[3] << 16), c[0] | (c[1] << 16), c[2] | (c[3] << 16));
		area[i] = hm->area[o];
		double l2 = 0;
		const int nn = hm->nnode[o];
		for(int j = 0; j < nn; j++) {
			const int p = hm->inpoel[(size_t)o*mw+j], q = hm->inpoel[(size_t)o*mw+(j+1)%nn];
			double sum = 0;
			for(int d = 0; d < 2; d++) { const double tt = hm->coords[2*(size_t)p+d] - hm->coords[2*(size_t)q+d]; sum += tt*tt; }
			if(l2 < sum) l2 = sum;
		}
		clength[i] = std::sqrt(l2);
	}

	// ---- boundary arrays (global boundary-face order; faces of other ranks carry cell -1)
	std::vector<int> bcell((size_t)nb);
	std::vector<double2> rcbp((size_t)nb);
	m->h_bentry.assign(nb, 0);
	for(int b = 0; b < nb; b++) {
		const int o = hm->intfac[4*(size_t)b];
		bcell[b] = is_own(g2d[o]) ? g2d[o] : -1;
		if(own_entry[b] >= 0) m->h_bentry[b] = own_entry[b];
		double mid[2];
		for(int d = 0; d < 2; d++) {
			double sum = 0;
			sum += hm->coords[2*(size_t)hm->intfac[4*(size_t)b+2]+d];
			sum += hm->coords[2*(size_t)hm->intfac[4*(size_t)b+3]+d];
			mid[d] = sum/2;
		}
		rcbp[b] = make_double2(2.0*mid[0] - rc[2*(size_t)o], 2.0*mid[1] - rc[2*(size_t)o+1]);
	}

	// ---- least-squares matrices: accumulate in reference face order over the GLOBAL mesh, invert as adj/det
	std::vector<double4> V((size_t)nown);
	{
		std::vector<double> A(4*(size_t)n, 0.0);
		for(int f = 0; f < nf; f++) {
			const int ie = hm->intfac[4*(size_t)f];
			double dr[2];
			if(f < nb && partner_of(f) >= 0) {
				// periodic face: the neighbour is the partner's cell displaced by the period
				const int fp = partner_of(f), je = hm->intfac[4*(size_t)fp];
				for(int d = 0; d < 2; d++) dr[d] = rc[2*(size_t)ie+d] - (rc[2*(size_t)je+d] + face_mid(f, d) - face_mid(fp, d));
			}
			else if(f < nb) { dr[0] = rc[2*(size_t)ie] - rcbp[f].x; dr[1] = rc[2*(size_t)ie+1] - rcbp[f].y; }
			else { const int je = hm->intfac[4*(size_t)f+1]; dr[0] = rc[2*(size_t)ie] - rc[2*(size_t)je]; dr[1] = rc[2*(size_t)ie+1] - rc[2*(size_t)je+1]; }
			double w2 = 0;
			for(int d = 0; d < 2; d++) w2 += dr[d]*dr[d];
			w2 = 1.0/w2;
			for(int p = 0; p < 2; p++) for(int q = 0; q < 2; q++) {
				A[4*(size_t)ie+2*p+q] += w2*dr[p]*dr[q];
				if(f >= nb) A[4*(size_t)hm->intfac[4*(size_t)f+1]+2*p+q] += w2*dr[p]*dr[q];
			}
		}
		for(int i = 0; i < nown; i++) {
			const int o = d2g[i];
			const double a = A[4*(size_t)o], b = A[4*(size_t)o+1], c = A[4*(size_t)o+2], d = A[4*(size_t)o+3];
			const double idet = 1.0/(a*d - c*b);
			V[i] = make_double4(d*idet, -b*idet, -c*idet, a*idet);
		}
	}

	// ---- upload
	DMesh &D = m->d;
	D.ncell = nown; D.nghost = ntot - nown; D.nbface = nb; D.naface = nf; D.ntile = ntile; D.TC = TC; D.HMAX = HMAX; D.EMAX = EMAX_LAYOUT; D.nstream = ns;
	D.nsend = (int)m->h_send_idx.size();
	int rcode;
#define UP(vec, field) if((rcode = upload(m, vec, &D.field)) != 0) return rcode;
	UP(cloc, cloc) UP(drc, rc) UP(area, area) UP(V, wlsV) UP(clength, clength)
	UP(tcell0, tcell0) UP(thoff, thoff) UP(thalo, thalo)
	UP(fsoff, fsoff) UP(tbnd, tbnd) UP(fLR, fLR) UP(fn, fn) UP(flen, flen) UP(fgr, fgr) UP(fgw, fgw) UP(fgln, fgln)
	UP(ford, ford) UP(m->h_fref, fref) UP(bcell, bcell) UP(m->h_bentry, bentry) UP(bslot, bslot) UP(rcbp, rcbp)
	UP(m->h_send_idx, send_idx)
	D.ntile_interior = ntile_interior;
	if(tile_order.empty()) D.tile_order = nullptr; else { UP(tile_order, tile_order) }
	{
		// one 48-byte record per tile with everything a kernel needs to start on it (three 16-byte pieces, fetched into
		// shared memory by cp.async ahead of time): in natural tile order and, on a partitioned mesh, interior tiles first
		std::vector<int4> tdesc(3*(size_t)ntile), tdesc_ord;
		for(int t = 0; t < ntile; t++) {
			tdesc[3*(size_t)t] = make_int4(t, tcell0[t], tcell0[t+1] - tcell0[t], thoff[t]);
			tdesc[3*(size_t)t+1] = make_int4(thoff[t+1] - thoff[t], fsoff[t], fsoff[t+1] - fsoff[t], tbnd[t].w);
			tdesc[3*(size_t)t+2] = make_int4(tbnd[t].x, tbnd[t].y, tbnd[t].z, 0);
		}
		UP(tdesc, tdesc)
		D.tdesc_ord = nullptr;
		if(!tile_order.empty()) {
			tdesc_ord.resize(3*(size_t)ntile);
			for(int i = 0; i < ntile; i++) for(int q = 0; q < 3; q++) tdesc_ord[3*(size_t)i+q] = tdesc[3*(size_t)tile_order[i]+q];
			UP(tdesc_ord, tdesc_ord)
		}
	}
	D.halo_src = nullptr;
	if(m->identity_perm) { D.new2old = nullptr; D.old2new = nullptr; }
	else {
		UP(d2g, new2old) D.old2new = nullptr;
		std::vector<int> halo_src(thalo.size());
		for(size_t q = 0; q < thalo.size(); q++) halo_src[q] = d2g[thalo[q]];
		UP(halo_src, halo_src)
	}
#undef UP
	m->h_tcell0 = tcell0; m->h_thoff = thoff; m->h_thalo = thalo;
	// per-tile send lists: the rows a tile's cells contribute to the neighbours' ghost blocks, as (tile-local cell,
	// peer rank, row inside this rank's block of the peer's ghost range). A producing kernel can push these rows to the
	// peers as soon as the tile is done (the per-rank list send_idx is the same set, grouped by peer).
	m->h_tsoff.assign((size_t)ntile + 1, 0);
	m->h_tsend.clear();
	if(!m->h_send_idx.empty()) {
		std::vector<std::array<int,4>> items;     // tile, local cell, peer, row
		int k = 0;
		for(int r = 0; r < nranks; r++)
			for(int q = 0; q < m->send_counts[r]; q++, k++) {
				const int c = m->h_send_idx[k];
				items.push_back({tile_of[c], c - tcell0[tile_of[c]], r, q});
			}
		std::sort(items.begin(), items.end());
		for(const auto &it : items) { m->h_tsoff[it[0]+1]++; m->h_tsend.push_back(it[1]); m->h_tsend.push_back(it[2]); m->h_tsend.push_back(it[3]); }
		for(int t = 0; t < ntile; t++) m->h_tsoff[t+1] += m->h_tsoff[t];
	}
	m->h_rc.resize(2*(size_t)nown);
	for(int i = 0; i < nown; i++) { m->h_rc[2*(size_t)i] = drc[i].x; m->h_rc[2*(size_t)i+1] = drc[i].y; }
	return 0;
}

} // namespace fvg

using namespace fvg;

extern "C" {

const char *fvg_last_error(void) { return last_error().c_str(); }

int fvg_device_count(int *count)
{
	if(!count) { set_error("fvg_device_count: null argument"); return FVG_ERR_INVALID; }
	*count = 0;
	FVG_CUDA(cudaGetDeviceCount(count));
	if(*count <= 0) { set_error("no CUDA device visible"); return FVG_ERR_CUDA; }
	return 0;
}

int fvg_mesh_create(const fvg_host_mesh *hm, const fvg_mesh_opts *opts, fvg_mesh **out)
{
	return fvg_mesh_create_part(hm, opts, nullptr, 0, 1, out);
}

int fvg_mesh_create_part(const fvg_host_mesh *hm, const fvg_mesh_opts *opts, const int *cell_rank, int rank, int nranks,
                         fvg_mesh **out)
{
	if(!hm || !out) { set_error("fvg_mesh_create: null argument"); return FVG_ERR_INVALID; }
	*out = nullptr;
	if(opts && opts->device >= 0) FVG_CUDA(cudaSetDevice(opts->device));
	fvg_mesh *m = new fvg_mesh;
	if(opts && opts->device == -2) m->device = -2;
	else {
		const cudaError_t e = cudaGetDevice(&m->device);
		if(e != cudaSuccess) { delete m; return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__); }
	}
	int rc;
	try { rc = build(hm, opts, cell_rank, rank, nranks, m); }
	catch(std::exception &ex) { set_error(std::string("fvg_mesh_create: ") + ex.what()); rc = FVG_ERR_INVALID; }
	if(rc != 0) { fvg_mesh_destroy(m); return rc; }
	*out = m;
	return 0;
}

void fvg_mesh_destroy(fvg_mesh *m)
{
	if(!m) return;
	for(void *p : m->allocs) cudaFree(p);
	delete m;
}

int fvg_mesh_get_info(const fvg_mesh *m, fvg_mesh_info *info)
{
	if(!m || !info) { set_error("fvg_mesh_get_info: null argument"); return FVG_ERR_INVALID; }
	info->nghost = m->d.nghost; info->rank = m->rank; info->nranks = m->nranks; info->nsend = m->d.nsend;
	info->ncell = m->d.ncell; info->nbface = m->d.nbface; info->naface = m->d.naface;
	info->ntile = m->d.ntile; info->tile_cells = m->d.TC; info->nstream = m->d.nstream;
	info->ncut_dup = m->ncut_dup; info->max_colours = m->max_colours; info->reorder = m->reorder;
	info->mean_neighbour_distance = m->mean_nbr_dist;
	info->entry_capacity = m->d.EMAX; info->halo_capacity = m->d.HMAX;
	info->bank_groups = m->bank_groups; info->bank_conflict_groups = m->bank_conflict_groups;
	return 0;
}

int fvg_mesh_permutation(const fvg_mesh *m, int *cell_new2old)
{
	if(!m || !cell_new2old) { set_error("fvg_mesh_permutation: null argument"); return FVG_ERR_INVALID; }
	std::memcpy(cell_new2old, m->h_new2old.data(), sizeof(int)*m->h_new2old.size());
	return 0;
}

int fvg_mesh_halo_lists(const fvg_mesh *m, int *send_counts, int *recv_counts, int *send_idx)
{
	if(!m) { set_error("fvg_mesh_halo_lists: null argument"); return FVG_ERR_INVALID; }
	if(send_counts) std::memcpy(send_counts, m->send_counts.data(), sizeof(int)*m->send_counts.size());
	if(recv_counts) std::memcpy(recv_counts, m->recv_counts.data(), sizeof(int)*m->recv_counts.size());
	if(send_idx && !m->h_send_idx.empty()) std::memcpy(send_idx, m->h_send_idx.data(), sizeof(int)*m->h_send_idx.size());
	return 0;
}

int fvg_mesh_tile_send_lists(const fvg_mesh *m, int *tile_off, int *cell_peer_row)
{
	if(!m || !tile_off) { set_error("fvg_mesh_tile_send_lists: null argument"); return FVG_ERR_INVALID; }
	std::memcpy(tile_off, m->h_tsoff.data(), sizeof(int)*m->h_tsoff.size());
	if(cell_peer_row && !m->h_tsend.empty()) std::memcpy(cell_peer_row, m->h_tsend.data(), sizeof(int)*m->h_tsend.size());
	return 0;
}

int fvg_mesh_tile_offsets(const fvg_mesh *m, int *tile_cell0)
{
	if(!m || !tile_cell0) { set_error("fvg_mesh_tile_offsets: null argument"); return FVG_ERR_INVALID; }
	std::memcpy(tile_cell0, m->h_tcell0.data(), sizeof(int)*m->h_tcell0.size());
	return 0;
}

int fvg_mesh_stream(const fvg_mesh *m, int *entry_face, int *entry_colour, int *entry_tile)
{
	if(!m) { set_error("fvg_mesh_stream: null argument"); return FVG_ERR_INVALID; }
	for(int e = 0; e < m->d.nstream; e++) {
		if(entry_face) entry_face[e] = m->h_fref[e] == INT_MIN ? -1 : (m->h_fref[e] >= 0 ? m->h_fref[e] : -1-m->h_fref[e]);
		if(entry_colour) entry_colour[e] = m->h_fcolour[e];
		if(entry_tile) entry_tile[e] = m->h_ftile[e];
	}
	return 0;
}

} // extern "C"
