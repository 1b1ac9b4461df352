/* The gradient + limiter pass (pass A) as a template, shared by the translation units that instantiate it
 * (kernels_cell.cu: the plain single-GPU forms; kernels_cell_modes.cu: multi-GPU, caller-ordered and looping forms). */
#pragma once
#include "engine.hpp"
#include "face_kernel.cuh"   // ld4 / st4 / extrapolation helpers

namespace fvg {

enum { GM_ZERO = 0, GM_GG = 1, GM_WLS = 2, GM_GIVEN = 3 };
enum { LM_NONE = 0, LM_BJ = 1, LM_VENKAT = 2 };

/// 32-byte row from shared memory, halves at double offsets o0 and o1 (0 and 2 in either order)
__device__ __forceinline__ void lds4h(const double *p, int o0, int o1, double v[4]) {
	const double2 a = *reinterpret_cast<const double2*>(p + o0), b = *reinterpret_cast<const double2*>(p + o1);
	v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

/// Shared-memory carve-up of the cell kernel
struct CellSmem {
	int sp, src, sgr, sW, scl, sV, sclen, bar, ring, total;
	__host__ __device__ CellSmem(int TC, int HMAX, int EMAX, bool mids, bool metrics, bool wls, bool venkat) {
		const int CAPC = TC + HMAX;
		int o = 0;
		sp = o; o += CAPC*32;
		src = o; o += CAPC*16;
		sgr = o; o += mids ? EMAX*16 : 0;
		sW = o; o += metrics ? EMAX*32 : 0;
		scl = o; o += TC*16;
		sV = o; o += wls ? TC*32 : 0;
		sclen = o; o += venkat ? (TC + 2)*8 : 0;
		bar = o; o += 16;
		ring = o; o += 2*48 + 16;             // descriptor records of the current and the next tile; evaluation number + ghost-row area
		total = o;
	}
};

/** Gradient + limiter pass: persistent CTAs (three per SM) walking over tiles ti = tile0 + blockIdx.x, + gridDim.x, ...
 * The tile's cell states and centres (own cells by TMA bulk copy, halo cells by cp.async gathers) and the face
 * midpoints of its stream are staged in shared memory; conserved states are converted to primitive ONCE per staged
 * cell (the reference converts the whole field in a separate pass, flow_spatial.cpp:697-699). Then one thread per own
 * cell gathers its <= 4 neighbours from shared memory: no scatter, no atomics. The staging buffers are single
 * (three CTAs per SM overlap each other's copies); the next tile's descriptor is fetched while the current tile
 * computes, so the only exposed latency per tile is that of its own copies.
 * Multi-GPU (A.dist): the first wave's prologue pushes the state rows the neighbours need, a tile that sees ghost
 * cells waits for the neighbours' state rows and reads them from the halo window, and every tile pushes the gradient
 * rows on its send list as soon as it has stored them (dist_dev.cuh).
 * Caller-ordered state (A.src_idx): own and halo rows are gathered through the permutation and the own rows are also
 * written in device order to A.ucopy, which is what the face pass reads - no separate permutation kernel.
 * MODE (bit flags; each feature is compiled in only where it is used - measured on the 10M-cell benchmark, the plain kernel
 * takes 0.490 ms, 0.527 ms with the never-taken permutation branches in it, 0.548 ms with the multi-GPU ones as well):
 *   CM_DIST  multi-GPU pushes and waits;  CM_PERM  caller-ordered state (src_idx / halo_src / ucopy);
 *   CM_LOOP  a resident-size grid walking the tiles instead of one tile per CTA (measured slower: 0.56 ms; the
 *            hardware's CTA scheduler balances uneven tiles better than a fixed stride, and carrying a tile loop costs
 *            registers the stencil loop needs). */
enum { CM_PLAIN = 0, CM_DIST = 1, CM_PERM = 2, CM_LOOP = 4 };
template <int GRAD, int LIM, bool PRIM_IN, int MODE>
__global__ void __launch_bounds__(CELL_BLOCK, FVG_CELL_MINB)
cell_kernel(const __grid_constant__ CellArgs A)
{
	constexpr bool LOOP = (MODE & CM_LOOP) != 0, PERM = (MODE & CM_PERM) != 0;
	const DistDev *const distd = (MODE & CM_DIST) ? A.dist.d : nullptr;
	const int *const src_idx = PERM ? A.src_idx : nullptr, *const halo_src = PERM ? A.halo_src : nullptr;
	double *const ucopy = PERM ? A.ucopy : nullptr;
	extern __shared__ __align__(1024) unsigned char smraw[];
	const DMesh &M = A.m;
	constexpr bool MIDS = LIM != LM_NONE || GRAD == GM_GG;
	constexpr bool METRICS = GRAD == GM_GG;
	const CellSmem S(M.TC, M.HMAX, M.EMAX, MIDS, METRICS, GRAD == GM_WLS, LIM == LM_VENKAT);
	double *const sp = reinterpret_cast<double*>(smraw + S.sp);
	double2 *const src = reinterpret_cast<double2*>(smraw + S.src);
	double2 *const sgr = reinterpret_cast<double2*>(smraw + S.sgr);
	const double2 *const sW = reinterpret_cast<const double2*>(smraw + S.sW);     // two 16-byte planes: weights, len*normal
	const uint4 *const scl = reinterpret_cast<const uint4*>(smraw + S.scl);
	const double4 *const sV = reinterpret_cast<const double4*>(smraw + S.sV);
	const double *const sclen = reinterpret_cast<const double*>(smraw + S.sclen);
	uint64_t *const bar = reinterpret_cast<uint64_t*>(smraw + S.bar);
	constexpr bool NEED_NBRS = GRAD == GM_GG || GRAD == GM_WLS || LIM != LM_NONE;

	const int tid = threadIdx.x, G = (int)gridDim.x;
	const int tend = A.tile1;
	int ti = A.tile0 + (int)blockIdx.x;
	// the tile's 48-byte descriptor record (DMesh::tdesc): the first one is loaded directly, those of a CTA's further
	// tiles (grids smaller than the tile count) travel into a two-slot shared-memory ring one tile ahead
	int4 *const ring = reinterpret_cast<int4*>(smraw + S.ring);
	if(tid == 0) mbar_init(bar, 1);
	pdl_launch_dependents();
	// (programmatic dependent launch) the tile descriptors are mesh data; the state read below may be the previous
	// kernel's output and the gradient rows written at the end are still being read by the previous face pass until it
	// completes. The one-tile-per-CTA forms are launched WITHOUT the attribute (plain stream order) and skip the wait:
	// griddepcontrol.wait costs every CTA a fraction of a microsecond, which 39 000 short CTAs cannot hide.
	if(A.pdl) pdl_wait();
	int4 r0 = make_int4(0, 0, 0, 0), r1 = r0, r2 = r0;
	if(ti < tend) {
		if(LOOP || A.ordered) { r0 = A.tdesc[3*(size_t)ti]; r1 = A.tdesc[3*(size_t)ti + 1]; r2 = A.tdesc[3*(size_t)ti + 2]; }
		else {
			// natural tile order, one tile per CTA: the compact per-tile arrays, whose cache lines neighbouring CTAs share
			// (measured 0.015 ms faster per launch than one 48-byte record per CTA)
			const int c0_ = M.tcell0[ti], c1_ = M.tcell0[ti+1], h0_ = M.thoff[ti], h1_ = M.thoff[ti+1], e0_ = M.fsoff[ti], e1_ = M.fsoff[ti+1];
			const int4 tb_ = M.tbnd[ti];
			r0 = make_int4(ti, c0_, c1_ - c0_, h0_); r1 = make_int4(h1_ - h0_, e0_, e1_ - e0_, tb_.w); r2 = make_int4(tb_.x, tb_.y, tb_.z, 0);
		}
	}
	// fused multi-GPU evaluation: the evaluation number, the window area of this evaluation's state rows, the prologue
	// (both kept in shared memory: they are needed once per partition-boundary tile only)
	// The evaluation number lives in device memory (DistCtl::k). Only the few CTAs that need it read it - those of the
	// first wave that push the state rows, and later the tiles that see a ghost cell (they wait and push): a load on
	// every CTA's critical path costs 0.13 ms per launch on the 39 000 one-tile CTAs of the benchmark.
	const double **const sghost = reinterpret_cast<const double**>(smraw + S.ring + 104);
	if(distd) {
		if(A.dist.first && (int)blockIdx.x < DIST_PROLOGUE_CTAS)
			dist_push_state_prologue(distd, A.dist.ctl->k, A.u, A.dist.force_push, src_idx);
	}
	else if(tid == 0) *sghost = A.gs_u.rows;
	__syncthreads();

	for(int it = 0; ti < tend; it++, ti += G) {
		const int t = r0.x, c0 = r0.y, nc = r0.z, h0 = r0.w, nh = r1.x, e0 = r1.y, ne = r1.z;
		const int4 tbq = make_int4(r2.x, r2.y, r2.z, r1.w);
		// the next tile's record travels while this tile is staged (it joins this tile's cp.async group)
		const bool have_next = LOOP && ti + G < tend;
		if(have_next) fetch_tile_desc(ring + 3*((it + 1) & 1), A.tdesc, ti + G);
		if(tid == 0) {
			if(it > 0) fence_proxy_async();       // the buffers were read through the generic proxy by the previous tile
			unsigned bytes = (unsigned)nc*((src_idx ? 0u : 32u) + 16u + 16u + (GRAD == GM_WLS ? 32u : 0u))
			               + (LIM == LM_VENKAT ? (unsigned)((nc + (c0 & 1) + 1) & ~1)*8u : 0u);
			if(MIDS) bytes += (unsigned)ne*16u;
			if(METRICS) bytes += (unsigned)ne*32u;
			mbar_expect_tx(bar, bytes);
			if(!src_idx) bulk_g2s(sp, A.u + 4*(size_t)c0, (unsigned)nc*32u, bar);
			bulk_g2s(src, M.rc + c0, (unsigned)nc*16u, bar);
			// the cells' stencil metadata rides along (consumed from shared memory: no registers held across the staging)
			bulk_g2s(smraw + S.scl, M.cloc + c0, (unsigned)nc*16u, bar);
			if(GRAD == GM_WLS) bulk_g2s(smraw + S.sV, M.wlsV + c0, (unsigned)nc*32u, bar);
			// 8-byte rows: copy whole 16-byte granules starting at the even cell below c0 (the array is padded by one entry)
			if(LIM == LM_VENKAT) bulk_g2s(smraw + S.sclen, M.clength + (c0 & ~1), (unsigned)((nc + (c0 & 1) + 1) & ~1)*8u, bar);
			if(MIDS) bulk_g2s(sgr, M.fgr + e0, (unsigned)ne*16u, bar);
			if(METRICS) { bulk_g2s(smraw + S.sW, M.fgw + e0, (unsigned)ne*16u, bar); bulk_g2s(smraw + S.sW + M.EMAX*16, M.fgln + e0, (unsigned)ne*16u, bar); }
		}
		if(tid == 32 && A.prefetch_distance > 0 && !A.ordered && t + A.prefetch_distance < M.ntile) {
			const int tp = t + A.prefetch_distance;
			const int pc0 = M.tcell0[tp], pnc = M.tcell0[tp+1] - pc0;
			const int pe0 = M.fsoff[tp], pne = M.fsoff[tp+1] - pe0;
			if(!src_idx) bulk_prefetch_l2(A.u + 4*(size_t)pc0, (unsigned)pnc*32u);
			bulk_prefetch_l2(M.rc + pc0, (unsigned)pnc*16u);
			bulk_prefetch_l2(M.cloc + pc0, (unsigned)pnc*16u);
			{ const int ph0 = M.thoff[tp] & ~3, ph1 = (M.thoff[tp+1] + 3) & ~3; if(ph1 > ph0) bulk_prefetch_l2(M.thalo + ph0, (unsigned)(ph1 - ph0)*4u); }
			if(GRAD == GM_WLS) bulk_prefetch_l2(M.wlsV + pc0, (unsigned)pnc*32u);
			if(MIDS) bulk_prefetch_l2(M.fgr + pe0, (unsigned)pne*16u);
			if(METRICS) { bulk_prefetch_l2(M.fgw + pe0, (unsigned)pne*16u); bulk_prefetch_l2(M.fgln + pe0, (unsigned)pne*16u); }
		}
		// caller-ordered state: the own rows are gathered through the permutation (16-byte pieces)
		if(src_idx) {
			for(int k = tid; k < nc*2; k += CELL_BLOCK) {
				const int row = k >> 1, piece = k & 1;
				cp_async16(sp + 4*row + 2*piece, A.u + 4*(size_t)src_idx[c0 + row] + 2*piece);
			}
		}
		// in-kernel receive of the state's ghost rows: a tile that sees ghost cells waits for the neighbours' rows (its
		// other copies are already in flight), then gathers those rows from the halo window
		const bool ghost_tile = (tbq.w >> 16) != 0;
		const bool ghost_win = ghost_tile && (distd ? (A.dist.wait & (1u << X_U)) != 0 : *sghost != nullptr);
		if(NEED_NBRS) {
			for(int k = tid; k < nh*3; k += CELL_BLOCK) {
				const int h = k/3, piece = k - 3*h;
				const size_t g = (size_t)M.thalo[h0 + h];
				const int row = nc + h;
				if(piece == 2) cp_async16(src + row, M.rc + g);
				else if(!(ghost_win && g >= (size_t)M.ncell)) {
					const size_t gs = halo_src ? (size_t)halo_src[h0 + h] : g;
					cp_async16(sp + 4*row + 2*piece, A.u + 4*gs + 2*piece);
				}
			}
			if(ghost_win) {
				const double *ghost_rows_u;
				if(distd) {
					const unsigned long long dk = A.dist.ctl->k;
					dist_wait(distd, 1u << X_U, dk);
					ghost_rows_u = A.dist.ghost[X_U][dk & 1ull];
				}
				else { ghost_wait(A.gs_u, A.gs_u.seq); ghost_rows_u = *sghost; }
				for(int k = tid; k < nh*2; k += CELL_BLOCK) {
					const int h = k >> 1, piece = k & 1;
					const size_t g = (size_t)M.thalo[h0 + h];
					if(g >= (size_t)M.ncell) cp_async16(sp + 4*(nc + h) + 2*piece, ghost_rows_u + 4*(g - (size_t)M.ncell) + 2*piece);
				}
			}
		}
		cp_async_commit();
		const int2 tb = make_int2(tbq.y, tbq.z);   // boundary entries of the tile: first (tile-local) and count
		const int grow0 = nc + nh;                 // their ghost cells are staged as rows grow0 .. grow0 + tb.y - 1
		cp_async_wait_all();
		mbar_wait(bar, (unsigned)(it & 1));
		__syncthreads();
		{
			// one pass over the staged rows: cell and halo states become primitive in place; the ghost cell of every
			// physical-boundary face gets its own row (state from the boundary condition applied to the conserved
			// state of the adjacent cell, flow_spatial.cpp:659-695; centre mirrored about the face midpoint,
			// aspatial.cpp:98-119), so that the stencil loop below needs no boundary branch at all
			const int nrows = NEED_NBRS ? grow0 + tb.y : nc;
			for(int k = tid; k < nrows; k += CELL_BLOCK) {
				if(k < grow0) {
					if(PRIM_IN) continue;
					// 32-byte rows, one per thread: threads 4..7 of every 8 take the halves in the opposite order, which
					// spreads a quarter warp over all 32 banks (plain row-order access is a 2-way conflict)
					const int hb = (tid >> 2) & 1;
					const double2 h0_ = *reinterpret_cast<const double2*>(sp + 4*k + 2*hb);
					const double2 h1_ = *reinterpret_cast<const double2*>(sp + 4*k + 2*(1 - hb));
					const double uc[4] = {hb ? h1_.x : h0_.x, hb ? h1_.y : h0_.y, hb ? h0_.x : h1_.x, hb ? h0_.y : h1_.y};
					if(ucopy && k < nc) st4(ucopy + 4*(size_t)(c0 + k), uc);
					double up[4];
					cons2prim(A.gas, uc, up);
					*reinterpret_cast<double2*>(sp + 4*k + 2*hb) = hb ? make_double2(up[2], up[3]) : make_double2(up[0], up[1]);
					*reinterpret_cast<double2*>(sp + 4*k + 2*(1 - hb)) = hb ? make_double2(up[0], up[1]) : make_double2(up[2], up[3]);
				} else {
					const int ge = e0 + tb.x + (k - grow0);
					const unsigned LR = M.fLR[ge];
					const int L = (int)(LR & 0xFFFFu);
					double pj[4];
					if(PRIM_IN) ld4(A.ug + 4*(size_t)M.fref[ge], pj);
					else {
						const double2 n = M.fn[ge];
						double ui[4], gs[4];
						ld4(A.u + 4*(size_t)(src_idx ? src_idx[c0 + L] : c0 + L), ui);
						ghost_state(A.gas, A.gas.bc[(LR >> 16) & 15u], ui, n.x, n.y, gs);
						cons2prim(A.gas, gs, pj);
					}
					const double2 mid = M.fgr[ge];
					const double2 rl = M.rc[c0 + L];
					*reinterpret_cast<double2*>(sp + 4*k) = make_double2(pj[0], pj[1]);
					*reinterpret_cast<double2*>(sp + 4*k + 2) = make_double2(pj[2], pj[3]);
					src[k] = make_double2(2.0*mid.x - rl.x, 2.0*mid.y - rl.y);
				}
			}
			__syncthreads();
		}

		for(int k = tid; k < nc; k += CELL_BLOCK) {
			const int i = c0 + k;
			const uint4 cl = scl[k];
			unsigned nb[4] = {cl.x & 0xFFFFu, cl.x >> 16, cl.y & 0xFFFFu, cl.y >> 16};
			const unsigned cf[4] = {cl.z & 0xFFFFu, cl.z >> 16, cl.w & 0xFFFFu, cl.w >> 16};
			const bool quad = nb[3] != NB_NONE;       // only the fourth slot can be empty (triangles)
			const double2 rci = src[k];
			// Threads 4..7 of every 8 work on the variables in the order (2,3,0,1): they read the second half of every
			// 32-byte state row first. Nothing below depends on which variable is which (gradient, limiter and their
			// inputs are per variable), so the permutation costs nothing and is undone by the store addresses; it makes
			// the own-row access conflict-free and spreads the neighbour gathers over all 8 bank groups instead of 4.
			const int hb = (tid >> 2) & 1;
			const int o0 = 2*hb, o1 = 2 - 2*hb;
			double pi[4];
			lds4h(sp + 4*k, o0, o1, pi);

			double acc[8] = {0,0,0,0,0,0,0,0};   // GG: gradient sums; WLS: right-hand side. Index d + 2*v
			double dmin[4] = {0,0,0,0}, dmax[4] = {0,0,0,0};
			const double ainv = GRAD == GM_GG ? frcp(M.area[i]) : 0.0;

			if(NEED_NBRS) {
				#pragma unroll
				for(int j = 0; j < 4; j++) {
					if(j == 3 && !quad) break;
					const int le = (int)(cf[j] & 0x7FFFu);
					const bool bndj = nb[j] == NB_BND;
					const unsigned nj = bndj ? (unsigned)(grow0 + le - tb.x) : nb[j];
					double pj[4];
					lds4h(sp + 4*nj, o0, o1, pj);
					const double2 rj = src[nj];
					if(GRAD == GM_WLS) {
						const double dx = rci.x - rj.x, dy = rci.y - rj.y;
						const double w = frcp(dx*dx + dy*dy);
						const double wx = w*dx, wy = w*dy;
						#pragma unroll
						for(int v = 0; v < 4; v++) {
							const double du = pi[v] - pj[v];
							acc[2*v] += wx*du;
							acc[2*v+1] += wy*du;
						}
					}
					if(GRAD == GM_GG) {
						// face value = own state * own weight + neighbour state * its weight; the face's len*normal points from the
						// entry's left to its right cell
						const bool isR = (cf[j] & 0x8000u) != 0;
						const double2 w = sW[le], ln = sW[M.EMAX + le];
						const double wi = isR ? w.y : w.x, wj = isR ? w.x : w.y;
						const double sx = isR ? -ln.x : ln.x, sy = isR ? -ln.y : ln.y;
						#pragma unroll
						for(int v = 0; v < 4; v++) {
							const double ut = pi[v]*wi + pj[v]*wj;
							acc[2*v] += ut*sx;
							acc[2*v+1] += ut*sy;
						}
					}
					if(LIM != LM_NONE && !(bndj && A.bnd_policy != 0)) {
						#pragma unroll
						for(int v = 0; v < 4; v++) {
							// plain selects: fmax/fmin on doubles cost three times as much for their NaN rules
							const double du = pj[v] - pi[v];
							dmax[v] = du > dmax[v] ? du : dmax[v];
							dmin[v] = du < dmin[v] ? du : dmin[v];
						}
					}
				}
			}

			double g[8];
			if(GRAD == GM_WLS) {
				const double4 V = sV[k];
				#pragma unroll
				for(int v = 0; v < 4; v++) {
					g[2*v]   = V.x*acc[2*v] + V.y*acc[2*v+1];
					g[2*v+1] = V.z*acc[2*v] + V.w*acc[2*v+1];
				}
			}
			else if(GRAD == GM_GG) { for(int q = 0; q < 8; q++) g[q] = acc[q]*ainv; }
			else if(GRAD == GM_GIVEN) { ld4(A.gin + 8*(size_t)i + 2*o0, g); ld4(A.gin + 8*(size_t)i + 2*o1, g+4); }
			else { for(int q = 0; q < 8; q++) g[q] = 0.0; }

			// GradBlock rows hold (d/dx, d/dy) of variables 0..3 in order: this thread's first two variables are 0,1 or 2,3
			if(A.gu) { st4(A.gu + 8*(size_t)i + 2*o0, g); st4(A.gu + 8*(size_t)i + 2*o1, g+4); }
			if(!A.lg) continue;

			if(LIM != LM_NONE) {
				// The limiter is the minimum over the faces of a ratio N/D with D > 0 (and of 1). The faces are
				// compared by cross-multiplication and only the winning ratio is divided: one reciprocal per
				// variable instead of one per face and variable. The reference divides per face and takes fmin
				// (limitedlinearreconstruction.cpp:150-170, 244-262); the selected face is the same up to ties.
				// Variable-outer order keeps the live state small (one variable's running minimum at a time).
				double eps2 = 0.0;
				if(LIM == LM_VENKAT) {
					const double kh = A.gas.limiter_param*sclen[k + (c0 & 1)];
					eps2 = kh*kh*kh;
				}
				double ddx[4], ddy[4];
				#pragma unroll
				for(int j = 0; j < 4; j++) {
					const double2 mid = sgr[(j == 3 && !quad) ? 0 : (cf[j] & 0x7FFFu)];
					ddx[j] = mid.x - rci.x; ddy[j] = mid.y - rci.y;
				}
				#pragma unroll
				for(int v = 0; v < 4; v++) {
					double bn = 1.0, bd = 1.0;
					#pragma unroll
					for(int j = 0; j < 4; j++) {
						if(j == 3 && !quad) break;
						const double uface = pi[v] + g[2*v]*ddx[j] + g[2*v+1]*ddy[j];
						const double dm = uface - pi[v];
						double n_, d_;
						if(LIM == LM_VENKAT) {
							const double dp = dm < 0.0 ? dmin[v] : dmax[v];
							const double dp2e = dp*dp + eps2, dpm = dp*dm;
							n_ = dp2e + 2.0*dpm;
							d_ = dp2e + dpm + 2.0*dm*dm;
						} else {
							// Barth-Jespersen: dmax/dm for dm > 0, dmin/dm for dm < 0 (ratios of like signs), else 1
							const bool pos = dm > 0.0;
							n_ = pos ? dmax[v] : -dmin[v];
							d_ = pos ? dm : -dm;
							if(dm == 0.0) { n_ = 1.0; d_ = 1.0; }
						}
						if(n_*bd < bn*d_) { bn = n_; bd = d_; }
					}
					const double lim = bn*frcp(bd);
					g[2*v] *= lim; g[2*v+1] *= lim;
				}
			}
			st4(A.lg + 8*(size_t)i + 2*o0, g); st4(A.lg + 8*(size_t)i + 2*o1, g+4);
		}
		// the staging buffers are free for the next tile once every thread is past the stencil loop; the same barrier
		// orders this tile's gradient stores before the push below reads them back
		// (a tile pushes rows exactly when it sees a ghost cell: the cells next to a cut face are the ones the neighbour needs)
		const bool pushes = distd != nullptr && A.dist.push != 0 && (LOOP || ghost_tile);
		if(have_next || pushes) __syncthreads();
		if(pushes) {
			const int4 rp = A.tdesc[3*(size_t)ti];      // (this tile's record again: nothing of it is held across the stencil loop)
			const unsigned long long dk = A.dist.ctl->k;
			if((A.dist.push & (1u << X_GU)) && A.gu) dist_push_tile(distd, X_GU, dk, rp.x, rp.y, A.gu);
			if((A.dist.push & (1u << X_LG)) && A.lg) dist_push_tile(distd, X_LG, dk, rp.x, rp.y, A.lg);
		}
		if(!LOOP) break;
		if(have_next) { const int4 *const rec = ring + 3*((it + 1) & 1); r0 = rec[0]; r1 = rec[1]; r2 = rec[2]; }
	}
	if(distd && A.dist.last) dist_finish_evaluation(distd, A.dist.ctl->k, false);
}

/// kernel launch with the programmatic-stream-serialization attribute (the kernel may start while its predecessor in
/// the stream drains; it calls griddepcontrol.wait before touching the predecessor's output)
template <typename Kern, typename Args>
static cudaError_t launch_pdl(Kern kernel, int grid, int block, size_t smem, cudaStream_t s, const Args &a)
{
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, a);
}

template <int GRAD, int LIM, bool PRIM_IN, int MODE>
static int launch_cell_grid(const CellArgs &b, int grid, cudaStream_t s)
{
	const CellSmem S(b.m.TC, b.m.HMAX, b.m.EMAX, LIM != LM_NONE || GRAD == GM_GG, GRAD == GM_GG, GRAD == GM_WLS, LIM == LM_VENKAT);
	if(S.total > 48*1024) {
		const cudaError_t ea = cudaFuncSetAttribute(cell_kernel<GRAD,LIM,PRIM_IN,MODE>,
			cudaFuncAttributeMaxDynamicSharedMemorySize, S.total);
		if(ea != cudaSuccess) return cuda_fail(ea, "cell_kernel smem attribute", __FILE__, __LINE__);
	}
	if(grid <= 0) {      // resident-size grid
		static int waves = -1;
		if(waves < 0) { const char *e = getenv("FVG_CELL_PERSIST"); waves = e ? atoi(e) : 1; if(waves < 1) waves = 1; }
		const int nt = b.tile1 - b.tile0;
		const int ctas = waves*resident_ctas((const void*)cell_kernel<GRAD,LIM,PRIM_IN,MODE>, CELL_BLOCK, (size_t)S.total);
		grid = nt < ctas ? nt : ctas;
	}
	CellArgs c = b;
	static int wait_always = -1;
	if(wait_always < 0) { const char *e = getenv("FVG_CELL_PDLWAIT"); wait_always = e ? atoi(e) : 0; }
	c.pdl = ((MODE & CM_LOOP) || wait_always) ? 1 : 0;
	cudaError_t el;
	if(c.pdl) el = launch_pdl(cell_kernel<GRAD,LIM,PRIM_IN,MODE>, grid, CELL_BLOCK, (size_t)S.total, s, c);
	else { cell_kernel<GRAD,LIM,PRIM_IN,MODE><<<grid, CELL_BLOCK, (size_t)S.total, s>>>(c); el = cudaGetLastError(); }
	if(el != cudaSuccess) return cuda_fail(el, "cell_kernel launch", __FILE__, __LINE__);
	return 0;
}


/// Fills in what every launch needs (tile range, tile sequence) and picks the feature set of the kernel
inline int cell_mode_of(const CellArgs &b, bool prim_in)
{
	static int persist = -1;
	if(persist < 0) { const char *e = getenv("FVG_CELL_PERSIST"); persist = e ? atoi(e) : 0; }
	if(prim_in) return CM_PLAIN;      // (the plug-in entry points with primitive input are single-GPU, device-ordered)
	int mode = (b.dist.d ? CM_DIST : 0) | ((b.src_idx || b.ucopy) ? CM_PERM : 0);
	if(persist > 0) mode = CM_LOOP | CM_DIST;
	return mode;
}
int launch_cell_kernel_modes(int grad, int lim, int mode, const CellArgs &b, cudaStream_t s);    // kernels_cell_modes.cu

} // namespace fvg
