/* Cell-centred kernels: gradients (Green-Gauss, weighted least squares), slope limiters
 * (Barth-Jespersen, Venkatakrishnan, WENO), the unfused plug-in kernels behind fvg_gradients /
 * fvg_face_values, and the small utility kernels. Reference restated (src/ relative):
 *   spatial/agradientschemes.cpp:62-214 (GG), 219-440 (WLS)
 *   spatial/limitedlinearreconstruction.cpp:28-105 (WENO), 117-176 (BJ), 179-268 (Venkatakrishnan)
 *   spatial/areconstruction.cpp:52-103, spatial/musclreconstruction.cpp:35-130
 *   spatial/flow_spatial.cpp:74-112, 131-310, 659-700; spatial/aoutput.cpp:28-63
 * The gradient + limiter pass is a cell-gather: one thread per cell walks its <= 4 faces, so there
 * is no scatter at all (the reference scatters from faces with omp atomics).
 */
#include "cell_kernel.cuh"

namespace fvg {

bool pdl_enabled()
{
	static int on = -1;
	if(on < 0) { const char *e = getenv("FVG_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
	return on != 0;
}

int resident_ctas(const void *kernel, int block, size_t smem)
{
	struct Key { const void *k; int dev; size_t smem; int ctas; };
	static std::vector<Key> cache;
	int dev = 0;
	cudaGetDevice(&dev);
	for(const Key &c : cache) if(c.k == kernel && c.dev == dev && c.smem == smem) return c.ctas;
	int sms = 0, per = 0;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, block, smem);
	const int ctas = sms*(per > 0 ? per : 1);
	cache.push_back({kernel, dev, smem, ctas});
	return ctas;
}


/** Gradient + limiter pass, one CTA per tile: the single-GPU, device-ordered form, kept as a kernel of its own.
 * It is the same algorithm as cell_kernel<.., CM_PLAIN> (cell_kernel.cuh) in the shape it had before that template grew its
 * multi-GPU / caller-order / looping features; on the 10M-cell benchmark this text runs in 0.490 ms and the template's
 * plain instantiation in 0.548 ms (same instruction mix, different code generation: A/B runs in profiles/r02_cell_kernel_ab.txt),
 * so the headline path keeps it. The tile's cell states and centres (own cells by TMA bulk
 * copy, halo cells by cp.async gathers) and the face midpoints of its stream are staged in shared
 * memory; conserved states are converted to primitive ONCE per staged cell (the reference converts the
 * whole field in a separate pass, flow_spatial.cpp:697-699). Then one thread per own cell gathers its
 * <= 4 neighbours from shared memory: no scatter, no atomics. */
template <int GRAD, int LIM, bool PRIM_IN, bool DIST, bool PERM = false>
__global__ void __launch_bounds__(CELL_BLOCK, FVG_CELL_MINB)
cell_kernel_plain(const CellArgs A)
{
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

	const int t = (int)blockIdx.x + A.tile0, tid = threadIdx.x;
	const int c0 = M.tcell0[t], nc = M.tcell0[t+1] - c0;
	const int h0 = M.thoff[t], nh = M.thoff[t+1] - h0;
	const int e0 = M.fsoff[t], ne = M.fsoff[t+1] - e0;
	constexpr bool NEED_NBRS = GRAD == GM_GG || GRAD == GM_WLS || LIM != LM_NONE;

	if(tid == 0) mbar_init(bar, 1);
	pdl_launch_dependents();
	// (PDL variant) the tile descriptors above are mesh data; the state read below may be the previous kernel's output
	// and the gradient rows written at the end are still being read by the previous face pass until it completes
	pdl_wait();
	__syncthreads();
	// fused multi-GPU evaluation (DIST): a few CTAs of the first wave push the state rows the neighbours need
	if(DIST && A.dist.first && (int)blockIdx.x < DIST_PROLOGUE_CTAS) dist_push_state_prologue(A.dist.d, A.dist.ctl->k, A.u, A.dist.force_push);
	if(tid == 0) {
		unsigned bytes = (unsigned)nc*((PERM ? 0u : 32u) + 16u + 16u + (GRAD == GM_WLS ? 32u : 0u)) + (LIM == LM_VENKAT ? (unsigned)((nc + (c0 & 1) + 1) & ~1)*8u : 0u);
		if(MIDS) bytes += (unsigned)ne*16u;
		if(METRICS) bytes += (unsigned)ne*32u;
		mbar_expect_tx(bar, bytes);
		if(!PERM) bulk_g2s(sp, A.u + 4*(size_t)c0, (unsigned)nc*32u, bar);
		bulk_g2s(src, M.rc + c0, (unsigned)nc*16u, bar);
		// the cells' stencil metadata rides along (consumed from shared memory: no registers held across the staging)
		bulk_g2s(smraw + S.scl, M.cloc + c0, (unsigned)nc*16u, bar);
		if(GRAD == GM_WLS) bulk_g2s(smraw + S.sV, M.wlsV + c0, (unsigned)nc*32u, bar);
		// 8-byte rows: copy whole 16-byte granules starting at the even cell below c0 (the array is padded by one entry)
		if(LIM == LM_VENKAT) bulk_g2s(smraw + S.sclen, M.clength + (c0 & ~1), (unsigned)((nc + (c0 & 1) + 1) & ~1)*8u, bar);
		if(MIDS) bulk_g2s(sgr, M.fgr + e0, (unsigned)ne*16u, bar);
		if(METRICS) { bulk_g2s(smraw + S.sW, M.fgw + e0, (unsigned)ne*16u, bar); bulk_g2s(smraw + S.sW + M.EMAX*16, M.fgln + e0, (unsigned)ne*16u, bar); }
	}
	if(tid == 32 && A.prefetch_distance > 0 && t + A.prefetch_distance < M.ntile) {
		const int tp = t + A.prefetch_distance;
		const int pc0 = M.tcell0[tp], pnc = M.tcell0[tp+1] - pc0;
		const int pe0 = M.fsoff[tp], pne = M.fsoff[tp+1] - pe0;
		if(!PERM) bulk_prefetch_l2(A.u + 4*(size_t)pc0, (unsigned)pnc*32u);
		else {
			// caller-ordered state: the rows themselves are scattered, their indices are not
			const int q0 = pc0 & ~3, q1 = min((pc0 + pnc + 3) & ~3, M.ncell & ~3);
			if(q1 > q0) bulk_prefetch_l2(A.src_idx + q0, (unsigned)(q1 - q0)*4u);
			const int ph0 = M.thoff[tp] & ~3, ph1 = (M.thoff[tp+1] + 3) & ~3;
			if(ph1 > ph0) bulk_prefetch_l2(A.halo_src + ph0, (unsigned)(ph1 - ph0)*4u);
		}
		bulk_prefetch_l2(M.rc + pc0, (unsigned)pnc*16u);
		bulk_prefetch_l2(M.cloc + pc0, (unsigned)pnc*16u);
		{ const int ph0 = M.thoff[tp] & ~3, ph1 = (M.thoff[tp+1] + 3) & ~3; if(ph1 > ph0) bulk_prefetch_l2(M.thalo + ph0, (unsigned)(ph1 - ph0)*4u); }
		if(GRAD == GM_WLS) bulk_prefetch_l2(M.wlsV + pc0, (unsigned)pnc*32u);
		if(MIDS) bulk_prefetch_l2(M.fgr + pe0, (unsigned)pne*16u);
		if(METRICS) { bulk_prefetch_l2(M.fgw + pe0, (unsigned)pne*16u); bulk_prefetch_l2(M.fgln + pe0, (unsigned)pne*16u); }
	}
	// caller-ordered state (A.src_idx: caller's row of every device cell): the own rows are gathered in 16-byte pieces
	// instead of one bulk copy, the halo rows through A.halo_src, and the conversion loop below leaves the own rows in
	// device order in A.ucopy for the face pass - no permutation kernel on either side of the evaluation
	if(PERM) {
		for(int k = tid; k < nc*2; k += CELL_BLOCK) {
			const int row = k >> 1, piece = k & 1;
			cp_async16(sp + 4*row + 2*piece, A.u + 4*(size_t)A.src_idx[c0 + row] + 2*piece);
		}
	}
	const int4 tbq = M.tbnd[t];
	// in-kernel receive of the state's ghost rows: a tile that sees ghost cells waits for the neighbours' rows (its
	// other copies are already in flight), then gathers those rows from the halo window
	const bool ghost_win = (tbq.w >> 16) != 0 && (DIST ? (A.dist.wait & (1u << X_U)) != 0 : A.gs_u.rows != nullptr);
	if(NEED_NBRS) {
		for(int k = tid; k < nh*3; k += CELL_BLOCK) {
			const int h = k/3, piece = k - 3*h;
			const size_t g = (size_t)M.thalo[h0 + h];
			const int row = nc + h;
			if(piece == 2) cp_async16(src + row, M.rc + g);
			else if(!(ghost_win && g >= (size_t)M.ncell)) cp_async16(sp + 4*row + 2*piece, A.u + 4*(PERM ? (size_t)A.halo_src[h0 + h] : g) + 2*piece);
		}
		if(ghost_win) {
			// (the evaluation number is read only here and where rows are pushed: a handful of tiles)
			const double *rows;
			if(DIST) { const unsigned long long dk = A.dist.ctl->k; dist_wait(A.dist.d, 1u << X_U, dk); rows = A.dist.ghost[X_U][dk & 1ull]; }
			else { ghost_wait(A.gs_u, A.gs_u.seq); rows = A.gs_u.rows; }
			for(int k = tid; k < nh*2; k += CELL_BLOCK) {
				const int h = k >> 1, piece = k & 1;
				const size_t g = (size_t)M.thalo[h0 + h];
				if(g >= (size_t)M.ncell) cp_async16(sp + 4*(nc + h) + 2*piece, rows + 4*(g - (size_t)M.ncell) + 2*piece);
			}
		}
		cp_async_commit();
	}
	const int2 tb = make_int2(tbq.y, tbq.z);   // boundary entries of the tile: first (tile-local) and count
	const int grow0 = nc + nh;                 // their ghost cells are staged as rows grow0 .. grow0 + tb.y - 1
	cp_async_wait_all();
	mbar_wait(bar, 0);
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
				const double2 h0 = *reinterpret_cast<const double2*>(sp + 4*k + 2*hb);
				const double2 h1 = *reinterpret_cast<const double2*>(sp + 4*k + 2*(1 - hb));
				const double uc[4] = {hb ? h1.x : h0.x, hb ? h1.y : h0.y, hb ? h0.x : h1.x, hb ? h0.y : h1.y};
				if(PERM && k < nc) st4(A.ucopy + 4*(size_t)(c0 + k), uc);
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
					ld4(A.u + 4*(size_t)(PERM ? A.src_idx[c0 + L] : c0 + L), ui);
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
	// a tile that sees a ghost cell has cells the neighbours need: its gradient rows go to their windows now
	if(DIST && A.dist.push != 0 && (M.tbnd[t].w >> 16) != 0) {
		__syncthreads();
		const unsigned long long dk = A.dist.ctl->k;
		if((A.dist.push & (1u << X_GU)) && A.gu) dist_push_tile(A.dist.d, X_GU, dk, t, M.tcell0[t], A.gu);
		if((A.dist.push & (1u << X_LG)) && A.lg) dist_push_tile(A.dist.d, X_LG, dk, t, M.tcell0[t], A.lg);
	}
}


template <int GRAD, int LIM, bool PRIM_IN, bool DIST, bool PERM = false>
static int launch_cell_plain(const CellArgs &b, int nt, cudaStream_t s)
{
	const CellSmem S(b.m.TC, b.m.HMAX, b.m.EMAX, LIM != LM_NONE || GRAD == GM_GG, GRAD == GM_GG, GRAD == GM_WLS, LIM == LM_VENKAT);
	if(S.total > 48*1024) {
		const cudaError_t ea = cudaFuncSetAttribute(cell_kernel_plain<GRAD,LIM,PRIM_IN,DIST,PERM>, cudaFuncAttributeMaxDynamicSharedMemorySize, S.total);
		if(ea != cudaSuccess) return cuda_fail(ea, "cell_kernel smem attribute", __FILE__, __LINE__);
	}
	cell_kernel_plain<GRAD,LIM,PRIM_IN,DIST,PERM><<<nt, CELL_BLOCK, S.total, s>>>(b);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "cell_kernel launch", __FILE__, __LINE__);
	return 0;
}

int launch_cell_kernel(int grad, int lim, bool prim_in, const CellArgs &a, cudaStream_t s)
{
	CellArgs b = a;
	if(b.tile1 < 0) b.tile1 = b.m.ntile;
	b.tdesc = (b.ordered && b.m.tdesc_ord) ? b.m.tdesc_ord : b.m.tdesc;
	const int nt = b.tile1 - b.tile0;
	if(nt <= 0) return 0;
	int mode = cell_mode_of(b, prim_in);
	if(mode == CM_PLAIN && b.ordered) mode = CM_DIST;       // (a tile sequence other than the natural one: the general kernel)
	static int fast_dist = -1;
	if(fast_dist < 0) { const char *e = getenv("FVG_CELL_FASTDIST"); fast_dist = e ? atoi(e) : 1; }
	if(mode == CM_DIST && b.dist.d && fast_dist && b.tile0 == 0 && nt == b.m.ntile) {
		// multi-GPU, device order: the one-tile-per-CTA kernel in its natural tile order (the partition-boundary tiles are
		// spread over the grid: their pushes travel while the rest computes; the face pass runs them last)
#define D(G,L) if(grad == G && lim == L) return launch_cell_plain<G,L,false,true>(b, nt, s);
		D(GM_ZERO,LM_NONE) D(GM_ZERO,LM_BJ) D(GM_ZERO,LM_VENKAT) D(GM_GG,LM_NONE) D(GM_GG,LM_BJ) D(GM_GG,LM_VENKAT)
		D(GM_WLS,LM_NONE) D(GM_WLS,LM_BJ) D(GM_WLS,LM_VENKAT)
#undef D
	}
	static int fast_perm = -1;
	if(fast_perm < 0) { const char *e = getenv("FVG_CELL_FASTPERM"); fast_perm = e ? atoi(e) : 1; }
	if(mode == CM_PERM && fast_perm && !prim_in && b.src_idx && b.halo_src && b.ucopy && !b.ordered) {
		// single GPU, caller-ordered state: the same one-tile-per-CTA text with the gathers through the permutation
#define D(G,L) if(grad == G && lim == L) return launch_cell_plain<G,L,false,false,true>(b, nt, s);
		D(GM_ZERO,LM_NONE) D(GM_ZERO,LM_BJ) D(GM_ZERO,LM_VENKAT) D(GM_GG,LM_NONE) D(GM_GG,LM_BJ) D(GM_GG,LM_VENKAT)
		D(GM_WLS,LM_NONE) D(GM_WLS,LM_BJ) D(GM_WLS,LM_VENKAT)
#undef D
	}
	if(mode != CM_PLAIN) return launch_cell_kernel_modes(grad, lim, mode, b, s);
#define C(G,L,P) if(grad == G && lim == L && prim_in == P) return launch_cell_plain<G,L,P,false>(b, nt, s);
	C(GM_ZERO,LM_NONE,false) C(GM_ZERO,LM_BJ,false) C(GM_ZERO,LM_VENKAT,false)
	C(GM_GG,LM_NONE,false) C(GM_GG,LM_BJ,false) C(GM_GG,LM_VENKAT,false)
	C(GM_WLS,LM_NONE,false) C(GM_WLS,LM_BJ,false) C(GM_WLS,LM_VENKAT,false)
	C(GM_ZERO,LM_NONE,true) C(GM_GG,LM_NONE,true) C(GM_WLS,LM_NONE,true)
	C(GM_GIVEN,LM_NONE,true) C(GM_GIVEN,LM_BJ,true) C(GM_GIVEN,LM_VENKAT,true)
#undef C
	set_error("cell kernel: unsupported gradient/limiter combination");
	return FVG_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------------
// WENO: weighted average of the cell's and its interior neighbours' gradients (limitedlinearreconstruction.cpp:28-105).
// Persistent CTAs over tiles; the unlimited gradient rows of the tile's own cells (one TMA bulk copy) and of its halo
// cells (cp.async gathers; ghost rows of other ranks from the halo window) are staged in shared memory, then one thread
// per cell averages its <= 5 stencil members out of shared memory. A 64-byte gradient row is four 16-byte chunks, one
// per variable; thread k takes the variables in the rotated order (c + k/2) mod 4, which makes "thread k reads row k"
// free of bank conflicts, and nothing in the weights couples the variables.

__global__ void __launch_bounds__(CELL_BLOCK)
weno_kernel(const __grid_constant__ WenoArgs A)
{
	extern __shared__ __align__(1024) unsigned char smraw[];
	const DMesh &M = A.m;
	double *const sg = reinterpret_cast<double*>(smraw);                                        // [TC + HMAX][8]
	const uint4 *const scl = reinterpret_cast<const uint4*>(smraw + (size_t)(M.TC + M.HMAX)*64);  // [TC]
	uint64_t *const bar = reinterpret_cast<uint64_t*>(smraw + (size_t)(M.TC + M.HMAX)*64 + (size_t)M.TC*16);
	const int tid = threadIdx.x, G = (int)gridDim.x;
	const double epsilon = 1.0e-5;
	if(tid == 0) mbar_init(bar, 1);
	pdl_launch_dependents();
	pdl_wait();
	__syncthreads();
	GhostSrc gsg = A.gs_gu;
	unsigned long long dk = 0;
	if(A.dist.d) {
		dk = A.dist.ctl->k;
		if(A.dist.wait & (1u << X_GU)) gsg.rows = A.dist.ghost[X_GU][dk & 1ull];
	}
	int it = 0;
	for(int ti = blockIdx.x; ti < M.ntile; ti += G, it++) {
		const int4 r0 = A.tdesc[3*(size_t)ti], r1 = A.tdesc[3*(size_t)ti + 1];
		const int t = r0.x, c0 = r0.y, nc = r0.z, h0 = r0.w, nh = r1.x;
		const bool ghost_win = gsg.rows != nullptr && (r1.w >> 16) != 0;
		if(tid == 0) {
			if(it > 0) fence_proxy_async();
			mbar_expect_tx(bar, (unsigned)nc*(64u + 16u));
			bulk_g2s(sg, A.gu + 8*(size_t)c0, (unsigned)nc*64u, bar);
			bulk_g2s(smraw + (size_t)(M.TC + M.HMAX)*64, M.cloc + c0, (unsigned)nc*16u, bar);
		}
		if(ghost_win) {
			if(A.dist.d) dist_wait(A.dist.d, 1u << X_GU, dk);
			else ghost_wait(gsg, gsg.seq);
		}
		for(int k = tid; k < nh*4; k += CELL_BLOCK) {
			const int h = k >> 2, piece = k & 3;
			const size_t g = (size_t)M.thalo[h0 + h];
			const double *const row = (ghost_win && g >= (size_t)M.ncell) ? gsg.rows + 8*(g - (size_t)M.ncell) : A.gu + 8*g;
			cp_async16(sg + 8*(nc + h) + 2*piece, row + 2*piece);
		}
		cp_async_commit();
		cp_async_wait_all();
		mbar_wait(bar, (unsigned)(it & 1));
		__syncthreads();
		for(int k = tid; k < nc; k += CELL_BLOCK) {
			const uint4 cl = scl[k];
			const unsigned nb[4] = {cl.x & 0xFFFFu, cl.x >> 16, cl.y & 0xFFFFu, cl.y >> 16};
			const int rot = (tid >> 1) & 3;
			double2 out[4];
			double wsum[4];
			#pragma unroll
			for(int c = 0; c < 4; c++) {
				const double2 g = *reinterpret_cast<const double2*>(sg + 8*k + 2*((c + rot) & 3));
				const double q = g.x*g.x + g.y*g.y + epsilon;
				const double q2 = q*q;
				const double w = A.lambda/(q2*q2);
				wsum[c] = w; out[c] = make_double2(w*g.x, w*g.y);
			}
			#pragma unroll
			for(int j = 0; j < 4; j++) {
				if(nb[j] >= NB_BND) continue;
				#pragma unroll
				for(int c = 0; c < 4; c++) {
					const double2 h = *reinterpret_cast<const double2*>(sg + 8*nb[j] + 2*((c + rot) & 3));
					const double q = h.x*h.x + h.y*h.y + epsilon;
					const double q2 = q*q;
					const double w = 1.0/(q2*q2);
					wsum[c] += w; out[c].x += w*h.x; out[c].y += w*h.y;
				}
			}
			double *const dst = A.lg + 8*(size_t)(c0 + k);
			#pragma unroll
			for(int c = 0; c < 4; c++)
				*reinterpret_cast<double2*>(dst + 2*((c + rot) & 3)) = make_double2(out[c].x/wsum[c], out[c].y/wsum[c]);
		}
		__syncthreads();
		if(A.dist.d && (A.dist.push & (1u << X_LG))) dist_push_tile(A.dist.d, X_LG, dk, t, c0, A.lg);
	}
	if(A.dist.d && A.dist.last) dist_finish_evaluation(A.dist.d, dk, false);
}

int launch_weno_kernel(const WenoArgs &a, cudaStream_t s)
{
	const size_t smem = (size_t)(a.m.TC + a.m.HMAX)*64 + (size_t)a.m.TC*16 + 16;
	if(smem > 48*1024) {
		const cudaError_t ea = cudaFuncSetAttribute(weno_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if(ea != cudaSuccess) return cuda_fail(ea, "weno_kernel smem attribute", __FILE__, __LINE__);
	}
	const int ctas = resident_ctas((const void*)weno_kernel, CELL_BLOCK, smem);
	WenoArgs b = a;
	b.tdesc = (b.ordered && b.m.tdesc_ord) ? b.m.tdesc_ord : b.m.tdesc;
	const cudaError_t el = launch_pdl(weno_kernel, a.m.ntile < ctas ? a.m.ntile : ctas, CELL_BLOCK, smem, s, b);
	if(el != cudaSuccess) return cuda_fail(el, "weno_kernel launch", __FILE__, __LINE__);
	return 0;
}

// ------------------------------------------------------------------------------------------------
// unfused face values (plug-in parity with SolutionReconstruction::compute_face_values), CTA per tile

__device__ __forceinline__ void extrapolate4g(const double pc[4], const double *g, double dx, double dy, double pf[4]) {
	double ga[4], gb[4];
	ld4(g, ga); ld4(g+4, gb);
	pf[0] = pc[0] + ga[0]*dx + ga[1]*dy;
	pf[1] = pc[1] + ga[2]*dx + ga[3]*dy;
	pf[2] = pc[2] + gb[0]*dx + gb[1]*dy;
	pf[3] = pc[3] + gb[2]*dx + gb[3]*dy;
}

__global__ void __launch_bounds__(FACE_BLOCK)
face_values_kernel(const FaceValArgs A)
{
	const DMesh &M = A.m;
	const int t = blockIdx.x;
	const int c0 = M.tcell0[t], nc = M.tcell0[t+1] - c0;
	for(int e = M.fsoff[t] + threadIdx.x; e < M.fsoff[t+1]; e += FACE_BLOCK) {
		// only the copy that lives in the left cell's tile writes (duplicates carry -1-f, padding INT_MIN)
		const int f = M.fref[e];
		if(f < 0) continue;
		const unsigned LR = M.fLR[e];
		const bool bnd = (LR >> 16) >= LR_BND;
		const size_t L = (size_t)tile_global(M, t, c0, nc, LR & 0xFFFFu);
		const size_t R = bnd ? 0 : (size_t)tile_global(M, t, c0, nc, LR >> 16);
		const double2 gr = M.fgr[e];
		const double2 rl = M.rc[L];
		double pl[4], pr[4], out[4];
		ld4(A.up + 4*L, pl);
		if(!A.muscl) {
			extrapolate4g(pl, A.g + 8*L, gr.x - rl.x, gr.y - rl.y, out);
			st4(A.ufl + 4*(size_t)f, out);
			if(!bnd) {
				const double2 rr = M.rc[R];
				ld4(A.up + 4*R, pr);
				extrapolate4g(pr, A.g + 8*R, gr.x - rr.x, gr.y - rr.y, out);
				st4(A.ufr + 4*(size_t)f, out);
			}
		}
		else {
			double2 rr;
			if(!bnd) { ld4(A.up + 4*R, pr); rr = M.rc[R]; }
			else { ld4(A.ug + 4*(size_t)f, pr); rr = M.rcbp[f]; }
			const double dx = rr.x - rl.x, dy = rr.y - rl.y;
			double g[8];
			ld4(A.g + 8*L, g); ld4(A.g + 8*L + 4, g+4);
			for(int k = 0; k < 4; k++) {
				const double dlr = pr[k] - pl[k];
				out[k] = pl[k] + muscl_term(2.0*(g[2*k]*dx + g[2*k+1]*dy) - dlr, dlr);
			}
			st4(A.ufl + 4*(size_t)f, out);
			if(!bnd) {
				ld4(A.g + 8*R, g); ld4(A.g + 8*R + 4, g+4);
				for(int k = 0; k < 4; k++) {
					const double dlr = pr[k] - pl[k];
					out[k] = pr[k] - muscl_term(2.0*(g[2*k]*dx + g[2*k+1]*dy) - dlr, dlr);
				}
				st4(A.ufr + 4*(size_t)f, out);
			}
		}
	}
}

int launch_face_values(const FaceValArgs &a, cudaStream_t s)
{
	face_values_kernel<<<a.m.ntile, FACE_BLOCK, 0, s>>>(a);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "face_values_kernel launch", __FILE__, __LINE__);
	return 0;
}

// ------------------------------------------------------------------------------------------------
// utilities

/// gather: dst[i] = src[idx[i]]; scatter: dst[idx[i]] (+)= src[i]; rows of `width` doubles
__global__ void permute_rows_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                    const int *__restrict__ idx, int n, int width, int gather, int accumulate)
{
	const long long k = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(k >= (long long)n*width) return;
	const int i = (int)(k/width), c = (int)(k%width);
	const int j = idx[i];
	if(gather) dst[k] = src[(size_t)j*width + c];
	else if(accumulate) dst[(size_t)j*width + c] += src[k];
	else dst[(size_t)j*width + c] = src[k];
}

int launch_permute_rows(const double *src, double *dst, const int *idx, int n, int width, bool gather,
                        bool accumulate, cudaStream_t s)
{
	const long long tot = (long long)n*width;
	if(tot == 0) return 0;
	permute_rows_kernel<<<(unsigned)((tot + 255)/256), 256, 0, s>>>(src, dst, idx, n, width, gather, accumulate);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "permute_rows launch", __FILE__, __LINE__);
	return 0;
}

/// sendbuf[k][:] = src[send_idx[k]][:]
__global__ void halo_pack_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                 const int *__restrict__ idx, int n, int width)
{
	const long long k = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(k >= (long long)n*width) return;
	const int i = (int)(k/width), c = (int)(k - (long long)i*width);
	dst[k] = src[(size_t)idx[i]*width + c];
}

int launch_halo_pack(const DMesh &m, const double *src, int width, double *dst, cudaStream_t s)
{
	const long long tot = (long long)m.nsend*width;
	if(tot == 0) return 0;
	halo_pack_kernel<<<(unsigned)((tot + 255)/256), 256, 0, s>>>(src, dst, m.send_idx, m.nsend, width);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "halo_pack launch", __FILE__, __LINE__);
	return 0;
}

__global__ void boundary_states_kernel(const DMesh M, const GasParams G,
                                       const double *__restrict__ ins, double *__restrict__ gs)
{
	const int b = blockIdx.x*blockDim.x + threadIdx.x;
	if(b >= M.nbface) return;
	const double2 n = M.fn[M.bentry[b]];
	double in[4], out[4];
	ld4(ins + 4*(size_t)b, in);
	ghost_state(G, G.bc[M.bslot[b]], in, n.x, n.y, out);
	st4(gs + 4*(size_t)b, out);
}

int launch_boundary_states(const DMesh &m, const GasParams &g, const double *ins,
                           double *gs, cudaStream_t s)
{
	if(m.nbface == 0) return 0;
	boundary_states_kernel<<<(m.nbface + 127)/128, 128, 0, s>>>(m, g, ins, gs);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "boundary_states launch", __FILE__, __LINE__);
	return 0;
}

/// ghost state of each boundary face from the adjacent CELL state (device order), conserved or primitive
__global__ void boundary_cell_ghosts_kernel(const DMesh M, const GasParams G,
                                            const double *__restrict__ u, double *__restrict__ ug, int prim_out)
{
	const int b = blockIdx.x*blockDim.x + threadIdx.x;
	if(b >= M.nbface) return;
	const double2 n = M.fn[M.bentry[b]];
	double in[4], out[4];
	ld4(u + 4*(size_t)M.bcell[b], in);
	ghost_state(G, G.bc[M.bslot[b]], in, n.x, n.y, out);
	if(prim_out) { double p[4]; cons2prim(G, out, p); st4(ug + 4*(size_t)b, p); }
	else st4(ug + 4*(size_t)b, out);
}

int launch_boundary_prim_ghosts(const DMesh &m, const GasParams &g, const double *u,
                                double *ug, bool prim_out, cudaStream_t s)
{
	if(m.nbface == 0) return 0;
	boundary_cell_ghosts_kernel<<<(m.nbface + 127)/128, 128, 0, s>>>(m, g, u, ug, prim_out ? 1 : 0);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "boundary_cell_ghosts launch", __FILE__, __LINE__);
	return 0;
}

__global__ void cons2prim_kernel(const GasParams G, const double *__restrict__ u, double *__restrict__ p, int n)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	double a[4], b[4];
	ld4(u + 4*(size_t)i, a);
	cons2prim(G, a, b);
	st4(p + 4*(size_t)i, b);
}

int launch_cons2prim(const GasParams &g, const double *u, double *p, int n, cudaStream_t s)
{
	if(n == 0) return 0;
	cons2prim_kernel<<<(n + 255)/256, 256, 0, s>>>(g, u, p, n);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "cons2prim launch", __FILE__, __LINE__);
	return 0;
}

/// Single-block, fixed-order sum of the per-tile partial norms (deterministic)
__global__ void final_norm_kernel(const double *__restrict__ partial, int n, double *__restrict__ out)
{
	__shared__ double s[1024];
	double acc = 0.0;
	for(int k = threadIdx.x; k < n; k += 1024) acc += partial[k];
	s[threadIdx.x] = acc;
	__syncthreads();
	for(int o = 512; o > 0; o >>= 1) {
		if(threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
		__syncthreads();
	}
	if(threadIdx.x == 0) out[0] = s[0];
}

int launch_final_norm(const double *partial, int n, double *out, cudaStream_t s)
{
	final_norm_kernel<<<1, 1024, 0, s>>>(partial, n, out);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "final_norm launch", __FILE__, __LINE__);
	return 0;
}

/// Cl, Cdp, Cdf numerators and the wetted length, summed in boundary-face order by one thread block
/// of one warp... the boundary is O(sqrt(N)) faces, so a single CTA with a fixed-order reduction.
__global__ void surface_kernel(const DMesh M, const GasParams G, const double aoa, const double *__restrict__ u,
                               const double *__restrict__ grad, const int slot, double *__restrict__ out4)
{
	__shared__ double s[4][256];
	double acc[4] = {0,0,0,0};
	const double wx = cos(aoa), wy = sin(aoa);
	for(int b = threadIdx.x; b < M.nbface; b += 256) {
		if(M.bslot[b] != slot) continue;
		const int e = M.bentry[b];
		const int c = M.bcell[b];
		const double2 n = M.fn[e];
		const double len = M.flen[e];
		double uc[4], g[8];
		ld4(u + 4*(size_t)c, uc);
		ld4(grad + 8*(size_t)c, g); ld4(grad + 8*(size_t)c + 4, g+4);
		const double p = pressure_cons(G, uc);
		const double cp = (p - G.pinf)*2.0;
		const double mu = viscosity_cons(G, uc);
		// velocity gradients from conserved-variable gradients: d(v_i)/dx_j
		double gv[2][2];
		for(int i = 0; i < 2; i++)
			for(int j = 0; j < 2; j++)
				gv[i][j] = (g[j + 2*(i+1)]*uc[0] - uc[i+1]*g[j])/(uc[0]*uc[0]);
		const double nn[2] = {n.x, n.y};
		const double tx = n.y, ty = -n.x;
		double fx = 0, fy = 0;
		for(int j = 0; j < 2; j++) { fx += (gv[0][j] + gv[j][0])*nn[j]; fy += (gv[1][j] + gv[j][1])*nn[j]; }
		const double tauw = mu*(fx*tx + fy*ty);
		const double cf = 2.0*tauw;
		acc[0] += len;
		acc[1] += cp*(n.x*(-wy) + n.y*wx)*len;
		acc[2] += cp*(n.x*wx + n.y*wy)*len;
		acc[3] += cf*(tx*wx + ty*wy)*len;
	}
	for(int k = 0; k < 4; k++) s[k][threadIdx.x] = acc[k];
	__syncthreads();
	for(int o = 128; o > 0; o >>= 1) {
		if(threadIdx.x < o) for(int k = 0; k < 4; k++) s[k][threadIdx.x] += s[k][threadIdx.x + o];
		__syncthreads();
	}
	if(threadIdx.x == 0) {
		out4[0] = s[1][0]/s[0][0]; out4[1] = s[2][0]/s[0][0]; out4[2] = s[3][0]/s[0][0]; out4[3] = s[0][0];
	}
}

int launch_surface_data(const DMesh &m, const GasParams &g, double aoa, const double *u, const double *grads,
                        int slot, double *out4, cudaStream_t s)
{
	surface_kernel<<<1, 256, 0, s>>>(m, g, aoa, u, grads, slot, out4);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "surface_kernel launch", __FILE__, __LINE__);
	return 0;
}

/// per-block partial sums of ((s - s_inf)/s_inf)^2 * area
__global__ void entropy_kernel(const DMesh M, const GasParams G, const double *__restrict__ u,
                               double *__restrict__ partial)
{
	__shared__ double s[256];
	const int i = blockIdx.x*256 + threadIdx.x;
	double v = 0.0;
	if(i < M.ncell) {
		double uc[4];
		ld4(u + 4*(size_t)i, uc);
		const double sinf = pressure_cons(G, G.uinf)/pow(G.uinf[0], G.g);
		const double se = (pressure_cons(G, uc)/pow(uc[0], G.g) - sinf)/sinf;
		v = se*se*M.area[i];
	}
	s[threadIdx.x] = v;
	__syncthreads();
	for(int o = 128; o > 0; o >>= 1) {
		if(threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
		__syncthreads();
	}
	if(threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

int launch_entropy(const DMesh &m, const GasParams &g, const double *u, double *out, cudaStream_t s)
{
	const int nblk = (m.ncell + 255)/256;
	double *partial = nullptr;
	FVG_CUDA(cudaMallocAsync(&partial, sizeof(double)*nblk, s));
	entropy_kernel<<<nblk, 256, 0, s>>>(m, g, u, partial);
	cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "entropy_kernel launch", __FILE__, __LINE__);
	const int rc = launch_final_norm(partial, nblk, out, s);
	FVG_CUDA(cudaFreeAsync(partial, s));
	return rc;
}

// ------------------------------------------------------------------------------------------------
// vector kernels of the matrix-free Jacobian-vector product (linalg/alinalg.cpp:143-230)

/// per-block partial sums of x_i^2 (fixed order inside the block; summed by final_norm_kernel)
__global__ void sumsq_kernel(const double *__restrict__ x, long long n, double *__restrict__ partial)
{
	__shared__ double s[256];
	double acc = 0.0;
	for(long long i = (long long)blockIdx.x*256 + threadIdx.x; i < n; i += (long long)gridDim.x*256) acc += x[i]*x[i];
	s[threadIdx.x] = acc;
	__syncthreads();
	for(int o = 128; o > 0; o >>= 1) {
		if(threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
		__syncthreads();
	}
	if(threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

/// aux = u + (eps/|x|) x; |x|^2 is read from device memory (no host round trip)
__global__ void perturb_kernel(const double *__restrict__ u, const double *__restrict__ x, const double *__restrict__ xnorm2,
                               double eps, long long n, double *__restrict__ aux)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	const double pertmag = eps/sqrt(xnorm2[0]);
	aux[i] = u[i] + pertmag*x[i];
}

/// y = mdt x + (res - yg)/(eps/|x|): res and yg hold -r(u) and -r(u + pert) as compute_residual leaves them
__global__ void jvp_combine_kernel(const double *__restrict__ x, const double *__restrict__ res, const double *__restrict__ yg,
                                   const double *__restrict__ mdt, const double *__restrict__ xnorm2, double eps,
                                   int ncell, int nvars, double *__restrict__ y)
{
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= (long long)ncell*nvars) return;
	const double pertmag = eps/sqrt(xnorm2[0]);
	y[i] = mdt[i/nvars]*x[i] + (-yg[i] + res[i])/pertmag;
}

int launch_sumsq(const double *x, long long n, double *partial, int nblk, double *out, cudaStream_t s)
{
	sumsq_kernel<<<nblk, 256, 0, s>>>(x, n, partial);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "sumsq_kernel launch", __FILE__, __LINE__);
	return launch_final_norm(partial, nblk, out, s);
}

int launch_perturb(const double *u, const double *x, const double *xnorm2, double eps, long long n, double *aux, cudaStream_t s)
{
	if(n == 0) return 0;
	perturb_kernel<<<(unsigned)((n + 255)/256), 256, 0, s>>>(u, x, xnorm2, eps, n, aux);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "perturb_kernel launch", __FILE__, __LINE__);
	return 0;
}

int launch_jvp_combine(const double *x, const double *res, const double *yg, const double *mdt, const double *xnorm2,
                       double eps, int ncell, int nvars, double *y, cudaStream_t s)
{
	const long long n = (long long)ncell*nvars;
	if(n == 0) return 0;
	jvp_combine_kernel<<<(unsigned)((n + 255)/256), 256, 0, s>>>(x, res, yg, mdt, xnorm2, eps, ncell, nvars, y);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "jvp_combine_kernel launch", __FILE__, __LINE__);
	return 0;
}

// ------------------------------------------------------------------------------------------------
// pointwise test hooks

template <int FLUX>
__global__ void pw_flux_kernel(const GasParams G, int n, const double *__restrict__ ul,
                               const double *__restrict__ ur, const double *__restrict__ nrm,
                               double *__restrict__ out)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	double a[4], b[4], f[4];
	for(int k = 0; k < 4; k++) { a[k] = ul[4*i+k]; b[k] = ur[4*i+k]; }
	inviscid_flux<FLUX>(G, a, b, nrm[2*i], nrm[2*i+1], f);
	for(int k = 0; k < 4; k++) out[4*i+k] = f[k];
}

int launch_pointwise_flux(int flux, const GasParams &g, int n, const double *ul, const double *ur,
                          const double *nrm, double *out, cudaStream_t s)
{
	if(n == 0) return 0;
	const int nb = (n + 127)/128;
	switch(flux) {
	case FLUX_LLF: pw_flux_kernel<FLUX_LLF><<<nb,128,0,s>>>(g,n,ul,ur,nrm,out); break;
	case FLUX_VANLEER: pw_flux_kernel<FLUX_VANLEER><<<nb,128,0,s>>>(g,n,ul,ur,nrm,out); break;
	case FLUX_AUSM: pw_flux_kernel<FLUX_AUSM><<<nb,128,0,s>>>(g,n,ul,ur,nrm,out); break;
	case FLUX_AUSMPLUS: pw_flux_kernel<FLUX_AUSMPLUS><<<nb,128,0,s>>>(g,n,ul,ur,nrm,out); break;
	case FLUX_ROE: pw_flux_kernel<FLUX_ROE><<<nb,128,0,s>>>(g,n,ul,ur,nrm,out); break;
	case FLUX_HLL: pw_flux_kernel<FLUX_HLL><<<nb,128,0,s>>>(g,n,ul,ur,nrm,out); break;
	case FLUX_HLLC: pw_flux_kernel<FLUX_HLLC><<<nb,128,0,s>>>(g,n,ul,ur,nrm,out); break;
	default: set_error("unknown flux id"); return FVG_ERR_INVALID;
	}
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "pw_flux launch", __FILE__, __LINE__);
	return 0;
}

__global__ void pw_bc_kernel(const GasParams G, int n, const double *__restrict__ ins,
                             const double *__restrict__ nrm, double *__restrict__ out)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	double a[4], g[4];
	for(int k = 0; k < 4; k++) a[k] = ins[4*i+k];
	ghost_state(G, G.bc[0], a, nrm[2*i], nrm[2*i+1], g);
	for(int k = 0; k < 4; k++) out[4*i+k] = g[k];
}

int launch_pointwise_bc(const GasParams &g, int n, const double *ins, const double *nrm, double *out,
                        cudaStream_t s)
{
	if(n == 0) return 0;
	pw_bc_kernel<<<(n + 127)/128, 128, 0, s>>>(g, n, ins, nrm, out);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "pw_bc launch", __FILE__, __LINE__);
	return 0;
}

template <bool ORDER2, bool CV>
__global__ void pw_visc_kernel(const GasParams G, int n, const double *__restrict__ nrm,
                               const double *__restrict__ rcl, const double *__restrict__ rcr,
                               const double *__restrict__ ucl, const double *__restrict__ ucr,
                               const double *__restrict__ gl, const double *__restrict__ gr,
                               const double *__restrict__ ul, const double *__restrict__ ur,
                               double *__restrict__ out)
{
	const int i = blockIdx.x*blockDim.x + threadIdx.x;
	if(i >= n) return;
	double a[4], b[4], c[4], d[4], g1[8], g2[8], f[4];
	for(int k = 0; k < 4; k++) { a[k] = ucl[4*i+k]; b[k] = ucr[4*i+k]; c[k] = ul[4*i+k]; d[k] = ur[4*i+k]; }
	for(int k = 0; k < 8; k++) { g1[k] = ORDER2 ? gl[8*i+k] : 0.0; g2[k] = ORDER2 ? gr[8*i+k] : 0.0; }
	viscous_face_flux<ORDER2,CV>(G, nrm[2*i], nrm[2*i+1], rcl[2*i], rcl[2*i+1], rcr[2*i], rcr[2*i+1],
	                             a, b, g1, g2, c, d, f);
	for(int k = 0; k < 4; k++) out[4*i+k] = f[k];
}

int launch_pointwise_visc(const GasParams &g, bool order2, bool cv, int n, const double *nrm,
                          const double *rcl, const double *rcr, const double *ucl, const double *ucr,
                          const double *gl, const double *gr, const double *ul, const double *ur,
                          double *out, cudaStream_t s)
{
	if(n == 0) return 0;
	const int nb = (n + 127)/128;
	if(order2 && cv) pw_visc_kernel<true,true><<<nb,128,0,s>>>(g,n,nrm,rcl,rcr,ucl,ucr,gl,gr,ul,ur,out);
	else if(order2) pw_visc_kernel<true,false><<<nb,128,0,s>>>(g,n,nrm,rcl,rcr,ucl,ucr,gl,gr,ul,ur,out);
	else if(cv) pw_visc_kernel<false,true><<<nb,128,0,s>>>(g,n,nrm,rcl,rcr,ucl,ucr,gl,gr,ul,ur,out);
	else pw_visc_kernel<false,false><<<nb,128,0,s>>>(g,n,nrm,rcl,rcr,ucl,ucr,gl,gr,ul,ur,out);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "pw_visc launch", __FILE__, __LINE__);
	return 0;
}

} // namespace fvg
