/* Device side of the fused multi-GPU evaluation (fvg_dist_*, dist.cu): the producing kernels push the rows their
 * neighbours need straight into the neighbours' peer-mapped windows (NVLink / NVSwitch stores) and the consuming
 * kernels wait on arrival flags inside the CTAs that reach a tile with ghost cells. No exchange kernel, no host round
 * trip, no collective library on the data path. Replaces the reference's L2TraceVector exchange
 * (src/linalg/tracevector.cpp:214-325) and the VecGhostUpdateBegin/End pairs of src/ode/aodesolver.cpp:212,247 and
 * src/spatial/flow_spatial.cpp:711-729.
 *
 * Protocol. Evaluation number k lives in DEVICE memory (DistCtl::k, advanced by the last CTA of the evaluation's last
 * kernel), so the same kernel arguments serve every evaluation and a whole evaluation can be replayed from a CUDA graph.
 * Three row types travel: X_U state (4 doubles), X_GU unlimited gradients (8), X_LG reconstruction gradients (8). A
 * window has, per type, two receive areas used by evaluation parity, and one arrival flag per (type, source rank) that
 * the source sets to k+1 once all its rows of evaluation k are in place:
 *   - state rows of evaluation k are pushed either by the PROLOGUE of the evaluation's first kernel (a handful of CTAs of
 *     its first wave, before anything waits) or, in a pseudo-time loop, by the step epilogue of evaluation k-1;
 *   - gradient rows are pushed by the tile that computed them, right after it has stored them.
 * A kernel never waits for rows that a CTA of the same launch pushes later than its first wave: waits only depend on
 * earlier kernels of the neighbour's stream or on prologues, which is what makes the scheme deadlock-free whatever the
 * residency of the grids. Parity areas: rows of evaluation k+1 overwrite those of k-1, which every neighbour has
 * finished reading before the sender can get there (its own evaluation k needed that neighbour's evaluation-k rows).
 * A wait that exceeds DistDev::spin_ns gives up, records k+1 in DistCtl::timeout and the API returns FVG_ERR_COMM at its
 * next status check (the results of that evaluation are not to be used).
 */
#pragma once
#include "async_copy.cuh"

namespace fvg {

constexpr int MAXRANKS = 16;
enum XType { X_U = 0, X_GU = 1, X_LG = 2, X_COUNT = 3 };
constexpr int NORM_SLOTS = 4;          ///< per-rank norm partials are kept for four steps (deferred gather, see norm kernel)

__host__ __device__ inline int xwidth(int type) { return type == X_U ? 4 : 8; }
/// doubles from the start of the row areas of a window with `ng` ghost rows to the area of (type, parity)
__host__ __device__ inline size_t xarea_off(int type, int parity, size_t ng) {
	return ng*(size_t)((type == X_U ? 0 : (type == X_GU ? 8 : 24)) + parity*xwidth(type));
}
constexpr size_t XAREA_DOUBLES_PER_GHOST = 40;

/// Window header (at the start of every rank's window allocation; mapped by the peers)
struct WinHdr {
	unsigned long long flag[X_COUNT][MAXRANKS];   ///< flag[t][r] = k+1: rank r's rows of type t for evaluation k have arrived
	unsigned long long norm_flag[MAXRANKS];       ///< norm_flag[r] = s+1: rank r's partial norm of step s has arrived
	double norm_val[NORM_SLOTS][MAXRANKS];
	unsigned long long pad[16];
};

/// Control words in local device memory (not mapped by peers)
struct DistCtl {
	unsigned long long k;            ///< current evaluation
	unsigned long long pushed_for;   ///< state rows already pushed (by a step epilogue) for evaluation pushed_for - 1
	unsigned long long timeout;      ///< nonzero: a wait gave up while waiting for this flag value
	unsigned long long step;         ///< pseudo-time steps taken (norm slots)
	unsigned long long hist0;        ///< step whose norm is entry 0 of the current loop's history array
	unsigned done;                   ///< CTAs of the evaluation's last kernel that have finished
	unsigned pro_arrive;             ///< prologue CTAs that have pushed their share
	unsigned arrive[X_COUNT][MAXRANKS];   ///< tiles that have pushed their rows of a type to a peer
};

struct DistDev {
	int nranks, rank, nghost, nsend;
	unsigned char *window;                  ///< local window: WinHdr, then the row areas
	unsigned char *peer[MAXRANKS];          ///< mapped windows of the peers (null: self / no traffic)
	int peer_nghost[MAXRANKS];              ///< ghost rows of peer r
	int peer_row0[MAXRANKS];                ///< first row of this rank's block inside peer r's ghost range
	int recv_off[MAXRANKS+1];               ///< this rank's ghost rows by source rank
	int send_off[MAXRANKS+1];               ///< this rank's send list by peer
	int ntile_send[MAXRANKS];               ///< tiles of this rank that send rows to peer r
	const int *tsoff;                       ///< [ntile+1] per-tile send lists ...
	const int *tsend;                       ///< ... triples (tile-local cell, peer, row inside my block of the peer's ghosts)
	const unsigned *tpeers;                 ///< [ntile] bit mask of the peers a tile sends to
	const int *send_idx;                    ///< [nsend] own cells by peer (prologue push)
	DistCtl *ctl;
	long long spin_ns;                      ///< give up a wait after this many nanoseconds
};

constexpr int DIST_PROLOGUE_CTAS = 32;

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
	asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ double2 ld_cg_f64x2(const double *p) {
	double2 v;
	asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
	return v;
}

/// local receive area of (type, evaluation k): [nghost][width]
__device__ __forceinline__ const double *dist_ghost_rows(const DistDev *d, int type, unsigned long long k) {
	return reinterpret_cast<const double*>(d->window + sizeof(WinHdr)) + xarea_off(type, (int)(k & 1ull), (size_t)d->nghost);
}
/// where this rank's row `row` of (type, evaluation k) goes in peer r's window
__device__ __forceinline__ double *dist_peer_row(const DistDev *d, int r, int type, unsigned long long k, int row) {
	return reinterpret_cast<double*>(d->peer[r] + sizeof(WinHdr)) + xarea_off(type, (int)(k & 1ull), (size_t)d->peer_nghost[r])
	       + (size_t)(d->peer_row0[r] + row)*xwidth(type);
}

/// Whole CTA: thread r waits until source rank r has published the rows of every type in `types` (bit mask) for
/// evaluation k. The barrier makes the acquired rows visible to all threads of the CTA.
__device__ __forceinline__ void dist_wait(const DistDev *d, unsigned types, unsigned long long k)
{
	const int r = threadIdx.x;
	if(r < d->nranks && d->recv_off[r+1] > d->recv_off[r]) {
		const WinHdr *const W = reinterpret_cast<const WinHdr*>(d->window);
		#pragma unroll
		for(int t = 0; t < X_COUNT; t++) {
			if(!((types >> t) & 1u)) continue;
			const unsigned long long *const f = &W->flag[t][r];
			if(ld_acquire_sys_u64(f) >= k + 1) continue;
			const unsigned long long t0 = global_timer_ns();
			unsigned spins = 0;
			while(ld_acquire_sys_u64(f) < k + 1) {
				__nanosleep(32);
				if((++spins & 255u) == 0 && (long long)(global_timer_ns() - t0) > d->spin_ns) {
					atomicMax(&d->ctl->timeout, k + 1);
					break;
				}
			}
		}
	}
	__syncthreads();
}

/// Whole CTA, after the tile's rows of `arr` ([.][width], device order) are stored and a CTA barrier has passed: copies
/// the rows on the tile's send list into the peers' windows and publishes the arrival of the last tile per peer.
__device__ __forceinline__ void dist_push_tile(const DistDev *d, int type, unsigned long long k, int t, int c0, const double *arr)
{
	const int s0 = d->tsoff[t], ns = d->tsoff[t+1] - s0;
	if(ns == 0) return;
	const int width = xwidth(type), w2 = width >> 1;
	for(int q = threadIdx.x; q < ns*w2; q += blockDim.x) {
		const int it = q/w2, c = q - it*w2;
		const int lc = d->tsend[3*(s0 + it)], r = d->tsend[3*(s0 + it) + 1], row = d->tsend[3*(s0 + it) + 2];
		const double2 v = ld_cg_f64x2(arr + (size_t)(c0 + lc)*width + 2*c);
		*reinterpret_cast<double2*>(dist_peer_row(d, r, type, k, row) + 2*c) = v;
	}
	// One system-scope fence per CTA, by the thread that signals: the CTA barrier orders every thread's stores before
	// it (fences are cumulative), and 255 fewer fences per tile travel over NVLink.
	__syncthreads();
	if(threadIdx.x == 0) {
		__threadfence_system();
		unsigned mask = d->tpeers[t];
		while(mask) {
			const int r = __ffs((int)mask) - 1;
			mask &= mask - 1;
			const unsigned prev = atomicAdd(&d->ctl->arrive[type][r], 1u);
			if(prev == (unsigned)d->ntile_send[r] - 1u) {
				d->ctl->arrive[type][r] = 0;
				__threadfence_system();      // (acquire side: the other tiles' fenced stores, seen through the counter)
				st_release_sys_u64(&reinterpret_cast<WinHdr*>(d->peer[r])->flag[type][d->rank], k + 1);
			}
		}
	}
}

/// Prologue push of the state rows of evaluation k (first DIST_PROLOGUE_CTAS CTAs of the evaluation's first kernel; the
/// others return at once). Skipped when a step epilogue has already pushed them. `u` is [ncell+.][4], device order (or
/// caller order with the row map src_idx: single-rank periodic meshes renumbered by the engine).
__device__ __forceinline__ void dist_push_state_prologue(const DistDev *d, unsigned long long k, const double *u, int force,
                                                         const int *src_idx = nullptr)
{
	const int np = (int)gridDim.x < DIST_PROLOGUE_CTAS ? (int)gridDim.x : DIST_PROLOGUE_CTAS;
	if((int)blockIdx.x >= np) return;
	if(!force && d->ctl->pushed_for == k + 1) return;
	const long long tot = 2ll*d->nsend;
	for(long long q = (long long)blockIdx.x*blockDim.x + threadIdx.x; q < tot; q += (long long)np*blockDim.x) {
		const int i = (int)(q >> 1), c = (int)(q & 1);
		int r = 0;
		while(i >= d->send_off[r+1]) r++;
		const int cell = d->send_idx[i];
		const double2 v = *reinterpret_cast<const double2*>(u + 4*(size_t)(src_idx ? src_idx[cell] : cell) + 2*c);
		*reinterpret_cast<double2*>(dist_peer_row(d, r, X_U, k, i - d->send_off[r]) + 2*c) = v;
	}
	__syncthreads();
	if(threadIdx.x == 0) {
		__threadfence_system();
		const unsigned prev = atomicAdd(&d->ctl->pro_arrive, 1u);
		if(prev == (unsigned)np - 1u) {
			d->ctl->pro_arrive = 0;
			__threadfence_system();
			for(int r = 0; r < d->nranks; r++)
				if(d->send_off[r+1] > d->send_off[r])
					st_release_sys_u64(&reinterpret_cast<WinHdr*>(d->peer[r])->flag[X_U][d->rank], k + 1);
		}
	}
}

/// End of the evaluation's last kernel (whole CTA, all its tiles done): the last CTA to get here advances the
/// evaluation counter; `state_pushed` says that this kernel's epilogue pushed the state rows of evaluation k+1.
__device__ __forceinline__ void dist_finish_evaluation(const DistDev *d, unsigned long long k, bool state_pushed)
{
	__syncthreads();
	if(threadIdx.x == 0) {
		__threadfence();
		const unsigned prev = atomicAdd(&d->ctl->done, 1u);
		if(prev == gridDim.x - 1u) {
			d->ctl->done = 0;
			if(state_pushed) d->ctl->pushed_for = k + 2;
			d->ctl->k = k + 1;
			__threadfence();
		}
	}
}

} // namespace fvg
