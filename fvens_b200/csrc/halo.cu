/* Peer-memory halo exchange over NVLink / NVSwitch: every rank owns a WINDOW (receive areas + arrival flags in
 * one cudaMalloc allocation) that the other ranks of the box map through CUDA IPC. An exchange is two small
 * kernels and no host round trip, no collective library:
 *   send  packs the rows each peer needs (device-mesh send lists) and stores them straight into that peer's
 *         window, then releases a sequence number into the peer's flag word (st.release.sys);
 *   recv  acquires the flags of the peers it expects rows from and copies the rows into the ghost block of the
 *         caller's array.
 * Replaces the reference's L2TraceVector::updateSharedFacesBegin/End (src/linalg/tracevector.cpp:214-325, MPI
 * Isend/Irecv of packed face states) and its VecGhostUpdateBegin/End calls (src/ode/aodesolver.cpp:212,247;
 * src/spatial/flow_spatial.cpp:711-729). A window has three receive areas used round robin by sequence number: a
 * rank may run one exchange ahead of a neighbour without overwriting rows the neighbour has not consumed yet (it
 * cannot get two ahead: its next receive needs the neighbour's next send), and with the IN-KERNEL RECEIVE
 * (fvg_halo_post + fvg_flow_ghost_source: no recv kernel, the consuming pass waits on the flags itself and reads the
 * window directly) the two most recent exchanges - state and gradients of one evaluation - stay readable while the
 * next state exchange is already arriving.
 */
#include "engine.hpp"
#include <cstring>
#include <memory>
#include <string>

struct fvg_halo {
	fvg_mesh *mesh = nullptr;
	int nranks = 1, rank = 0, max_width = 0;
	size_t area_doubles = 0;                 ///< doubles per parity buffer (nghost * max_width)
	unsigned char *window = nullptr;         ///< local window: header (nranks flags + 1 error word, uint64, padded to 256 B), then [3][area] doubles
	size_t hdr = 0;                          ///< header bytes (the same on every rank)
	std::vector<unsigned long long> peer_area;   ///< area_doubles of each peer's window (its parity-buffer stride)
	unsigned long long *d_peer_area = nullptr;
	std::vector<unsigned char*> peer;        ///< mapped windows of the peers (nullptr for self / ranks without traffic)
	std::vector<int> send_off, recv_off;     ///< row offsets of each peer's block in my send list / my ghost range
	std::vector<int> peer_row0;              ///< row offset of MY block inside peer r's ghost range
	// device copies of the per-peer tables
	unsigned char **d_peer = nullptr;
	int *d_send_off = nullptr, *d_recv_off = nullptr, *d_peer_row0 = nullptr;
	unsigned *d_arrive = nullptr;            ///< per-peer CTA arrival counters of the send kernel
	unsigned long long seq = 0;
	bool connected = false;
};

namespace fvg {

constexpr int HALO_CTAS_PER_PEER = 4;
constexpr unsigned long long HALO_NBUF = 3;   ///< receive areas per window, used round robin by sequence number
constexpr int HALO_THREADS = 512;

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
	asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

/** grid = nranks * HALO_CTAS_PER_PEER. CTAs [r*K, (r+1)*K) serve peer r: they copy rows send_idx[send_off[r] ..
 * send_off[r+1]) of src into the peer's window at row peer_row0[r], 16 bytes per store; the last of the K CTAs to
 * finish publishes the sequence number in the peer's flag word for this rank. */
__device__ __forceinline__ void
halo_send_block(const int block, const double *__restrict__ src, const int *__restrict__ send_idx, const int *__restrict__ send_off,
                 const int *__restrict__ peer_row0, unsigned char *const *__restrict__ peer, unsigned *__restrict__ arrive,
                 const unsigned long long *__restrict__ peer_area, const int width, const size_t hdr, const int rank,
                 const unsigned long long seq)
{
	const int r = block/HALO_CTAS_PER_PEER, sub = block - r*HALO_CTAS_PER_PEER;
	const int k0 = send_off[r], nrow = send_off[r+1] - k0;
	if(nrow == 0 || peer[r] == nullptr) return;
	double *const dst = reinterpret_cast<double*>(peer[r] + hdr) + (seq % HALO_NBUF)*peer_area[r] + (size_t)peer_row0[r]*width;
	const int w2 = width/2;                   // widths are even (4 or 8): rows move as 16-byte pieces
	const long long tot = (long long)nrow*w2;
	for(long long q = (long long)sub*HALO_THREADS + threadIdx.x; q < tot; q += (long long)HALO_CTAS_PER_PEER*HALO_THREADS) {
		const int row = (int)(q/w2), c = (int)(q - (long long)row*w2);
		const double2 v = *reinterpret_cast<const double2*>(src + (size_t)send_idx[k0 + row]*width + 2*c);
		*reinterpret_cast<double2*>(dst + (size_t)row*width + 2*c) = v;
	}
	__threadfence_system();
	__syncthreads();
	if(threadIdx.x == 0) {
		const unsigned prev = atomicAdd(&arrive[r], 1u);
		if(prev == HALO_CTAS_PER_PEER - 1) {
			arrive[r] = 0;
			__threadfence_system();
			st_release_sys(reinterpret_cast<unsigned long long*>(peer[r]) + rank, seq);
		}
	}
}

/** Waits until every peer that sends rows has published `seq`, then copies the window's rows into the ghost block
 * of dst (rows [ncell, ncell + nghost)). A bounded spin: if a peer never arrives the error word is set and the copy
 * proceeds (the caller reads the status; nothing hangs). */
__device__ __forceinline__ void
halo_recv_block(const int block, const int nblocks, double *__restrict__ dst, const unsigned char *__restrict__ window, const int *__restrict__ recv_off,
                 const int ncell, const int nghost, const int width, const size_t area_doubles, const size_t hdr, const int nranks,
                 const unsigned long long seq, const long long spin_limit)
{
	const unsigned long long *const flags = reinterpret_cast<const unsigned long long*>(window);
	unsigned long long *const err = const_cast<unsigned long long*>(flags) + nranks;
	if((int)threadIdx.x < nranks) {
		const int r = threadIdx.x;
		if(recv_off[r+1] > recv_off[r]) {
			long long spins = 0;
			while(ld_acquire_sys(flags + r) < seq) {
				__nanosleep(64);
				if(++spins > spin_limit) { atomicExch(err, seq); break; }
			}
		}
	}
	__syncthreads();
	const double *const srcw = reinterpret_cast<const double*>(window + hdr) + (seq % HALO_NBUF)*area_doubles;
	const long long tot = (long long)nghost*width/2;
	double2 *const out = reinterpret_cast<double2*>(dst + (size_t)ncell*width);
	const double2 *const in = reinterpret_cast<const double2*>(srcw);
	for(long long q = (long long)block*HALO_THREADS + threadIdx.x; q < tot; q += (long long)nblocks*HALO_THREADS)
		out[q] = in[q];
}

struct HaloArgs {
	const double *src; double *dst;
	const int *send_idx, *send_off, *peer_row0, *recv_off;
	unsigned char *const *peer; unsigned *arrive; const unsigned long long *peer_area;
	const unsigned char *window;
	int width, rank, nranks, ncell, nghost, nsend_blocks, nrecv_blocks;
	size_t hdr, area_doubles;
	unsigned long long seq; long long spin_limit;
};

/// One launch: the first nsend_blocks CTAs push this rank's rows into the neighbours' windows, the remaining CTAs wait
/// for the neighbours' rows and unpack them. All CTAs are co-resident (at most 4*nranks + 16), so the waiting ones
/// cannot starve the pushing ones. Either part may be empty (nsend_blocks = 0: receive only; nrecv_blocks = 0: send only).
__global__ void __launch_bounds__(HALO_THREADS)
halo_kernel(const HaloArgs a)
{
	if((int)blockIdx.x < a.nsend_blocks)
		halo_send_block(blockIdx.x, a.src, a.send_idx, a.send_off, a.peer_row0, a.peer, a.arrive, a.peer_area, a.width, a.hdr, a.rank, a.seq);
	else
		halo_recv_block(blockIdx.x - a.nsend_blocks, a.nrecv_blocks, a.dst, a.window, a.recv_off, a.ncell, a.nghost, a.width,
		                a.area_doubles, a.hdr, a.nranks, a.seq, a.spin_limit);
}

} // namespace fvg

using namespace fvg;

extern "C" {

int fvg_halo_create(fvg_mesh *mesh, int max_width, fvg_halo **out)
{
	if(!mesh || !out || max_width < 2 || (max_width & 1)) { set_error("fvg_halo_create: bad argument (width must be even)"); return FVG_ERR_INVALID; }
	*out = nullptr;
	if(mesh->device < 0) { set_error("fvg_halo_create: host-only mesh"); return FVG_ERR_INVALID; }
	FVG_CUDA(cudaSetDevice(mesh->device));
	std::unique_ptr<fvg_halo> h(new fvg_halo);
	h->mesh = mesh; h->nranks = mesh->nranks; h->rank = mesh->rank; h->max_width = max_width;
	h->area_doubles = (size_t)std::max(mesh->d.nghost, 1)*max_width;
	h->hdr = (((size_t)h->nranks + 1)*sizeof(unsigned long long) + 255)/256*256;
	const size_t bytes = h->hdr + HALO_NBUF*h->area_doubles*sizeof(double);
	FVG_CUDA(cudaMalloc((void**)&h->window, bytes));
	FVG_CUDA(cudaMemset(h->window, 0, bytes));
	h->send_off.assign(h->nranks + 1, 0); h->recv_off.assign(h->nranks + 1, 0);
	for(int r = 0; r < h->nranks; r++) {
		h->send_off[r+1] = h->send_off[r] + mesh->send_counts[r];
		h->recv_off[r+1] = h->recv_off[r] + mesh->recv_counts[r];
	}
	FVG_CUDA(cudaDeviceSynchronize());
	*out = h.release();
	return 0;
}

int fvg_halo_ipc_handle(fvg_halo *h, void *handle64)
{
	if(!h || !handle64) { set_error("fvg_halo_ipc_handle: null argument"); return FVG_ERR_INVALID; }
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	cudaIpcMemHandle_t mh;
	FVG_CUDA(cudaIpcGetMemHandle(&mh, h->window));
	std::memcpy(handle64, &mh, 64);
	return 0;
}

int fvg_halo_connect(fvg_halo *h, const void *handles, const int *all_recv_counts)
{
	if(!h || !handles || !all_recv_counts) { set_error("fvg_halo_connect: null argument"); return FVG_ERR_INVALID; }
	if(h->connected) { set_error("fvg_halo_connect: already connected"); return FVG_ERR_INVALID; }
	FVG_CUDA(cudaSetDevice(h->mesh->device));
	const int n = h->nranks;
	h->peer.assign(n, nullptr); h->peer_row0.assign(n, 0); h->peer_area.assign(n, 0);
	for(int r = 0; r < n; r++) {
		long long ng = 0;
		for(int q = 0; q < n; q++) ng += all_recv_counts[(size_t)r*n + q];
		h->peer_area[r] = (unsigned long long)std::max(ng, 1ll)*h->max_width;
		// my block inside peer r's ghost range starts after the blocks of the ranks below me
		int off = 0;
		for(int q = 0; q < h->rank; q++) off += all_recv_counts[(size_t)r*n + q];
		h->peer_row0[r] = off;
		if(r == h->rank) continue;
		if(all_recv_counts[(size_t)r*n + h->rank] != h->mesh->send_counts[r]) {
			set_error("fvg_halo_connect: send/receive counts of two ranks disagree"); return FVG_ERR_COMM;
		}
		if(h->mesh->send_counts[r] == 0) continue;
		cudaIpcMemHandle_t mh;
		std::memcpy(&mh, static_cast<const unsigned char*>(handles) + 64*(size_t)r, 64);
		void *p = nullptr;
		const cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
		if(e != cudaSuccess) { cuda_fail(e, "cudaIpcOpenMemHandle", __FILE__, __LINE__); return FVG_ERR_COMM; }
		h->peer[r] = static_cast<unsigned char*>(p);
	}
	FVG_CUDA(cudaMalloc((void**)&h->d_peer, sizeof(unsigned char*)*n));
	FVG_CUDA(cudaMalloc((void**)&h->d_send_off, sizeof(int)*(n+1)));
	FVG_CUDA(cudaMalloc((void**)&h->d_recv_off, sizeof(int)*(n+1)));
	FVG_CUDA(cudaMalloc((void**)&h->d_peer_row0, sizeof(int)*n));
	FVG_CUDA(cudaMalloc((void**)&h->d_arrive, sizeof(unsigned)*n));
	FVG_CUDA(cudaMalloc((void**)&h->d_peer_area, sizeof(unsigned long long)*n));
	FVG_CUDA(cudaMemcpy(h->d_peer_area, h->peer_area.data(), sizeof(unsigned long long)*n, cudaMemcpyHostToDevice));
	FVG_CUDA(cudaMemcpy(h->d_peer, h->peer.data(), sizeof(unsigned char*)*n, cudaMemcpyHostToDevice));
	FVG_CUDA(cudaMemcpy(h->d_send_off, h->send_off.data(), sizeof(int)*(n+1), cudaMemcpyHostToDevice));
	FVG_CUDA(cudaMemcpy(h->d_recv_off, h->recv_off.data(), sizeof(int)*(n+1), cudaMemcpyHostToDevice));
	FVG_CUDA(cudaMemcpy(h->d_peer_row0, h->peer_row0.data(), sizeof(int)*n, cudaMemcpyHostToDevice));
	FVG_CUDA(cudaMemset(h->d_arrive, 0, sizeof(unsigned)*n));
	h->connected = true;
	return 0;
}

static int halo_launch(fvg_halo *h, const double *src, double *dst, int width, bool do_send, bool do_recv, void *stream, const char *who)
{
	if(!h || (do_send && !src) || (do_recv && !dst) || !h->connected) { set_error(std::string(who) + ": not connected / null argument"); return FVG_ERR_INVALID; }
	if(width < 2 || (width & 1) || width > h->max_width) { set_error(std::string(who) + ": width must be even and within the window's width"); return FVG_ERR_INVALID; }
	if(do_send) h->seq++;
	HaloArgs a;
	a.src = src; a.dst = dst; a.send_idx = h->mesh->d.send_idx; a.send_off = h->d_send_off; a.peer_row0 = h->d_peer_row0;
	a.recv_off = h->d_recv_off; a.peer = h->d_peer; a.arrive = h->d_arrive; a.peer_area = h->d_peer_area; a.window = h->window;
	a.width = width; a.rank = h->rank; a.nranks = h->nranks; a.ncell = h->mesh->d.ncell; a.nghost = h->mesh->d.nghost;
	a.hdr = h->hdr; a.area_doubles = h->area_doubles; a.seq = h->seq;
	a.spin_limit = 30000000ll;                 // about two seconds of 64 ns naps
	a.nsend_blocks = (do_send && h->mesh->d.nsend > 0) ? h->nranks*HALO_CTAS_PER_PEER : 0;
	a.nrecv_blocks = (do_recv && a.nghost > 0) ? std::max(1, std::min(16, (a.nghost*width/2 + HALO_THREADS - 1)/HALO_THREADS)) : 0;
	if(a.nsend_blocks + a.nrecv_blocks == 0) return 0;
	halo_kernel<<<a.nsend_blocks + a.nrecv_blocks, HALO_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "halo kernel launch", __FILE__, __LINE__);
	return 0;
}

int fvg_halo_send(fvg_halo *h, const double *d_arr, int width, void *stream)
{
	return halo_launch(h, d_arr, nullptr, width, true, false, stream, "fvg_halo_send");
}

int fvg_halo_post(fvg_halo *h, const double *d_arr, int width, void *stream, unsigned long long *token)
{
	if(!token) { set_error("fvg_halo_post: null argument"); return FVG_ERR_INVALID; }
	const int rc = halo_launch(h, d_arr, nullptr, width, true, false, stream, "fvg_halo_post");
	if(rc == 0) *token = h->seq;
	return rc;
}

int fvg_halo_ghost_source(fvg_halo *h, unsigned long long token, fvg::GhostSrc *out)
{
	if(!h || !out || !h->connected || token == 0 || token > h->seq || token + HALO_NBUF <= h->seq + 1) {
		set_error("fvg_halo_ghost_source: no such exchange in the window any more"); return FVG_ERR_INVALID;
	}
	out->rows = reinterpret_cast<const double*>(h->window + h->hdr) + (token % HALO_NBUF)*h->area_doubles;
	out->flags = reinterpret_cast<const unsigned long long*>(h->window);
	out->recv_off = h->d_recv_off;
	out->seq = token;
	out->nranks = h->nranks;
	return 0;
}

int fvg_halo_recv(fvg_halo *h, double *d_arr, int width, void *stream)
{
	return halo_launch(h, nullptr, d_arr, width, false, true, stream, "fvg_halo_recv");
}

int fvg_halo_exchange(fvg_halo *h, double *d_arr, int width, void *stream)
{
	return halo_launch(h, d_arr, d_arr, width, true, true, stream, "fvg_halo_exchange");
}

int fvg_halo_status(fvg_halo *h, unsigned long long *h_timed_out_seq)
{
	if(!h || !h_timed_out_seq) { set_error("fvg_halo_status: null argument"); return FVG_ERR_INVALID; }
	FVG_CUDA(cudaMemcpy(h_timed_out_seq, h->window + (size_t)h->nranks*sizeof(unsigned long long),
	                    sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	return 0;
}

void fvg_halo_destroy(fvg_halo *h)
{
	if(!h) return;
	for(unsigned char *p : h->peer) if(p) cudaIpcCloseMemHandle(p);
	cudaFree(h->d_peer); cudaFree(h->d_send_off); cudaFree(h->d_recv_off); cudaFree(h->d_peer_row0); cudaFree(h->d_arrive); cudaFree(h->d_peer_area);
	cudaFree(h->window);
	delete h;
}

} // extern "C"
