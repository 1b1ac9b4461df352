/* Fused multi-GPU evaluation: FlowFV::compute_residual and SteadyForwardEulerSolver::solve on a mesh partitioned over
 * the GPUs of one NVLink / NVSwitch box, one process per GPU. Reference: the ghost updates and trace exchanges of
 * src/spatial/flow_spatial.cpp:711-788 and src/linalg/tracevector.cpp:214-325, the pseudo-time loop with its ghost
 * update and MPI_Allreduce of src/ode/aodesolver.cpp:136-282 (:212, :227).
 *
 * A residual evaluation is the same TWO kernels as on one GPU (three with WENO): the producing kernels push the rows
 * the neighbours need into their peer-mapped windows and the consuming kernels wait for them tile by tile
 * (dist_dev.cuh); the tiles that see no ghost cell run first, so the NVLink latency hides behind them. The evaluation
 * number lives on the device, so a whole evaluation is captured once into a CUDA graph and replayed. The norm of the
 * pseudo-time step is reduced through the same windows: every rank stores its partial sum into every rank's window and
 * all of them add the N values in rank order - no collective library, and the same bits on every rank.
 */
#include "engine.hpp"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace fvg;

struct fvg_dist {
	fvg_flow *flow = nullptr;
	fvg_mesh *mesh = nullptr;
	int nranks = 1, rank = 0;
	unsigned char *window = nullptr;
	size_t window_bytes = 0;
	std::vector<unsigned char*> peer;
	DistDev h{};                       ///< host copy of the device descriptor
	DistDev *d_dev = nullptr;
	DistCtl *d_ctl = nullptr;
	int *d_tsoff = nullptr, *d_tsend = nullptr;
	unsigned *d_tpeers = nullptr;
	bool connected = false;
	const double *pushed_ptr = nullptr;   ///< array whose send rows a step epilogue has pushed as the next evaluation's state
	bool use_graph = true;
	double *d_u2 = nullptr, *d_u1 = nullptr, *d_hist = nullptr;
	int hist_cap = 0;
	double *d_norm2 = nullptr;         ///< [1] global norm^2 of the most recently gathered step
	struct Graph { int kind; const void *a, *b, *c; int f0, f1, f2; double x; cudaStream_t s; cudaGraphExec_t exec; long long launches; };
	cudaStream_t own_stream = nullptr;   ///< stream of fvg_dist_forward_euler_solve (graphs cannot be captured on the default stream)
	std::vector<Graph> graphs;           ///< most recently used last; at most graph_cap entries (FVG_GRAPH_CACHE)
	size_t graph_cap = 32;
	long long evaluations = 0, graph_replays = 0, graph_evictions = 0;
};

namespace fvg {

/** One CTA. (1) Sums this rank's per-(tile, warp) partials of r_E^2 * area in index order, (2) stores the sum into
 * slot [step mod 4][rank] of every rank's window and releases norm_flag[rank] = step + 1 there, (3) if gather_lag >= 0:
 * waits for the partial sums of step (step - gather_lag) from all ranks, adds them in rank order and writes the total to
 * out[0] and, if hist is given, to hist[that step]. lag 1 (pseudo-time loop) gathers the previous step, whose values have
 * long arrived, so no rank ever waits for a slower one here; lag 0 is the immediate all-reduce.
 * push == 0: gather only (the loop's last step). The step counter is device-resident (graph replay). */
__global__ void __launch_bounds__(1024)
dist_norm_kernel(const DistDev *d, const double *__restrict__ partial, int n, int push, int gather_lag,
                 double *__restrict__ out, double *__restrict__ hist)
{
	__shared__ double s[1024];
	__shared__ double vals[MAXRANKS];
	const int tid = threadIdx.x;
	const unsigned long long step = d->ctl->step;
	const long long hist0 = (long long)d->ctl->hist0;      // hist[0] belongs to this step (start of the caller's loop)
	if(push) {
		double acc = 0.0;
		for(int k = tid; k < n; k += 1024) acc += partial[k];
		s[tid] = acc;
		__syncthreads();
		for(int o = 512; o > 0; o >>= 1) {
			if(tid < o) s[tid] += s[tid + o];
			__syncthreads();
		}
		if(tid < d->nranks) {
			unsigned char *const w = tid == d->rank ? d->window : d->peer[tid];
			WinHdr *const W = reinterpret_cast<WinHdr*>(w);
			W->norm_val[step % NORM_SLOTS][d->rank] = s[0];
			__threadfence_system();
			st_release_sys_u64(&W->norm_flag[d->rank], step + 1);
		}
	}
	const long long g = (long long)step - (push ? gather_lag : 1);
	if(gather_lag >= 0 && g >= 0 && (!hist || g >= hist0)) {
		const WinHdr *const W = reinterpret_cast<const WinHdr*>(d->window);
		if(tid < d->nranks) {
			const unsigned long long *const f = &W->norm_flag[tid];
			const unsigned long long t0 = global_timer_ns();
			unsigned spins = 0;
			while(ld_acquire_sys_u64(f) < (unsigned long long)g + 1) {
				__nanosleep(32);
				if((++spins & 255u) == 0 && (long long)(global_timer_ns() - t0) > d->spin_ns) { atomicMax(&d->ctl->timeout, (unsigned long long)g + 1); break; }
			}
			vals[tid] = *reinterpret_cast<const volatile double*>(&W->norm_val[g % NORM_SLOTS][tid]);
		}
		__syncthreads();
		if(tid == 0) {
			double tot = 0.0;
			for(int r = 0; r < d->nranks; r++) tot += vals[r];
			if(out) out[0] = tot;
			if(hist && g >= hist0) hist[g - hist0] = tot;
		}
	}
	if(tid == 0 && push) d->ctl->step = step + 1;
}

} // namespace fvg

static int dist_fail(const std::string &msg, int code) { set_error(msg); return code; }

/// roles of the kernels of one evaluation for this flow's numerics
static void set_roles(fvg_dist *D, bool step, bool force_push)
{
	fvg_flow *f = D->flow;
	const FlowPlan &P = f->plan;
	fvg_flow::DistRoles R;
	R.active = true;
	const DistDev *dev = D->d_dev;
	const unsigned U = 1u << X_U, GU = 1u << X_GU, LG = 1u << X_LG;
	const bool weno = P.order2 && P.recon == FVG_RECON_WENO, muscl = P.order2 && P.recon == FVG_RECON_VANALBADA;
	const bool limited = P.recon == FVG_RECON_BARTHJESPERSEN || P.recon == FVG_RECON_VENKATAKRISHNAN;
	const bool visc = P.visc != VISC_NONE;
	R.cell.d = R.weno.d = R.face.d = dev;
	R.cell.ctl = R.weno.ctl = R.face.ctl = D->d_ctl;
	for(int t = 0; t < X_COUNT; t++)
		for(int par = 0; par < 2; par++) {
			const double *area = reinterpret_cast<const double*>(D->window + sizeof(WinHdr)) + xarea_off(t, par, (size_t)D->mesh->d.nghost);
			R.cell.ghost[t][par] = R.weno.ghost[t][par] = R.face.ghost[t][par] = area;
		}
	R.face.wait = U; R.face.last = 1; R.face.push = step ? U : 0u; R.face.force_push = force_push ? 1 : 0;
	R.face.visc_type = X_LG;
	if(!P.order2) R.face.first = 1;
	else {
		R.cell.first = 1; R.cell.wait = U; R.cell.force_push = force_push ? 1 : 0;
		if(weno) {
			R.cell.push = GU;
			R.weno.wait = GU; R.weno.push = LG;
			R.face.wait |= LG;
			if(visc) { R.face.wait |= GU; R.face.visc_type = X_GU; }
		} else if(muscl) {
			R.cell.push = GU;
			R.face.wait |= GU;
			R.face.visc_type = X_GU;
		} else {
			R.cell.push = LG;
			R.face.wait |= LG;
			if(visc && limited) { R.cell.push |= GU; R.face.wait |= GU; R.face.visc_type = X_GU; }
		}
	}
	f->roles = R;
}

/// launches of one evaluation (residual, or the fused step when unew is given)
static int enqueue_evaluation(fvg_dist *D, const double *u, double *res, int accumulate, int gettimesteps, double *dtm,
                              double cfl, double *unew, bool force_push, cudaStream_t s)
{
	fvg_flow *f = D->flow;
	// single-rank periodic mesh renumbered by the engine: caller-ordered arrays, permutation fused into the passes
	// (the fused step always runs on device-ordered ping-pong buffers)
	const int *perm = (D->mesh->identity_perm || unew) ? nullptr : D->mesh->d.new2old;
	if(perm && !f->d_uperm) { const int ra = flow_dev_alloc(f, &f->d_uperm, 4*(size_t)D->mesh->d.ncell); if(ra != 0) return ra; }
	set_roles(D, unew != nullptr, force_push);
	int rc = 0;
	const double *uface = u;
	// per-pass timing (fvg_flow_timing): events on the launching stream, direct launches only (not inside a captured graph)
	cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
	cudaStreamIsCapturing(s, &cap);
	const bool timed = f->timing && cap == cudaStreamCaptureStatusNone;
	auto mark = [&]() { if(timed) { cudaEvent_t e; if(cudaEventCreate(&e) == cudaSuccess) { f->ev.push_back(e); cudaEventRecord(e, s); } } };
	mark();
	if(perm) {
		if(f->plan.order2) rc = run_gradient_pass(f, u, s, 0, -1, perm, f->d_uperm);
		else { rc = launch_permute_rows(u, f->d_uperm, perm, D->mesh->d.ncell, 4, true, false, s); f->launches++; }
		uface = f->d_uperm;
	}
	else rc = run_gradient_pass(f, u, s);
	mark();
	if(rc == 0) rc = unew ? run_face_pass(f, uface, EP_STEP, 0, 1, nullptr, nullptr, cfl, unew, s)
	                      : run_face_pass(f, uface, EP_RESIDUAL, accumulate, gettimesteps, res, dtm, 0.0, nullptr, s, 0, -1, perm);
	mark();
	f->roles = fvg_flow::DistRoles();
	return rc;
}

static int enqueue_norm(fvg_dist *D, int push, int lag, double *out, double *hist, cudaStream_t s)
{
	fvg_flow *f = D->flow;
	dist_norm_kernel<<<1, 1024, 0, s>>>(D->d_dev, f->d_partial, f->mesh->d.ntile*(FACE_BLOCK/32), push, lag, out, hist);
	const cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return cuda_fail(e, "dist_norm_kernel launch", __FILE__, __LINE__);
	f->launches++;
	return 0;
}

/// Runs `body` (kernel launches on stream s) through a CUDA graph captured on first use for this key and replayed
/// afterwards. The legacy default stream cannot be captured: there (or with graphs switched off) the launches are direct.
template <typename Body>
static int run_graphed(fvg_dist *D, const fvg_dist::Graph &key, cudaStream_t s, Body body)
{
	if(!D->use_graph || D->flow->timing || s == nullptr || s == cudaStreamLegacy) return body();
	for(size_t i = 0; i < D->graphs.size(); i++) {
		const fvg_dist::Graph &g = D->graphs[i];
		if(g.kind == key.kind && g.a == key.a && g.b == key.b && g.c == key.c && g.f0 == key.f0 && g.f1 == key.f1 && g.f2 == key.f2 &&
		   g.x == key.x && g.s == key.s) {
			FVG_CUDA(cudaGraphLaunch(g.exec, s));
			D->graph_replays++;
			D->flow->launches += g.launches;
			if(i + 1 != D->graphs.size()) std::rotate(D->graphs.begin() + (long)i, D->graphs.begin() + (long)i + 1, D->graphs.end());
			return 0;
		}
	}
	const long long launches0 = D->flow->launches;
	cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
	if(e != cudaSuccess) { cudaGetLastError(); return body(); }      // e.g. the stream is already being captured by the caller
	const int rc = body();
	cudaGraph_t graph = nullptr;
	e = cudaStreamEndCapture(s, &graph);
	if(rc != 0) { if(graph) cudaGraphDestroy(graph); return rc; }
	if(e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
	fvg_dist::Graph g = key;
	e = cudaGraphInstantiate(&g.exec, graph, 0);
	cudaGraphDestroy(graph);
	if(e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
	g.launches = D->flow->launches - launches0;      // kernels per replay
	// a caller that hands in new arrays every call would otherwise grow the cache without bound: the least recently used
	// graph goes (destroying an executable graph whose launch is still in flight is allowed: the work completes first)
	while(D->graphs.size() >= std::max<size_t>(D->graph_cap, 1)) {
		cudaGraphExecDestroy(D->graphs.front().exec);
		D->graphs.erase(D->graphs.begin());
		D->graph_evictions++;
	}
	D->graphs.push_back(g);
	FVG_CUDA(cudaGraphLaunch(g.exec, s));
	return 0;
}

extern "C" {

int fvg_dist_create(fvg_flow *flow, fvg_dist **out)
{
	if(!flow || !out) return dist_fail("fvg_dist_create: null argument", FVG_ERR_INVALID);
	*out = nullptr;
	fvg_mesh *m = flow->mesh;
	if(m->device < 0) return dist_fail("fvg_dist_create: host-only mesh", FVG_ERR_INVALID);
	if(m->nranks > MAXRANKS) return dist_fail("fvg_dist_create: at most 16 ranks per box", FVG_ERR_UNSUPPORTED);
	if(!m->identity_perm && m->nranks > 1) return dist_fail("fvg_dist_create: subdomain meshes are device-ordered", FVG_ERR_INVALID);
	FVG_CUDA(cudaSetDevice(m->device));
	// (a failed allocation below leaves through FVG_CUDA: the deleter frees what had been allocated)
	std::unique_ptr<fvg_dist, void(*)(fvg_dist*)> D(new fvg_dist, fvg_dist_destroy);
	D->flow = flow; D->mesh = m; D->nranks = m->nranks; D->rank = m->rank;
	D->window_bytes = sizeof(WinHdr) + (size_t)std::max(m->d.nghost, 1)*XAREA_DOUBLES_PER_GHOST*sizeof(double);
	FVG_CUDA(cudaMalloc((void**)&D->window, D->window_bytes));
	FVG_CUDA(cudaMemset(D->window, 0, D->window_bytes));
	FVG_CUDA(cudaMalloc((void**)&D->d_ctl, sizeof(DistCtl)));
	FVG_CUDA(cudaMemset(D->d_ctl, 0, sizeof(DistCtl)));
	FVG_CUDA(cudaMalloc((void**)&D->d_dev, sizeof(DistDev)));
	FVG_CUDA(cudaMalloc((void**)&D->d_norm2, sizeof(double)));
	if(const char *e = getenv("FVG_GRAPH")) D->use_graph = e[0] != '0';
	if(const char *e = getenv("FVG_GRAPH_CACHE")) D->graph_cap = (size_t)std::max(1, atoi(e));
	FVG_CUDA(cudaDeviceSynchronize());
	*out = D.release();
	return 0;
}

int fvg_dist_ipc_handle(fvg_dist *D, void *handle64)
{
	if(!D || !handle64) return dist_fail("fvg_dist_ipc_handle: null argument", FVG_ERR_INVALID);
	cudaIpcMemHandle_t mh;
	FVG_CUDA(cudaIpcGetMemHandle(&mh, D->window));
	std::memcpy(handle64, &mh, 64);
	return 0;
}

int fvg_dist_connect(fvg_dist *D, const void *handles, const int *all_recv_counts)
{
	if(!D || !handles || !all_recv_counts) return dist_fail("fvg_dist_connect: null argument", FVG_ERR_INVALID);
	if(D->connected) return dist_fail("fvg_dist_connect: already connected", FVG_ERR_INVALID);
	fvg_mesh *m = D->mesh;
	FVG_CUDA(cudaSetDevice(m->device));
	const int n = D->nranks;
	DistDev &H = D->h;
	std::memset(&H, 0, sizeof(H));
	H.nranks = n; H.rank = D->rank; H.nghost = m->d.nghost; H.nsend = m->d.nsend;
	H.window = D->window;
	D->peer.assign(n, nullptr);
	for(int r = 0; r < n; r++) {
		long long ng = 0;
		for(int q = 0; q < n; q++) ng += all_recv_counts[(size_t)r*n + q];
		H.peer_nghost[r] = (int)ng;
		int off = 0;
		for(int q = 0; q < D->rank; q++) off += all_recv_counts[(size_t)r*n + q];
		H.peer_row0[r] = off;
		H.send_off[r+1] = H.send_off[r] + m->send_counts[r];
		H.recv_off[r+1] = H.recv_off[r] + m->recv_counts[r];
		if(r == D->rank) { H.peer[r] = D->window; continue; }       // (periodic rows of this rank's own cells)
		if(all_recv_counts[(size_t)r*n + D->rank] != m->send_counts[r])
			return dist_fail("fvg_dist_connect: send/receive counts of two ranks disagree", FVG_ERR_COMM);
		// every rank is mapped (not only the halo neighbours): the norm reduction stores into all windows
		cudaIpcMemHandle_t mh;
		std::memcpy(&mh, static_cast<const unsigned char*>(handles) + 64*(size_t)r, 64);
		void *p = nullptr;
		const cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
		if(e != cudaSuccess) { cuda_fail(e, "cudaIpcOpenMemHandle", __FILE__, __LINE__); return FVG_ERR_COMM; }
		D->peer[r] = static_cast<unsigned char*>(p);
		H.peer[r] = D->peer[r];
	}
	// per-tile send lists and peer masks on the device
	const int ntile = m->d.ntile;
	std::vector<unsigned> tpeers((size_t)ntile, 0u);
	std::vector<int> ntile_send(n, 0);
	for(int t = 0; t < ntile; t++) {
		for(int q = m->h_tsoff[t]; q < m->h_tsoff[t+1]; q++) tpeers[t] |= 1u << m->h_tsend[3*(size_t)q + 1];
		for(int r = 0; r < n; r++) if((tpeers[t] >> r) & 1u) ntile_send[r]++;
	}
	for(int r = 0; r < n; r++) H.ntile_send[r] = ntile_send[r];
	auto up = [&](const void *src, size_t bytes, void **dst) -> int {
		FVG_CUDA(cudaMalloc(dst, std::max<size_t>(bytes, 8)));
		if(bytes) FVG_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
		return 0;
	};
	int rc;
	if((rc = up(m->h_tsoff.data(), sizeof(int)*m->h_tsoff.size(), (void**)&D->d_tsoff)) != 0) return rc;
	if((rc = up(m->h_tsend.data(), sizeof(int)*m->h_tsend.size(), (void**)&D->d_tsend)) != 0) return rc;
	if((rc = up(tpeers.data(), sizeof(unsigned)*tpeers.size(), (void**)&D->d_tpeers)) != 0) return rc;
	H.tsoff = D->d_tsoff; H.tsend = D->d_tsend; H.tpeers = D->d_tpeers; H.send_idx = m->d.send_idx;
	H.ctl = D->d_ctl;
	double timeout_ms = 20000.0;
	if(const char *e = getenv("FVG_HALO_TIMEOUT_MS")) timeout_ms = atof(e);
	H.spin_ns = (long long)(timeout_ms*1e6);
	FVG_CUDA(cudaMemcpy(D->d_dev, &H, sizeof(H), cudaMemcpyHostToDevice));
	D->connected = true;
	return 0;
}

static int check_ready(fvg_dist *D, const char *who)
{
	if(!D) return dist_fail(std::string(who) + ": null argument", FVG_ERR_INVALID);
	if(!D->connected) return dist_fail(std::string(who) + ": fvg_dist_connect has not been called", FVG_ERR_INVALID);
	return 0;
}

int fvg_dist_residual(fvg_dist *D, const double *d_u, double *d_res, int accumulate, int gettimesteps, double *d_dtm, void *stream)
{
	int rc = check_ready(D, "fvg_dist_residual");
	if(rc != 0) return rc;
	if(!d_u || !d_res || (gettimesteps && !d_dtm)) return dist_fail("fvg_dist_residual: null argument", FVG_ERR_INVALID);
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const bool force = d_u != D->pushed_ptr;
	fvg_dist::Graph key{0, d_u, d_res, d_dtm, accumulate, gettimesteps, force ? 1 : 0, 0.0, s, nullptr, 0};
	rc = run_graphed(D, key, s, [&]() { return enqueue_evaluation(D, d_u, d_res, accumulate, gettimesteps, d_dtm, 0.0, nullptr, force, s); });
	D->evaluations++;
	return rc;
}

int fvg_dist_euler_step(fvg_dist *D, const double *d_u, double *d_unew, double cfl, double *d_resnorm2, void *stream)
{
	int rc = check_ready(D, "fvg_dist_euler_step");
	if(rc != 0) return rc;
	if(!d_u || !d_unew || d_u == d_unew) return dist_fail("fvg_dist_euler_step: the step needs two distinct state arrays", FVG_ERR_INVALID);
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const bool force = d_u != D->pushed_ptr;
	fvg_dist::Graph key{1, d_u, d_unew, d_resnorm2, 0, 0, force ? 1 : 0, cfl, s, nullptr, 0};
	rc = run_graphed(D, key, s, [&]() {
		int r = enqueue_evaluation(D, d_u, nullptr, 0, 1, nullptr, cfl, d_unew, force, s);
		// immediate reduction over the ranks: the caller's scalar is this step's global sum
		if(r == 0) r = enqueue_norm(D, 1, d_resnorm2 ? 0 : -1, d_resnorm2, nullptr, s);
		return r;
	});
	D->pushed_ptr = d_unew;
	D->evaluations++;
	return rc;
}

int fvg_dist_invalidate_state(fvg_dist *D)
{
	if(!D) return dist_fail("fvg_dist_invalidate_state: null argument", FVG_ERR_INVALID);
	D->pushed_ptr = nullptr;
	return 0;
}

int fvg_dist_status(fvg_dist *D, unsigned long long *h_timed_out)
{
	if(!D || !h_timed_out) return dist_fail("fvg_dist_status: null argument", FVG_ERR_INVALID);
	DistCtl c;
	FVG_CUDA(cudaMemcpy(&c, D->d_ctl, sizeof(c), cudaMemcpyDeviceToHost));
	*h_timed_out = c.timeout;
	if(c.timeout != 0) return dist_fail("a halo wait timed out: a neighbour rank did not deliver its rows of evaluation " +
	                                    std::to_string(c.timeout - 1) + " (results since then are not valid)", FVG_ERR_COMM);
	return 0;
}

int fvg_dist_counters(fvg_dist *D, long long *evaluations, long long *graph_replays, unsigned long long *device_evaluation)
{
	if(!D) return dist_fail("fvg_dist_counters: null argument", FVG_ERR_INVALID);
	if(evaluations) *evaluations = D->evaluations;
	if(graph_replays) *graph_replays = D->graph_replays;
	if(device_evaluation) {
		DistCtl c;
		FVG_CUDA(cudaMemcpy(&c, D->d_ctl, sizeof(c), cudaMemcpyDeviceToHost));
		*device_evaluation = c.k;
	}
	return 0;
}

int fvg_dist_forward_euler_solve(fvg_dist *D, double *d_u, double cfl, double tol, int maxiter, int check_every,
                                 int *h_steps, double *h_hist)
{
	int rc = check_ready(D, "fvg_dist_forward_euler_solve");
	if(rc != 0) return rc;
	if(!d_u || !h_steps) return dist_fail("fvg_dist_forward_euler_solve: null argument", FVG_ERR_INVALID);
	if(check_every < 1) check_every = 1;
	*h_steps = 0;
	if(maxiter <= 0) return 0;
	fvg_flow *f = D->flow;
	FVG_CUDA(cudaSetDevice(D->mesh->device));
	const size_t nrow = (size_t)D->mesh->d.ncell + (size_t)D->mesh->d.nghost, nown = (size_t)D->mesh->d.ncell;
	if(!D->d_u1 && (rc = flow_dev_alloc(f, &D->d_u1, 4*nrow)) != 0) return rc;
	if(!D->d_u2 && (rc = flow_dev_alloc(f, &D->d_u2, 4*nrow)) != 0) return rc;
	if(D->hist_cap < maxiter) {
		if((rc = flow_dev_alloc(f, &D->d_hist, (size_t)maxiter)) != 0) return rc;
		D->hist_cap = maxiter;
	}
	cudaStream_t s = nullptr;
	if(D->use_graph) {
		if(!D->own_stream) FVG_CUDA(cudaStreamCreateWithFlags(&D->own_stream, cudaStreamNonBlocking));
		s = D->own_stream;
		FVG_CUDA(cudaDeviceSynchronize());       // earlier work of the caller on other streams
	}
	double *cur = D->d_u1, *nxt = D->d_u2;
	const int *perm = D->mesh->identity_perm ? nullptr : D->mesh->d.new2old;
	cudaError_t e = cudaSuccess;
	if(perm) { rc = launch_permute_rows(d_u, cur, perm, (int)nown, 4, true, false, s); f->launches++; }
	else {
		e = cudaMemcpyAsync(cur, d_u, 4*nown*sizeof(double), cudaMemcpyDeviceToDevice, s);
		if(e != cudaSuccess) rc = cuda_fail(e, "state copy", __FILE__, __LINE__);
	}
	D->pushed_ptr = nullptr;
	std::vector<double> hist((size_t)maxiter);
	int step = 0, status = FVG_OK;
	double initres = 1.0;
	if(rc == 0) {
		// the device's step counter numbers the norm slots; entry 0 of the history belongs to its current value
		DistCtl c;
		e = cudaMemcpy(&c, D->d_ctl, sizeof(c), cudaMemcpyDeviceToHost);
		if(e == cudaSuccess) e = cudaMemcpy(&D->d_ctl->hist0, &c.step, sizeof(c.step), cudaMemcpyHostToDevice);
		if(e != cudaSuccess) rc = cuda_fail(e, "control words", __FILE__, __LINE__);
	}
	double *const histbase = D->d_hist;
	while(rc == 0) {
		const int batch = std::min(check_every, maxiter - step);
		for(int k = 0; k < batch && rc == 0; k++) {
			const bool force = cur != D->pushed_ptr;
			fvg_dist::Graph key{2, cur, nxt, D->d_hist, 0, 0, force ? 1 : 0, cfl, s, nullptr, 0};
			rc = run_graphed(D, key, s, [&]() {
				int r = enqueue_evaluation(D, cur, nullptr, 0, 1, nullptr, cfl, nxt, force, s);
				// deferred gather: this step's kernel adds up the PREVIOUS step's partial sums (all long there)
				if(r == 0) r = enqueue_norm(D, 1, 1, nullptr, histbase, s);
				return r;
			});
			D->pushed_ptr = nxt;
			D->evaluations++;
			std::swap(cur, nxt);
		}
		if(rc != 0) break;
		// the batch's last step is gathered by a reduction-only launch
		rc = enqueue_norm(D, 0, 1, nullptr, histbase, s);
		if(rc != 0) break;
		e = cudaMemcpyAsync(hist.data() + step, D->d_hist + step, sizeof(double)*batch, cudaMemcpyDeviceToHost, s);
		if(e == cudaSuccess) e = cudaStreamSynchronize(s);
		if(e != cudaSuccess) { rc = cuda_fail(e, "norm read-back", __FILE__, __LINE__); break; }
		unsigned long long to = 0;
		if((rc = fvg_dist_status(D, &to)) != 0) break;
		bool stop = false;
		for(int k = 0; k < batch; k++) {
			const double resi = std::sqrt(hist[step+k]);
			hist[step+k] = resi;
			if(step + k == 0) initres = resi;
			if(stop) continue;
			if(!std::isfinite(resi)) { status = FVG_ERR_NUMERICAL; stop = true; }
			else if(!(resi/initres > tol)) stop = true;
		}
		step += batch;
		if(step >= maxiter) { if(status == FVG_OK) status = FVG_ERR_TOLERANCE; stop = true; }
		if(stop) break;
	}
	if(rc == 0) {
		*h_steps = step;
		if(h_hist) std::memcpy(h_hist, hist.data(), sizeof(double)*step);
		if(perm) { rc = launch_permute_rows(cur, d_u, perm, (int)nown, 4, false, false, s); f->launches++; }
		else e = cudaMemcpyAsync(d_u, cur, 4*nown*sizeof(double), cudaMemcpyDeviceToDevice, s);
		if(e == cudaSuccess) e = cudaStreamSynchronize(s);
		if(rc == 0 && e != cudaSuccess) rc = cuda_fail(e, "state copy", __FILE__, __LINE__);
	}
	D->pushed_ptr = nullptr;     // the caller's array is not the one whose rows the neighbours hold
	if(rc != 0) return rc;
	if(status == FVG_ERR_NUMERICAL) set_error("forward Euler: residual norm is not finite");
	if(status == FVG_ERR_TOLERANCE) set_error("forward Euler: exceeded max iterations");
	return status;
}

/* test hook: copies the receive area of (type, parity) - [nghost][width] doubles - and the header's flags to the host */
int fvg_dist_debug_window(fvg_dist *D, int type, int parity, double *h_rows, unsigned long long *h_flags)
{
	if(!D || type < 0 || type >= X_COUNT) return dist_fail("fvg_dist_debug_window: bad argument", FVG_ERR_INVALID);
	FVG_CUDA(cudaDeviceSynchronize());
	const size_t ng = (size_t)D->mesh->d.nghost;
	if(h_rows && ng) FVG_CUDA(cudaMemcpy(h_rows, D->window + sizeof(WinHdr) + xarea_off(type, parity, ng)*sizeof(double),
	                                     ng*xwidth(type)*sizeof(double), cudaMemcpyDeviceToHost));
	if(h_flags) FVG_CUDA(cudaMemcpy(h_flags, D->window, sizeof(unsigned long long)*X_COUNT*MAXRANKS, cudaMemcpyDeviceToHost));
	return 0;
}

void fvg_dist_destroy(fvg_dist *D)
{
	if(!D) return;
	cudaDeviceSynchronize();
	for(const fvg_dist::Graph &g : D->graphs) if(g.exec) cudaGraphExecDestroy(g.exec);
	if(D->own_stream) cudaStreamDestroy(D->own_stream);
	for(unsigned char *p : D->peer) if(p) cudaIpcCloseMemHandle(p);
	cudaFree(D->d_tsoff); cudaFree(D->d_tsend); cudaFree(D->d_tpeers);
	cudaFree(D->d_dev); cudaFree(D->d_ctl); cudaFree(D->d_norm2);
	cudaFree(D->window);
	delete D;
}

} // extern "C"
