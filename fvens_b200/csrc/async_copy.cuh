/* Asynchronous global->shared staging primitives for sm_100a: 1-D bulk copies through the TMA unit
 * (cp.async.bulk, completion counted on an mbarrier; SASS: UBLKCP) for contiguous ranges, and 16-byte
 * cp.async (SASS: LDGSTS) for gathered rows. Neither occupies registers while in flight, which is the
 * point: the tile kernels issue all of a tile's traffic up front and compute out of shared memory.
 */
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace fvg {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
	// make the initialised barrier visible to the async (TMA) proxy
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

/// Spins (with back-off inside try_wait) until the barrier's phase with the given parity completes.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"FVG_WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra FVG_DONE_%=;\n"
		"bra FVG_WAIT_%=;\n"
		"FVG_DONE_%=:\n"
		"}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

/// Orders this thread's earlier generic-proxy accesses to shared memory (made visible to it by a barrier) before
/// the async-proxy writes of bulk copies it issues next: needed when a staging buffer is recycled.
__device__ __forceinline__ void fence_proxy_async() {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/// Contiguous copy of `bytes` (multiple of 16, both addresses 16-byte aligned) global -> shared.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/// 2-D tiled copy through a tensor map (box rows starting at row `row`, all columns) global -> shared. The map's
/// swizzle mode decides where the 16-byte chunks of a row land (see rows32_chunk / rows64_chunk in face_kernel.cuh);
/// rows past the end of the global array arrive as zeros and are counted in the transaction bytes.
__device__ __forceinline__ void tensor_rows_g2s(void *dst_smem, const void *tmap, int row, uint64_t *bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
	             :: "r"(smem_u32(dst_smem)), "l"(tmap), "r"(0), "r"(row), "r"(smem_u32(bar)) : "memory");
}

/// Asks the TMA unit to pull `bytes` (multiple of 16, 16-byte aligned) of global memory into L2; no destination,
/// no completion tracking. Used to warm L2 with the operands of the tile that will run a wave later.
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, unsigned bytes) {
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src_gmem), "r"(bytes) : "memory");
}

/// 16-byte asynchronous copy global -> shared (L2 only: the data is consumed from shared memory)
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}


// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start while
// its predecessor in the stream is still running; it must not touch what the predecessor writes (or overwrite what it
// reads) before pdl_wait(), which returns once the predecessor has completed and its writes are visible.
// pdl_launch_dependents() lets the successor's CTAs be scheduled as soon as every CTA of this grid has started. Without
// the launch attribute (FVG_PDL=0) both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

} // namespace fvg
